"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the authoring container only (the reference tree is not shipped to the GPU box):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference's util/fourier.py and util/resampling.py as they are (a stub
``soundfile`` module satisfies util/resampling.py:5; nothing of it is called), runs them on
small seeded inputs and stores inputs + outputs as .npz.  The reference has no tests or
golden vectors of its own for this path; these files are what pins the
oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

Back-ends exercised: util.fourier.np_rfft_pick (the only CPU back-end runnable here:
pyfftw is not installed, torch has no CUDA device in the container), util.fourier.istft,
util.resampling.speed_to_pos / sinc_wrapper / sinc_wrapper_mt (numba).
"""
import logging
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference(root):
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    sys.path.insert(0, root)
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    from util import fourier, resampling  # noqa: E402
    return fourier, resampling


def synth(n, seed, sr=44100.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = 0.25 * np.sin(2 * np.pi * 1000.0 * t) + 0.1 * np.sin(2 * np.pi * (sr / 4.3) * t) \
        + 0.05 * rng.standard_normal(n)
    return x.astype(np.float32)


def wow_curve(duration, sr, hop, depth=0.01, freq=0.5556):
    k = int(duration * sr / hop)
    times = np.linspace(0, duration, k)
    speeds = 1 + depth * np.sin(2 * np.pi * freq * times)
    return np.stack((times, speeds), -1)


def main(root):
    fourier, resampling = load_reference(root)
    from scipy import signal as dsp

    # ---- STFT (util/fourier.py:136-157 through np_rfft_pick) -------------------------
    stft_cases = [
        # name, L, n_fft, step, window, zeropad, seed
        ("fft1024_hop256", 5000, 1024, 256, "blackmanharris", 1, 1),
        ("fft512_hop32_hann", 3000, 512, 32, "hann", 1, 2),
        ("fft256_hop64_zp4", 2500, 256, 64, "blackmanharris", 4, 3),
        ("fft4096_hop1024", 6000, 4096, 1024, "blackmanharris", 1, 4),
        ("fft64_hop48_ragged", 1000, 64, 48, "hann", 1, 5),
        ("fft1024_short", 700, 1024, 256, "blackmanharris", 1, 6),
        ("fft2048_hop512_zp2", 5000, 2048, 512, "blackmanharris", 2, 7),
    ]
    out = {}
    for name, n, n_fft, step, wname, zp, seed in stft_cases:
        x = synth(n, seed)
        window = dsp.get_window(wname, n_fft).astype(np.float32)
        s = fourier.np_rfft_pick(n_fft, step, window, x, zp)
        out[f"{name}__x"] = x
        out[f"{name}__S"] = np.asarray(s)
        out[f"{name}__meta"] = np.array([n_fft, step, zp])
        out[f"{name}__window"] = np.array(wname)
    # strided column view of an interleaved (frames, channels) array, as the GUIs pass it
    inter = np.stack([synth(4000, 8), synth(4000, 9)], axis=1)
    window = dsp.get_window("blackmanharris", 1024).astype(np.float32)
    out["interleaved__x"] = inter
    out["interleaved__S"] = np.asarray(fourier.np_rfft_pick(1024, 256, window, inter[:, 1], 1))
    np.savez_compressed(os.path.join(HERE, "stft.npz"), **out)

    # ---- iSTFT (util/fourier.py:314-437) ----------------------------------------------
    out = {}
    for name, n, n_fft, hop, seed in [("rt512_32", 4000, 512, 32, 11), ("rt1024_256", 6000, 1024, 256, 12),
                                      ("rt4096_1024", 9000, 4096, 1024, 13)]:
        x = synth(n, seed)
        ypad = fourier.fix_length(x, n + n_fft // 2)
        window = dsp.get_window("blackmanharris", n_fft).astype(np.float32)
        s128 = np.asarray(fourier.np_rfft_pick(n_fft, hop, window, ypad, 1))
        s64 = s128.astype(np.complex64)
        y64 = fourier.istft(s64.copy(), length=n, hop_length=hop)        # float32 path
        y128 = fourier.istft(s128.copy(), length=n, hop_length=hop)      # float64 path
        out[f"{name}__x"] = x
        out[f"{name}__S"] = s64
        out[f"{name}__y_from_c64"] = y64
        out[f"{name}__y_from_c128"] = y128
        out[f"{name}__meta"] = np.array([n_fft, hop, n])
    # modified spectrum (gain mask), no length given
    x = synth(3000, 14)
    window = dsp.get_window("blackmanharris", 512).astype(np.float32)
    s = np.asarray(fourier.np_rfft_pick(512, 128, window, x, 1)).astype(np.complex64)
    s[40:90, 5:15] *= 0.25
    out["masked__S"] = s
    out["masked__y"] = fourier.istft(s.copy(), hop_length=128)
    out["masked__meta"] = np.array([512, 128, 0])
    np.savez_compressed(os.path.join(HERE, "istft.npz"), **out)

    # ---- speed_to_pos (util/resampling.py:93-137) ------------------------------------------
    out = {}
    sr = 44100
    curve = wow_curve(0.5, sr, 256)                     # 86 points, +-1 %
    n_in = 21900                                        # inside the curve's span: end test fires
    out["wow__sampletimes"] = curve[:, 0] * sr
    out["wow__speeds"] = curve[:, 1]
    out["wow__n_in"] = np.array(n_in)
    out["wow__pos"] = resampling.speed_to_pos(curve[:, 0] * sr, curve[:, 1], n_in)
    curve2 = wow_curve(0.5, sr, 256, depth=0.05, freq=3.0)
    n_in2 = 30000                                       # longer than the curve: end test never fires
    pos2 = resampling.speed_to_pos(curve2[:, 0] * sr, curve2[:, 1], n_in2)
    out["wow5_noend__sampletimes"] = curve2[:, 0] * sr
    out["wow5_noend__speeds"] = curve2[:, 1]
    out["wow5_noend__n_in"] = np.array(n_in2)
    out["wow5_noend__pos_full_len"] = np.array(len(pos2))   # includes the uninitialised tail
    out["wow5_noend__pos"] = pos2                            # compare on the filled prefix only
    st = np.array([0.0, 8000.0])
    sp = np.array([0.5, 2.0])                               # the ramp of test_sinc (:271-273)
    out["ramp__sampletimes"] = st
    out["ramp__speeds"] = sp
    out["ramp__n_in"] = np.array(8000)
    out["ramp__pos"] = resampling.speed_to_pos(st, sp, 8000)
    np.savez_compressed(os.path.join(HERE, "positions.npz"), **out)

    # ---- sinc_core / sinc_wrapper(_mt) (util/resampling.py:21-90) ---------------------------
    out = {}
    x = synth(n_in, 21, sr)
    pos = out_pos = resampling.speed_to_pos(curve[:, 0] * sr, curve[:, 1], n_in)
    for nt in (8, 50, 128):
        out[f"wow_nt{nt}__y"] = resampling.sinc_wrapper(pos, x, 0, nt)
    y_mt = np.empty(len(pos), np.float32)
    resampling.sinc_wrapper_mt(y_mt, pos, x, 0, 50)
    out["wow_nt50__y_mt"] = y_mt
    out["wow_nt50__mt_threads"] = np.array(os.cpu_count())
    out["wow__x"] = x
    out["wow__pos"] = out_pos
    xr = synth(8000, 22, sr)
    pr = resampling.speed_to_pos(st, sp, 8000)
    out["ramp__x"] = xr
    out["ramp__pos"] = pr
    out["ramp_nt50__y"] = resampling.sinc_wrapper(pr, xr, 0, 50)
    # positions running past the end of a short signal and starting inside the first NT samples
    xs = synth(600, 23, sr)
    ps = np.cumsum(np.full(700, 0.97)) - 0.4
    out["edges__x"] = xs
    out["edges__pos"] = ps
    out["edges_nt50__y"] = resampling.sinc_wrapper(ps, xs, 0, 50)
    out["edges_nt128__y"] = resampling.sinc_wrapper(ps, xs, 0, 128)
    # exact-integer positions (shift == 0 -> np.sinc(0) branch) and repeated positions (period clamp)
    pi_ = np.concatenate([np.arange(100, 400, dtype=np.float64), np.full(5, 400.0),
                          np.arange(400, 500, 2, dtype=np.float64)])
    out["integer__pos"] = pi_
    out["integer_nt50__y"] = resampling.sinc_wrapper(pi_, xs, 0, 50)
    np.savez_compressed(os.path.join(HERE, "sinc.npz"), **out)

    for f in ("stft", "istft", "positions", "sinc"):
        p = os.path.join(HERE, f + ".npz")
        print(f, os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
