"""GPU parity tests: the sm_100a kernels, called through the C ABI of libpar_b200.so (via the
Python mirror of the reference's module surface, and directly with device pointers), against the
CPU oracle on the committed golden inputs and on seeded synthetic inputs.

Bars (BASELINE.json north_star, SURVEY.md 8c "parity metric"):
  * float results: ||a-b||_2 / ||b||_2 <= 1e-6 per frame / per channel and
    max|a-b| / max|b| <= 1e-6 against the float64 oracle;
  * integer / index results and the float64 read positions: identical.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu

TOL = 1e-6


@pytest.fixture(scope="module")
def par():
    import pyaudiorestoration_b200 as p
    from pyaudiorestoration_b200 import _lib
    if not os.path.exists(p.library_path()):
        p.build()
    _lib.require_device()
    return p


@pytest.fixture(scope="module")
def fourier(par):
    from pyaudiorestoration_b200.util import fourier
    return fourier


@pytest.fixture(scope="module")
def resampling(par):
    from pyaudiorestoration_b200.util import resampling
    return resampling


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def rel_l2(a, b, axis=None):
    num = np.sqrt(np.sum(np.abs(a - b) ** 2, axis=axis))
    den = np.maximum(np.sqrt(np.sum(np.abs(b) ** 2, axis=axis)), 1e-300)
    return num / den


def rel_max(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def synth(n, seed, sr=96000.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = 0.25 * np.sin(2 * np.pi * 1000.0 * t) + 0.1 * np.sin(2 * np.pi * (sr / 4.3) * t) \
        + 0.05 * rng.standard_normal(n)
    return x.astype(np.float32)


def wow_curve(duration, sr, hop, depth=0.01, freq=0.5556):
    k = int(duration * sr / hop)
    times = np.linspace(0, duration, k)
    return np.stack((times, 1 + depth * np.sin(2 * np.pi * freq * times)), -1)


# ----------------------------------------------------------------------------------------- STFT
def _check_stft(s, x, n_fft, step, window, zp):
    truth = onp.stft_f64(x, n_fft, step, window, zp).T          # (F, T) complex128
    assert s.shape == truth.shape
    assert s.dtype == np.complex64
    per_frame = rel_l2(s.astype(np.complex128), truth, axis=0)
    scale = np.sqrt(np.mean(np.abs(truth) ** 2))
    # frames that are (numerically) silent carry no relative information
    live = np.sqrt(np.mean(np.abs(truth) ** 2, axis=0)) > 1e-6 * scale
    assert np.all(per_frame[live] <= TOL), float(per_frame[live].max())
    assert rel_max(s, truth) <= TOL


def test_stft_golden_cases(golden_dir, fourier):
    z = _load(golden_dir, "stft")
    names = sorted(k[:-6] for k in z.files if k.endswith("__meta"))
    assert len(names) >= 7
    for name in names:
        n_fft, step, zp = (int(v) for v in z[name + "__meta"])
        window = str(z[name + "__window"])
        x = z[name + "__x"]
        s = fourier.stft(x, n_fft, step, window, zp)
        _check_stft(s, x, n_fft, step, window, zp)
        # and against the reference's own float32 numpy back-end output
        ref = z[name + "__S"]
        assert s.shape == ref.shape
        assert rel_l2(s, ref) <= 2 * TOL, name
        assert s.flags.f_contiguous


def test_stft_strided_column_view(golden_dir, fourier):
    z = _load(golden_dir, "stft")
    inter = z["interleaved__x"]
    s = fourier.stft(inter[:, 1], 1024, 256)
    _check_stft(s, np.ascontiguousarray(inter[:, 1]), 1024, 256, "blackmanharris", 1)


def test_column_of_a_wide_interleaved_array(fourier, resampling, monkeypatch):
    """signal[:, c] of an 8-channel (frames, channels) array, as the GUIs pass it: the strided span goes up in
    staged pieces and is de-interleaved on the device (no per-sample DMA rows); several pipeline chunks."""
    monkeypatch.setenv("PAR_B200_CHUNK_BYTES", str(1 << 18))
    sig = np.stack([synth(60000, 300 + c) for c in range(8)], axis=1)
    assert sig.strides == (32, 4)
    for c in (0, 5):
        s = fourier.stft(sig[:, c], 1024, 256)
        _check_stft(s, np.ascontiguousarray(sig[:, c]), 1024, 256, "blackmanharris", 1)
    pos = np.linspace(300.0, 59000.0, 40000)
    y = resampling.sinc_wrapper(pos, sig[:, 3], 0, 50)
    ref = oracle.sinc_c(pos, np.ascontiguousarray(sig[:, 3]), 50)
    assert rel_l2(y.astype(np.float64), ref.astype(np.float64)) <= TOL and rel_max(y, ref) <= TOL


@pytest.mark.parametrize("n_fft,hop", [(32, 8), (64, 16), (128, 32), (256, 64), (512, 32), (1024, 256),
                                       (2048, 512), (4096, 1024), (8192, 2048), (16384, 4096),
                                       (32768, 8192), (4096, 1000), (4096, 4096), (1024, 3000)])
def test_stft_sizes(fourier, n_fft, hop):
    x = synth(max(5 * n_fft + 123, 20000), n_fft + hop)
    s = fourier.stft(x, n_fft, hop)
    _check_stft(s, x, n_fft, hop, "blackmanharris", 1)


@pytest.mark.parametrize("n_fft,hop", [(65536, 16384), (131072, 32768), (262144, 131072), (524288, 1048576),
                                       (1048576, 262144)])
def test_stft_large_sizes(fourier, n_fft, hop):
    """Four-step path (N*Z > 32768): the GUIs go up to 1048576 (util/widgets.py:333-335)."""
    x = synth(3 * n_fft + 4567, n_fft % 1000 + 1)
    s = fourier.stft(x, n_fft, hop)
    _check_stft(s, x, n_fft, hop, "blackmanharris", 1)
    mag = fourier.get_mag(x, n_fft, hop, "hann")
    truth = np.abs(onp.stft_f64(x, n_fft, hop, "hann").T) + 1e-7
    assert rel_max(mag, truth) <= TOL


def test_stft_large_zeropad_and_multichannel(fourier):
    sig = np.stack([synth(200000, 81), synth(200000, 82)], axis=1)
    multi = fourier.stft_multi(sig, 16384, 4096, zeropad=4)              # 65536-point transform
    for c in range(2):
        _check_stft(multi[c], np.ascontiguousarray(sig[:, c]), 16384, 4096, "blackmanharris", 4)
    short = synth(40000, 83)                                             # shorter than the window: both edges reflect
    _check_stft(fourier.stft(short, 65536, 16384), short, 65536, 16384, "blackmanharris", 1)


@pytest.mark.parametrize("n_fft,hop,zp", [(256, 64, 4), (1024, 256, 2), (2048, 512, 16), (64, 16, 8)])
def test_stft_zeropad(fourier, n_fft, hop, zp):
    x = synth(9000, 31 + zp)
    s = fourier.stft(x, n_fft, hop, "hann", zp)
    _check_stft(s, x, n_fft, hop, "hann", zp)


def test_stft_short_and_edge_inputs(fourier):
    # shorter than the window but longer than the reflect pad
    for n, n_fft, hop in [(700, 1024, 256), (513, 1024, 1), (2049, 4096, 1024), (17, 32, 8)]:
        x = synth(n, n)
        _check_stft(fourier.stft(x, n_fft, hop), x, n_fft, hop, "blackmanharris", 1)
    with pytest.raises(ValueError):
        fourier.stft(np.zeros((10, 2), np.float32))
    # float64 / int16 inputs are accepted and computed in float32 like the reference's torch path
    x = synth(5000, 3)
    s64 = fourier.stft(x.astype(np.float64), 512, 128)
    assert np.array_equal(s64, fourier.stft(x, 512, 128))
    # step=None -> n_fft // 2 (util/fourier.py:63)
    assert fourier.stft(x, 512, None).shape == (257, 5000 // 256 + 1)


def test_get_mag_matches_abs_of_stft(fourier):
    x = synth(50000, 5)
    mag = fourier.get_mag(x, 4096, 1024)
    truth = np.abs(onp.stft_f64(x, 4096, 1024).T) + 1e-7
    assert mag.dtype == np.float32 and mag.shape == truth.shape
    assert np.all(rel_l2(mag.astype(np.float64), truth, axis=0) <= TOL)
    assert rel_max(mag, truth) <= TOL
    mag2 = fourier.get_mag(x, 1024, 256, "hann", 4)
    truth2 = np.abs(onp.stft_f64(x, 1024, 256, "hann", 4).T) + 1e-7
    assert rel_max(mag2, truth2) <= TOL


def test_stft_linearity_full_size(fourier):
    """Size-independent property at a BASELINE-sized frame count: STFT(a*x + y) = a*STFT(x) +
    STFT(y) and Parseval against the time-domain frame energy on sampled frames."""
    n = 96000 * 60
    x, y = synth(n, 100), synth(n, 101)
    a = np.float32(0.37)
    sx, sy = fourier.stft(x, 4096, 1024), fourier.stft(y, 4096, 1024)
    sz = fourier.stft(a * x + y, 4096, 1024)
    assert sx.shape == (2049, n // 1024 + 1)
    lin = a * sx + sy
    assert rel_l2(sz, lin) <= 2 * TOL
    # spot-check frames across the whole length against the oracle
    win = onp.get_window_f32("blackmanharris", 4096).astype(np.float64)
    xp = np.pad(x, 2048, mode="reflect").astype(np.float64)
    for t in (0, 1, 2, 1000, 2811, sx.shape[1] - 2, sx.shape[1] - 1):
        truth = np.fft.rfft(xp[t * 1024: t * 1024 + 4096] * win) / 64.0
        assert rel_l2(sx[:, t], truth) <= TOL, t


# ----------------------------------------------------------------------------------------- iSTFT
def test_istft_golden(golden_dir, fourier):
    z = _load(golden_dir, "istft")
    for name in ("rt512_32", "rt1024_256", "rt4096_1024"):
        n_fft, hop, n = (int(v) for v in z[name + "__meta"])
        s = z[name + "__S"]
        keep = s.copy()
        y = fourier.istft(s, hop_length=hop, length=n)
        assert np.array_equal(s, keep)          # no in-place scaling (reference bug not reproduced)
        truth = z[name + "__y_from_c128"]       # float64 path of the reference on the same spectrum
        ref32 = z[name + "__y_from_c64"]
        assert y.dtype == np.float32 and y.shape == ref32.shape
        assert rel_l2(y, truth) <= TOL, name
        assert rel_max(y, truth) <= TOL, name
        assert rel_max(y, ref32) <= 2 * TOL, name
        y128 = fourier.istft(s.astype(np.complex128), hop_length=hop, length=n)
        assert y128.dtype == np.float64
    s = z["masked__S"]
    y = fourier.istft(s, hop_length=128)
    assert y.shape == z["masked__y"].shape
    assert rel_max(y, z["masked__y"]) <= 2 * TOL
    assert rel_max(y, onp.istft_ref(s.astype(np.complex128), hop_length=128)) <= TOL


@pytest.mark.parametrize("n_fft,hop", [(64, 16), (512, 32), (1024, 256), (4096, 1024), (16384, 4096), (2048, 1024)])
def test_stft_istft_round_trip(fourier, n_fft, hop):
    """fix_length -> stft -> istft(length=n) reconstructs the input, the way
    dropout_healer_gui.py:129-164 chains them."""
    n = 6 * n_fft + 777
    x = synth(n, n_fft)
    ypad = fourier.fix_length(x, n + n_fft // 2)
    s = fourier.stft(ypad, n_fft, hop)
    y = fourier.istft(s, hop_length=hop, length=n)
    assert y.shape == (n,)
    assert np.max(np.abs(y - x)) <= 1e-6 * np.max(np.abs(x)) * 2


def test_istft_fused_equals_two_pass(par):
    """The fused inverse transform + overlap-add (no scratch) and the frames-to-scratch + gather pair give identical
    bits: both accumulate every output sample in ascending frame order.  Sizes with sub-warp slots (n_fft 64 ... 512),
    one-warp and multi-warp slots, hop not dividing n_fft, several channels, short and over-long `length`."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent("""
        import os, sys, numpy as np
        sys.path.insert(0, %r)
        from pyaudiorestoration_b200.util import fourier
        rng = np.random.default_rng(3)
        out = {}
        for n_fft, hop, n in ((64, 16, 5000), (512, 32, 40000), (512, 96, 20011), (1024, 256, 50000), (4096, 1024, 120000),
                              (16384, 4096, 200000), (32768, 8192, 150000)):
            x = rng.standard_normal(n).astype(np.float32)
            S = np.array(fourier.stft(fourier.fix_length(x, n + n_fft // 2), n_fft, hop))
            out[f"{n_fft}_{hop}"] = fourier.istft(S, hop_length=hop, length=n)
            out[f"{n_fft}_{hop}_long"] = fourier.istft(S, hop_length=hop, length=n + 3 * n_fft)
            out[f"{n_fft}_{hop}_free"] = fourier.istft(S, hop_length=hop)
        np.savez(sys.argv[1], **out)
    """ % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        res = {}
        for mode in ("0", "1"):
            env = dict(os.environ, PAR_B200_ISTFT_TWO_PASS=mode)
            path = os.path.join(d, f"r{mode}.npz")
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env)
            res[mode] = np.load(path)
        for k in res["0"].files:
            assert np.array_equal(res["0"][k], res["1"][k]), k


def test_istft_length_rules(fourier):
    x = synth(5000, 77)
    s = fourier.stft(x, 512, 128)
    full = fourier.istft(s, hop_length=128)
    assert full.shape == onp.istft_ref(s, hop_length=128).shape
    # `length` beyond the natural length: the tail is covered by the last frame(s) only, where the
    # window sum-square is ~1e-9 and y / wss amplifies float32 rounding by 1/w (in the reference's
    # float32 path just the same).  The 1e-6 bar applies where the envelope is well conditioned.
    nfr = min(s.shape[1], int(np.ceil((6000 + 512) / 128)))
    wss = onp.window_sumsquare("blackmanharris", nfr, 128, 512, dtype=np.float64)[256:]
    wss = onp.fix_length(wss, 6000)
    good = wss > 1e-2 * wss.max()
    longer = fourier.istft(s, hop_length=128, length=6000)
    assert longer.shape == (6000,)
    ref = onp.istft_ref(s.astype(np.complex128), hop_length=128, length=6000)
    assert rel_max(longer[good], ref[good]) <= TOL
    assert rel_max(longer, ref) <= 1e-3
    assert np.array_equal(longer[len(full) + 256:], np.zeros(6000 - len(full) - 256, np.float32))
    shorter = fourier.istft(s, hop_length=128, length=1000)
    ref = onp.istft_ref(s.astype(np.complex128), hop_length=128, length=1000)
    assert rel_max(shorter, ref) <= TOL


# ----------------------------------------------------------------------------------------- positions
def test_speed_to_pos_golden_bit_exact(golden_dir, resampling):
    z = _load(golden_dir, "positions")
    for name in ("wow", "ramp"):
        pos = resampling.speed_to_pos(z[name + "__sampletimes"], z[name + "__speeds"], int(z[name + "__n_in"]))
        assert pos.dtype == np.float64
        assert np.array_equal(pos, z[name + "__pos"]), name
    pos = resampling.speed_to_pos(z["wow5_noend__sampletimes"], z["wow5_noend__speeds"], int(z["wow5_noend__n_in"]))
    ref = z["wow5_noend__pos"]
    assert len(pos) < len(ref)
    assert np.array_equal(pos, ref[: len(pos)])


def test_speed_to_pos_matches_oracle_long(resampling):
    sr, dur = 96000, 20.0
    curve = wow_curve(dur, sr, 1024)
    n_in = int(sr * dur)
    pos = resampling.speed_to_pos(curve[:, 0] * sr, curve[:, 1], n_in)
    ref = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], n_in)
    assert len(pos) == len(ref)
    assert np.array_equal(pos, ref)
    # tuples are accepted (the reference's test_sinc passes tuples, util/resampling.py:271-273)
    pos = resampling.speed_to_pos((0, 8000), (.5, 2), 8000)
    assert np.array_equal(pos, oracle.speed_to_pos_c(np.array([0., 8000.]), np.array([.5, 2.]), 8000))


def test_positions_quotient_is_the_ieee_quotient(par):
    """The positions kernels replace j/(n-1) by q0 = j*rcp, r = fma(-q0, n-1, j), q = fma(r, rcp, q0);
    exhaustive device check against the IEEE division for every segment length up to 40000."""
    from pyaudiorestoration_b200 import _lib
    assert _lib.lib().par_selftest_positions_quotient(40000, _lib.device()) == 0


def test_speed_to_pos_many_segment_lengths(resampling):
    rng = np.random.default_rng(3)
    lens = np.concatenate([[2, 3, 4, 5, 1023, 1024, 1025, 2047, 2048, 4095, 4096, 8191], rng.integers(2, 6000, 150)])
    st = np.concatenate([[0.0], np.cumsum(lens.astype(np.float64))])
    sp = 1 + 0.3 * rng.uniform(-1, 1, len(st))
    pos = resampling.speed_to_pos(st, sp, 10 ** 9)
    ref = oracle.speed_to_pos_c(st, sp, 10 ** 9)
    assert len(pos) == len(ref) and np.array_equal(pos, ref)


# ----------------------------------------------------------------------------------------- resampler
def _check_sinc(y, pos, x, nt):
    truth = oracle.sinc_c(pos, x, nt)       # float64 arithmetic, float32 store
    assert y.dtype == np.float32 and y.shape == truth.shape
    scale = max(np.max(np.abs(truth)), 1e-30)
    assert np.max(np.abs(y - truth)) <= TOL * scale, float(np.max(np.abs(y - truth)) / scale)
    assert rel_l2(y.astype(np.float64), truth.astype(np.float64)) <= TOL


def test_sinc_golden(golden_dir, resampling):
    z = _load(golden_dir, "sinc")
    checks = [("wow_nt8__y", "wow__pos", "wow__x", 8), ("wow_nt50__y", "wow__pos", "wow__x", 50),
              ("wow_nt128__y", "wow__pos", "wow__x", 128), ("ramp_nt50__y", "ramp__pos", "ramp__x", 50),
              ("edges_nt50__y", "edges__pos", "edges__x", 50), ("edges_nt128__y", "edges__pos", "edges__x", 128),
              ("integer_nt50__y", "integer__pos", "edges__x", 50)]
    for yk, pk, xk, nt in checks:
        y = resampling.sinc_wrapper(z[pk], z[xk], 0, nt)
        ref = z[yk]
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(y - ref)) <= TOL * scale, (yk, float(np.max(np.abs(y - ref)) / scale))
        _check_sinc(y, z[pk], z[xk], nt)


@pytest.mark.parametrize("nt", [1, 2, 8, 50, 100, 128, 512])
def test_sinc_quality_range(resampling, nt):
    sr = 96000
    x = synth(sr * 2, nt)
    curve = wow_curve(2.0, sr, 1024, depth=0.03, freq=2.0)
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], len(x))
    y = resampling.sinc_wrapper(pos, x, 0, nt)
    _check_sinc(y, pos, x, nt)


def test_sinc_speed_extremes(resampling):
    x = synth(40000, 9)
    for speeds in [(0.5, 2.0), (2.0, 0.5), (0.25, 0.25), (3.9, 3.9), (1.0, 1.0)]:
        pos = oracle.speed_to_pos_c(np.array([0.0, 30000.0]), np.array(speeds, dtype=np.float64), len(x))
        y = resampling.sinc_wrapper(pos, x, 0, 50)
        _check_sinc(y, pos, x, 50)


def test_sinc_mt_fills_strided_output(resampling):
    """sinc_wrapper_mt(output[:, c], sample_at, signal[:, c], 0, NT) exactly as run() calls it
    (util/resampling.py:227): strided input and output views of interleaved arrays."""
    sig = np.stack([synth(30000, 1), synth(30000, 2)], axis=1)
    pos = oracle.speed_to_pos_c(np.array([0.0, 30000.0]), np.array([0.98, 1.03]), 30000)
    out = np.zeros((len(pos), 2), np.float32)
    for c in range(2):
        resampling.sinc_wrapper_mt(out[:, c], pos, sig[:, c], 0, 50)
        _check_sinc(np.ascontiguousarray(out[:, c]), pos, np.ascontiguousarray(sig[:, c]), 50)


def test_sinc_empty_and_degenerate(resampling):
    x = synth(1000, 4)
    assert len(resampling.sinc_wrapper(np.zeros(0), x, 0, 8)) == 0
    # positions far outside the signal give zeros (empty tap slice in the reference)
    y = resampling.sinc_wrapper(np.array([5000.0, 5001.0, 1e12]), x, 0, 8)
    assert np.array_equal(y, np.zeros(3, np.float32))
    # a single position: the reference leaves period_to unbound; fc=1 is used here
    y1 = resampling.sinc_wrapper(np.array([500.25]), x, 0, 8)
    assert np.isfinite(y1).all()


def test_linear_mode(resampling):
    x = synth(20000, 6)
    pos = oracle.speed_to_pos_c(np.array([0.0, 20000.0]), np.array([0.9, 1.2]), 20000)
    pos = np.concatenate([[-3.0, 0.0, 0.5], pos, [19999.0, 19999.5, 30000.0]])
    out = resampling.resample_channels(x[:, None], pos, [0], "Linear")
    ref = onp.linear_resample(pos, x)
    assert np.array_equal(out[:, 0], ref)


# ----------------------------------------------------------------------------------------- run()
class _Prog:
    class _Sig:
        def __init__(self):
            self.values = []

        def emit(self, v):
            self.values.append(v)

    def __init__(self):
        self.notifyProgress = self._Sig()


@pytest.mark.parametrize("mode", ["Sinc", "Linear"])
def test_run_writes_reference_wav(tmp_path, resampling, mode):
    from pyaudiorestoration_b200.util import io_ops
    sr = 48000
    sig = np.stack([synth(sr, 11, sr), synth(sr, 12, sr), synth(sr, 13, sr)], axis=1)
    curve = wow_curve(1.0, sr, 512, depth=0.02, freq=3.0)
    prog = _Prog()
    name = str(tmp_path / "take.wav")
    ret = resampling.run([name], signal_data=[(sig, sr)], speed_curve=curve, resampling_mode=mode,
                         sinc_quality=50, use_channels=[0, 2, 7], prog_sig=prog, suffix="_x")
    assert ret is None
    out, sr2, ch = io_ops.read_file(str(tmp_path / "take_res_x.wav"))
    assert sr2 == sr and ch == 2
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], len(sig))
    assert out.shape == (len(pos), 2)
    for o, c in enumerate((0, 2)):
        x = np.ascontiguousarray(sig[:, c])
        if mode == "Sinc":
            _check_sinc(np.ascontiguousarray(out[:, o]), pos, x, 50)
        else:
            assert np.array_equal(out[:, o], onp.linear_resample(pos, x))
    assert prog.notifyProgress.values[0] == 0 and prog.notifyProgress.values[-1] == 100


def test_run_lag_curve_and_file_input(tmp_path, resampling):
    from pyaudiorestoration_b200.util import io_ops
    sr = 44100
    sig = np.stack([synth(20000, 21, sr), synth(20000, 22, sr)], axis=1)
    src = str(tmp_path / "src.wav")
    io_ops.write_float_wav(src, sig, sr)
    lag = np.array([[0.0, 0.0], [0.2, 0.001], [0.45, -0.002]])
    resampling.run([src], lag_curve=lag, resampling_mode="Sinc", sinc_quality=20)
    out, _, ch = io_ops.read_file(str(tmp_path / "src_res.wav"))
    pos = onp.lag_to_positions(lag, sr, len(sig))
    assert ch == 2 and out.shape == (len(pos), 2)
    _check_sinc(np.ascontiguousarray(out[:, 1]), pos, np.ascontiguousarray(sig[:, 1]), 20)


# ----------------------------------------------------------------------------------------- device-pointer ABI
def test_device_pointer_calls_multichannel(par):
    """PAR_DEVICE_PTRS entry points on torch-owned device memory and torch's current stream:
    2-channel planar STFT (complex + magnitude), positions, 2-channel sinc -- the bench path."""
    import torch
    from pyaudiorestoration_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", _lib.device())
    sr, dur, n_fft, hop, nt = 96000, 3.0, 4096, 1024, 128
    n = int(sr * dur)
    xs = np.stack([synth(n, 1234), synth(n, 1235)])
    x = torch.from_numpy(xs).to(dev)
    T = int(L.par_stft_num_frames(n, n_fft, hop))
    F = n_fft // 2 + 1
    S = torch.empty((2, T, F), dtype=torch.complex64, device=dev)
    win = onp.get_window_f32("blackmanharris", n_fft)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = L.par_stft_f32(x.data_ptr(), n, 1, 2, n, n_fft, hop, 1, win.ctypes.data, S.data_ptr(), F, T * F,
                        _lib.PAR_DEVICE_PTRS, dev.index, stream)
    _lib.check(rc, "par_stft_f32")
    M = torch.empty((2, T, F), dtype=torch.float32, device=dev)
    rc = L.par_stft_f32(x.data_ptr(), n, 1, 2, n, n_fft, hop, 1, win.ctypes.data, M.data_ptr(), F, T * F,
                        _lib.PAR_DEVICE_PTRS | _lib.PAR_OUT_MAGNITUDE, dev.index, stream)
    _lib.check(rc, "par_stft_f32 mag")
    curve = wow_curve(dur, sr, hop)
    st, sp = np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1])
    ref_pos = oracle.speed_to_pos_c(st, sp, n)
    pos = torch.empty(len(ref_pos) + 4096, dtype=torch.float64, device=dev)
    m = np.zeros(1, np.int64)
    rc = L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos.data_ptr(), pos.numel(),
                                m.ctypes.data, _lib.PAR_DEVICE_PTRS, dev.index, stream)
    _lib.check(rc, "par_speed_to_pos_f64")
    m = int(m[0])
    assert m == len(ref_pos)
    out = torch.empty((2, m), dtype=torch.float32, device=dev)
    rc = L.par_sinc_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, 2, n, nt, out.data_ptr(), 1, m,
                                 _lib.PAR_DEVICE_PTRS, dev.index, stream)
    _lib.check(rc, "par_sinc_resample_f32")
    torch.cuda.synchronize(dev)
    assert np.array_equal(pos[:m].cpu().numpy(), ref_pos)
    S, M, out = S.cpu().numpy(), M.cpu().numpy(), out.cpu().numpy()
    for c in range(2):
        truth = onp.stft_f64(xs[c], n_fft, hop)
        assert np.all(rel_l2(S[c].astype(np.complex128), truth, axis=1) <= TOL)
        assert rel_max(M[c], np.abs(truth) + 1e-7) <= TOL
        _check_sinc(out[c], ref_pos, xs[c], nt)
    assert L.par_kernel_launch_count() > 0


# ----------------------------------------------------------------------------------------- in-place layouts / fused call
def test_stft_multi_interleaved_matches_per_channel(fourier):
    sig = np.stack([synth(30000, 41), synth(30000, 42), synth(30000, 43)], axis=1)      # (frames, channels)
    multi = fourier.stft_multi(sig, 1024, 256)
    assert multi.shape == (3, 513, 30000 // 256 + 1)
    for c in range(3):
        assert np.array_equal(multi[c], fourier.stft(sig[:, c], 1024, 256))
        _check_stft(multi[c], np.ascontiguousarray(sig[:, c]), 1024, 256, "blackmanharris", 1)
    mags = fourier.stft_multi(sig[:, ::2], 1024, 256, magnitude=True)                   # strided channel subset
    assert np.array_equal(mags[1], fourier.get_mag(sig[:, 2], 1024, 256))
    planar = np.ascontiguousarray(sig.T)
    assert np.array_equal(fourier.stft_multi(planar.T, 1024, 256), multi)               # planar memory, same view


@pytest.mark.parametrize("mode", ["Sinc", "Linear"])
def test_varispeed_fused_matches_two_step(resampling, mode):
    sr = 96000
    sig = np.stack([synth(sr * 3, 51), synth(sr * 3, 52)], axis=1)
    curve = wow_curve(3.0, sr, 1024, depth=0.02, freq=1.3)
    out = resampling.varispeed(sig, sr, curve, None, mode, 50)
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], len(sig))
    assert out.shape == (len(pos), 2) and out.dtype == np.float32
    two_step = resampling.resample_channels(sig, pos, [0, 1], mode, 50)
    assert np.array_equal(out, two_step)
    for c in range(2):
        x = np.ascontiguousarray(sig[:, c])
        if mode == "Sinc":
            _check_sinc(np.ascontiguousarray(out[:, c]), pos, x, 50)
        else:
            assert np.array_equal(out[:, c], onp.linear_resample(pos, x))
    # channel subsets / reordering go through per-run buffers
    sub = resampling.varispeed(sig, sr, curve, [1], mode, 50)
    assert np.array_equal(sub[:, 0], out[:, 1])
    swapped = resampling.resample_channels(sig, pos, [1, 0], mode, 50)
    assert np.array_equal(swapped[:, ::-1], out)


def test_varispeed_capacity_error(par):
    from pyaudiorestoration_b200 import _lib
    L = _lib.lib()
    st, sp = np.array([0.0, 8000.0]), np.array([0.5, 2.0])
    x = synth(8000, 5)
    out = np.zeros(100, np.float32)
    m = np.zeros(1, np.int64)
    rc = L.par_varispeed_f32(st.ctypes.data, sp.ctypes.data, 2, x.ctypes.data, 8000, 1, 1, 0, 1, 50,
                             out.ctypes.data, 100, 1, 0, m.ctypes.data, 0, _lib.device(), None)
    assert rc == _lib.PAR_ECAPACITY and m[0] == len(oracle.speed_to_pos_c(st, sp, 8000))


def test_host_pipeline_chunking_is_invisible(fourier, resampling, monkeypatch):
    """Host-pointer calls stream chunks through upload / kernel / download streams; results must
    not depend on the chunk size (PAR_B200_CHUNK_BYTES forces many small chunks)."""
    sr = 96000
    sig = np.stack([synth(sr * 4, 61), synth(sr * 4, 62)], axis=1)
    curve = wow_curve(4.0, sr, 1024, depth=0.05, freq=0.9)
    ref_s = fourier.stft(sig[:, 1], 1024, 256)
    ref_m = fourier.stft_multi(sig, 4096, 1024, magnitude=True)
    ref_v = resampling.varispeed(sig, sr, curve, None, "Sinc", 50).copy()
    ref_l = resampling.varispeed(sig[:, 0], sr, curve, None, "Linear").copy()
    ref_s, ref_m = ref_s.copy(), ref_m.copy()
    for chunk in ("4096", "70000", "1000000"):
        monkeypatch.setenv("PAR_B200_CHUNK_BYTES", chunk)
        assert np.array_equal(fourier.stft(sig[:, 1], 1024, 256), ref_s)
        assert np.array_equal(fourier.stft_multi(sig, 4096, 1024, magnitude=True), ref_m)
        assert np.array_equal(resampling.varispeed(sig, sr, curve, None, "Sinc", 50), ref_v)
        assert np.array_equal(resampling.varispeed(sig[:, 0], sr, curve, None, "Linear"), ref_l)
    monkeypatch.delenv("PAR_B200_CHUNK_BYTES")
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], len(sig))
    _check_sinc(np.ascontiguousarray(ref_v[:, 1]), pos, np.ascontiguousarray(sig[:, 1]), 50)


# ----------------------------------------------------------------------------------------- time shards
@pytest.mark.parametrize("world", [1, 3])
def test_time_shards_reassemble_to_the_single_gpu_result(par, fourier, resampling, world):
    """every rank transforms / resamples its chunk + halo through the range entry
    points; concatenating the ranks' results equals the unsharded call bit for bit."""
    import torch
    from pyaudiorestoration_b200 import _lib, dist as pdist
    dev = torch.device("cuda", _lib.device())
    sr, n_fft, hop, nt = 96000, 1024, 256, 50
    n = sr * 2 + 123
    xs = np.stack([synth(n, 71), synth(n, 72)])                       # (channels, n) planar
    win = onp.get_window_f32("blackmanharris", n_fft)
    curve = wow_curve(n / sr, sr, hop, depth=0.03, freq=1.7)
    pos_np = resampling.speed_to_pos(curve[:, 0] * sr, curve[:, 1], n)
    pos = torch.from_numpy(np.ascontiguousarray(pos_np)).to(dev)
    full_S = fourier.stft_multi(xs.T, n_fft, hop)                     # (C, F, T)
    full_y = resampling.resample_channels(xs.T, pos_np, [0, 1], "Sinc", nt)
    full_l = resampling.resample_channels(xs.T, pos_np, [0, 1], "Linear")
    S_parts, y_parts, l_parts, o_prev = [], [], [], 0
    for rank in range(world):
        sh = pdist.TimeShard(n, n_fft, hop, nt, rank, world)
        buf = sh.local_buffer(2, dev)
        buf.copy_(torch.from_numpy(xs[:, sh.origin:sh.origin + sh.local_len]))   # chunk + halos as exchange_halos leaves them
        S_parts.append(sh.stft(buf, win).cpu().numpy())
        o0, o1 = sh.output_range(pos)
        assert o0 == o_prev
        o_prev = o1
        # positions: the rank's own slice (par_speed_to_pos_range_f64) equals the same slice of the full array
        ps, p0, m_glob = sh.positions(curve[:, 0] * sr, curve[:, 1], dev)
        assert m_glob == len(pos_np) and p0 <= o0 and p0 + len(ps) >= min(o1 + 1, m_glob)
        assert np.array_equal(ps.cpu().numpy(), pos_np[p0:p0 + len(ps)])
        assert sh.output_range(ps, p0, m_glob) == (o0, o1)
        y_parts.append(sh.resample(buf, ps, "Sinc", pos_origin=p0, m=m_glob).cpu().numpy())
        l_parts.append(sh.resample(buf, pos, "Linear", (o0, o1)).cpu().numpy())
    assert o_prev == len(pos_np)
    S = np.concatenate(S_parts, axis=1).transpose(0, 2, 1)
    assert np.array_equal(S, full_S)
    assert np.array_equal(np.concatenate(y_parts, axis=1).T, full_y)
    assert np.array_equal(np.concatenate(l_parts, axis=1).T, full_l)


@pytest.mark.parametrize("workload,channels,points", [("cfg2", 2, 100000), ("cfg3", 1, 100000)])
def test_parity_at_baseline_sizes_sampled(par, workload, channels, points):
    """BASELINE configs[1] (600 s, 96 kHz) and configs[2] (3600 s, 192 kHz; positions up to 6.9e8, 675 k curve
    segments) at FULL length: the read positions at 10^5 random outputs bit for bit against the serial float64
    recurrence, the resampled values of those outputs against the float64 oracle on the same taps, and a set of STFT
    frames -- bench.sampled_parity, the check bench.py also prints as `parity`."""
    import torch
    import bench
    from pyaudiorestoration_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", _lib.device())
    sr, dur, _, _ = bench.WORKLOADS[workload]
    n = int(sr * dur)
    x = torch.empty((channels, n), dtype=torch.float32, device=dev)
    for c in range(channels):
        bench.device_synth(torch, x[c], sr, 77 + c)
    curve = bench.wow_curve(dur, sr)
    st, sp = np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1])
    cap = int(n * 1.02) + 4096
    pos = torch.empty(cap, dtype=torch.float64, device=dev)
    out = torch.empty((channels, cap), dtype=torch.float32, device=dev)
    T, F = int(L.par_stft_num_frames(n, bench.N_FFT, bench.HOP)), bench.N_FFT // 2 + 1
    S = torch.empty((channels, T, F), dtype=torch.complex64, device=dev)
    win = bench.get_window()
    m_box = np.zeros(1, np.int64)
    _lib.check(L.par_stft_f32(x.data_ptr(), n, 1, channels, n, bench.N_FFT, bench.HOP, 1, win.ctypes.data, S.data_ptr(), F, T * F,
                              _lib.PAR_DEVICE_PTRS, dev.index, None), "stft")
    _lib.check(L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos.data_ptr(), cap, m_box.ctypes.data,
                                      _lib.PAR_DEVICE_PTRS, dev.index, None), "positions")
    m = int(m_box[0])
    _lib.check(L.par_sinc_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, channels, n, bench.NT, out.data_ptr(), 1, cap,
                                       _lib.PAR_DEVICE_PTRS, dev.index, None), "sinc")
    torch.cuda.synchronize(dev)
    res = bench.sampled_parity(torch, x, pos, m, out, S, curve, sr, n, n_points=points, n_frames=32, seed=5)
    del x, pos, out, S
    torch.cuda.empty_cache()
    L.par_release_cached_memory(dev.index)
    assert res["output_count_equal"] and res["positions_bit_exact"], res
    assert res["sinc_points"] >= 0.9 * points * channels
    assert res["sinc_rel_l2"] <= TOL and res["sinc_rel_max"] <= TOL, res
    assert res["stft_rel_l2"] <= TOL, res


def test_shared_segment_sums_give_the_same_position_slices(par, resampling):
    """Time-sharded positions with the per-segment totals computed in slices and shared (what
    dist.TimeShard.positions does over an all-gather) equal the all-local range call and the global array,
    bit for bit -- also when the curve runs past the end of the signal and the chunks are shorter than a
    curve segment (the windowed search must not walk into the segments behind the end)."""
    import torch
    from pyaudiorestoration_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", _lib.device())
    for sr, dur, hop, n_in, world in ((48000, 6.0, 1024, 48000 * 6, 3), (44100, 4.0, 4096, int(44100 * 2.5), 7)):
        curve = wow_curve(dur, sr, hop, depth=0.05)
        st = np.ascontiguousarray(curve[:, 0] * sr)
        sp = np.ascontiguousarray(curve[:, 1])
        full = resampling.speed_to_pos(st, sp, n_in)
        k, n_seg = len(st), len(st) - 1
        per = -(-n_seg // world)
        sums = torch.zeros(per * world, dtype=torch.float64, device=dev)
        for r in range(world):
            a, b = min(r * per, n_seg), min(r * per + per, n_seg)
            got_n = np.empty(n_seg, dtype=np.int64)
            _lib.check(L.par_segment_sums_f64(st.ctypes.data, sp.ctypes.data, k, a, b, sums[r * per:].data_ptr(),
                                              got_n.ctypes.data if r % 2 else None, _lib.PAR_DEVICE_PTRS, dev.index, None),
                       "par_segment_sums_f64")
            if r % 2:
                seg_n = got_n                                  # the segment lengths come back from the call (optional)
        bounds = np.linspace(0, n_in, world + 1)
        for r in range(world):
            lo = -np.inf if r == 0 else bounds[r] - 1.0
            hi = np.inf if r == world - 1 else bounds[r + 1] + 1.0
            got = []
            for shared in (True, False):
                out = torch.full((len(full) + 16,), np.nan, dtype=torch.float64, device=dev)
                box = np.zeros(3, dtype=np.int64)
                if shared:
                    rc = L.par_speed_to_pos_range_sums_f64(st.ctypes.data, sp.ctypes.data, k, float(n_in), lo, hi, sums.data_ptr(),
                                                           seg_n.ctypes.data if r % 2 else None,
                                                           out.data_ptr(), out.numel(), box[0:].ctypes.data, box[1:].ctypes.data,
                                                           box[2:].ctypes.data, _lib.PAR_DEVICE_PTRS, dev.index, None)
                else:
                    rc = L.par_speed_to_pos_range_f64(st.ctypes.data, sp.ctypes.data, k, float(n_in), lo, hi, out.data_ptr(),
                                                      out.numel(), box[0:].ctypes.data, box[1:].ctypes.data, box[2:].ctypes.data,
                                                      _lib.PAR_DEVICE_PTRS, dev.index, None)
                _lib.check(rc, "positions range")
                o0, cnt, m = (int(v) for v in box)
                assert m == len(full) and cnt > 0
                got.append((o0, out[:cnt].cpu().numpy()))
                assert np.array_equal(got[-1][1], full[o0:o0 + cnt])
                # the slice covers every position inside the window
                inside = np.nonzero((full >= lo) & (full <= hi))[0]
                assert len(inside) == 0 or (o0 <= inside[0] and inside[-1] < o0 + cnt)
            assert got[0][0] == got[1][0] and np.array_equal(got[0][1], got[1][1])


def test_range_entry_points_reject_uncovered_slices(par):
    import torch
    from pyaudiorestoration_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", _lib.device())
    x = torch.zeros(10000, device=dev)
    out = torch.zeros((100, 513), dtype=torch.complex64, device=dev)
    win = onp.get_window_f32("hann", 1024)
    # frames 10..19 need samples from 10*256-512; a slice starting at 4096 does not cover them
    rc = L.par_stft_range_f32(x.data_ptr(), 10000, 4096, 100000, 1, 10000, 1024, 256, 1, win.ctypes.data, 10, 10,
                              out.data_ptr(), 513, 0, _lib.PAR_DEVICE_PTRS, dev.index, None)
    assert rc == -1 and b"cover" in L.par_last_error()
    rc = L.par_stft_range_f32(x.data_ptr(), 10000, 4096, 100000, 1, 10000, 1024, 256, 1, win.ctypes.data, 10, 10,
                              out.data_ptr(), 513, 0, 0, dev.index, None)
    assert rc == -3          # host pointers are not supported by the range entries


# ----------------------------------------------------------------------------------------- BASELINE configs 1 and 4
def test_cfg1_flutter_stft_matches_reference(golden_dir, fourier):
    """samples/flutter.flac, n_fft 4096 / hop 1024: against frames of the reference's own output."""
    z = _load(golden_dir, "flutter")
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    s = fourier.stft(x, 4096, 1024)
    assert s.shape == tuple(z["shape"])
    _check_stft(s, x, 4096, 1024, "blackmanharris", 1)
    assert rel_l2(s[:, z["frames"]], z["S"]) <= TOL
    energy = np.sum(np.abs(s.astype(np.complex128)) ** 2, axis=0)
    assert np.max(np.abs(energy - z["frame_energy"]) / z["frame_energy"]) <= 2 * TOL


def test_cfg4_dropout_heal_and_locate(golden_dir, fourier):
    """dropouts_sample excerpt: heal against the unmodified reference's output; locator peaks
    (integer frame indices) bit-exact against the CPU oracle."""
    from pyaudiorestoration_b200 import dropouts
    z = _load(golden_dir, "dropouts")
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    sr, fft_size, hop = int(z["sr"]), int(z["fft_size"]), int(z["hop"])
    drops = [dropouts.Dropout(*m) for m in z["markers"].tolist()]
    y = dropouts.heal(x[:, None], sr, drops, fft_size, hop, channels=[0])[:, 0]
    ref = z["healed"]
    assert y.dtype == np.float32 and y.shape == ref.shape
    assert rel_max(y, ref) <= TOL and rel_l2(y.astype(np.float64), ref.astype(np.float64)) <= TOL
    # locator: magnitudes from the GPU, peaks must equal the ones found on the CPU magnitudes
    mag = fourier.get_mag(x, fft_size, hop, "blackmanharris", 1)
    cpu_mag = onp.to_mag(onp.stft_ref(x, fft_size, hop))
    dur = len(x) / sr
    for sens in (2.0, 4.0, 6.0):
        peaks, found = dropouts.locate(mag, sr, fft_size, hop, 0.1, dur - 0.1, 768.0, 13723.0, sensitivity=sens)
        want = onp.locate_peaks_ref(cpu_mag, sr, fft_size, hop, 0.1, dur - 0.1, 768.0, 13723.0, sens)
        assert peaks.dtype.kind == "i" and np.array_equal(peaks, want)
        assert len(found) == len(peaks)
    assert len(want) >= 10                                       # the excerpt really has dropouts to find
    # batch tool: band-wise valleys on the Hann spectrogram (dropouts_gui.py:264-288)
    mag_h = fourier.get_mag(x, fft_size, hop, "hann")
    cpu_h = onp.to_mag(onp.stft_ref(x, fft_size, hop, "hann"))
    got = dropouts.heuristic_band_peaks(dropouts.to_dB(np.array(mag_h)), sr, fft_size, 100, 15000, 5)
    want_h = onp.heuristic_peaks_ref(cpu_h, sr, fft_size, 100, 15000, 5)
    assert len(got) == len(want_h) == 4
    for g, w in zip(got, want_h):
        assert np.array_equal(g[4], w)
    assert any(len(g[4]) for g in got)                           # band edges no longer overflow: valleys are found
    out = dropouts.heuristic(x[:40000, None], sr, fft_size, hop)
    assert out.shape == (40000, 1) and np.isfinite(out).all() and np.any(out[:, 0] != x[:40000])


def test_cfg4_max_mono(fourier):
    from pyaudiorestoration_b200 import dropouts
    sig = np.stack([synth(30000, 91, 44100.0), synth(30000, 92, 44100.0)], axis=1)
    got = dropouts.max_mono(sig, 512, 32)
    want = onp.max_mono_ref(sig, 512, 32)
    for k in ("max", "min"):
        # a cell flips channel only if |L| and |R| agree to float32 rounding; allow a handful of such cells
        assert rel_l2(got[k].astype(np.float64), want[k].astype(np.float64)) <= 5e-6


def test_cfg4_full_file_against_the_reference_run(golden_dir, fourier):
    """BASELINE config 4 on the WHOLE samples/dropouts_sample.flac against outputs of the UNMODIFIED reference
    (tests/golden/make_golden_dropouts_full.py): heal with all 32 markers, the Alt-drag locator's integer frame
    indices (bit-exact) and markers, max/min-mono, and the heuristic batch tool's per-band peaks and output."""
    from pyaudiorestoration_b200 import dropouts
    z = _load(golden_dir, "dropouts_full")
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    sr, fft_size, hop = int(z["sr"]), int(z["fft_size"]), int(z["hop"])
    # D1
    drops = [dropouts.Dropout(*m) for m in z["markers"].tolist()]
    assert len(drops) == 32
    regs = [dropouts.marker_region(d, sr, fft_size, hop) for d in drops]
    assert np.array_equal(np.array(regs, dtype=np.int64), z["regions"])
    y = dropouts.heal(x[:, None], sr, drops, fft_size, hop, channels=[0])[:, 0]
    assert y.dtype == np.float32 and y.shape == z["healed"].shape
    assert rel_max(y, z["healed"]) <= TOL and rel_l2(y.astype(np.float64), z["healed"].astype(np.float64)) <= TOL
    # D2: GPU magnitudes in, the reference's own peak indices out
    (t0, f0), (t1, f1) = z["loc_corners"]
    f_lo, f_hi = max(min(f0, f1), 1), min(max(f0, f1), sr // 2 - 1)
    mag = fourier.get_mag(x, fft_size, hop, "blackmanharris", 1)
    peaks, found = dropouts.locate(mag, sr, fft_size, hop, min(t0, t1), max(t0, t1), f_lo, f_hi,
                                   sensitivity=float(z["loc_sensitivity"]), width_ms=float(z["loc_width_ms"]))
    assert peaks.dtype.kind == "i" and np.array_equal(peaks, z["loc_peaks"])
    got = np.array([(m.t - m.width / 2, m.f - m.height / 2, m.t + m.width / 2, m.f + m.height / 2) for m in found])
    assert got.shape == z["loc_markers"].shape and np.allclose(got, z["loc_markers"], rtol=0, atol=1e-9)
    # D3a
    d = int(z["mm_right_delay"])
    pcm_r = np.zeros_like(z["pcm"])
    pcm_r[d:] = (z["pcm"][:-d].astype(np.int32) * 4 // 5).astype(np.int16)
    right = (pcm_r.astype(np.float64) / 32768.0).astype(np.float32)
    mm = dropouts.max_mono(np.stack([x, right], axis=1), fft_size, hop)
    for k, ref in (("max", z["mm_max"]), ("min", z["mm_min"])):
        # a cell flips channel only where |L| and |R| agree to float32 rounding; allow a handful of such cells
        assert rel_l2(mm[k].astype(np.float64), ref.astype(np.float64)) <= 5e-6
    # D3b: per-band valley indices (bit-exact) and the corrected signal
    mag_h = fourier.get_mag(x, fft_size, hop, "hann")
    bands = dropouts.heuristic_band_peaks(dropouts.to_dB(np.array(mag_h)), sr, fft_size, 100, 15000, 5)
    assert len(bands) == 4
    for i, b in enumerate(bands):
        assert len(b[4]) > 0 and np.array_equal(b[4], z[f"heur_band_peaks_{i}"])
    out = dropouts.heuristic(x[:, None], sr, fft_size, hop)[:, 0]
    assert np.sum(out != x) > 1000                                  # the tool really changed the signal
    # the correction curve comes from dB MEANS over bands, dominated by cells far below the loud ones, where the
    # single-precision transforms of both sides carry ~1e-3 relative error: the audio agrees to a few 1e-6, the integer
    # valley indices above exactly
    assert rel_l2(out.astype(np.float64), z["heur_out"].astype(np.float64)) <= 5e-6 and rel_max(out, z["heur_out"]) <= 5e-5


def test_device_resident_masks_match_the_host_mask_path_and_the_reference(golden_dir, fourier):
    """stft -> mask -> istft with the spectrogram resident on the device (one upload, one
    download): the noise gate against the unmodified renoiser_gui run, the heal and max/min-mono operators against
    the reference-run fixtures (above) AND against this repo's host-mask path, multi-channel and strided inputs."""
    from pyaudiorestoration_b200 import _lib, dropouts
    L = _lib.lib()
    z = _load(golden_dir, "gate")
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    fft_size, hop = int(z["fft_size"]), int(z["hop"])
    launches0 = L.par_kernel_launch_count()
    y = dropouts.noise_gate(x, z["profile_db"], float(z["gain_db"]), fft_size, hop)[:, 0]
    assert 4 <= L.par_kernel_launch_count() - launches0 <= 6        # de-interleave, stft, gate, istft frames + overlap-add
    # the same gate applied on the host to the SAME device spectrogram (one more round trip): identical decisions
    n = len(x)
    spec = np.array(fourier.stft(fourier.fix_length(x, n + fft_size // 2), n_fft=fft_size, step=hop))
    mask = np.where(20 * np.log10(np.abs(spec).astype(np.float64) + 1e-7) > z["profile_db"][:, None], 0.0, float(z["gain_db"]))
    y_host = fourier.istft((spec * np.power(10, mask / 20).astype(np.float32)).astype(np.complex64), length=n, hop_length=hop)
    assert rel_l2(y.astype(np.float64), y_host.astype(np.float64)) <= TOL and rel_max(y, y_host) <= TOL
    # against the unmodified reference run: a hard gate flips the few cells whose level sits within the float32 rounding of
    # EITHER transform (both are single precision; quiet cells carry ~1e-3 relative error) of the profile
    assert rel_l2(y.astype(np.float64), z["gated"].astype(np.float64)) <= 5e-5
    # heal: device operator == host mask path, 3 channels, middle one untouched
    d = _load(golden_dir, "dropouts")
    xd = (d["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    sr = int(d["sr"])
    drops = [dropouts.Dropout(*m) for m in d["markers"].tolist()]
    sig = np.stack([xd, 0.5 * xd[::-1], synth(len(xd), 8, sr)], axis=1)
    dev = dropouts.heal(sig, sr, drops, 512, 32, channels=[0, 2])
    host = dropouts.heal(sig, sr, drops, 512, 32, channels=[0, 2], on_device=False)
    assert np.array_equal(dev[:, 1], sig[:, 1])
    for c in (0, 2):
        assert rel_l2(dev[:, c].astype(np.float64), host[:, c].astype(np.float64)) <= TOL and rel_max(dev[:, c], host[:, c]) <= TOL
    assert rel_l2(dev[:, 0].astype(np.float64), d["healed"].astype(np.float64)) <= TOL
    # max / min mono: device select == host select
    st = np.stack([synth(30000, 91, 44100.0), synth(30000, 92, 44100.0)], axis=1)
    a, b = dropouts.max_mono(st, 512, 32), dropouts.max_mono(st, 512, 32, on_device=False)
    for k in ("max", "min"):
        assert rel_l2(a[k].astype(np.float64), b[k].astype(np.float64)) <= 5e-6
    # a marker the device operator does not take (one bin high) falls back to the host mask, not to an error
    thin = [dropouts.Dropout(t=1.0, width=0.02, f=1000.0, height=10.0, surrounding=0.5)]
    assert dropouts._device_regions(thin, sr, 512, 32, len(xd)) is None


# ----------------------------------------------------------------------------------------- trackers (8f rank 1)
def _wow_signal(sr=44100, dur=3.0, seed=5):
    rng = np.random.default_rng(seed)
    t = np.arange(int(sr * dur)) / sr
    inst = 3150.0 * (1 + 0.006 * np.sin(2 * np.pi * 0.8 * t))
    phase = 2 * np.pi * np.cumsum(inst) / sr
    x = 0.3 * np.sin(phase) + 0.05 * np.sin(2.31 * phase) + 0.01 * rng.standard_normal(len(t))
    return x.astype(np.float32), sr


def test_trackers_match_reference_classes(golden_dir, fourier):
    """Device trackers vs (a) the oracle on the very same GPU magnitudes: Peak / Peak Track identical
    (index work + fixed float32/float64 op sequence), Center of Gravity to 1e-9; (b) the golden output
    of the unmodified reference classes on CPU magnitudes: same trace within the spectrogram's 1e-6."""
    from pyaudiorestoration_b200.util import wow_detection
    z = _load(golden_dir, "trackers")
    x, sr = _wow_signal()
    fft_size, hop, sr2, zp = (int(v) for v in z["params"])
    assert sr == sr2
    trail = [tuple(r) for r in z["trail"]]
    mag = fourier.get_mag(x, fft_size, hop, "blackmanharris", zp)
    for key, name in (("peak", "Peak"), ("peak_track", "Peak Track"), ("cog", "Center of Gravity")):
        tr = wow_detection.wow_detectors[name](mag, x, list(trail), fft_size * zp, hop, sr, 1.0, "Linear")
        t_ref, f_ref = onp.track_ref(key, np.array(mag), trail, fft_size * zp, hop, sr, 1.0)
        assert np.array_equal(tr.times, t_ref) and np.array_equal(tr.times, z[key + "__times"])
        if key == "cog":
            assert np.max(np.abs(tr.freqs - f_ref) / f_ref) <= 1e-9
        else:
            assert np.array_equal(tr.freqs, f_ref), key
        assert np.max(np.abs(tr.freqs - z[key + "__freqs"]) / z[key + "__freqs"]) <= 2e-6, key
        # fused: spectrogram never materialised
        t2, f2 = wow_detection.trace_signal(x, trail, fft_size, hop, sr, name, 1.0, "blackmanharris", zp)
        assert np.array_equal(t2, tr.times) and np.array_equal(f2, tr.freqs), key
    # the trace follows the synthetic wow: +-0.6 % around 3150 Hz at 0.8 Hz
    assert 10 < np.std(tr.freqs) < 16


def test_trackers_on_strided_signal_and_host_copy(fourier):
    from pyaudiorestoration_b200.util import wow_detection
    x, sr = _wow_signal(dur=2.0, seed=9)
    stereo = np.stack([x, x[::-1]], axis=1)
    trail = [(0.2, 3150.0), (1.8, 3150.0)]
    t1, f1 = wow_detection.trace_signal(stereo[:, 0], trail, 2048, 128, sr, "Peak", 2.0)
    mag = np.array(fourier.get_mag(stereo[:, 0], 2048, 128))            # an ordinary (bins, frames) C array
    tr = wow_detection.PeakTracker(mag, None, trail, 2048, 128, sr, 2.0)
    assert np.array_equal(t1, tr.times) and np.array_equal(f1, tr.freqs)
    assert abs(np.mean(f1) - 3150.0) < 5.0


@pytest.mark.parametrize("channels", [3, 4, 8, 9])
def test_sinc_channel_groups(resampling, channels):
    """Channel groups of 8 / 4 / 2 / 1 share one set of tap weights per output sample; every channel
    must equal the single-channel result bit for bit."""
    sr = 48000
    sig = np.stack([synth(sr, 200 + c, sr) for c in range(channels)], axis=1)
    curve = wow_curve(1.0, sr, 512, depth=0.04, freq=2.5)
    out = resampling.varispeed(sig, sr, curve, None, "Sinc", 64)
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], len(sig))
    assert out.shape == (len(pos), channels)
    for c in (0, channels // 2, channels - 1):
        single = resampling.sinc_wrapper(pos, sig[:, c], 0, 64)
        assert np.array_equal(out[:, c], single), c
    _check_sinc(np.ascontiguousarray(out[:, channels - 1]), pos, np.ascontiguousarray(sig[:, channels - 1]), 64)


@pytest.mark.parametrize("nt,channels,speed", [(128, 2, 1.0), (50, 1, 1.0), (64, 4, 1.3), (128, 3, 0.7), (300, 2, 1.0),
                                               (8, 2, 1.0), (128, 2, 3.2)])
def test_both_sinc_kernels_give_the_same_bits(par, nt, channels, speed):
    """The resampler has two kernels (csrc/resample.cu: `sinc_kernel`, every warp sets up and interpolates; and
    `sinc_kernel_ws`, set-up warps feeding interpolating warps).  Records, units and tap arithmetic are shared, so their
    outputs must be identical arrays -- wow, constant fast / slow speeds (all low-passed / all unpaired outputs), a span
    wider than the two-CTA kernel stages (3.2x), small and large tap tables -- and both must match the oracle."""
    from pyaudiorestoration_b200 import _lib
    L = _lib.lib()
    sr = 48000
    n = sr * 2 + 77
    sig = np.stack([synth(n, 300 + c, sr) for c in range(channels)])                  # planar (channels, n)
    t = np.arange(0, n, 512, dtype=np.float64)
    t = np.append(t, float(n))
    sp = speed * (1.0 + 0.02 * np.sin(2 * np.pi * 3.0 * t / sr))
    pos = oracle.speed_to_pos_c(t, sp, n)
    m = len(pos)
    outs = []
    for flag in (_lib.PAR_SINC_KERNEL_TILED, _lib.PAR_SINC_KERNEL_WS):
        out = np.full((channels, m), np.nan, np.float32)
        rc = L.par_sinc_resample_f32(pos.ctypes.data, m, sig.ctypes.data, n, 1, channels, n, nt, out.ctypes.data, 1, m,
                                     flag, _lib.device(), None)
        _lib.check(rc, "par_sinc_resample_f32")
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    _check_sinc(outs[1][channels - 1], pos, sig[channels - 1], nt)


# ----------------------------------------------------------------------------------------- re-entrancy
def test_concurrent_calls_from_threads(fourier, resampling):
    """The reference enters this path from QThread workers while the GUI thread calls stft/istft
    (util/qt_threads.py:19-35): concurrent calls from several host threads must give exactly the
    results of the same calls made one after the other."""
    import threading
    sr = 48000
    sigs = [np.stack([synth(sr * 2, 300 + 2 * k, sr), synth(sr * 2, 301 + 2 * k, sr)], axis=1) for k in range(4)]
    curve = wow_curve(2.0, sr, 512, depth=0.02, freq=1.1)

    def work(k):
        s = fourier.stft(sigs[k][:, 0], 2048, 512)
        m = fourier.get_mag(sigs[k][:, 1], 1024, 128, "hann", 2)
        y = resampling.varispeed(sigs[k], sr, curve, None, "Sinc", 32)
        r = fourier.istft(s, hop_length=512, length=sr * 2)
        return [np.array(v) for v in (s, m, y, r)]
    serial = [work(k) for k in range(4)]
    results, errors = [None] * 4, []

    def runner(k):
        try:
            for _ in range(3):
                results[k] = work(k)
        except Exception as e:                      # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=runner, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        for a, b in zip(results[k], serial[k]):
            assert np.array_equal(a, b)


def test_release_cached_memory(par, fourier):
    import torch
    from pyaudiorestoration_b200 import _lib
    x = synth(2_000_000, 17)
    a = np.array(fourier.stft(x, 4096, 1024))
    free0, _ = torch.cuda.mem_get_info(_lib.device())
    _lib.release_cached_memory()
    free1, _ = torch.cuda.mem_get_info(_lib.device())
    assert free1 >= free0
    assert np.array_equal(np.array(fourier.stft(x, 4096, 1024)), a)      # and everything still works afterwards
