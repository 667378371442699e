"""Host logic of the Python mirror on a machine without a GPU: the compute entry points of the C ABI
are replaced by a TEST DOUBLE (tests/abi_double.py, oracle arithmetic behind the real pointer /
stride / pitch conventions of include/par_b200.h), so what is tested here is everything the Python
layer does around the library -- coercion, layouts, channel runs, views, files, progress."""
import os

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

from abi_double import AbiDouble


@pytest.fixture
def double(monkeypatch):
    from pyaudiorestoration_b200 import _lib
    fake = AbiDouble(_lib.lib())
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "require_device", lambda: 1)
    return fake


def synth(n, seed, sr=44100.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    return (0.25 * np.sin(2 * np.pi * 1000.0 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)


def wow_curve(duration, sr, hop, depth=0.02, freq=3.0):
    k = int(duration * sr / hop)
    times = np.linspace(0, duration, k)
    return np.stack((times, 1 + depth * np.sin(2 * np.pi * freq * times)), -1)


def test_stft_views_and_layouts(double):
    from pyaudiorestoration_b200.util import fourier
    inter = np.stack([synth(9000, 1), synth(9000, 2)], axis=1)
    s = fourier.stft(inter[:, 1], 1024, 256)                      # strided column view: passed in place
    name, n, stride = double.calls[-1][:3]
    assert (name, n, stride) == ("par_stft_f32", 9000, 2)
    want = onp.stft_f64(np.ascontiguousarray(inter[:, 1]), 1024, 256).T
    assert s.shape == want.shape and s.dtype == np.complex64 and s.flags.f_contiguous
    assert np.max(np.abs(s - want)) <= 1e-6 * np.max(np.abs(want))
    assert fourier.stft(inter[:, 1].astype(np.float64), 1024, 256).shape == s.shape   # other dtypes are converted
    assert double.calls[-1][2] == 1
    mag = fourier.get_mag(inter[:, 0], 512, 128, "hann", 2)
    assert mag.dtype == np.float32 and mag.shape == (513, 9000 // 128 + 1)
    multi = fourier.stft_multi(inter, 1024, 256)
    assert double.calls[-1][2:5] == (2, 2, 1)                     # frame stride 2, 2 channels, channel stride 1
    assert np.array_equal(multi[1], s)
    with pytest.raises(ValueError):
        fourier.stft(inter)
    with pytest.raises(ValueError):
        fourier.stft(np.zeros(0, np.float32))


def test_istft_round_trip_and_lengths(double):
    from pyaudiorestoration_b200.util import fourier
    x = synth(6000, 3)
    n_fft, hop = 512, 128
    s = fourier.stft(fourier.fix_length(x, len(x) + n_fft // 2), n_fft, hop)
    y = fourier.istft(s, hop_length=hop, length=len(x))
    assert y.dtype == np.float32 and np.max(np.abs(y - x)) < 2e-6
    assert fourier.istft(s, hop_length=hop).shape == onp.istft_ref(s, hop_length=hop).shape
    assert fourier.istft(s.astype(np.complex128), hop_length=hop, length=100).dtype == np.float64
    # only the frames the reference would use are handed to the library (util/fourier.py:373-381)
    fourier.istft(s, hop_length=hop, length=1000)
    assert double.calls[-1][2] == min(s.shape[1], int(np.ceil((1000 + n_fft) / hop)))


@pytest.mark.parametrize("mode", ["Sinc", "Linear"])
def test_run_files_progress_and_channel_runs(double, tmp_path, mode):
    from pyaudiorestoration_b200.util import io_ops, resampling
    sr = 22050
    sig = np.stack([synth(sr, 10 + c, sr) for c in range(4)], axis=1)
    curve = wow_curve(1.0, sr, 256)
    emitted = []
    prog = type("P", (), {"notifyProgress": type("S", (), {"emit": staticmethod(emitted.append)})()})()
    src = str(tmp_path / "tape.wav")
    resampling.run([src], signal_data=[(sig, sr)], speed_curve=curve, resampling_mode=mode, sinc_quality=16,
                   use_channels=[0, 1, 3, 9], prog_sig=prog, suffix="_t")
    out, sr2, ch = io_ops.read_file(str(tmp_path / "tape_res_t.wav"))
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], len(sig))
    assert (sr2, ch) == (sr, 3) and out.shape == (len(pos), 3)
    # channels 0,1 travel as one run (interleaved in place: frame stride 4, channel stride 1), channel 3 alone
    runs = [c for c in double.calls if c[0] == "par_varispeed_f32"]
    assert [(c[3], c[4], c[5]) for c in runs] == [(4, 2, 1), (4, 1, 1)]
    for o, c in enumerate((0, 1, 3)):
        x = np.ascontiguousarray(sig[:, c])
        want = oracle.sinc_c(pos, x, 16) if mode == "Sinc" else onp.linear_resample(pos, x)
        assert np.array_equal(out[:, o], want)
    assert emitted[0] == 0 and emitted[-1] == 100 and all(0 <= v <= 100 for v in emitted)
    # lag-curve mode reads the file and resamples every channel
    io_ops.write_float_wav(src, sig[:, :2], sr)
    lag = np.array([[0.0, 0.0], [0.5, 0.002], [0.9, -0.001]])
    resampling.run([src], lag_curve=lag, resampling_mode=mode, sinc_quality=8)
    out2, _, ch2 = io_ops.read_file(str(tmp_path / "tape_res.wav"))
    assert ch2 == 2 and out2.shape[0] == len(onp.lag_to_positions(lag, sr, len(sig)))
    with pytest.raises(ValueError):
        resampling.run([src], signal_data=[(sig, sr)])


def test_sinc_wrappers_fill_strided_outputs(double):
    from pyaudiorestoration_b200.util import resampling
    sig = np.stack([synth(5000, 20), synth(5000, 21)], axis=1)
    pos = oracle.speed_to_pos_c(np.array([0.0, 5000.0]), np.array([0.97, 1.04]), 5000)
    out = np.zeros((len(pos), 2), np.float32)
    resampling.sinc_wrapper_mt(out[:, 1], pos, sig[:, 1], 0, 12)
    assert np.array_equal(out[:, 1], oracle.sinc_c(pos, np.ascontiguousarray(sig[:, 1]), 12)) and not out[:, 0].any()
    y = resampling.sinc_wrapper(pos, sig[:, 0], 0, 12)
    n_arr, win = np.arange(-12, 13, dtype="float32"), np.hanning(25).astype("float32")
    y2 = np.empty(len(pos), "float32")
    resampling.sinc_core(pos, sig[:, 0], 0, y2, win, n_arr)
    assert np.array_equal(y, y2)
    assert np.array_equal(resampling.speed_to_pos((0, 5000), (0.97, 1.04), 5000), pos)


def test_dropout_heal_against_the_reference_output(double, golden_dir):
    """dropouts.heal (product host logic) + double == the unmodified reference's healed audio to 1e-6."""
    from pyaudiorestoration_b200 import dropouts
    z = np.load(os.path.join(golden_dir, "dropouts.npz"))
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    sr, fft_size, hop = int(z["sr"]), int(z["fft_size"]), int(z["hop"])
    drops = [dropouts.Dropout(*m) for m in z["markers"].tolist()]
    y = dropouts.heal(x[:, None], sr, drops, fft_size, hop)[:, 0]
    assert np.max(np.abs(y - z["healed"])) <= 1e-6 * np.max(np.abs(z["healed"]))
    mag = onp.to_mag(onp.stft_ref(x, fft_size, hop)).astype(np.float32)
    peaks, found = dropouts.locate(mag, sr, fft_size, hop, 0.1, len(x) / sr - 0.1, 768.0, 13723.0, sensitivity=4.0)
    assert np.array_equal(peaks, onp.locate_peaks_ref(mag, sr, fft_size, hop, 0.1, len(x) / sr - 0.1, 768.0, 13723.0, 4.0))
    assert len(found) == len(peaks) and all(d.width > 0 for d in found)


def test_trackers_host_side(double, golden_dir):
    from pyaudiorestoration_b200.util import wow_detection
    z = np.load(os.path.join(golden_dir, "trackers.npz"))
    rng = np.random.default_rng(5)
    sr, dur = 44100, 3.0
    t = np.arange(int(sr * dur)) / sr
    phase = 2 * np.pi * np.cumsum(3150.0 * (1 + 0.006 * np.sin(2 * np.pi * 0.8 * t))) / sr
    x = (0.3 * np.sin(phase) + 0.05 * np.sin(2.31 * phase) + 0.01 * rng.standard_normal(len(t))).astype(np.float32)
    fft_size, hop, _, zp = (int(v) for v in z["params"])
    spec = onp.to_mag(onp.stft_ref(x, fft_size, hop)).astype(np.float32)       # (bins, frames), C order
    trail = [tuple(r) for r in z["trail"]]
    for key, name in (("peak", "Peak"), ("peak_track", "Peak Track"), ("cog", "Center of Gravity")):
        tr = wow_detection.wow_detectors[name](spec, x, list(trail), fft_size * zp, hop, sr, 1.0, "Linear")
        assert np.array_equal(tr.times, z[key + "__times"]) and np.array_equal(tr.freqs, z[key + "__freqs"]), key
    # the host-side members of the registry against the unmodified reference classes (same spectrogram, same trail):
    # identical scipy calls on identical inputs
    for key, name in (("zero_crossing", "Zero-Crossing"), ("correlation", "Correlation")):
        tr = wow_detection.wow_detectors[name](spec, x[:, None], list(trail), fft_size * zp, hop, sr, 1.0, "Linear")
        assert np.array_equal(tr.times, z[key + "__times"]), key
        np.testing.assert_allclose(tr.freqs, z[key + "__freqs"], rtol=1e-12, atol=0, err_msg=key)
    free = wow_detection.wow_detectors["Freehand Draw"](spec, x, list(trail), fft_size * zp, hop, sr)
    assert np.array_equal(free.freqs, np.interp(free.times, [p[0] for p in trail], [p[1] for p in trail]))
    assert set(wow_detection.wow_detectors) == {"Peak", "Peak Track", "Center of Gravity", "Zero-Crossing", "Partials",
                                                "Freehand Draw", "Correlation", "Sine Regression"}
    # sine regression of a speed curve (pyrespeeder_gui.py:177): recovers a 0.5556 Hz wow of 1 %
    tt = np.linspace(0, 20, 2000)
    curve = np.stack((tt, 1 + 0.01 * np.sin(2 * np.pi * 0.5556 * tt + 0.3)), -1)
    amp, omega, phase, off = wow_detection.trace_sine_reg(curve, 2.0, 18.0, rpm=33.333)
    assert abs(abs(amp) - 0.01) < 1e-6 and abs(omega / (2 * np.pi) - 0.5556) < 1e-6 and off == 0
    assert list(wow_detection.zero_crossings(np.array([1.0, -1.0, -2.0, 3.0]))) == [0, 2]
