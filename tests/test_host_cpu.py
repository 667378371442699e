"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/par_b200.h
declares, the host-only entry points are bit-exact, the product refuses to compute without a
device, and the host helpers of the Python mirror behave like the reference's."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import pyaudiorestoration_b200 as p
    from pyaudiorestoration_b200 import _lib
    if not os.path.exists(p.library_path()):
        p.build()
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "par_b200.h")).read()
    declared = set(re.findall(r"PAR_API\s+[\w\s\*]+?\b(par_\w+)\s*\(", hdr))
    assert len(declared) >= 15
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(lib.EXPORTS)
    assert b"sm_100a" in L.par_version()


def test_no_device_means_loud_failure(lib):
    from pyaudiorestoration_b200.util import fourier, resampling
    if lib.lib().par_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(RuntimeError):
        fourier.stft(np.zeros(4096, np.float32))
    with pytest.raises(RuntimeError):
        resampling.sinc_wrapper(np.arange(10.0), np.zeros(100, np.float32), 0, 8)
    # the raw ABI reports PAR_ECUDA rather than computing on the host
    x = np.zeros(4096, np.float32)
    w = np.ones(1024, np.float32)
    out = np.zeros((5, 513), np.complex64)
    rc = lib.lib().par_stft_f32(x.ctypes.data, 4096, 1, 1, 0, 1024, 1024, 1, w.ctypes.data, out.ctypes.data, 513, 0,
                                0, 0, None)
    assert rc == -2 and b"CUDA" in lib.lib().par_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyaudiorestoration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, os.path.join(dirpath, f)


def test_speed_segments_bit_exact(lib, golden_dir):
    """par_speed_segments is host-only (serial float64, no FMA contraction): compare with the
    oracle's restatement of util/resampling.py:111-118 and with the reference's output length."""
    from oracle import oracle_np as onp
    z = np.load(os.path.join(golden_dir, "positions.npz"))
    L = lib.lib()
    for name in ("wow", "wow5_noend", "ramp"):
        st = np.ascontiguousarray(z[name + "__sampletimes"], dtype=np.float64)
        sp = np.ascontiguousarray(z[name + "__speeds"], dtype=np.float64)
        seg = np.zeros(len(st) - 1, np.int64)
        tot = np.zeros(1, np.int64)
        assert L.par_speed_segments(st.ctypes.data, sp.ctypes.data, len(st), seg.ctypes.data, tot.ctypes.data) == 0
        assert np.array_equal(seg, onp.speed_segments(st, sp))
        assert tot[0] == seg.sum()
    # the reference's filled prefix is exactly sum(n) long when its end test never fires
    assert tot[0] >= 0
    rng = np.random.default_rng(0)
    st = np.cumsum(rng.uniform(200, 3000, 500))
    sp = 1 + 0.2 * rng.standard_normal(500).clip(-3, 3)
    seg = np.zeros(499, np.int64)
    assert L.par_speed_segments(st.ctypes.data, sp.ctypes.data, 500, seg.ctypes.data, None) == 0
    assert np.array_equal(seg, onp.speed_segments(st, sp))
    assert L.par_speed_segments(st.ctypes.data, sp.ctypes.data, 1, seg.ctypes.data, None) == -1


def test_stft_num_frames(lib):
    L = lib.lib()
    from oracle import oracle_np as onp
    for n, n_fft, hop in [(186291, 4096, 1024), (57600000, 4096, 1024), (700, 1024, 256), (1000, 64, 48), (5, 32, 1)]:
        assert L.par_stft_num_frames(n, n_fft, hop) == onp.n_frames(n, n_fft, hop) == n // hop + 1


def test_wav_round_trip_and_pcm(tmp_path):
    from pyaudiorestoration_b200.util import io_ops
    from scipy.io import wavfile
    rng = np.random.default_rng(1)
    sig = rng.uniform(-1, 1, (1234, 3)).astype(np.float32)
    p = str(tmp_path / "a.wav")
    io_ops.write_float_wav(p, sig, 48000)
    back, sr, ch = io_ops.read_file(p)
    assert sr == 48000 and ch == 3 and np.array_equal(back, sig)
    sr2, sc = wavfile.read(p)                      # an independent reader agrees
    assert sr2 == 48000 and np.array_equal(sc, sig)
    for dtype, scale in ((np.int16, 32768.0), (np.int32, 2147483648.0), (np.uint8, None)):
        q = str(tmp_path / f"pcm_{np.dtype(dtype).name}.wav")
        if dtype == np.uint8:
            data = rng.integers(0, 256, (500, 2)).astype(np.uint8)
            want = (data.astype(np.float32) - 128) / 128
        else:
            info = np.iinfo(dtype)
            data = rng.integers(info.min, info.max, (500, 2)).astype(dtype)
            want = (data.astype(np.float64) / scale).astype(np.float32)
        wavfile.write(q, 22050, data)
        got, sr3, ch3 = io_ops.read_file(q)
        assert sr3 == 22050 and ch3 == 2 and np.array_equal(got, want)
    io_ops.write_file(str(tmp_path / "b.flac"), sig[:, :1], 8000, 1, suffix="_x")
    assert os.path.exists(str(tmp_path / "b_x.wav"))
    with pytest.raises(ValueError):
        io_ops.read_wav(__file__)


def test_host_helpers_match_reference_semantics():
    from pyaudiorestoration_b200.util import fourier, resampling
    from oracle import oracle_np as onp
    x = np.arange(10.0)
    assert np.array_equal(fourier.fix_length(x, 14), onp.fix_length(x, 14))
    assert np.array_equal(fourier.fix_length(x, 4), x[:4])
    assert np.array_equal(fourier.fix_length(x.reshape(5, 2), 7, axis=0)[5:], np.zeros((2, 2)))
    assert np.allclose(fourier.fft_freqs(8, 8000.0), [0, 1000, 2000, 3000, 4000])
    assert np.array_equal(fourier.to_mag(np.array([3 + 4j])), [5.0000001])
    assert fourier.dtype_r2c(np.float32) == np.complex64 and fourier.dtype_c2r(np.complex128) == np.float64
    assert fourier.pad_center(np.ones(4), 8).tolist() == [0, 0, 1, 1, 1, 1, 0, 0]
    with pytest.raises(fourier.ParameterError):
        fourier.pad_center(np.ones(4), 2)
    wss = fourier.window_sumsquare("hann", 6, hop_length=64, n_fft=256)
    assert np.allclose(wss, onp.window_sumsquare("hann", 6, 64, 256), atol=1e-6)
    assert resampling.find_cutoff(np.array([1.0, 2.0, 5.0, 7.0]), 5) == (2,)
    assert resampling.find_cutoff(np.array([1.0, 2.0]), 5) is None
    lag = np.array([[0.0, 0.0], [0.2, 0.001], [0.45, -0.002]])
    assert np.array_equal(resampling.lag_to_pos(lag, 44100, 20000), onp.lag_to_positions(lag, 44100, 20000))
    assert resampling._channel_runs(None, [0, 1, 2, 5, 7, 8]) == [[0, 0, 3], [3, 5, 1], [4, 7, 2]]
    with fourier.timed_log("x"):
        pass
