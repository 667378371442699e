"""Drop-in mirror of the reference's ``util/resampling.py`` module surface, computed on a B200.

Same names, argument meaning and side effects as the reference (citations are
``util/resampling.py:<line>`` of HENDRIX-ZT2/pyaudiorestoration): ``run`` turns a speed curve (or
a lag curve) into fractional read positions, resamples the selected channels with the windowed
sinc (or linear) interpolator and writes ``<stem>_res<suffix>.wav`` as an IEEE-float WAV.  The
positions expansion and both interpolators run in the sm_100a kernels of ``libpar_b200.so``
(``include/par_b200.h``); there is no numba / CPU path -- a missing library or device raises
``RuntimeError``.

Deliberate differences:

* ``speed_to_pos`` returns only the filled prefix of the reference's ``np.empty`` buffer (the
  reference leaks an uninitialised tail when its end test never fires, :108-109/:137);
* ``sinc_wrapper_mt`` does not split the output into ``os.cpu_count()`` thread chunks, so the
  "last element reuses the previous period" rule (:76-77) applies once, at the end of the array,
  exactly like ``sinc_wrapper`` (in the reference the result formally depends on the host's core
  count);
* with ``signal_data`` given and ``use_channels`` empty the reference hits an unbound
  ``num_channels`` (:216); here all channels are resampled, as the comment at :215 intends.
"""
import logging
import os
from time import time

import numpy as np

from .. import _lib
from . import io_ops
from .timing import log_duration


# ------------------------------------------------------------------------------------ helpers
def find_cutoff(array, cutoff):
    """First index whose value is >= cutoff, as a 1-tuple, or None (util/resampling.py:14-18)."""
    hit = np.nonzero(np.asarray(array) >= cutoff)[0]
    if len(hit):
        return (int(hit[0]),)
    return None


def _resample_into(output, sample_at, signal, nt, sinc=True, aligned_edges=False):
    """output[i] (1-D float32 view, any stride) = interpolated signal at sample_at[i].  The signal
    view is passed as it lies in memory (strided column views are uploaded as one interleaved
    span); a strided output goes through a contiguous pinned buffer."""
    L = _lib.lib()
    _lib.require_device()
    pos = np.ascontiguousarray(sample_at, dtype=np.float64)
    keep, sig_ptr, sig_stride = _lib.f32_layout(signal)
    n_in = len(keep)
    m = len(pos)
    if len(output) != m:
        raise ValueError("output and sample_at must have the same length")
    direct = (isinstance(output, np.ndarray) and output.dtype == np.float32 and output.ndim == 1
              and output.flags.c_contiguous and output.flags.writeable)
    tmp = output if direct else _lib.pinned_empty((max(m, 1),), np.float32)[:m]
    if m:
        flags = _lib.PAR_SINC_ALIGNED_EDGES if aligned_edges else 0
        if sinc:
            rc = L.par_sinc_resample_f32(pos.ctypes.data, m, sig_ptr, n_in, sig_stride, 1, 0, int(nt),
                                         tmp.ctypes.data, 1, 0, flags, _lib.device(), None)
            _lib.check(rc, "par_sinc_resample_f32")
        else:
            rc = L.par_linear_resample_f32(pos.ctypes.data, m, sig_ptr, n_in, sig_stride, 1, 0,
                                           tmp.ctypes.data, 1, 0, 0, _lib.device(), None)
            _lib.check(rc, "par_linear_resample_f32")
    if not direct:
        output[:] = tmp
    del keep
    return output


def _channel_runs(signal, use_channels):
    """Split the selected channels into runs of consecutive channel indices: each run is one
    library call reading ``signal[:, c0:c0+k]`` in place."""
    runs = []
    for o, c in enumerate(use_channels):
        if runs and runs[-1][1] + runs[-1][2] == c:
            runs[-1][2] += 1
        else:
            runs.append([o, c, 1])
    return runs


# ------------------------------------------------------------------------------------ public API
def sinc_wrapper(sample_at, signal, lowpass, NT):
    """util/resampling.py:21-27.  ``lowpass`` is unused, as in the reference (:79 derives the
    cut-off from the local period)."""
    output = np.empty(len(sample_at), "float32")
    _resample_into(output, sample_at, signal, NT)
    return output


def sinc_wrapper_mt(output, sample_at, signal, lowpass, NT):
    """util/resampling.py:30-46: fills ``output`` in place.  The thread fan-out of the reference
    is replaced by one kernel launch over the whole output range."""
    _resample_into(output, sample_at, signal, NT)


def sinc_core(sample_at, signal, lowpass, output, win_func, N):
    """util/resampling.py:51-90 signature: ``win_func`` must be ``np.hanning(2*NT+1)`` and ``N``
    ``arange(-NT, NT+1)`` as ``sinc_wrapper`` builds them; NT is taken from ``len(N)``."""
    nt = (len(N) - 1) // 2
    if len(win_func) != 2 * nt + 1:
        raise ValueError("win_func and N must both have 2*NT+1 entries")
    _resample_into(output, sample_at, signal, nt)


def speed_to_pos(sampletimes, speeds, num_input_samples):
    """Read positions from a speed curve (util/resampling.py:93-137): float64 ndarray.

    The error-diffused integer segment lengths and the carried segment offsets follow the
    reference's serial float64 operation order (host, in the library); the per-segment
    ``cumsum(1/speed)`` expansion runs on the GPU with the same operation order, so the result is
    bit-identical to the reference on its filled prefix."""
    L = _lib.lib()
    _lib.require_device()
    st = np.ascontiguousarray(sampletimes, dtype=np.float64)
    sp = np.ascontiguousarray(speeds, dtype=np.float64)
    if st.ndim != 1 or st.shape != sp.shape or len(st) < 2:
        raise ValueError("sampletimes and speeds must be 1-D arrays of the same length >= 2")
    k = len(st)
    seg_n = np.empty(k - 1, dtype=np.int64)
    total = np.zeros(1, dtype=np.int64)
    _lib.check(L.par_speed_segments(st.ctypes.data, sp.ctypes.data, k, seg_n.ctypes.data, total.ctypes.data),
               "par_speed_segments")
    cap = int(total[0])
    pos = _lib.pinned_empty((max(cap, 1),), np.float64)
    m = np.zeros(1, dtype=np.int64)
    rc = L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, k, float(num_input_samples),
                                pos.ctypes.data, cap, m.ctypes.data, 0, _lib.device(), None)
    _lib.check(rc, "par_speed_to_pos_f64")
    return pos[:int(m[0])]


def lag_to_pos(lag_curve, sr, num_input_samples):
    """Positions of the lag-curve mode of ``run`` (util/resampling.py:189-206): host numpy
    (a K-point ``np.interp``; the curve has a handful of points)."""
    sampletimes = lag_curve[:, 0] * sr
    lags = lag_curve[:, 1] * sr
    num_output_samples = num_input_samples + abs(lags[-1])
    sample_at = np.interp(np.arange(num_output_samples), sampletimes, sampletimes - lags)
    trim_end = find_cutoff(sample_at, num_input_samples)
    if trim_end is not None:
        logging.debug(f"Trimmed to sample {trim_end[0]}")
        sample_at = sample_at[:trim_end[0]]
    np.clip(sample_at, 0, None, out=sample_at)
    return sample_at


def _mode_name(resampling_mode):
    if resampling_mode not in ("Sinc", "Linear"):
        raise ValueError(f"unknown resampling_mode {resampling_mode!r}")
    return resampling_mode == "Sinc"


def resample_channels(signal, sample_at, use_channels, resampling_mode="Sinc", sinc_quality=50, prog_sig=None):
    """The "Resampling" phase of ``run`` (util/resampling.py:217-231) for given read positions:
    ``signal`` (L, C) float32 -> ``(len(sample_at), len(use_channels))`` float32, interleaved like the
    reference's output array (:222).  Consecutive selected channels go through ONE library call:
    the interleaved input is uploaded as it lies in memory, the tap weights of an output sample are
    computed once and applied to every channel, and the kernel writes the interleaved output."""
    L = _lib.lib()
    _lib.require_device()
    sinc = _mode_name(resampling_mode)
    signal = np.asarray(signal)
    if signal.ndim == 1:
        signal = signal[:, None]
    use_channels = list(use_channels)
    num_channels = len(use_channels)
    m = len(sample_at)
    output = _lib.pinned_empty((max(m, 1), max(num_channels, 1)), np.float32)[:m, :num_channels]
    pos = np.ascontiguousarray(sample_at, dtype=np.float64)
    done = 0
    for o, c0, k in _channel_runs(signal, use_channels):
        keep, ptr, fs, cs = _lib.f32_layout_2d(signal[:, c0:c0 + k])
        # a run that does not cover every output column goes through its own contiguous buffer
        out = output if k == num_channels else _lib.pinned_empty((max(m, 1), k), np.float32)[:m]
        if m:
            args = (pos.ctypes.data, m, ptr, keep.shape[0], fs, k, cs)
            tail = (out.ctypes.data, k, 1, 0, _lib.device(), None)
            if sinc:
                _lib.check(L.par_sinc_resample_f32(*args, int(sinc_quality), *tail), "par_sinc_resample_f32")
            else:
                _lib.check(L.par_linear_resample_f32(*args, *tail), "par_linear_resample_f32")
        if out is not output:
            output[:, o:o + k] = out
        done += k
        if prog_sig:
            prog_sig.notifyProgress.emit(done / num_channels * 100)
    return output


def varispeed(signal, sr, speed_curve, use_channels=None, resampling_mode="Sinc", sinc_quality=50, prog_sig=None):
    """The "Preparing" + "Resampling" phases of ``run`` for a speed curve (util/resampling.py:181-184,
    :217-231) in one library call per run of consecutive channels: the read positions are expanded on
    the device (bit-identical to ``speed_to_pos``) and never copied to the host.
    Returns the ``(M, len(use_channels))`` float32 output array."""
    L = _lib.lib()
    _lib.require_device()
    sinc = _mode_name(resampling_mode)
    signal = np.asarray(signal)
    if signal.ndim == 1:
        signal = signal[:, None]
    if use_channels is None:
        use_channels = range(signal.shape[1])
    use_channels = list(use_channels)
    num_channels = len(use_channels)
    speed_curve = np.asarray(speed_curve, dtype=np.float64)
    st = np.ascontiguousarray(speed_curve[:, 0] * sr)
    sp = np.ascontiguousarray(speed_curve[:, 1])
    k = len(st)
    seg_n = np.empty(max(k - 1, 1), dtype=np.int64)
    total = np.zeros(1, dtype=np.int64)
    _lib.check(L.par_speed_segments(st.ctypes.data, sp.ctypes.data, k, seg_n.ctypes.data, total.ctypes.data),
               "par_speed_segments")
    cap = int(total[0])
    full = _lib.pinned_empty((max(cap, 1), max(num_channels, 1)), np.float32)
    m = np.zeros(1, dtype=np.int64)
    done = 0
    for o, c0, kk in _channel_runs(signal, use_channels):
        keep, ptr, fs, cs = _lib.f32_layout_2d(signal[:, c0:c0 + kk])
        out = full if kk == num_channels else _lib.pinned_empty((max(cap, 1), kk), np.float32)
        rc = L.par_varispeed_f32(st.ctypes.data, sp.ctypes.data, k, ptr, keep.shape[0], fs, kk, cs,
                                 _lib.PAR_MODE_SINC if sinc else _lib.PAR_MODE_LINEAR, int(sinc_quality),
                                 out.ctypes.data, cap, kk, 1, m.ctypes.data, 0, _lib.device(), None)
        _lib.check(rc, "par_varispeed_f32")
        if out is not full:
            full[:int(m[0]), o:o + kk] = out[:int(m[0])]
        done += kk
        if prog_sig:
            prog_sig.notifyProgress.emit(done / max(num_channels, 1) * 100)
    return full[:int(m[0]), :num_channels]


def run(filenames, signal_data=None, speed_curve=None, resampling_mode="Linear", sinc_quality=50, use_channels=(),
        prog_sig=None, lag_curve=None, suffix=""):
    """util/resampling.py:162-240, same arguments, progress signals, log lines and output files.
    Returns None."""
    if prog_sig:
        prog_sig.notifyProgress.emit(0)
    if signal_data is None:
        signal_data = [None for filename in filenames]
    for filename, sig_data in zip(filenames, signal_data):
        with log_duration("Preparing"):
            logging.info(f"Resampling '{os.path.basename(filename)}'... {resampling_mode}, {sinc_quality}, {use_channels}")
            if sig_data:
                signal, sr = sig_data
            else:
                signal, sr, _ = io_ops.read_file(filename)
            signal = np.asarray(signal)
            if signal.ndim == 1:
                signal = signal[:, None]
            sample_at = None
            if speed_curve is None and lag_curve is not None:
                sample_at = lag_to_pos(np.asarray(lag_curve), sr, len(signal))
            elif speed_curve is None:
                raise ValueError("run needs a speed_curve or a lag_curve")
        if use_channels:
            channels = [channel for channel in use_channels if channel < signal.shape[1]]
        else:
            channels = tuple(range(signal.shape[1]))
        with log_duration("Resampling"):
            if sample_at is None:
                output = varispeed(signal, sr, speed_curve, channels, resampling_mode, sinc_quality, prog_sig)
            else:
                output = resample_channels(signal, sample_at, channels, resampling_mode, sinc_quality, prog_sig)
        with log_duration("Writing"):
            out_file_path = f"{os.path.splitext(filename)[0]}_res{suffix}.wav"
            io_ops.write_float_wav(out_file_path, output, sr)
            if prog_sig:
                prog_sig.notifyProgress.emit(100)
    logging.info("Done!")


def timefunc(correct, s, func, *args, **kwargs):
    """Benchmark helper of the reference (util/resampling.py:243-256): min of 2x5 runs, ms."""
    print(s.ljust(20), end=" ")
    res = func(*args, **kwargs)
    if correct is not None:
        assert np.allclose(res, correct), (res, correct)
    best = float("inf")
    for _ in range(2):
        t0 = time()
        for _ in range(5):
            func(*args, **kwargs)
        best = min(best, (time() - t0) / 5)
    print('{:>5.0f} ms'.format(best * 1000))
    return res
