"""Drop-in mirror of the reference's ``util/fourier.py`` module surface, computed on a B200.

Same names, argument meaning and return conventions as the reference (citations are
``util/fourier.py:<line>`` of HENDRIX-ZT2/pyaudiorestoration); the transforms themselves run in
the sm_100a kernels of ``libpar_b200.so`` (``include/par_b200.h``).  Differences, all deliberate
(SURVEY.md 8b):

* no back-end fall-through and no CPU path: the reference logs and swallows back-end errors and
  may return ``None`` (:67-75); here a failure raises ``RuntimeError``;
* ``stft`` always returns an F-contiguous ``(n_freqs, n_frames)`` complex64 ndarray (the
  reference returns a torch tensor, a complex64 or a complex128 ndarray depending on which
  back-end ran); for ``zeropad > 1`` it follows the numpy/pyfftw frame placement (left aligned)
  and the ``1/sqrt(n_fft)`` scaling of :157;
* ``istft`` does not scale its argument in place (the reference does, :359).
"""
import contextlib
import logging
import time

import numpy as np
from scipy import signal as dsp

from .. import _lib

# Constrain STFT block sizes to 256 KB (kept for API compatibility, util/fourier.py:21)
MAX_MEM_BLOCK = 2 ** 8 * 2 ** 10


class ParameterError(ValueError):
    """The reference raises an undefined ``ParameterError`` name (:264); this is it."""


def to_mag(spectrum):
    """util/fourier.py:23-24."""
    return abs(spectrum) + .0000001


def _prep_args(n_fft, step, zeropad):
    n_fft = int(n_fft)
    step = max(n_fft // 2, 1) if step is None else int(step)
    zeropad = 1 if zeropad is None else int(zeropad)
    return n_fft, step, zeropad


def _stft_call(x, n_fft, step, window_name, zeropad, magnitude):
    n_fft, step, zeropad = _prep_args(n_fft, step, zeropad)       # :62-63
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError('x must be 1D')                          # :64-65
    if x.dtype.kind == 'c':
        raise ValueError('x must be real')
    if len(x) < 1:
        raise ValueError('x must not be empty')
    # strided float32 column views (signal[:, ch] of an interleaved array) are passed as they are
    keep, ptr, stride = _lib.f32_layout(x)
    L = _lib.lib()
    _lib.require_device()
    window = np.ascontiguousarray(dsp.get_window(window_name, n_fft), dtype=np.float32)   # :66
    n_frames = int(L.par_stft_num_frames(len(x), n_fft, step))
    n_freqs = (n_fft * zeropad) // 2 + 1
    out = _lib.pinned_empty((n_frames, n_freqs), np.float32 if magnitude else np.complex64)
    flags = _lib.PAR_OUT_MAGNITUDE if magnitude else 0
    rc = L.par_stft_f32(ptr, len(x), stride, 1, 0, n_fft, step, zeropad, window.ctypes.data,
                        out.ctypes.data, n_freqs, 0, flags, _lib.device(), None)
    _lib.check(rc, "par_stft_f32")
    del keep
    return out.T      # (n_freqs, n_frames), F-contiguous like the reference's numpy path (:147)


def stft_multi(signal, n_fft=1024, step=512, window_name='blackmanharris', zeropad=1, magnitude=False):
    """Extension (no reference counterpart): the transform of EVERY channel of a (frames, channels)
    array in one library call / one upload.  Returns ``(channels, n_freqs, n_steps)``; ``out[c]``
    equals ``stft(signal[:, c], ...)`` (or ``get_mag`` with ``magnitude=True``)."""
    n_fft, step, zeropad = _prep_args(n_fft, step, zeropad)
    signal = np.asarray(signal)
    if signal.ndim != 2:
        raise ValueError('signal must be (frames, channels)')
    keep, ptr, fs, cs = _lib.f32_layout_2d(signal)
    frames, channels = keep.shape
    if frames < 1 or channels < 1:
        raise ValueError('signal must not be empty')
    L = _lib.lib()
    _lib.require_device()
    window = np.ascontiguousarray(dsp.get_window(window_name, n_fft), dtype=np.float32)
    n_frames = int(L.par_stft_num_frames(frames, n_fft, step))
    n_freqs = (n_fft * zeropad) // 2 + 1
    out = _lib.pinned_empty((channels, n_frames, n_freqs), np.float32 if magnitude else np.complex64)
    flags = _lib.PAR_OUT_MAGNITUDE if magnitude else 0
    rc = L.par_stft_f32(ptr, frames, fs, channels, cs, n_fft, step, zeropad, window.ctypes.data,
                        out.ctypes.data, n_freqs, n_frames * n_freqs, flags, _lib.device(), None)
    _lib.check(rc, "par_stft_f32")
    del keep
    return out.transpose(0, 2, 1)


SPEC_OPS = {"gate": _lib.PAR_SPEC_GATE, "max": _lib.PAR_SPEC_SELECT_MAX, "min": _lib.PAR_SPEC_SELECT_MIN,
            "max_min": _lib.PAR_SPEC_SELECT_BOTH, "heal": _lib.PAR_SPEC_HEAL}


def stft_mask_istft(signal, op, n_fft=512, step=32, window_name='blackmanharris', params=None, gain_db=0.0):
    """Extension (no single reference counterpart): ``istft(op(stft(fix_length(signal, n + n_fft//2))), length=n,
    hop_length=step)`` for every channel of a ``(frames, channels)`` array with the spectrogram RESIDENT ON THE DEVICE --
    one upload and one download instead of the two host round trips of the reference's tools
    (dropout_healer_gui.py:129-164, dropouts_gui.py:148-159, renoiser_gui.py:310-317).

    op: ``"gate"`` (params = per-bin threshold in dB, gain_db), ``"max"`` / ``"min"`` / ``"max_min"`` (stereo in, 1 or 2
    channels out), ``"heal"`` (params = int64 ``(n, 5)`` marker regions ``(frame_b, frame_a, frame_surrounding, bin_l,
    bin_u)``).  Returns ``(frames, channels_out)`` float32."""
    n_fft, step, _ = _prep_args(n_fft, step, 1)
    signal = np.asarray(signal)
    if signal.ndim == 1:
        signal = signal[:, None]
    keep, ptr, fs, cs = _lib.f32_layout_2d(signal)
    frames, channels = keep.shape
    if frames < 1:
        raise ValueError('signal must not be empty')
    L = _lib.lib()
    _lib.require_device()
    window = np.ascontiguousarray(dsp.get_window(window_name, n_fft), dtype=np.float32)
    code = SPEC_OPS[op]
    n_out = 2 if op == "max_min" else (1 if op in ("max", "min") else channels)
    if op == "gate":
        p = np.ascontiguousarray(params, dtype=np.float64)
        pptr, pn = p.ctypes.data, len(p)
    elif op == "heal":
        p = np.ascontiguousarray(params, dtype=np.int64).reshape(-1, 5)
        pptr, pn = (p.ctypes.data if len(p) else None), len(p)
    else:
        p, pptr, pn = None, None, 0
    out = _lib.pinned_empty((frames, n_out), np.float32)
    rc = L.par_spectral_process_f32(ptr, frames, fs, channels, cs, n_fft, step, window.ctypes.data, window.ctypes.data, code,
                                    pptr, pn, float(gain_db), out.ctypes.data, n_out, 1, 0, _lib.device(), None)
    _lib.check(rc, "par_spectral_process_f32")
    del keep, p
    return out


def stft(x, n_fft=1024, step=512, window_name='blackmanharris', zeropad=1):
    """Compute the STFT (util/fourier.py:37-75).

    x : 1-D real array-like (any stride / real dtype).  Returns ndarray (n_freqs, n_steps)
    complex64, ``n_freqs = n_fft*zeropad//2 + 1``, ``n_steps = len(x)//step + 1``.

    Size restriction (the reference accepts any ``n_fft``): ``n_fft * zeropad`` must be a power of two in
    [32, 1048576] -- every size the reference's GUIs offer (util/widgets.py:333-335); anything else raises
    ``RuntimeError`` (``PAR_EUNSUPPORTED``), there is no CPU fallback.
    """
    with timed_log("b200"):
        return _stft_call(x, n_fft, step, window_name, zeropad, False)


def get_mag(*args, **kwargs):
    """Magnitude spectrum ``abs(stft(...)) + 1e-7`` (util/fourier.py:27-29), fused into the
    transform kernel: float32, half the output bytes of the complex result."""
    def _args(x, n_fft=1024, step=512, window_name='blackmanharris', zeropad=1):
        return x, n_fft, step, window_name, zeropad
    with timed_log("b200"):
        return _stft_call(*_args(*args, **kwargs), True)


@contextlib.contextmanager
def timed_log(method_name):
    """util/fourier.py:85-89."""
    start = time.time()
    yield
    logging.info(f"{method_name} {time.time() - start:0.2f}s")


def dtype_r2c(d, default=np.complex64):
    """util/fourier.py:169-199."""
    mapping = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}
    dt = np.dtype(d)
    if dt.kind == 'c':
        return dt
    return np.dtype(mapping.get(dt, default))


def dtype_c2r(d, default=np.float32):
    """util/fourier.py:202-233."""
    mapping = {np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64,
               np.dtype(complex): float}
    dt = np.dtype(d)
    if dt.kind == 'f':
        return dt
    return np.dtype(mapping.get(dt, default))


def pad_center(data, size, axis=-1, **kwargs):
    """util/fourier.py:236-277."""
    kwargs.setdefault('mode', 'constant')
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ParameterError(f'Target size ({size:d}) must be at least input size ({n:d})')
    return np.pad(data, lengths, **kwargs)


def tiny(x):
    """util/fourier.py:280-311."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.float32
    return np.finfo(dtype).tiny


def istft(stft_matrix, hop_length=None, win_length=None, window_name='blackmanharris', center=True,
          dtype=None, length=None):
    """Inverse STFT (util/fourier.py:314-437): irfft of every column, synthesis window,
    overlap-add, division by the window sum-square envelope, centre trim / ``length`` fix.

    The arithmetic is float32 on the device for every input dtype; a complex128 argument still
    yields a float64 array like the reference, but carries float32 accuracy.
    Size restriction: ``n_fft = 2 * (n_freqs - 1)`` must be a power of two in [32, 32768] (``RuntimeError`` otherwise).
    """
    stft_matrix = np.asarray(stft_matrix)
    if stft_matrix.ndim != 2:
        raise ValueError('stft_matrix must be 2D (n_freqs, n_frames)')
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    hop_length = int(hop_length)
    window = dsp.get_window(window_name, win_length, fftbins=True)
    window = pad_center(window, n_fft)
    if length:
        padded_length = length + int(n_fft) if center else length
        n_frames = min(stft_matrix.shape[1], int(np.ceil(padded_length / hop_length)))
    else:
        n_frames = stft_matrix.shape[1]
    expected_signal_len = n_fft + hop_length * (n_frames - 1)
    if dtype is None:
        dtype = dtype_c2r(stft_matrix.dtype)
    if length is None:
        start = int(n_fft // 2) if center else 0
        out_len = expected_signal_len - 2 * start
    else:
        start = int(n_fft // 2) if center else 0
        out_len = int(length)
    if n_frames < 1 or out_len <= 0:
        return np.zeros(max(out_len, 0), dtype=dtype)
    # memory image the library expects: frames contiguous, bins fastest
    frames = np.ascontiguousarray(stft_matrix[:, :n_frames].T, dtype=np.complex64)
    L = _lib.lib()
    _lib.require_device()
    win32 = np.ascontiguousarray(window, dtype=np.float32)
    y = _lib.pinned_empty((out_len,), np.float32)
    n_freqs = n_fft // 2 + 1
    rc = L.par_istft_f32(frames.ctypes.data, n_fft, n_frames, n_freqs, 1, 0, hop_length,
                         win32.ctypes.data, start, out_len, y.ctypes.data, 1, 0, 0, _lib.device(), None)
    _lib.check(rc, "par_istft_f32")
    return y if np.dtype(dtype) == np.float32 else y.astype(dtype)


def fix_length(data, size, axis=-1, **kwargs):
    """util/fourier.py:440-478."""
    kwargs.setdefault('mode', 'constant')
    n = data.shape[axis]
    if n > size:
        slices = [slice(None)] * data.ndim
        slices[axis] = slice(0, size)
        return data[tuple(slices)]
    elif n < size:
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (0, size - n)
        return np.pad(data, lengths, **kwargs)
    return data


def window_sumsquare(window_name, n_frames, hop_length=512, win_length=None, n_fft=2048,
                     dtype=np.float32, norm=None):
    """Sum-square envelope of a window at a hop length (util/fourier.py:492-546); host numpy,
    kept for API compatibility -- ``istft`` computes its envelope inside the overlap-add kernel."""
    if win_length is None:
        win_length = n_fft
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = dsp.get_window(window_name, win_length)
    win_sq = normalize(win_sq, norm=norm) ** 2
    win_sq = pad_center(win_sq, n_fft)
    for i in range(n_frames):
        sample = i * hop_length
        x[sample:min(n, sample + n_fft)] += win_sq[:max(0, min(n_fft, n - sample))]
    return x


def normalize(S, norm=np.inf, axis=0, threshold=None, fill=None):
    """util/fourier.py:549-674 (host numpy helper of window_sumsquare)."""
    if threshold is None:
        threshold = tiny(S)
    elif threshold <= 0:
        raise ParameterError(f'threshold={threshold} must be strictly positive')
    if fill not in [None, False, True]:
        raise ParameterError(f'fill={fill} must be None or boolean')
    if not np.all(np.isfinite(S)):
        raise ParameterError('Input must be finite')
    mag = np.abs(S).astype(float)
    fill_norm = 1
    if norm is None:
        return S
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == -np.inf:
        length = np.min(mag, axis=axis, keepdims=True)
    elif norm == 0:
        if fill is True:
            raise ParameterError('Cannot normalize with norm=0 and fill=True')
        length = np.sum(mag > 0, axis=axis, keepdims=True, dtype=mag.dtype)
    elif np.issubdtype(type(norm), np.number) and norm > 0:
        length = np.sum(mag ** norm, axis=axis, keepdims=True) ** (1. / norm)
        fill_norm = (mag.size if axis is None else mag.shape[axis]) ** (-1. / norm)
    else:
        raise ParameterError(f'Unsupported norm: {norm!r}')
    small_idx = length < threshold
    Snorm = np.empty_like(S)
    if fill is None:
        length[small_idx] = 1.0
        Snorm[:] = S / length
    elif fill:
        length[small_idx] = np.nan
        Snorm[:] = S / length
        Snorm[np.isnan(Snorm)] = fill_norm
    else:
        length[small_idx] = np.inf
        Snorm[:] = S / length
    return Snorm


def fft_freqs(n_fft, fs):
    """Frequencies of the DFT bins (util/fourier.py:690-700)."""
    return np.arange(0, (n_fft // 2 + 1)) / float(n_fft) * float(fs)
