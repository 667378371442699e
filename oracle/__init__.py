"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle_np.py / oracle.c headers).

``oracle.np_`` = numpy restatements; ``oracle.c`` = ctypes bindings of liboracle.so (the
float64 sinc interpolator with the reference's thread fan-out, and speed_to_pos).
The product package never imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import oracle_np as np_  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i64, dp, fp = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p
        L.oracle_sinc_f64.argtypes = [dp, i64, fp, i64, i64, ctypes.c_int, fp, fp, i64]
        L.oracle_sinc_f64.restype = None
        L.oracle_sinc_mt.argtypes = [dp, i64, fp, i64, i64, ctypes.c_int, fp, fp, i64,
                                     ctypes.c_int]
        L.oracle_sinc_mt.restype = ctypes.c_int
        L.oracle_speed_to_pos.argtypes = [dp, dp, i64, ctypes.c_double, dp, i64]
        L.oracle_speed_to_pos.restype = i64
        L.oracle_speed_to_pos_at.argtypes = [dp, dp, i64, ctypes.c_double, dp, i64, dp]
        L.oracle_speed_to_pos_at.restype = i64
        L.oracle_sinc_windows.argtypes = [dp, dp, i64, fp, i64, ctypes.c_int, fp, fp]
        L.oracle_sinc_windows.restype = None
        _LIB = L
    return _LIB


def sinc_c(sample_at, signal, nt, nthreads=1):
    """util/resampling.py:21-46 (sinc_wrapper / sinc_wrapper_mt) through oracle.c."""
    sample_at = np.ascontiguousarray(sample_at, dtype=np.float64)
    signal = np.ascontiguousarray(signal, dtype=np.float32)
    win = np.hanning(2 * nt + 1).astype(np.float32)
    out = np.empty(len(sample_at), dtype=np.float32)
    L = lib()
    if nthreads <= 1:
        L.oracle_sinc_f64(sample_at.ctypes.data, len(sample_at), signal.ctypes.data,
                          len(signal), 1, int(nt), win.ctypes.data, out.ctypes.data, 1)
    else:
        rc = L.oracle_sinc_mt(sample_at.ctypes.data, len(sample_at), signal.ctypes.data,
                              len(signal), 1, int(nt), win.ctypes.data, out.ctypes.data, 1,
                              int(nthreads))
        if rc != 0:
            raise RuntimeError("oracle_sinc_mt: thread creation failed")
    return out


def speed_to_pos_c(sampletimes, speeds, num_input_samples):
    """util/resampling.py:93-137 through oracle.c (filled prefix only)."""
    st = np.ascontiguousarray(sampletimes, dtype=np.float64)
    sp = np.ascontiguousarray(speeds, dtype=np.float64)
    cap = int(np.mean(sp) * (st[-1] - st[0]) * 1.01) + 4096
    out = np.empty(cap, dtype=np.float64)
    n = lib().oracle_speed_to_pos(st.ctypes.data, sp.ctypes.data, len(sp),
                                  float(num_input_samples), out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("oracle_speed_to_pos: capacity exceeded")
    return out[:n].copy()


def speed_to_pos_at_c(sampletimes, speeds, num_input_samples, idx):
    """The positions util/resampling.py:93-137 would return at the ascending output indices ``idx`` (NaN past the
    end), and the total number of valid positions -- without materialising the whole array."""
    st = np.ascontiguousarray(sampletimes, dtype=np.float64)
    sp = np.ascontiguousarray(speeds, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    if len(idx) > 1 and np.any(np.diff(idx) < 0):
        raise ValueError("idx must ascend")
    out = np.empty(len(idx), dtype=np.float64)
    m = lib().oracle_speed_to_pos_at(st.ctypes.data, sp.ctypes.data, len(sp), float(num_input_samples),
                                     idx.ctypes.data, len(idx), out.ctypes.data)
    return out, int(m)


def sinc_windows_c(p, pnext, windows, nt):
    """util/resampling.py:51-90 for isolated outputs: row j of ``windows`` holds the input samples around output j,
    ``p[j]`` / ``pnext[j]`` its read position and the next one, relative to the row's first sample."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    pnext = np.ascontiguousarray(pnext, dtype=np.float64)
    windows = np.ascontiguousarray(windows, dtype=np.float32)
    win = np.hanning(2 * nt + 1).astype(np.float32)
    out = np.empty(len(p), dtype=np.float32)
    lib().oracle_sinc_windows(p.ctypes.data, pnext.ctypes.data, len(p), windows.ctypes.data, windows.shape[1], int(nt),
                              win.ctypes.data, out.ctypes.data)
    return out
