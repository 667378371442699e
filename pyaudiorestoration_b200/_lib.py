"""ctypes binding of libpar_b200.so (include/par_b200.h).  Loading is lazy; a missing library or
device is a loud RuntimeError -- never a silent CPU path."""
import ctypes
import os
import subprocess
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpar_b200.so")
_lock = threading.Lock()
_lib = None

PAR_DEVICE_PTRS = 1 << 0
PAR_OUT_MAGNITUDE = 1 << 1
PAR_SINC_ALIGNED_EDGES = 1 << 2
PAR_SINC_KERNEL_TILED = 1 << 3
PAR_SINC_KERNEL_WS = 1 << 4
PAR_ECAPACITY = -4
PAR_MODE_LINEAR = 0
PAR_MODE_SINC = 1
PAR_TRACE_PEAK, PAR_TRACE_PEAK_TRACK, PAR_TRACE_COG = 0, 1, 2
PAR_SPEC_GATE, PAR_SPEC_SELECT_MAX, PAR_SPEC_SELECT_MIN, PAR_SPEC_SELECT_BOTH, PAR_SPEC_HEAL = 0, 1, 2, 3, 4

EXPORTS = (
    "par_last_error", "par_version", "par_device_count", "par_kernel_launch_count",
    "par_last_kernel_ms", "par_selftest_positions_quotient", "par_release_cached_memory", "par_host_alloc", "par_host_free", "par_stft_num_frames", "par_stft_f32",
    "par_istft_f32", "par_speed_segments", "par_speed_to_pos_f64", "par_sinc_resample_f32",
    "par_linear_resample_f32", "par_varispeed_f32", "par_stft_range_f32", "par_resample_range_f32", "par_speed_to_pos_range_f64", "par_segment_sums_f64", "par_speed_to_pos_range_sums_f64", "par_spectral_process_f32", "par_trace_f32", "par_stft_trace_f32",
)


def library_path():
    return _SO


def build(verbose=False):
    """Compile libpar_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], stdout=out)
    return _SO


def _declare(L):
    c = ctypes
    i64, vp, u32, i32, dbl = c.c_int64, c.c_void_p, c.c_uint, c.c_int, c.c_double
    L.par_last_error.restype = c.c_char_p
    L.par_version.restype = c.c_char_p
    L.par_device_count.restype = i32
    L.par_kernel_launch_count.restype = i64
    L.par_last_kernel_ms.restype = dbl
    L.par_selftest_positions_quotient.restype = i64
    L.par_selftest_positions_quotient.argtypes = [i64, i32]
    L.par_release_cached_memory.restype = i32
    L.par_release_cached_memory.argtypes = [i32]
    L.par_host_alloc.restype = vp
    L.par_host_alloc.argtypes = [i64]
    L.par_host_free.restype = None
    L.par_host_free.argtypes = [vp]
    L.par_stft_num_frames.restype = i64
    L.par_stft_num_frames.argtypes = [i64, i32, i32]
    L.par_stft_f32.restype = i32
    L.par_stft_f32.argtypes = [vp, i64, i64, i32, i64, i32, i32, i32, vp, vp, i64, i64, u32, i32, vp]
    L.par_istft_f32.restype = i32
    L.par_istft_f32.argtypes = [vp, i32, i64, i64, i32, i64, i32, vp, i64, i64, vp, i64, i64, u32, i32, vp]
    L.par_speed_segments.restype = i32
    L.par_speed_segments.argtypes = [vp, vp, i64, vp, vp]
    L.par_speed_to_pos_f64.restype = i32
    L.par_speed_to_pos_f64.argtypes = [vp, vp, i64, dbl, vp, i64, vp, u32, i32, vp]
    L.par_sinc_resample_f32.restype = i32
    L.par_sinc_resample_f32.argtypes = [vp, i64, vp, i64, i64, i32, i64, i32, vp, i64, i64, u32, i32, vp]
    L.par_linear_resample_f32.restype = i32
    L.par_linear_resample_f32.argtypes = [vp, i64, vp, i64, i64, i32, i64, vp, i64, i64, u32, i32, vp]
    L.par_stft_range_f32.restype = i32
    L.par_stft_range_f32.argtypes = [vp, i64, i64, i64, i32, i64, i32, i32, i32, vp, i64, i64, vp, i64, i64, u32, i32, vp]
    L.par_resample_range_f32.restype = i32
    L.par_resample_range_f32.argtypes = [vp, i64, i64, i64, i64, i64, vp, i64, i64, i64, i32, i64, i32, i32, vp, i64,
                                         i64, u32, i32, vp]
    L.par_speed_to_pos_range_f64.restype = i32
    L.par_speed_to_pos_range_f64.argtypes = [vp, vp, i64, dbl, dbl, dbl, vp, i64, vp, vp, vp, u32, i32, vp]
    L.par_spectral_process_f32.restype = i32
    L.par_spectral_process_f32.argtypes = [vp, i64, i64, i32, i64, i32, i32, vp, vp, i32, vp, i64, dbl, vp, i64, i64, u32, i32, vp]
    L.par_segment_sums_f64.restype = i32
    L.par_segment_sums_f64.argtypes = [vp, vp, i64, i64, i64, vp, vp, u32, i32, vp]
    L.par_speed_to_pos_range_sums_f64.restype = i32
    L.par_speed_to_pos_range_sums_f64.argtypes = [vp, vp, i64, dbl, dbl, dbl, vp, vp, vp, i64, vp, vp, vp, u32, i32, vp]
    L.par_trace_f32.restype = i32
    L.par_trace_f32.argtypes = [vp, i32, i64, i64, i64, i64, i32, dbl, dbl, i32, vp, u32, i32, vp]
    L.par_stft_trace_f32.restype = i32
    L.par_stft_trace_f32.argtypes = [vp, i64, i64, i32, i32, i32, vp, i64, i64, dbl, dbl, i32, vp, u32, i32, vp]
    L.par_varispeed_f32.restype = i32
    L.par_varispeed_f32.argtypes = [vp, vp, i64, vp, i64, i64, i32, i64, i32, i32, vp, i64, i64, i64, vp, u32,
                                    i32, vp]


def lib():
    """The loaded library.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(_SO):
                    raise RuntimeError(
                        f"{_SO} is missing: build it with pyaudiorestoration_b200.build() "
                        "(nvcc, sm_100a). There is no CPU fallback.")
                L = ctypes.CDLL(_SO)
                _declare(L)
                _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().par_last_error()
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def device():
    """CUDA device used by the Python layer: $PAR_B200_DEVICE, else $LOCAL_RANK, else 0."""
    for k in ("PAR_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(k)
        if v not in (None, ""):
            return int(v)
    return 0


def require_device():
    n = lib().par_device_count()
    if n <= 0:
        raise RuntimeError("pyaudiorestoration_b200: no CUDA device visible; this path has no CPU fallback")
    return n


# ---- pinned host buffers --------------------------------------------------------------------------
class _Pinned:
    """A cudaHostAlloc'ed block exposed through the array interface; returned to a small
    free-list when the last ndarray viewing it dies."""
    _free = {}
    _free_lock = threading.Lock()
    _MAX_CACHED = 4

    def __init__(self, nbytes):
        self.nbytes = int(max(nbytes, 16))
        self.ptr = None
        with _Pinned._free_lock:
            lst = _Pinned._free.get(self.nbytes)
            if lst:
                self.ptr = lst.pop()
        if self.ptr is None:
            self.ptr = lib().par_host_alloc(self.nbytes)
            if not self.ptr:
                check(-2, "par_host_alloc")
        weakref.finalize(self, _Pinned._release, self.ptr, self.nbytes)

    @staticmethod
    def _release(ptr, nbytes):
        with _Pinned._free_lock:
            lst = _Pinned._free.setdefault(nbytes, [])
            if len(lst) < _Pinned._MAX_CACHED:
                lst.append(ptr)
                return
        try:
            lib().par_host_free(ptr)
        except Exception:
            pass

    def array(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        assert n <= self.nbytes
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        # keep this object (and with it the allocation) alive as long as any view exists
        buf._pinned_owner = self
        return arr


def release_cached_memory():
    """Return the library's cached device scratch and pinned host buffers to the driver."""
    with _Pinned._free_lock:
        cached, _Pinned._free = _Pinned._free, {}
    for ptrs in cached.values():
        for p in ptrs:
            lib().par_host_free(p)
    check(lib().par_release_cached_memory(device()), "par_release_cached_memory")


def pinned_empty(shape, dtype):
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    dtype = np.dtype(dtype)
    return _Pinned(int(np.prod(shape)) * dtype.itemsize).array(shape, dtype)


def f32_layout(a):
    """(array to keep alive, data pointer, element stride) of a 1-D real array for the C ABI:
    float32 views with a positive element stride are passed as they lie in memory (no host copy --
    the library uploads the interleaved span and reads it strided); anything else is converted."""
    a = np.asarray(a)
    if a.ndim != 1:
        raise ValueError("expected a 1-D array")
    ok = a.dtype == np.float32 and (len(a) <= 1 or (a.strides[0] > 0 and a.strides[0] % 4 == 0))
    if not ok:
        a = np.ascontiguousarray(a, dtype=np.float32)
    stride = a.strides[0] // 4 if len(a) > 1 else 1
    return a, a.ctypes.data, max(int(stride), 1)


def f32_layout_2d(a):
    """(array, pointer, frame stride, channel stride) of a (frames, channels) float32 array, in
    elements; copies only when the strides are not positive multiples of 4 bytes."""
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("expected a (frames, channels) array")
    def good(st, n):
        return n <= 1 or (st > 0 and st % 4 == 0)
    if a.dtype != np.float32 or not good(a.strides[0], a.shape[0]) or not good(a.strides[1], a.shape[1]):
        a = np.ascontiguousarray(a, dtype=np.float32)
    fs = a.strides[0] // 4 if a.shape[0] > 1 else max(a.shape[1], 1)
    cs = a.strides[1] // 4 if a.shape[1] > 1 else 1
    return a, a.ctypes.data, max(int(fs), 1), max(int(cs), 0)
