"""TEST DOUBLE of libpar_b200's compute entry points, for the CPU-only host-logic tests.

The product has no CPU path (see test_host_cpu.py::test_no_device_means_loud_failure).  To exercise
the Python mirror of the reference API -- argument coercion, strides / pitches / channel runs handed
to the C ABI, file naming, progress signals, result views -- on a machine without a GPU, the tests
monkeypatch ``pyaudiorestoration_b200._lib.lib`` with this object: it decodes the raw pointers and
strides exactly as include/par_b200.h specifies them and computes with the CPU oracle.  It lives in
tests/ only and is never importable from the package.  Host-only entry points
(par_speed_segments, par_stft_num_frames, par_host_alloc ...) are forwarded to the real library.
"""
import ctypes

import numpy as np
from numpy.lib.stride_tricks import as_strided

import oracle
from oracle import oracle_np as onp

PAR_OUT_MAGNITUDE = 2


def _view(ptr, dtype, shape, strides_elems):
    dtype = np.dtype(dtype)
    if int(np.prod(shape)) == 0:
        return np.zeros(shape, dtype)
    extent = 1 + sum((s - 1) * abs(st) for s, st in zip(shape, strides_elems))
    buf = (ctypes.c_char * (extent * dtype.itemsize)).from_address(int(ptr))
    base = np.frombuffer(buf, dtype=dtype, count=extent)
    return as_strided(base, shape=shape, strides=[st * dtype.itemsize for st in strides_elems])


def _ptr(p):
    return p if isinstance(p, int) else (p.value if hasattr(p, "value") else int(p or 0))


class AbiDouble:
    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):                      # host-only entries and diagnostics
        return getattr(self._real, name)

    def par_device_count(self):
        return 1

    # pinned allocations need the CUDA driver: plain malloc here
    _libc = ctypes.CDLL(None)
    _libc.malloc.restype = ctypes.c_void_p
    _libc.malloc.argtypes = [ctypes.c_size_t]
    _libc.free.argtypes = [ctypes.c_void_p]

    def par_host_alloc(self, nbytes):
        return self._libc.malloc(max(int(nbytes), 16))

    def par_host_free(self, p):
        self._libc.free(p)

    def par_stft_f32(self, x, n, x_stride, n_ch, x_ch_stride, n_fft, hop, zeropad, window, out, out_pitch,
                     out_ch_stride, flags, device, stream):
        self.calls.append(("par_stft_f32", n, x_stride, n_ch, x_ch_stride, out_pitch, out_ch_stride, flags))
        win = _view(_ptr(window), np.float32, (n_fft,), (1,))
        T = n // hop + 1
        F = n_fft * zeropad // 2 + 1
        mag = bool(flags & PAR_OUT_MAGNITUDE)
        for c in range(n_ch):
            xc = np.ascontiguousarray(_view(_ptr(x) + 4 * c * x_ch_stride, np.float32, (n,), (x_stride,)))
            xp = np.pad(xc, n_fft // 2, mode="reflect").astype(np.float64)
            idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
            S = np.fft.rfft(xp[idx] * win.astype(np.float64)[None, :], n=n_fft * zeropad, axis=1) / np.sqrt(n_fft)
            if mag:
                o = _view(_ptr(out) + 4 * c * out_ch_stride, np.float32, (T, F), (out_pitch, 1))
                o[...] = (np.abs(S) + 1e-7).astype(np.float32)
            else:
                o = _view(_ptr(out) + 8 * c * out_ch_stride, np.complex64, (T, F), (out_pitch, 1))
                o[...] = S.astype(np.complex64)
        return 0

    def par_istft_f32(self, S, n_fft, n_frames, s_pitch, n_ch, s_ch_stride, hop, window, start, length, y, y_stride,
                      y_ch_stride, flags, device, stream):
        self.calls.append(("par_istft_f32", n_fft, n_frames, s_pitch, hop, start, length))
        F = n_fft // 2 + 1
        win = _view(_ptr(window), np.float32, (n_fft,), (1,)).astype(np.float64)
        for c in range(n_ch):
            frames = _view(_ptr(S) + 8 * c * s_ch_stride, np.complex64, (n_frames, F), (s_pitch, 1)).astype(np.complex128)
            t = np.fft.irfft(frames * np.sqrt(n_fft), n=n_fft, axis=1) * win[None, :]
            total = n_fft + hop * (n_frames - 1)
            acc, wss = np.zeros(total), np.zeros(total)
            for i in range(n_frames):
                acc[i * hop:i * hop + n_fft] += t[i]
                wss[i * hop:i * hop + n_fft] += win ** 2
            ok = wss > np.finfo(np.float32).tiny
            acc[ok] /= wss[ok]
            seg = np.zeros(length)
            avail = max(0, min(length, total - start))
            seg[:avail] = acc[start:start + avail]
            _view(_ptr(y) + 4 * c * y_ch_stride, np.float32, (length,), (y_stride,))[...] = seg.astype(np.float32)
        return 0

    def par_spectral_process_f32(self, x, n, x_stride, n_ch, x_ch_stride, n_fft, hop, window, syn_window, op, params,
                                 n_params, gain_db, y, y_stride, y_ch_stride, flags, device, stream):
        """stft -> mask -> istft of include/par_b200.h, with the oracle's transforms and float64 masks."""
        self.calls.append(("par_spectral_process_f32", n, x_stride, n_ch, x_ch_stride, n_fft, hop, op, n_params, y_stride,
                           y_ch_stride))
        F = n_fft // 2 + 1
        specs = []
        for c in range(n_ch):
            xc = np.ascontiguousarray(_view(_ptr(x) + 4 * c * x_ch_stride, np.float32, (n,), (x_stride,)))
            pad = onp.fix_length(xc, n + n_fft // 2)
            specs.append(np.array(onp.stft_f64(pad, n_fft, hop)).T)                  # (F, T) complex128
        db = lambda s_: 20 * np.log10(np.abs(s_) + 1e-7)                              # noqa: E731
        if op == 0:                                                                   # gate
            thr = _view(_ptr(params), np.float64, (F,), (1,))
            outs = [s_ * np.where(db(s_) > thr[:, None], 1.0, np.float32(10 ** (gain_db / 20))) for s_ in specs]
        elif op in (1, 2, 3):                                                         # select max / min / both
            l, r = specs
            mx, mn = np.where(np.abs(l) > np.abs(r), l, r), np.where(np.abs(l) < np.abs(r), l, r)
            outs = [mx] if op == 1 else ([mn] if op == 2 else [mx, mn])
        else:                                                                         # heal
            regs = _view(_ptr(params), np.int64, (n_params, 5), (5, 1)) if n_params else np.zeros((0, 5), np.int64)
            outs = []
            for s_ in specs:
                d, gain = db(s_), np.zeros(s_.shape)
                for fb, fa, ar, bl, bu in regs:
                    before, after = np.mean(d[bl:bu, fb - ar:fb], axis=1), np.mean(d[bl:bu, fa:fa + ar], axis=1)
                    yv = (np.linspace(fb, fa, num=fa - fb) - fb) / (fa - fb)
                    target = before[:, None] * (1 - yv)[None, :] + after[:, None] * yv[None, :]
                    boost = np.clip(target - d[bl:bu, fb:fa], gain[bl:bu, fb:fa], 255)
                    gain[bl:bu, fb:fa] = boost
                outs.append(s_ * 10 ** (gain / 20))
        for c, s_ in enumerate(outs):
            yy = onp.istft_ref(s_, hop_length=hop, length=n)
            _view(_ptr(y) + 4 * c * y_ch_stride, np.float32, (n,), (y_stride,))[...] = np.asarray(yy).astype(np.float32)
        return 0

    def par_speed_to_pos_f64(self, st, sp, k, n_in, pos, cap, m, flags, device, stream):
        stv = _view(_ptr(st), np.float64, (k,), (1,))
        spv = _view(_ptr(sp), np.float64, (k,), (1,))
        p = oracle.speed_to_pos_c(stv, spv, n_in)
        _view(_ptr(m), np.int64, (1,), (1,))[0] = len(p)
        if len(p) > cap:
            return -4
        _view(_ptr(pos), np.float64, (len(p),), (1,))[...] = p
        return 0

    def _resample(self, sinc, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out, out_stride, out_ch_stride):
        p = np.array(_view(_ptr(pos), np.float64, (m,), (1,)))
        for c in range(n_ch):
            x = np.ascontiguousarray(_view(_ptr(signal) + 4 * c * sig_ch_stride, np.float32, (n_in,), (sig_stride,)))
            y = oracle.sinc_c(p, x, nt) if sinc else onp.linear_resample(p, x)
            _view(_ptr(out) + 4 * c * out_ch_stride, np.float32, (m,), (out_stride,))[...] = y
        return 0

    def par_sinc_resample_f32(self, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out, out_stride,
                              out_ch_stride, flags, device, stream):
        self.calls.append(("par_sinc_resample_f32", m, n_in, sig_stride, n_ch, sig_ch_stride, nt, out_stride, out_ch_stride))
        return self._resample(True, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out, out_stride, out_ch_stride)

    def par_linear_resample_f32(self, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, out, out_stride,
                                out_ch_stride, flags, device, stream):
        self.calls.append(("par_linear_resample_f32", m, n_in, sig_stride, n_ch, sig_ch_stride, out_stride, out_ch_stride))
        return self._resample(False, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, 1, out, out_stride, out_ch_stride)

    def par_varispeed_f32(self, st, sp, k, signal, n_in, sig_stride, n_ch, sig_ch_stride, mode, nt, out, out_cap,
                          out_stride, out_ch_stride, m, flags, device, stream):
        self.calls.append(("par_varispeed_f32", k, n_in, sig_stride, n_ch, sig_ch_stride, mode, nt, out_cap, out_stride,
                           out_ch_stride))
        stv = _view(_ptr(st), np.float64, (k,), (1,))
        spv = _view(_ptr(sp), np.float64, (k,), (1,))
        p = np.ascontiguousarray(oracle.speed_to_pos_c(stv, spv, n_in))
        _view(_ptr(m), np.int64, (1,), (1,))[0] = len(p)
        if len(p) > out_cap:
            return -4
        return self._resample(mode == 1, p.ctypes.data, len(p), signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out,
                              out_stride, out_ch_stride)

    def par_trace_f32(self, mag, num_bins, n_frames, pitch, frame0, count, fft_size, sr, tolerance_st, mode, freqs,
                      flags, device, stream):
        self.calls.append(("par_trace_f32", num_bins, n_frames, pitch, frame0, count, fft_size, mode))
        spec = _view(_ptr(mag), np.float32, (n_frames, num_bins), (pitch, 1)).T
        f = _view(_ptr(freqs), np.float64, (count,), (1,))
        tol = tolerance_st / 12
        if mode == 2:
            fft_freqs = np.arange(0, (fft_size // 2 + 1)) / float(fft_size) * float(sr)
            nl, nu = onp._bin_limits(f[0], tol, num_bins, fft_size, sr)
            for i in range(count):
                w = np.hanning(nu - nl) * spec[nl:nu, frame0 + i]
                f[i] = 2 ** (np.sum(w * np.log2(fft_freqs[nl:nu])) / np.sum(w))
                nl, nu = onp._bin_limits(f[i], tol, num_bins, fft_size, sr)
        else:
            first = float(f[0])
            for i in range(count):
                centre, t = (f[i], tol) if mode == 0 else (first, tol / 2 if i > 2 else tol)
                nl, nu = onp._bin_limits(centre, t, num_bins, fft_size, sr)
                f[i] = onp._get_peak(spec, frame0 + i, nl, nu, fft_size, sr)
        return 0
