"""CPU oracle for the STFT + varispeed-resample hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing in ``pyaudiorestoration_b200`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker / the reported baseline.

Every function restates one function of the reference (HENDRIX-ZT2/pyaudiorestoration,
paths relative to the reference root) and cites the lines it follows.  The reference
holds no golden vectors or tests for this path, so the oracle is
pinned by fixtures generated from the reference itself, imported and run unmodified in
the authoring container: ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``
(checked by ``tests/test_oracle_golden.py``).

Third-party arithmetic the reference delegates to (not vendored in the reference):
numpy pocketfft (``np.fft.rfft/irfft``, numpy 2.3.5 in this image, unpinned in the
reference's requirements.txt), ``scipy.signal.get_window`` (scipy 1.18.1 here) and
``np.sinc`` / ``np.hanning``.  The oracle calls the same numpy/scipy entry points.
"""
import numpy as np
from scipy import signal as dsp

__all__ = [
    "get_window_f32", "reflect_pad", "stft_ref", "stft_f64", "to_mag", "istft_ref",
    "window_sumsquare", "fix_length", "speed_to_pos", "speed_segments",
    "sinc_resample", "hanning_f32", "linear_resample", "lag_to_positions", "noise_gate_ref",
]


# --------------------------------------------------------------------------- STFT

def get_window_f32(window_name, n_fft):
    """Periodic scipy window cast to float32 -- util/fourier.py:66."""
    return dsp.get_window(window_name, int(n_fft)).astype(np.float32)


def reflect_pad(x, n_fft):
    """util/fourier.py:78-82 (estimate_and_center): np.pad(x, n_fft//2, 'reflect')."""
    return np.pad(x, int(n_fft // 2), mode="reflect")


def n_frames(length, n_fft, step):
    """Frame count of the centred transform -- util/fourier.py:81."""
    return (length + 2 * (n_fft // 2) - n_fft) // step + 1


def stft_ref(x, n_fft=1024, step=512, window_name="blackmanharris", zeropad=1):
    """The reference's numpy back-end, same operations in the same precision.

    util/fourier.py:37-75 (stft: int coercion, 1-D check, float32 periodic window) and
    :136-157 (np_rfft_pick: reflect pad, per-frame ``np.fft.rfft(window * frame, n=N*Z)``
    for n_fft > 512, the vectorised segment_array path :160-166 otherwise, then
    ``/ sqrt(n_fft)``).  Result is (N*Z/2+1, T), complex128 after the division (numpy 2.x).
    """
    n_fft = int(n_fft)
    step = max(n_fft // 2, 1) if step is None else int(step)
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("x must be 1D")
    window = get_window_f32(window_name, n_fft)
    xp = reflect_pad(x, n_fft)
    t = (len(xp) - n_fft) // step + 1
    nz = n_fft * zeropad
    if n_fft > 512:
        cdt = np.complex64 if xp.dtype == np.float32 else (
            np.complex128 if xp.dtype == np.float64 else np.complex64)
        out = np.empty((nz // 2 + 1, t), dtype=cdt, order="F")
        for i in range(t):
            out[:, i] = np.fft.rfft(window * xp[i * step: i * step + n_fft], n=nz)
    else:
        fft_in = np.zeros((nz, t), dtype=np.float32)
        for i in range(t):
            fft_in[:n_fft, i] = window * xp[i * step: i * step + n_fft]
        out = np.fft.rfft(fft_in, axis=0)
    return out / np.sqrt(n_fft)


def stft_f64(x, n_fft=1024, step=512, window_name="blackmanharris", zeropad=1):
    """fp64 "truth" of the same transform (SURVEY.md 8c restatement rules): float32-rounded
    window and float32 input, all arithmetic in float64, left-aligned frame with zeros
    appended (util/fourier.py:164-166), scale 1/sqrt(n_fft) (:157).  Returns C-order
    (T, F) complex128 -- i.e. the memory image of the reference's F-ordered (F, T)."""
    n_fft = int(n_fft)
    step = max(n_fft // 2, 1) if step is None else int(step)
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("x must be 1D")
    window = get_window_f32(window_name, n_fft).astype(np.float64)
    xp = reflect_pad(np.asarray(x, dtype=np.float32), n_fft).astype(np.float64)
    t = (len(xp) - n_fft) // step + 1
    nz = n_fft * zeropad
    idx = np.arange(n_fft)[None, :] + step * np.arange(t)[:, None]
    out = np.empty((t, nz // 2 + 1), dtype=np.complex128)
    blk = max(1, (1 << 24) // nz)
    for s in range(0, t, blk):
        e = min(t, s + blk)
        out[s:e] = np.fft.rfft(xp[idx[s:e]] * window[None, :], n=nz, axis=1)
    out /= np.sqrt(n_fft)
    return out


def to_mag(spectrum):
    """util/fourier.py:23-24."""
    return abs(spectrum) + .0000001


# --------------------------------------------------------------------------- iSTFT

def fix_length(data, size, axis=-1, **kwargs):
    """util/fourier.py:440-478."""
    kwargs.setdefault("mode", "constant")
    n = data.shape[axis]
    if n > size:
        sl = [slice(None)] * data.ndim
        sl[axis] = slice(0, size)
        return data[tuple(sl)]
    if n < size:
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (0, size - n)
        return np.pad(data, lengths, **kwargs)
    return data


def window_sumsquare(window_name, n_frames_, hop_length, n_fft, dtype=np.float32):
    """util/fourier.py:492-546 with win_length == n_fft, norm=None; fill loop :481-489."""
    n = n_fft + hop_length * (n_frames_ - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = dsp.get_window(window_name, n_fft) ** 2
    for i in range(n_frames_):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def istft_ref(stft_matrix, hop_length=None, window_name="blackmanharris", length=None,
              dtype=None):
    """util/fourier.py:314-437 with win_length=None, center=True.  Does NOT reproduce the
    in-place ``stft_matrix *= sqrt(n_fft)`` side effect (:359) -- works on a copy.
    Accumulates in the output dtype exactly like the reference (:383-404)."""
    stft_matrix = np.array(stft_matrix)
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    stft_matrix = stft_matrix * np.asarray(np.sqrt(n_fft), dtype=stft_matrix.real.dtype)
    if hop_length is None:
        hop_length = int(n_fft // 4)
    window = dsp.get_window(window_name, n_fft, fftbins=True)[:, None]
    if length:
        padded = length + int(n_fft)
        nfr = min(stft_matrix.shape[1], int(np.ceil(padded / hop_length)))
    else:
        nfr = stft_matrix.shape[1]
    if dtype is None:
        dtype = np.float32 if stft_matrix.dtype == np.complex64 else np.float64
    y = np.zeros(n_fft + hop_length * (nfr - 1), dtype=dtype)
    ncol = max((2 ** 18) // (stft_matrix.shape[0] * stft_matrix.itemsize), 1)
    frame = 0
    for s in range(0, nfr, ncol):
        e = min(s + ncol, nfr)
        ytmp = window * np.fft.irfft(stft_matrix[:, s:e], axis=0)
        for f in range(e - s):
            a = (frame + f) * hop_length
            y[a:a + n_fft] += ytmp[:, f]
        frame += e - s
    wss = window_sumsquare(window_name, nfr, hop_length, n_fft, dtype=dtype)
    nz = wss > np.finfo(wss.dtype).tiny
    y[nz] /= wss[nz]
    if length is None:
        return y[n_fft // 2:-(n_fft // 2)]
    return fix_length(y[n_fft // 2:], length)


# --------------------------------------------------------------------------- positions

def speed_segments(sampletimes, speeds):
    """Integer segment lengths of util/resampling.py:111-118: error-diffused
    ``n = int(round(period * mean(speeds[i:i+2]) + err))`` (Python round = half-even)."""
    sampletimes = np.asarray(sampletimes, dtype=np.float64)
    speeds = np.asarray(speeds, dtype=np.float64)
    periods = np.diff(sampletimes)
    err = 0.0
    ns = np.empty(len(speeds) - 1, dtype=np.int64)
    for i in range(len(speeds) - 1):
        inerr = periods[i] * ((speeds[i] + speeds[i + 1]) / 2.0) + err
        n = int(round(inerr))
        err = inerr - n
        ns[i] = n
    return ns


def speed_to_pos(sampletimes, speeds, num_input_samples):
    """util/resampling.py:93-137.  Returns only the filled prefix: where the reference's end
    test (:129) never fires it returns its np.empty buffer with an uninitialised tail;
    parity is defined on ``pos[0:sum(n)]``."""
    sampletimes = np.asarray(sampletimes, dtype=np.float64)
    speeds = np.asarray(speeds, dtype=np.float64)
    ns = speed_segments(sampletimes, speeds)
    offset = sampletimes[0]
    out = np.empty(int(ns.sum()), dtype=np.float64)
    o = 0
    for i, n in enumerate(ns):
        n = int(n)
        with np.errstate(divide="ignore", invalid="ignore"):
            block_speeds = np.arange(n) / (n - 1) * (speeds[i + 1] - speeds[i]) + speeds[i]
            sample_at = np.cumsum(1 / block_speeds) + offset
        offset = sample_at[-1]
        out[o:o + n] = sample_at
        if out[o] <= num_input_samples <= out[o + n - 1]:
            end = o + int(np.argmin(np.abs(sample_at - num_input_samples)))
            return out[:end]
        o += n
    return out[:o]


def lag_to_positions(lag_curve, sr, n_in):
    """Positions for the lag-curve mode of run -- util/resampling.py:189-206."""
    sampletimes = lag_curve[:, 0] * sr
    lags = lag_curve[:, 1] * sr
    num_out = n_in + abs(lags[-1])
    sample_at = np.interp(np.arange(num_out), sampletimes, sampletimes - lags)
    hit = np.nonzero(sample_at >= n_in)[0]
    if len(hit):
        sample_at = sample_at[:hit[0]]
    np.clip(sample_at, 0, None, out=sample_at)
    return sample_at


# --------------------------------------------------------------------------- resampler

def hanning_f32(nt):
    """util/resampling.py:24,36."""
    return np.hanning(2 * nt + 1).astype(np.float32)


def sinc_resample(sample_at, signal, nt, block=4096):
    """util/resampling.py:51-90 (sinc_core) as a single thread sees it, vectorised over
    output samples, all arithmetic in float64, store float32.

    Reproduced quirks: half-even rounding of the position (:69), taps
    ``lower = max(0, ind-NT) .. upper = min(ind+NT, len_in)`` -- the +NT tap dropped
    (:71-72); weights/window always indexed from 0 even when ``lower`` was clamped
    (start-edge misalignment, :89-90); ``period_to`` of the last element reuses the
    previous one (:76-77); ``fc = min(1/period_to, 1)`` (:79); empty slice -> 0.
    """
    sample_at = np.asarray(sample_at, dtype=np.float64)
    signal = np.asarray(signal)
    m = len(sample_at)
    len_in = len(signal)
    out = np.zeros(m, dtype=np.float32)
    if m == 0:
        return out
    win = hanning_f32(nt).astype(np.float64)[: 2 * nt]
    n_arr = np.arange(-nt, nt + 1, dtype=np.float32).astype(np.float64)[: 2 * nt]
    period = np.empty(m, dtype=np.float64)
    period[:-1] = np.maximum(0.000000000001, sample_at[1:] - sample_at[:-1])
    period[-1] = period[-2] if m > 1 else 0.0
    sig64 = signal.astype(np.float64)
    k = np.arange(2 * nt)
    for s in range(0, m, block):
        e = min(m, s + block)
        p = sample_at[s:e]
        ind = np.rint(p).astype(np.int64)
        lower = np.maximum(0, ind - nt)
        upper = np.minimum(ind + nt, len_in)
        cnt = np.maximum(upper - lower, 0)
        with np.errstate(divide="ignore"):
            fc = np.minimum(1.0 / period[s:e], 1.0)
        shift = p - ind
        si = np.sinc((n_arr[None, :] - shift[:, None]) * fc[:, None]) * fc[:, None]
        idx = lower[:, None] + k[None, :]
        valid = k[None, :] < cnt[:, None]
        sig = np.where(valid, sig64[np.clip(idx, 0, max(len_in - 1, 0))], 0.0) if len_in else \
            np.zeros(idx.shape)
        out[s:e] = np.sum(sig * si * win[None, :], axis=1).astype(np.float32)
    return out


def linear_resample(sample_at, signal):
    """The "Linear" mode of run -- util/resampling.py:228-229."""
    return np.interp(sample_at, np.arange(len(signal)), signal, left=0.0, right=0.0).astype(
        np.float32)


# --------------------------------------------------------------------------- dropout tools (config 4)

def time_2_frame(t, sr, hop):
    """dropout_healer_gui.py:99-101."""
    return int(t * sr / hop)


def freq_2_bin(f, fft_size, sr):
    """dropout_healer_gui.py:107-109."""
    return max(1, min(fft_size // 2, int(round(f * fft_size / sr))))


def heal_regions(markers, sr, fft_size, hop):
    """Integer regions dropout_healer_gui.py:136-141 derives from markers given as rows
    ``(t, width, f, height, surrounding)``: ``(frame_b, frame_a, frame_surrounding, bin_l, bin_u)``."""
    out = []
    for t, width, f, height, surrounding in np.asarray(markers, dtype=np.float64).tolist():
        out.append((time_2_frame(t - (width / 2), sr, hop), time_2_frame(t + (width / 2), sr, hop),
                    max(1, time_2_frame(width * surrounding, sr, hop)),
                    freq_2_bin(f - (height / 2), fft_size, sr), freq_2_bin(f + (height / 2), fft_size, sr)))
    return np.array(out, dtype=np.int64).reshape(-1, 5)


def heal_ref(x, sr, markers, fft_size, hop):
    """One channel of dropout_healer_gui.py:124-164 on the CPU: fix_length :129, stft (numpy back-end)
    :132, dB :133, per-marker means :144-145, bilinear target :148-154, clipped gain :156-160,
    ``S *= 10**(gain/20)`` :162, istft :164.  Pinned by tests/golden/dropouts.npz, which holds the
    output of the unmodified reference method."""
    from scipy.interpolate import RegularGridInterpolator
    x = np.asarray(x, dtype=np.float32)
    n = len(x)
    y_pad = fix_length(x, n + fft_size // 2)
    spec = np.array(stft_ref(y_pad, fft_size, hop))
    spec_db = 20 * np.log10(to_mag(spec))
    gain = np.zeros(spec.shape, dtype=float)
    for frame_b, frame_a, around, bin_l, bin_u in heal_regions(markers, sr, fft_size, hop).tolist():
        mag_before = np.mean(spec_db[bin_l:bin_u, frame_b - around:frame_b], axis=1)
        mag_after = np.mean(spec_db[bin_l:bin_u, frame_a:frame_a + around], axis=1)
        fp_frames = np.linspace(frame_b, frame_a, num=frame_a - frame_b)
        fp_bins = np.linspace(bin_l, bin_u, num=bin_u - bin_l)
        interp = RegularGridInterpolator(((frame_b, frame_a), fp_bins), (mag_before, mag_after))
        mp_bins, mp_frames = np.meshgrid(fp_bins, fp_frames)
        fp_db = np.swapaxes(interp((mp_frames, mp_bins)), 0, 1)
        g = fp_db - spec_db[bin_l:bin_u, frame_b:frame_a]
        np.clip(g, gain[bin_l:bin_u, frame_b:frame_a], 255, out=g)
        gain[bin_l:bin_u, frame_b:frame_a] = g
    spec = spec * np.power(10, gain / 20)
    return istft_ref(spec, hop_length=hop, length=n).astype(np.float32)


def locate_peaks_ref(magnitude, sr, fft_size, hop, t_0, t_1, f_lower, f_upper, sensitivity):
    """Integer frame indices of dropout_healer_gui.py:188-204: dB, band/time window, mean over the
    band, ``find_peaks(-vol, prominence=10-sensitivity)``."""
    import scipy.signal
    imdata = 20 * np.log10(np.array(magnitude))
    frame_b, frame_a = time_2_frame(t_0, sr, hop), time_2_frame(t_1, sr, hop)
    bin_l, bin_u = freq_2_bin(f_lower, fft_size, sr), freq_2_bin(f_upper, fft_size, sr)
    vol = np.mean(imdata[bin_l:bin_u, frame_b:frame_a], axis=0)
    peaks, _ = scipy.signal.find_peaks(-vol, height=None, threshold=None, distance=None,
                                       prominence=10.0 - sensitivity, wlen=None, rel_height=0.5, plateau_size=None)
    return peaks


def max_mono_ref(signal, fft_size, hop):
    """dropouts_gui.py:137-163 on the CPU: ``{"max": y, "min": y}``."""
    n = len(signal)
    y_pad = fix_length(np.asarray(signal), n + fft_size // 2, axis=0)
    d_l = np.array(stft_ref(y_pad[:, 0], fft_size, hop))
    d_r = np.array(stft_ref(y_pad[:, 1], fft_size, hop))
    out = {}
    for name, mask in (("max", np.abs(d_l) > np.abs(d_r)), ("min", np.abs(d_l) < np.abs(d_r))):
        out[name] = istft_ref(np.where(mask, d_l, d_r), hop_length=hop, length=n)
    return out


def noise_gate_ref(signal, profile_db, gain_db, fft_size, hop):
    """renoiser_gui.py:273-278 (get_mask_fac) + :303-319 (run_resample) for one channel: pad, STFT, scale every cell whose
    ``to_dB(to_mag(S))`` does not exceed the profile by ``to_fac(gain)`` (as float32), iSTFT."""
    x = np.asarray(signal)
    n = len(x)
    pad = fix_length(x, n + fft_size // 2, axis=0)
    spec = np.array(stft_ref(pad, fft_size, hop))
    gain_mask = np.where(20 * np.log10(np.array(to_mag(spec))) > np.expand_dims(np.asarray(profile_db), axis=1), 0.0, gain_db)
    fac = np.power(10, gain_mask / 20).astype(np.float32)
    return istft_ref(spec * fac, hop_length=hop, length=n)


def heuristic_peaks_ref(magnitude, sr, fft_size, f_lower, f_upper, num_bands):
    """Per-band integer peak lists of dropouts_gui.py:251, :264-288, top band first.  Band edges are converted
    to ``int`` before the multiplication (numpy 1.x promotion; under numpy 2 the reference's ``uint16 * int``
    wraps and every band is empty -- see tests/golden/make_golden_dropouts_full.py)."""
    import scipy.signal
    imdata = 20 * np.log10(np.array(magnitude))
    bands = np.logspace(np.log2(f_lower), np.log2(f_upper), num=num_bands, endpoint=True, base=2, dtype=np.uint16)
    pairs = list(zip(bands[:-1], bands[1:]))
    out = []
    for f_lower_band, f_upper_band in reversed(pairs):
        bin_lower = int(int(f_lower_band) * fft_size / sr)
        bin_upper = int(int(f_upper_band) * fft_size / sr)
        vol = np.mean(imdata[bin_lower:bin_upper], axis=0)
        peaks, _ = scipy.signal.find_peaks(-vol, prominence=5, rel_height=0.5)
        out.append(peaks)
    return out


# --------------------------------------------------------------------------- frequency trackers (8f rank 1)

def _track_setup(spectrum, trail, fft_size, hop, sr):
    """Track.__init__ / sample_trail / ensure_frames -- util/wow_detection.py:32-89."""
    num_bins, frame_1 = spectrum.shape
    trail = sorted(trail, key=lambda tup: tup[0])
    times_raw = [d[0] for d in trail]
    freqs_raw = [d[1] for d in trail]
    frame_0 = 0
    if times_raw[0]:
        frame_0 = max(frame_0, int(times_raw[0] * sr / hop))
    if times_raw[-1]:
        frame_1 = min(frame_1, int(times_raw[-1] * sr / hop))
    times = np.linspace(frame_0 * hop / sr, frame_1 * hop / sr, frame_1 - frame_0)
    freqs = np.interp(times, times_raw, freqs_raw)
    return num_bins, frame_0, times, freqs


def _bin_limits(freq, tolerance, num_bins, fft_size, sr, min_bins=4):
    """freq_plus_tolerance :106-116 + set_bin_limits :91-104 + freq_2_bin :79-80."""
    logfreq = np.log2(freq)
    fl = max(1.0, np.power(2, (logfreq - tolerance)))
    fu = min(sr / 2, np.power(2, (logfreq + tolerance)))

    def f2b(f):
        return max(1, min(num_bins - 1, int(round(f * fft_size / sr))))
    nl, nu = f2b(fl), f2b(fu)
    while (nu - nl) < min_bins:
        nl -= 1
        nu += 1
    return nl, nu


def _get_peak(spectrum, frame, nl, nu, fft_size, sr):
    """get_peak :119-134 (rectangular window) + is_peak :136-139 + parabolic (util/correlation.py:42-46)."""
    fft_frame = spectrum[:, frame]
    peak = nl + np.argmax(fft_frame[nl:nu] * np.ones(nu - nl))
    if fft_frame[peak - 1] < fft_frame[peak] > fft_frame[peak + 1]:
        f = fft_frame
        peak = 1 / 2. * (f[peak - 1] - f[peak + 1]) / (f[peak - 1] - 2 * f[peak] + f[peak + 1]) + peak
    return peak / fft_size * sr


def _interp_nans(y):
    nans = np.isnan(y)
    if nans.any() and (~nans).any():
        y[nans] = np.interp(nans.nonzero()[0], (~nans).nonzero()[0], y[~nans])
    return y


def track_ref(mode, spectrum, trail, fft_size, hop, sr, tolerance_st=1):
    """``(times, freqs)`` of PeakTracker.trace (:294-302, mode 'peak'), PeakTrackTracker.trace (:305-327,
    'peak_track') or CenterOfGravity.trace (:256-291, 'cog') on a (bins, frames) magnitude array."""
    spectrum = np.asarray(spectrum)
    num_bins, frame_0, times, freqs = _track_setup(spectrum, trail, fft_size, hop, sr)
    tol = tolerance_st / 12
    if mode == "peak":
        for i, raw in enumerate(freqs):
            nl, nu = _bin_limits(raw, tol, num_bins, fft_size, sr)
            freqs[i] = _get_peak(spectrum, frame_0 + i, nl, nu, fft_size, sr)
    elif mode == "peak_track":
        freq = freqs[0]
        for i in range(len(freqs)):
            nl, nu = _bin_limits(freq, tol / 2 if i > 2 else tol, num_bins, fft_size, sr)
            freqs[i] = _get_peak(spectrum, frame_0 + i, nl, nu, fft_size, sr)
    elif mode == "cog":
        fft_freqs = np.arange(0, (fft_size // 2 + 1)) / float(fft_size) * float(sr)
        nl, nu = _bin_limits(freqs[0], tol, num_bins, fft_size, sr)
        for i in range(len(freqs)):
            weighted = np.hanning(nu - nl) * spectrum[nl:nu, frame_0 + i]
            with np.errstate(divide="ignore", invalid="ignore"):
                freqs[i] = 2 ** (np.sum(weighted * np.log2(fft_freqs[nl:nu])) / np.sum(weighted))
            nl, nu = _bin_limits(freqs[i], tol, num_bins, fft_size, sr)
    else:
        raise ValueError(mode)
    return times, _interp_nans(freqs)
