"""Device-side mirror of the reference's frequency trackers (``util/wow_detection.py``; SURVEY.md 8f
rank 1): the consumers of ``get_mag`` that turn a spectrogram and a hand-drawn trail into the speed
curve ``util.resampling.run`` applies.

Same constructor as the reference's ``Track`` (:32-60) and the same registry ``wow_detectors``
(:453-456).  For the three spectral trackers, Peak, Peak Track and Center of Gravity, the per-frame
work (band limits, arg-max, parabolic refinement, centre of gravity) runs in ``csrc/track.cu``; the
other five modes (Zero-Crossing, Correlation, Partials, Freehand Draw, Sine Regression) are small host
scipy routines in the reference and stay on the host here.
``trace_signal`` goes one step further: the magnitudes of the traced frames are computed on the
device, traced there and discarded -- the spectrogram (11 GB per channel at BASELINE config 3)
never crosses PCIe, only one float64 per frame comes back.
"""
import logging

import numpy as np
from scipy import signal as dsp

from .. import _lib
from . import fourier


def interp_nans(y):
    """util/wow_detection.py:14-22."""
    nans = np.isnan(y)
    if nans.any() and (~nans).any():
        y[nans] = np.interp(nans.nonzero()[0], (~nans).nonzero()[0], y[~nans])


def sample_trail(trail, num_frames, hop, sr):
    """Track.sample_trail / ensure_frames (:62-89): ``(frame_0, times, freqs)`` of a drawn trail
    ``[(time, freq), ...]`` sampled at every spectrogram frame it spans."""
    trail = sorted(trail, key=lambda tup: tup[0])
    times_raw = [d[0] for d in trail]
    freqs_raw = [d[1] for d in trail]
    frame_0, frame_1 = 0, num_frames
    if times_raw[0]:
        frame_0 = max(frame_0, int(times_raw[0] * sr / hop))
    if times_raw[-1]:
        frame_1 = min(frame_1, int(times_raw[-1] * sr / hop))
    if frame_0 == frame_1:
        logging.warning("No point in tracing just one FFT")
    times = np.linspace(frame_0 * hop / sr, frame_1 * hop / sr, frame_1 - frame_0)
    return frame_0, times, np.interp(times, times_raw, freqs_raw)


class Track:
    """Base of the device trackers: ``Track(spectrum, signal, trail, fft_size, hop, sr, tolerance_st,
    adaptation_mode)`` like the reference; results in ``.times`` / ``.freqs``."""
    name = ""
    tooltip = ""
    mode = None

    def __init__(self, spectrum, signal, trail, fft_size, hop, sr, tolerance_st=1, adaptation_mode="Linear",
                 dB_cutoff=75):
        self.fft_size, self.hop, self.sr = int(fft_size), int(hop), sr
        self.spectrum, self.signal = spectrum, signal
        self.fft_freqs = fourier.fft_freqs(fft_size, sr)
        self.num_bins, num_frames = spectrum.shape
        self.frame_0, self.times, self.freqs = sample_trail(list(trail), num_frames, hop, sr)
        self.frame_1 = self.frame_0 + len(self.freqs)
        self.tolerance = tolerance_st / 12
        self.tolerance_st = tolerance_st
        self.min_bins = 4
        self.trace()
        interp_nans(self.freqs)

    def trace(self):
        L = _lib.lib()
        _lib.require_device()
        spec = np.asarray(self.spectrum)
        if spec.dtype != np.float32:
            spec = spec.astype(np.float32)
        # the library wants frames contiguous: the (bins, frames) arrays get_mag returns are views of exactly that
        frames = spec.T if spec.T.flags.c_contiguous else np.ascontiguousarray(spec.T)
        freqs = np.ascontiguousarray(self.freqs, dtype=np.float64)
        rc = L.par_trace_f32(frames.ctypes.data, self.num_bins, frames.shape[0], frames.strides[0] // 4, self.frame_0,
                             len(freqs), self.fft_size, float(self.sr), float(self.tolerance_st), self.mode,
                             freqs.ctypes.data, 0, _lib.device(), None)
        _lib.check(rc, "par_trace_f32")
        self.freqs = freqs


class CenterOfGravity(Track):
    name = 'Center of Gravity'
    mode = _lib.PAR_TRACE_COG


class PeakTracker(Track):
    name = 'Peak'
    tooltip = "Tracks the mouse input to the loudest peak frequency"
    mode = _lib.PAR_TRACE_PEAK


class PeakTrackTracker(Track):
    name = 'Peak Track'
    tooltip = "Follows the first peak frequency established"
    mode = _lib.PAR_TRACE_PEAK_TRACK


# ---- the registry's host-side members -------------------------------------------------------------
# The reference's remaining trackers are scipy code on small arrays (one waveform band-pass, or a few hundred bins
# per frame); they are outside the device path and stay on the host, restated here so that ``wow_detectors`` offers
# the same eight modes to the GUIs (util/widgets.py:19, pyrespeeder_gui.py:11).

def nan_helper(y):
    """util/wow_detection.py:14-16: NaN mask and an index function for ``np.interp``."""
    return np.isnan(y), lambda z: z.nonzero()[0]


def zero_crossings(a):
    """Indices i with a sign change between a[i] and a[i+1] (util/wow_detection.py:448-450)."""
    pos = np.asarray(a) > 0
    return np.flatnonzero(pos[1:] != pos[:-1])


def parabolic(f, x):
    """Vertex of the parabola through f[x-1], f[x], f[x+1] (util/correlation.py:42-46): (abscissa, ordinate)."""
    xv = 0.5 * (f[x - 1] - f[x + 1]) / (f[x - 1] - 2 * f[x] + f[x + 1]) + x
    return xv, f[x] - 0.25 * (f[x - 1] - f[x + 1]) * (xv - x)


def xcorr(a, b, mode="full"):
    """Cross-correlation of the two inputs scaled to unit norm (util/correlation.py:6-13)."""
    return dsp.correlate(a / np.linalg.norm(a), b / np.linalg.norm(b), mode=mode, method="auto")


def _bandpass(data, lowcut, highcut, fs, order):
    """util/filters.py:7-25: zero-phase Butterworth band / high / low pass, whichever corner lies inside (0, fs/2)."""
    low, high = lowcut / (0.5 * fs), highcut / (0.5 * fs)
    lo_ok, hi_ok = 0 < low < 1, 0 < high < 1
    if lo_ok and hi_ok:
        sos = dsp.butter(order, [low, high], btype="band", output="sos")
    elif lo_ok:
        sos = dsp.butter(order, low, btype="high", output="sos")
    elif hi_ok:
        sos = dsp.butter(order, high, btype="low", output="sos")
    else:
        return data
    return dsp.sosfiltfilt(sos, data)


class HostTrack(Track):
    """A tracker whose ``trace`` needs no device work."""

    def band(self, freq):
        """``freq`` -/+ the tolerance on a log2 scale (Track.freq_plus_tolerance, :109-117)."""
        lf = np.log2(freq)
        return np.power(2, lf - self.tolerance), np.power(2, lf + self.tolerance)

    def bin_limits(self, f_lo, f_hi):
        """Track.set_bin_limits (:97-107): clamped, rounded bins at least ``min_bins`` apart."""
        to_bin = lambda f: max(1, min(self.num_bins - 1, int(round(f * self.fft_size / self.sr))))   # noqa: E731
        lo, hi = to_bin(max(1.0, f_lo)), to_bin(min(self.sr / 2, f_hi))
        while hi - lo < self.min_bins:
            lo, hi = lo - 1, hi + 1
        return lo, hi

    def trace(self):
        pass


class ZeroCrossingTracker(HostTrack):
    name = 'Zero-Crossing'
    tooltip = "Track the distance between zero-crossings of the waveform. Good for flutter detection of clean signals"

    def trace(self):
        """util/wow_detection.py:334-358: band-pass channel 0 around the trail, turn the spacing of its zero crossings
        into a frequency, smooth with a Hann kernel of 10 ms and resample onto the frame times."""
        f_lo, _ = self.band(np.min(self.freqs))
        _, f_hi = self.band(np.max(self.freqs))
        s0, s1 = int(self.times[0] * self.sr), int(self.times[-1] * self.sr)
        crossings = zero_crossings(_bandpass(np.asarray(self.signal)[s0:s1, 0], f_lo, f_hi, self.sr, order=3))
        gaps = np.diff(crossings).astype(np.float32)
        size = int(self.sr / 100 / np.mean(gaps))
        kernel = dsp.get_window("hann", size) / size * 2
        smooth = np.convolve(np.pad(gaps, size, mode="reflect"), kernel, mode="same")[size:-size]
        self.freqs[:] = np.interp(self.times, crossings[:len(smooth)] / self.sr + self.times[0], self.sr / 2 / smooth)


class PartialsTracker(HostTrack):
    name = 'Partials'

    def trace(self):
        """util/wow_detection.py:378-387 hands the waveform to ``librosa.piptrack`` and plots the result; librosa is
        not a dependency of either code base, so the mode reports that instead of silently doing nothing."""
        import librosa  # noqa: F401
        raise NotImplementedError("the reference's Partials mode only plots librosa.piptrack; nothing to trace")


class FreehandTracker(HostTrack):
    name = 'Freehand Draw'


class CorrelationTracker(HostTrack):
    name = 'Correlation'
    tooltip = "Compare the spectra for each segment and track the offsets between"

    def trace(self):
        """util/wow_detection.py:400-436: every frame's band is resampled onto a 4x finer log2-frequency grid
        (quadratic spline), neighbouring frames are cross-correlated under a Hann window, the sub-sample lag of the
        correlation peak is the frame-to-frame pitch change; the changes are summed and mapped back to Hz around the
        band's centre.  Like the reference it reads spectrogram columns 0 .. len(freqs)-1 (not offset by the trail's
        first frame) and compares the last frame with a column of ones."""
        from scipy.interpolate import interp1d
        f_lo, f_hi = min(self.freqs), max(self.freqs)
        lo, hi = self.bin_limits(f_lo, f_hi)
        fine = (hi - lo) * 4
        log_f = np.log2(self.fft_freqs[lo:hi])
        grid = np.linspace(log_f[0], log_f[-1], fine)
        count = len(self.freqs)
        resampled = np.ones((fine, count + 1))
        for i in range(count):
            resampled[:, i] = interp1d(log_f, self.spectrum[lo:hi, i], kind="quadratic")(grid)
        wind = np.hanning(fine)
        changes = np.ones(count)
        for i in range(count):
            res = xcorr(resampled[:, i] * wind, resampled[:, i + 1] * wind, mode="same")
            changes[i] = fine // 2 - parabolic(res, np.argmax(res))[0]
        drift = np.cumsum(changes) / fine * (log_f[-1] - log_f[0])
        np.power(2, np.log2((f_lo + f_hi) / 2) + drift, self.freqs)


class SineRegression(HostTrack):
    name = 'Sine Regression'
    tooltip = "Perform a regression on an area of the master speed curve to yield a sine fit"


def fit_sin(tt, yy, assumed_freq=None):
    """Least-squares sine through (tt, yy) (util/wow_detection.py:190-228): the start values come from the largest
    bin of the real FFT (optionally weighted towards ``assumed_freq``), the fit from ``scipy.optimize.curve_fit``.
    Returns the reference's dictionary."""
    from scipy.optimize import curve_fit
    tt, yy = np.array(tt), np.array(yy)
    step = tt[1] - tt[0]
    ff = np.fft.rfftfreq(len(tt), step)
    spec = np.fft.rfft(yy)[1:]
    if assumed_freq:
        expected = int(round(assumed_freq * (len(yy) + 1) * step))
        spec *= np.interp(np.arange(len(spec)), (0, expected, len(spec)), (0, 1, 0))
    peak = np.argmax(np.abs(spec)) + 1
    guess = np.array([np.std(yy) * 2. ** 0.5, 2. * np.pi * ff[peak], np.angle(spec[peak]), np.mean(yy)])

    def model(t, A, w, p, c):
        return A * np.sin(w * t + p) + c
    popt, pcov = curve_fit(model, tt, yy, p0=guess)
    A, w, p, c = popt
    f = w / (2. * np.pi)
    return {"amp": A, "omega": w, "phase": p, "offset": c, "freq": f, "period": 1. / f,
            "fitfunc": lambda t: A * np.sin(w * t + p) + c, "maxcov": np.max(pcov), "rawres": (guess, popt, pcov)}


def trace_sine_reg(speed_curve, t0, t1, rpm=None):
    """util/wow_detection.py:231-253: sine fit of the master speed curve between t0 and t1 -> (amp, omega, phase, 0)."""
    times, speeds = speed_curve[:, 0], speed_curve[:, 1]
    step = times[1] - times[0]
    a, b = int(t0 / step), int(t1 / step)
    try:
        assumed = float(rpm) / 60
    except (TypeError, ValueError):
        assumed = None
    res = fit_sin(times[a:b], speeds[a:b], assumed_freq=assumed)
    return res["amp"], res["omega"], res["phase"], 0


wow_detectors = {cls.name: cls for cls in (CenterOfGravity, PeakTracker, PeakTrackTracker, ZeroCrossingTracker,
                                           PartialsTracker, FreehandTracker, CorrelationTracker, SineRegression)}


def trace_signal(signal, trail, fft_size, hop, sr, mode="Peak", tolerance_st=1, window_name="blackmanharris", zeropad=1):
    """Fused ``get_mag`` + tracker: ``(times, freqs)`` of the trace of a 1-D signal along ``trail``
    without materialising the spectrogram on the host (or anywhere beyond the traced frames)."""
    L = _lib.lib()
    _lib.require_device()
    if wow_detectors[mode].mode is None:
        raise ValueError(f"trace_signal runs the spectral trackers (Peak, Peak Track, Center of Gravity), not {mode!r}")
    keep, ptr, stride = _lib.f32_layout(signal)
    n = len(keep)
    fft_size, hop, zeropad = int(fft_size), int(hop), int(zeropad)
    num_frames = int(L.par_stft_num_frames(n, fft_size, hop))
    frame_0, times, freqs = sample_trail(list(trail), num_frames, hop, sr)
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    window = np.ascontiguousarray(dsp.get_window(window_name, fft_size), dtype=np.float32)
    rc = L.par_stft_trace_f32(ptr, n, stride, fft_size, hop, zeropad, window.ctypes.data, frame_0, len(freqs), float(sr),
                              float(tolerance_st), wow_detectors[mode].mode, freqs.ctypes.data, 0, _lib.device(), None)
    _lib.check(rc, "par_stft_trace_f32")
    interp_nans(freqs)
    del keep
    return times, freqs
