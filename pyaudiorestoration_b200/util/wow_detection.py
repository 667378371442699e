"""Device-side mirror of the reference's frequency trackers (``util/wow_detection.py``; SURVEY.md 8f
rank 1): the consumers of ``get_mag`` that turn a spectrogram and a hand-drawn trail into the speed
curve ``util.resampling.run`` applies.

Same constructor as the reference's ``Track`` (:32-60) and the same registry ``wow_detectors``
(:453-456) for the three spectral trackers, Peak, Peak Track and Center of Gravity; the per-frame
work (band limits, arg-max, parabolic refinement, centre of gravity) runs in ``csrc/track.cu``.
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


wow_detectors = {cls.name: cls for cls in (CenterOfGravity, PeakTracker, PeakTrackTracker)}


def trace_signal(signal, trail, fft_size, hop, sr, mode="Peak", tolerance_st=1, window_name="blackmanharris", zeropad=1):
    """Fused ``get_mag`` + tracker: ``(times, freqs)`` of the trace of a 1-D signal along ``trail``
    without materialising the spectrogram on the host (or anywhere beyond the traced frames)."""
    L = _lib.lib()
    _lib.require_device()
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
