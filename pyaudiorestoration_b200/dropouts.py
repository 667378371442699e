"""Headless numeric bodies of the reference's dropout tools (BASELINE config 4).

In the reference these live inside Qt/vispy classes; here they are plain functions over arrays.  The
transforms (``stft`` / ``get_mag`` / ``istft``) run on the B200 through ``util.fourier``; the work
between them is small per-marker numpy/scipy arithmetic on the host, written to produce the same
integer regions and peak indices as the reference:

* ``heal``          <- dropout_healer_gui.py:111-166  (Canvas.resample_files)
* ``marker_region`` <- dropout_healer_gui.py:99-109, :136-141  (time_2_frame / freq_2_bin rules)
* ``locate``        <- dropout_healer_gui.py:185-242  (Alt-drag batch detection)
* ``max_mono``      <- dropouts_gui.py:137-163        (process_max_mono)
* ``noise_gate``    <- renoiser_gui.py:273-278, :303-319  (get_mask_fac / run_resample)
* ``heuristic``     <- dropouts_gui.py:241-323        (process_heuristic)
"""
import logging
from collections import namedtuple

import numpy as np
import scipy.signal
from scipy.interpolate import RegularGridInterpolator

from .util import fourier

Dropout = namedtuple("Dropout", "t width f height surrounding")
Dropout.__doc__ = """A dropout marker as util/markers.py:366-387 stores it: centre time [s], width [s],
centre frequency [Hz], height [Hz], fraction of the width to average on either side."""


def to_dB(a):
    """util/units.py:27-28.  Evaluated in float64: the reference's CPU back-end hands its tools float64 magnitudes
    (util/fourier.py:157 promotes), and a float32 decibel value near -100 dB would carry an error of several 1e-6 dB
    that the gain curves of the dropout tools turn into the same RELATIVE error of the audio."""
    return 20 * np.log10(np.asarray(a, dtype=np.float64))


def to_fac(a):
    """util/units.py:31-32."""
    return np.power(10, a / 20)


def dropout_from_corners(a, b, surrounding=0.5):
    """Marker from two (time, frequency) corners, like ``DropoutSample(canvas, a, b, surrounding)``."""
    return Dropout(t=(a[0] + b[0]) / 2, width=abs(a[0] - b[0]), f=(a[1] + b[1]) / 2, height=abs(a[1] - b[1]),
                   surrounding=surrounding)


def time_2_frame(t, sr, hop):
    """dropout_healer_gui.py:99-101: truncation towards zero, not rounding."""
    return int(t * sr / hop)


def freq_2_bin(f, fft_size, sr):
    """dropout_healer_gui.py:107-109: rounded, clamped to [1, fft_size/2]."""
    return max(1, min(fft_size // 2, int(round(f * fft_size / sr))))


def marker_region(drop, sr, fft_size, hop):
    """Integer region of a marker: ``(frame_b, frame_a, frame_surrounding, bin_l, bin_u)``."""
    frame_b = time_2_frame(drop.t - (drop.width / 2), sr, hop)
    frame_a = time_2_frame(drop.t + (drop.width / 2), sr, hop)
    frame_surrounding = max(1, time_2_frame(drop.width * drop.surrounding, sr, hop))
    bin_l = freq_2_bin(drop.f - (drop.height / 2), fft_size, sr)
    bin_u = freq_2_bin(drop.f + (drop.height / 2), fft_size, sr)
    return frame_b, frame_a, frame_surrounding, bin_l, bin_u


def heal_gain_db(spectrum_db, drops, sr, fft_size, hop):
    """The dB boost mask of dropout_healer_gui.py:134-160 for one channel's spectrogram in dB."""
    gain = np.zeros(spectrum_db.shape, dtype=float)
    for drop in drops:
        frame_b, frame_a, around, bin_l, bin_u = marker_region(drop, sr, fft_size, hop)
        before = np.mean(spectrum_db[bin_l:bin_u, frame_b - around:frame_b], axis=1)
        after = np.mean(spectrum_db[bin_l:bin_u, frame_a:frame_a + around], axis=1)
        frames = np.linspace(frame_b, frame_a, num=frame_a - frame_b)
        bins = np.linspace(bin_l, bin_u, num=bin_u - bin_l)
        # bilinear blend between the spectrum just before and just after the gap, per bin
        blend = RegularGridInterpolator(((frame_b, frame_a), bins), (before, after))
        grid_bins, grid_frames = np.meshgrid(bins, frames)
        target = np.swapaxes(blend((grid_frames, grid_bins)), 0, 1)
        boost = target - spectrum_db[bin_l:bin_u, frame_b:frame_a]
        # never less than what an earlier marker already asked for, never more than 255 dB
        np.clip(boost, gain[bin_l:bin_u, frame_b:frame_a], 255, out=boost)
        gain[bin_l:bin_u, frame_b:frame_a] = boost
    return gain


def _device_regions(drops, sr, fft_size, hop, n):
    """The markers' integer regions as the ``(n, 5)`` int64 table ``par_spectral_process_f32`` takes, or None if a
    region does not fit the device operator's contract (outside the spectrogram, thinner than two bins)."""
    n_frames = (n + fft_size // 2) // hop + 1
    regs = np.array([marker_region(d, sr, fft_size, hop) for d in drops], dtype=np.int64).reshape(-1, 5)
    for frame_b, frame_a, around, bin_l, bin_u in regs:
        if frame_b < 0 or frame_a <= frame_b or frame_a > n_frames or around < 1 or bin_u - bin_l < 2 or bin_u > fft_size // 2 + 1:
            return None
    return regs


def heal(signal, sr, drops, fft_size=512, hop=32, channels=None, on_device=True):
    """Repair marked dropouts (dropout_healer_gui.py:111-166): per channel STFT, interpolate the
    magnitude across every marker from its surroundings, inverse STFT.  ``signal`` is
    ``(frames, channels)``; returns the same shape and dtype (untouched channels are left as the
    reference leaves them: uninitialised there, copied through here).

    The whole chain -- pad, transform, per-marker gain, inverse transform -- runs on the device in one call
    (``fourier.stft_mask_istft``: one upload, one download, the spectrogram never visits the host).
    ``on_device=False``, or a marker that does not fit the device operator, takes the transforms to the device but
    builds the gain mask with numpy/scipy on the host exactly like the reference (``heal_gain_db``)."""
    signal = np.asarray(signal)
    if signal.ndim == 1:
        signal = signal[:, None]
    n = len(signal)
    if channels is None:
        channels = range(signal.shape[1])
    channels = list(channels)
    output = np.array(signal, copy=True)
    regs = _device_regions(drops, sr, fft_size, hop, n) if on_device else None
    if regs is not None and channels:
        healed = fourier.stft_mask_istft(signal[:, channels], "heal", fft_size, hop, params=regs)
        output[:, channels] = healed
        return output
    y_pad = fourier.fix_length(signal, n + fft_size // 2, axis=0)
    for channel in channels:
        spectrum = np.array(fourier.stft(y_pad[:, channel], n_fft=fft_size, step=hop))
        gain = heal_gain_db(to_dB(fourier.to_mag(spectrum)), drops, sr, fft_size, hop)
        spectrum = (spectrum * to_fac(gain)).astype(np.complex64)
        output[:, channel] = fourier.istft(spectrum, length=n, hop_length=hop)
    return output


def band_volume(magnitude, frame_b, frame_a, bin_l, bin_u):
    """Mean dB level over a band per frame (dropout_healer_gui.py:188-195)."""
    return np.mean(to_dB(np.asarray(magnitude))[bin_l:bin_u, frame_b:frame_a], axis=0)


def locate(magnitude, sr, fft_size, hop, t_0, t_1, f_lower, f_upper, sensitivity=4.0, width_ms=20.0):
    """Batch dropout detection (dropout_healer_gui.py:185-242) on a magnitude spectrogram (``get_mag``
    output, ``(bins, frames)``): valleys of the band volume with prominence ``10 - sensitivity``.

    Returns ``(peaks, markers)``: the integer frame indices (relative to the frame of ``t_0``) found by
    ``scipy.signal.find_peaks`` -- bit-exact against the CPU path is BASELINE's bar for this config --
    and one ``Dropout`` per peak with the reference's parabola-refined width."""
    frame_b = time_2_frame(t_0, sr, hop)
    frame_a = time_2_frame(t_1, sr, hop)
    bin_l = freq_2_bin(f_lower, fft_size, sr)
    bin_u = freq_2_bin(f_upper, fft_size, sr)
    vol = band_volume(magnitude, frame_b, frame_a, bin_l, bin_u)
    half_width = width_ms / 1000 / 2
    frames_half_width = time_2_frame(half_width, sr, hop)
    vol_lt = scipy.signal.savgol_filter(vol, frames_half_width * 12, 5)
    vol_st = scipy.signal.savgol_filter(vol, frames_half_width, 5)
    peaks, _ = scipy.signal.find_peaks(-vol, prominence=10.0 - sensitivity, rel_height=0.5)
    found = []
    for f_peak in peaks:
        t_center = (frame_b + f_peak) / sr * hop
        try:
            # width of a parabola through the short-term curve where it meets the long-term curve
            f_qw = time_2_frame(half_width / 4, sr, hop)
            xp = np.arange(f_peak - f_qw, f_peak + f_qw)
            parabola = np.poly1d(np.polyfit(xp, vol_st[f_peak - f_qw:f_peak + f_qw], 2))
            f_hw = time_2_frame(half_width, sr, hop)
            lo, hi = f_peak - f_hw, f_peak + f_hw
            crossings = scipy.signal.argrelmin(np.abs(parabola(np.arange(lo, hi)) - vol_lt[lo:hi]))[0]
            assert len(crossings) == 2
            half_width = (crossings[1] - crossings[0]) / sr * hop      # carried over to later peaks, as in the reference
        except Exception:                                                # noqa: BLE001 -- the reference logs and goes on
            logging.debug(f"Could not refine width at peak {f_peak}")
        found.append(dropout_from_corners((t_center - half_width, f_lower), (t_center + half_width, f_upper)))
    return peaks, found


def max_mono(signal, fft_size=512, hop=32, on_device=True):
    """dropouts_gui.py:137-163: per time-frequency cell keep the louder (``max``) or the quieter
    (``min``) of the two channels of a stereo take.  Returns ``{"max": y, "min": y}``.  One device call: both
    channels are transformed once, both selections are made and inverted there (``on_device=False``: select on the
    host between a device STFT and two device iSTFTs, like the reference)."""
    signal = np.asarray(signal)
    if signal.ndim != 2 or signal.shape[1] != 2:
        raise ValueError("expects stereo input")
    n = len(signal)
    if on_device:
        both = fourier.stft_mask_istft(signal, "max_min", fft_size, hop)
        return {"max": np.ascontiguousarray(both[:, 0]), "min": np.ascontiguousarray(both[:, 1])}
    y_pad = fourier.fix_length(signal, n + fft_size // 2, axis=0)
    left, right = (np.array(s) for s in fourier.stft_multi(y_pad, n_fft=fft_size, step=hop))
    out = {}
    for name, mask in (("max", np.abs(left) > np.abs(right)), ("min", np.abs(left) < np.abs(right))):
        out[name] = fourier.istft(np.where(mask, left, right), length=n, hop_length=hop)
    return out


def noise_gate(signal, profile_db, gain_db, fft_size=512, hop=32, channels=None):
    """The spectral gate of renoiser_gui.py:273-278 + :303-319 (get_mask_fac / run_resample): every cell whose level
    ``to_dB(|S| + 1e-7)`` does not exceed ``profile_db[bin]`` (the tool's ``final_profile``: noise profile + gain +
    control curve + overhead) is scaled by ``to_fac(gain_db)``; all on the device in one call.  ``signal`` is
    ``(frames, channels)``; returns ``(frames, len(channels))`` float32."""
    signal = np.asarray(signal)
    if signal.ndim == 1:
        signal = signal[:, None]
    if channels is None:
        channels = range(signal.shape[1])
    profile_db = np.asarray(profile_db, dtype=np.float64)
    if profile_db.shape != (fft_size // 2 + 1,):
        raise ValueError("profile_db needs one level per bin")
    return fourier.stft_mask_istft(signal[:, list(channels)], "gate", fft_size, hop, params=profile_db, gain_db=gain_db)


def heuristic_band_peaks(imdata_db, sr, fft_size, f_lower=100, f_upper=15000, num_bands=5):
    """The per-band valley detection of dropouts_gui.py:251, :269-288: integer band edges
    (``np.logspace(..., dtype=uint16)``), band bins by truncation, ``find_peaks(-vol, prominence=5)``.
    Returns ``[(f_lo, f_hi, bin_lo, bin_hi, peaks)]`` from the top band down.

    The reference multiplies the ``np.uint16`` band edge by ``fft_size`` directly.  Under numpy >= 2 (NEP 50) that
    product wraps in uint16 and every band comes out empty (the tool then returns its input unchanged); under the
    numpy 1.x the reference was written for it promotes to a Python-sized integer.  This follows the intended
    arithmetic: the edge is converted to ``int`` first (``tests/golden/make_golden_dropouts_full.py`` records both
    behaviours of the unmodified reference)."""
    bands = np.logspace(np.log2(f_lower), np.log2(f_upper), num=num_bands, endpoint=True, base=2, dtype=np.uint16)
    out = []
    for f_lo, f_hi in reversed(list(zip(bands[:-1], bands[1:]))):
        bin_lo = int(int(f_lo) * fft_size / sr)
        bin_hi = int(int(f_hi) * fft_size / sr)
        vol = np.mean(imdata_db[bin_lo:bin_hi], axis=0)
        peaks, _ = scipy.signal.find_peaks(-vol, prominence=5, rel_height=0.5)
        out.append((int(f_lo), int(f_hi), bin_lo, bin_hi, peaks))
    return out


def heuristic(signal, sr, fft_size=512, hop=32, max_width=0.03, max_slope=0.5, num_bands=5, bottom_freedom=1.0,
              f_lower=100, f_upper=15000):
    """dropouts_gui.py:241-323 (process_heuristic): band-wise valley detection on the Hann
    magnitude spectrogram, a linear patch across each accepted valley, the resulting gain applied
    to a band-passed copy of the signal.  Returns the corrected ``(frames, channels)`` array."""
    from scipy.signal import butter, sosfiltfilt
    signal = np.array(signal, copy=True)
    if signal.ndim == 1:
        signal = signal[:, None]
    d = int(max_width / 1.5 * sr / hop)
    for channel in range(signal.shape[1]):
        imdata = to_dB(np.array(fourier.get_mag(signal[:, channel], fft_size, hop, "hann")))
        n_frames = imdata.shape[1]
        correction = np.ones(n_frames) * 1000
        for f_lo, f_hi, bin_lo, bin_hi, peaks in heuristic_band_peaks(imdata, sr, fft_size, f_lower, f_upper, num_bands):
            vol = np.mean(imdata[bin_lo:bin_hi], axis=0)
            gain_curve = np.zeros(n_frames)
            for p in peaks:
                if 2 * d < p < n_frames - 2 * d - 1:
                    left = np.mean(vol[p - 2 * d:p - d])
                    right = np.mean(vol[p + d:p + 2 * d])
                    if abs((left - right) / (2 * d)) < max_slope:
                        patch = np.interp(range(2 * d + 1), (0, 2 * d), (left, right))
                        gain_curve[p - d:p + d + 1] = patch - vol[p - d:p + d + 1]
            correction = np.clip(to_fac(gain_curve), 1, correction * bottom_freedom)
            x = signal[:, channel]
            extra = x * np.interp(np.linspace(0, 1, len(x)), np.linspace(0, 1, n_frames), correction - 1)
            low, high = f_lo / (0.5 * sr), f_hi / (0.5 * sr)
            if 0 < low < 1 and 0 < high < 1:
                extra = sosfiltfilt(butter(3, [low, high], btype="band", output="sos"), extra)
            elif 0 < low < 1:
                extra = sosfiltfilt(butter(3, low, btype="high", output="sos"), extra)
            elif 0 < high < 1:
                extra = sosfiltfilt(butter(3, high, btype="low", output="sos"), extra)
            signal[:, channel] += extra.astype(signal.dtype)
    return signal
