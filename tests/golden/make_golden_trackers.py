"""Golden fixture for the frequency trackers: runs the UNMODIFIED tracker classes of the
reference's util/wow_detection.py (Peak, Peak Track, Center of Gravity, Zero-Crossing, Correlation) on a float32 magnitude spectrogram of a
seeded synthetic wow signal.  matplotlib (imported at module level by the reference, never used by these
classes) is an inert stub.

Run in the authoring container only:   python tests/golden/make_golden_trackers.py [/root/reference]
Output: tests/golden/trackers.npz (trail, parameters, one traced frequency array per tracker; the test
regenerates the signal from the seed).
"""
import logging
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def wow_signal(sr=44100, dur=3.0, seed=5):
    """A 3150 Hz pilot tone with +-0.6 % wow at 0.8 Hz, a weaker second partial and noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(int(sr * dur)) / sr
    inst = 3150.0 * (1 + 0.006 * np.sin(2 * np.pi * 0.8 * t))
    phase = 2 * np.pi * np.cumsum(inst) / sr
    x = 0.3 * np.sin(phase) + 0.05 * np.sin(2.31 * phase) + 0.01 * rng.standard_normal(len(t))
    return x.astype(np.float32), sr


def main(ref):
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    for name in ("matplotlib", "matplotlib.pyplot", "soundfile"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, ref)
    from util import fourier, wow_detection
    x, sr = wow_signal()
    fft_size, hop, zeropad = 4096, 256, 1
    window = __import__("scipy.signal").signal.get_window("blackmanharris", fft_size).astype(np.float32)
    spec = fourier.to_mag(np.asarray(fourier.np_rfft_pick(fft_size, hop, window, x, zeropad))).astype(np.float32)
    trail = [(0.35, 3140.0), (1.2, 3165.0), (2.0, 3135.0), (2.7, 3160.0)]
    out = {"trail": np.array(trail), "params": np.array([fft_size, hop, sr, zeropad]), "tolerance_st": np.array(1.0),
           "spec_checksum": np.array([float(spec.astype(np.float64).sum()), float(spec[300, 100])])}
    for key, name in (("peak", "Peak"), ("peak_track", "Peak Track"), ("cog", "Center of Gravity")):
        tr = wow_detection.wow_detectors[name](spec, x, list(trail), fft_size * zeropad, hop, sr, 1.0, "Linear")
        out[key + "__times"] = np.asarray(tr.times)
        out[key + "__freqs"] = np.asarray(tr.freqs)
        print(name, len(tr.freqs), tr.freqs[:3], float(np.std(tr.freqs)))
    # the two host-side trackers of the registry that take a waveform / compare neighbouring frames
    # (Zero-Crossing reads signal[s0:s1, 0]: a (samples, channels) array)
    for key, name in (("zero_crossing", "Zero-Crossing"), ("correlation", "Correlation")):
        tr = wow_detection.wow_detectors[name](spec, x[:, None], list(trail), fft_size * zeropad, hop, sr, 1.0, "Linear")
        out[key + "__times"] = np.asarray(tr.times)
        out[key + "__freqs"] = np.asarray(tr.freqs)
        print(name, len(tr.freqs), tr.freqs[:3], float(np.std(tr.freqs)))
    np.savez_compressed(os.path.join(HERE, "trackers.npz"), **out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
