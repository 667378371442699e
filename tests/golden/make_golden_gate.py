"""Reference-run fixture for the spectral noise gate: the UNMODIFIED
``renoiser_gui.Canvas.run_resample`` + ``get_mask_fac`` (renoiser_gui.py:273-278, :303-319) on an excerpt of
``samples/dropouts_sample.flac``, headless behind inert stubs of the GUI imports (PyQt5, matplotlib, resampy, vispy
and the reference's own spectrum / widgets / markers / ... modules); ``util.fourier`` and ``util.units`` are real.

Run in the authoring container only:   python tests/golden/make_golden_gate.py [/root/reference]

The profile is built the way the tool does (load_noise_profile :239-250 + redraw_plot :288-290): the mean dB spectrum
of a noise-only stretch + gain + overhead, flat control curve.  Output: tests/golden/gate.npz.
"""
import logging
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

FFT_SIZE, HOP = 512, 32
T0, T1 = 2.0, 3.5
GAIN_DB, OVERHEAD_DB = -12.0, 3.0


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main(ref):
    from pyaudiorestoration_b200.util import flac
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    pcm, sr, _ = flac.decode_flac(open(os.path.join(ref, "samples", "dropouts_sample.flac"), "rb").read(), verify_md5=True)
    excerpt = pcm[int(T0 * sr):int(T1 * sr), :1].astype(np.int16)
    x = (excerpt.astype(np.float64) / 32768.0).astype(np.float32)
    captured = {}

    class SoundFile:
        def __init__(self, path, mode="r", samplerate=None, channels=None, subtype=None):
            self.path = path

        def write(self, data):
            captured[os.path.basename(self.path)] = np.array(data)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    stub("soundfile", SoundFile=SoundFile)
    stub("resampy")
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.patches = stub("matplotlib.patches")
    stub("matplotlib.backend_bases", MouseEvent=object)
    stub("matplotlib.backends")
    stub("matplotlib.backends.backend_qt5", NavigationToolbar2QT=object)
    stub("matplotlib.backends.backend_qt5agg", FigureCanvasQTAgg=object)
    qt = stub("PyQt5")
    qt.QtWidgets = stub("PyQt5.QtWidgets", QMainWindow=object)
    qt.QtCore = stub("PyQt5.QtCore", pyqtSignal=lambda *a, **k: None, QTimer=object)
    sys.path.insert(0, ref)
    import util  # the reference's package  # noqa: E402
    for name, attrs in (("util.undo", {"AddAction": object, "MoveAction": object, "MergeAction": object}),
                        ("util.spectrum", {"SpectrumCanvas": object}),
                        ("util.qt_threads", {}), ("util.markers", {}), ("util.vispy_ext", {}),
                        ("util.widgets", {"MainWindow": object, "ParamWidget": object, "vbox2": None}),
                        ("util.wow_detection", {"wow_detectors": {}}),
                        ("util.config", {"logging_setup": lambda: None})):
        setattr(util, name.split(".")[1], stub(name, **attrs))
    import renoiser_gui as g  # noqa: E402
    from util import fourier as ref_fourier
    from util.units import to_dB

    # load_noise_profile + redraw_plot: mean dB spectrum of the first 0.25 s, + gain + overhead
    noise_profile = np.average(to_dB(np.array(ref_fourier.get_mag(x[:int(0.25 * sr), 0], FFT_SIZE, HOP, "blackmanharris", zeropad=1))), axis=1)
    final_profile = noise_profile + GAIN_DB + 0.0 + OVERHEAD_DB
    ns = types.SimpleNamespace
    fake = ns(spectra=[ns(audio_path="x.wav", signal=x, sr=sr)], sr=sr, fft_size=FFT_SIZE, hop=HOP, final_profile=final_profile,
              props=ns(files_widget=ns(files=[ns(channel_widget=ns(channels=[0]))])),
              parent=ns(props=ns(noise_widget=ns(gain=GAIN_DB))))
    fake.get_mask_fac = types.MethodType(g.Canvas.get_mask_fac, fake)
    g.Canvas.run_resample(fake)
    (name, y), = captured.items()
    out = os.path.join(HERE, "gate.npz")
    np.savez_compressed(out, pcm=excerpt[:, 0], sr=np.array(sr), fft_size=np.array(FFT_SIZE), hop=np.array(HOP),
                        profile_db=final_profile, gain_db=np.array(GAIN_DB), gated=y[:, 0].astype(np.float32))
    changed = float(np.mean(np.abs(y[:, 0] - x[:, 0]) > 1e-6))
    print(out, os.path.getsize(out) // 1024, "KiB; output", y.shape, y.dtype, "written as", name, "; samples changed:", round(changed, 3))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
