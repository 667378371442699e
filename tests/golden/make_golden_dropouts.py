"""Golden fixture for BASELINE config 4 (dropout repair): runs the UNMODIFIED heal body of the reference,
``dropout_healer_gui.Canvas.resample_files`` (dropout_healer_gui.py:111-166), on an excerpt of its own
sample ``samples/dropouts_sample.flac`` with the markers of ``samples/dropouts_sample.drop``.

Run in the authoring container only:   python tests/golden/make_golden_dropouts.py [/root/reference]

The method lives in a vispy/PyQt5 canvas class, so the GUI-only imports are satisfied with inert stub
modules (PyQt5, vispy, matplotlib and the reference's own util.spectrum/widgets/markers/undo/
qt_threads/config); ``util.fourier``, ``util.units`` and ``util.io_ops`` are the reference's real
modules.  ``soundfile`` is a stub whose reader is this repo's FLAC decoder (checked against the
STREAMINFO MD5) and whose writer captures the array the reference would have written.  The numeric
body -- fix_length, stft (numpy back-end), dB gain interpolation, istft -- runs as it is.

Output: tests/golden/dropouts.npz with the int16 excerpt, the marker list, the integer regions the
reference derives from them and the healed float32 audio.
"""
import json
import logging
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

T0, T1 = 2.0, 6.5          # excerpt of the sample, seconds
FFT_SIZE, OVERLAP = 512, 16


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main(ref):
    from pyaudiorestoration_b200.util import flac
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    pcm, sr, bps = flac.decode_flac(open(os.path.join(ref, "samples", "dropouts_sample.flac"), "rb").read(),
                                    verify_md5=True)
    s0, s1 = int(T0 * sr), int(T1 * sr)
    excerpt = pcm[s0:s1, :1].astype(np.int16)
    captured = {}

    class SoundFile:
        def __init__(self, path, mode="r", samplerate=None, channels=None, subtype=None):
            self.path, self.mode, self.samplerate, self.channels = path, mode, samplerate or sr, channels or 1

        def read(self, always_2d=True, dtype="float32"):
            return (excerpt.astype(np.float64) / 32768.0).astype(np.float32)

        def write(self, data):
            captured[self.path] = np.array(data)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    stub("soundfile", SoundFile=SoundFile)
    stub("matplotlib")
    stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    qt = stub("PyQt5")
    qt.QtWidgets = stub("PyQt5.QtWidgets")
    qt.QtCore = stub("PyQt5.QtCore")
    sys.path.insert(0, ref)
    import util  # the reference's package  # noqa: E402
    for name, attrs in (("util.undo", {"AddAction": object}),
                        ("util.spectrum", {"SpectrumCanvas": object}),
                        ("util.qt_threads", {}),
                        ("util.widgets", {"MainWindow": object, "ParamWidget": object}),
                        ("util.markers", {"DropoutSample": object}),
                        ("util.config", {"logging_setup": lambda: None})):
        setattr(util, name.split(".")[1], stub(name, **attrs))
    import dropout_healer_gui as g  # noqa: E402

    drop = json.load(open(os.path.join(ref, "samples", "dropouts_sample.drop")))
    surrounding = float(drop.get("surrounding", 0.5))
    marks = []
    for entry in drop["dropouts"]:
        a0, a1, b0, b1 = entry[:4]                       # (t, f) corners; the file's 6-tuples do not load any more
        if a0 - T0 > 0.2 and b0 - T0 < (T1 - T0) - 0.2:
            a, b = (a0 - T0, a1), (b0 - T0, b1)
            marks.append(types.SimpleNamespace(t=(a[0] + b[0]) / 2, width=abs(a[0] - b[0]), f=(a[1] + b[1]) / 2,
                                               height=abs(a[1] - b[1]), surrounding=surrounding))
    hop = FFT_SIZE // OVERLAP
    ns = types.SimpleNamespace
    fake = ns(props=ns(files_widget=ns(files=[ns(channel_widget=ns(channels=[0]))]),
                       output_widget=ns(bump_index=lambda: None, suffix="")),
              filenames=["x.flac", "x.flac"], markers=marks, fft_size=FFT_SIZE, hop=hop, sr=sr)
    for meth in ("time_2_frame", "frame_2_time", "freq_2_bin"):
        setattr(fake, meth, types.MethodType(getattr(g.Canvas, meth), fake))
    g.Canvas.resample_files(fake, ["x.flac"])
    (path, healed), = captured.items()
    regions = []
    for m in marks:
        regions.append((fake.time_2_frame(m.t - m.width / 2), fake.time_2_frame(m.t + m.width / 2),
                        max(1, fake.time_2_frame(m.width * m.surrounding)),
                        fake.freq_2_bin(m.f - m.height / 2), fake.freq_2_bin(m.f + m.height / 2)))
    out = os.path.join(HERE, "dropouts.npz")
    np.savez_compressed(out, pcm=excerpt[:, 0], sr=np.array(sr), fft_size=np.array(FFT_SIZE), hop=np.array(hop),
                        markers=np.array([(m.t, m.width, m.f, m.height, m.surrounding) for m in marks]),
                        regions=np.array(regions, dtype=np.int64), healed=healed[:, 0].astype(np.float32))
    print(out, os.path.getsize(out) // 1024, "KiB;", len(marks), "markers; output", healed.shape, healed.dtype,
          "written to", path)

    # ---- BASELINE config 1: samples/flutter.flac, util.fourier.stft n_fft=4096 hop=1024 on the CPU.
    # pyfftw is not installed, so the reference's stft() lands in its numpy back-end (util/fourier.py:67-75).
    from util import fourier as ref_fourier  # the reference's module
    pcm1, sr1, _ = flac.decode_flac(open(os.path.join(ref, "samples", "flutter.flac"), "rb").read(), verify_md5=True)
    x1 = (pcm1[:, 0].astype(np.float64) / 32768.0).astype(np.float32)
    s1_full = np.asarray(ref_fourier.stft(x1, n_fft=4096, step=1024))
    frames = np.array([0, 1, 2, 60, 120, s1_full.shape[1] - 2, s1_full.shape[1] - 1])
    out1 = os.path.join(HERE, "flutter.npz")
    np.savez_compressed(out1, pcm=pcm1[:, 0].astype(np.int16), sr=np.array(sr1), shape=np.array(s1_full.shape),
                        frames=frames, S=s1_full[:, frames].astype(np.complex64),
                        frame_energy=np.sum(np.abs(s1_full) ** 2, axis=0))
    print(out1, os.path.getsize(out1) // 1024, "KiB; STFT", s1_full.shape, s1_full.dtype)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
