"""Reference-run fixtures for BASELINE config 4 on the WHOLE ``samples/dropouts_sample.flac``:

* D1  ``dropout_healer_gui.Canvas.resample_files`` (dropout_healer_gui.py:111-166) with ALL markers of
      ``samples/dropouts_sample.drop``                                   -> ``healed``
* D2  ``dropout_healer_gui.Canvas.on_mouse_release`` with an Alt-drag (dropout_healer_gui.py:168-242):
      the batch locator, driven by a fake mouse event over the whole file -> ``loc_peaks`` (the integer
      frame indices ``scipy.signal.find_peaks`` returned inside the handler), ``loc_markers``
* D3  ``dropouts_gui.MainWindow.process_max_mono`` (dropouts_gui.py:137-163) on a stereo take made of the
      sample and a delayed, attenuated copy                               -> ``mm_max``, ``mm_min``
      ``dropouts_gui.MainWindow.process_heuristic`` (dropouts_gui.py:241-323) on the sample
                                                                          -> ``heur_*``

Run in the authoring container only:   python tests/golden/make_golden_dropouts_full.py [/root/reference]

All methods run UNMODIFIED and unbound on ``types.SimpleNamespace`` stand-ins for the Qt objects; the
GUI-only imports (PyQt5, vispy, matplotlib, librosa and the reference's own spectrum / widgets / markers /
undo / qt_threads / config modules) are inert stubs, ``soundfile`` is a stub whose reader hands out the decoded
sample and whose writer captures what the reference would have written.  ``util.fourier``, ``util.units``,
``util.io_ops``, ``util.filters`` are the reference's real modules (numpy back-end of ``stft``).

``process_heuristic`` computes its band edges as ``np.uint16 * int`` (dropouts_gui.py:251, :283-284).  Under the
numpy this image has (2.3.5, NEP 50) that product wraps in uint16, every band comes out empty and the method
returns its input unchanged (``heur_np2_unchanged``).  The reference predates NEP 50 (requirements.txt does not
pin numpy): under numpy 1.x the same expression promotes to a Python-sized integer.  The fixture therefore also
runs the method with ``np.logspace`` wrapped to hand out int64 band edges -- the ONLY change, made in the harness,
not in the reference -- and records the per-band peaks ``find_peaks`` returned and the corrected signal
(``heur_band_peaks_*``, ``heur_out``).  The product follows that intended arithmetic.
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

FFT_SIZE, OVERLAP = 512, 16
SENSITIVITY, WIDTH_MS = 4.0, 20.0           # widgets.DropoutWidget defaults are GUI state: fixed here
HEUR = dict(max_width=0.03, max_slope=0.5, num_bands=5, bottom_freedom=1.0, f_upper=15000, f_lower=100)


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
    mono16 = pcm[:, :1].astype(np.int16)
    files = {"mono.flac": mono16}
    # a deterministic stereo take for max/min-mono: the sample and a copy delayed by 37 samples at 0.8 gain
    right = np.zeros_like(mono16)
    right[37:] = (mono16[:-37].astype(np.int32) * 4 // 5).astype(np.int16)
    files["stereo.flac"] = np.concatenate([mono16, right], axis=1)
    captured = {}

    class SoundFile:
        def __init__(self, path, mode="r", samplerate=None, channels=None, subtype=None):
            self.path, self.mode, self.samplerate = path, mode, samplerate or sr
            self.channels = channels or (files[os.path.basename(path)].shape[1] if mode == "r" else 1)

        def read(self, always_2d=True, dtype="float32"):
            return (files[os.path.basename(self.path)].astype(np.float64) / 32768.0).astype(np.float32)

        def write(self, data):
            captured[os.path.basename(self.path)] = np.array(data)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    stub("soundfile", SoundFile=SoundFile)
    stub("librosa")
    stub("matplotlib")
    stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    qt = stub("PyQt5")
    qt.QtWidgets = stub("PyQt5.QtWidgets", QMainWindow=object)
    qt.QtCore = stub("PyQt5.QtCore")
    sys.path.insert(0, ref)
    import util  # the reference's package  # noqa: E402

    class DropoutSample:                      # records what markers.DropoutSample(canvas, a, b) is given
        def __init__(self, canvas, a, b, surrounding=0.5):
            self.a, self.b = a, b

    for name, attrs in (("util.undo", {"AddAction": lambda m: m}),
                        ("util.spectrum", {"SpectrumCanvas": object}),
                        ("util.qt_threads", {}),
                        ("util.widgets", {"MainWindow": object, "ParamWidget": object}),
                        ("util.markers", {"DropoutSample": DropoutSample}),
                        ("util.config", {"logging_setup": lambda: None, "load_config": lambda: {}}),
                        ("util.correlation", {"xcorr": None})):
        setattr(util, name.split(".")[1], stub(name, **attrs))
    import scipy.signal
    import dropout_healer_gui as g  # noqa: E402
    import dropouts_gui as dg       # noqa: E402
    from util import fourier as ref_fourier  # the reference's module

    found_peaks = []
    real_find_peaks = scipy.signal.find_peaks

    def recording_find_peaks(*a, **k):
        r = real_find_peaks(*a, **k)
        found_peaks.append(np.array(r[0], dtype=np.int64))
        return r
    scipy.signal.find_peaks = recording_find_peaks

    hop = FFT_SIZE // OVERLAP
    ns = types.SimpleNamespace
    out = {}

    # ---- D1: heal with all markers ----
    drop = json.load(open(os.path.join(ref, "samples", "dropouts_sample.drop")))
    surrounding = float(drop.get("surrounding", 0.5))
    marks = []
    for entry in drop["dropouts"]:
        a0, a1, b0, b1 = entry[:4]                       # (t, f) corners; the file's 6-tuples do not load any more
        marks.append(ns(t=(a0 + b0) / 2, width=abs(a0 - b0), f=(a1 + b1) / 2, height=abs(a1 - b1), surrounding=surrounding))
    pushed = []
    fake = ns(props=ns(files_widget=ns(files=[ns(channel_widget=ns(channels=[0]))]),
                       output_widget=ns(bump_index=lambda: None, suffix=""),
                       dropout_widget=ns(width=WIDTH_MS, sensitivity=SENSITIVITY),
                       undo_stack=ns(push=pushed.append)),
              filenames=["mono.flac", "mono.flac"], markers=marks, fft_size=FFT_SIZE, hop=hop, sr=sr)
    for meth in ("time_2_frame", "frame_2_time", "freq_2_bin"):
        setattr(fake, meth, types.MethodType(getattr(g.Canvas, meth), fake))
    g.Canvas.resample_files(fake, ["mono.flac"])
    healed = captured.pop("mono_drops.wav")
    regions = [(fake.time_2_frame(m.t - m.width / 2), fake.time_2_frame(m.t + m.width / 2),
                max(1, fake.time_2_frame(m.width * m.surrounding)),
                fake.freq_2_bin(m.f - m.height / 2), fake.freq_2_bin(m.f + m.height / 2)) for m in marks]
    out.update(markers=np.array([(m.t, m.width, m.f, m.height, m.surrounding) for m in marks]),
               regions=np.array(regions, dtype=np.int64), healed=healed[:, 0].astype(np.float32))

    # ---- D2: the Alt-drag locator over (0.2 s, 300 Hz) .. (7.0 s, 9000 Hz) ----
    x = (mono16[:, 0].astype(np.float64) / 32768.0).astype(np.float32)
    mag = np.array(ref_fourier.get_mag(x, FFT_SIZE, hop, "blackmanharris", 1))
    key = (FFT_SIZE, 0, hop, 1)
    spec = ns(fft_storage={key: mag}, key=key, sr=sr)
    # Spectrum.get_times_freqs (util/spectrum.py:173-178) compiled from the reference's source: the module itself
    # cannot be imported (vispy is not installed)
    import ast
    tree = ast.parse(open(os.path.join(ref, "util", "spectrum.py")).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_times_freqs")
    scope = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "util/spectrum.py", "exec"), scope)
    spec.get_times_freqs = types.MethodType(scope["get_times_freqs"], spec)
    corner_a, corner_b = (0.2, 300.0), (7.0, 9000.0)
    fake.spectra = [spec]
    fake.px_to_spectrum = lambda px: px                     # the fake event already carries spectrum coordinates
    event = ns(trail=lambda: [corner_a], button=1, pos=corner_b, modifiers=("Alt",))
    found_peaks.clear()
    g.Canvas.on_mouse_release(fake, event)
    loc_peaks, = found_peaks
    loc_markers = np.array([(m.a[0], m.a[1], m.b[0], m.b[1]) for m in pushed[-1]], dtype=np.float64)
    out.update(loc_corners=np.array([corner_a, corner_b]), loc_sensitivity=np.array(SENSITIVITY), loc_width_ms=np.array(WIDTH_MS),
               loc_peaks=loc_peaks, loc_markers=loc_markers)

    # ---- D3a: max / min mono ----
    win = ns(file_names=["stereo.flac"], names_to_full_paths={"stereo.flac": "stereo.flac"})
    dg.MainWindow.process_max_mono(win, FFT_SIZE, hop)
    out.update(mm_right_delay=np.array(37), mm_max=captured.pop("stereomax.wav").astype(np.float32),
               mm_min=captured.pop("stereomin.wav").astype(np.float32))

    # ---- D3b: heuristic, as it runs under this numpy, and with integer band edges (numpy 1.x arithmetic) ----
    win = ns(file_names=["mono.flac"], names_to_full_paths={"mono.flac": "mono.flac"}, dropout_widget=ns(**HEUR))
    found_peaks.clear()
    dg.MainWindow.process_heuristic(win, FFT_SIZE, hop)
    np2 = captured.pop("mono_out.wav")
    out["heur_np2_unchanged"] = np.array(bool(np.array_equal(np2[:, 0], x)))
    out["heur_np2_band_peak_counts"] = np.array([len(p) for p in found_peaks])

    real_logspace = np.logspace

    def int_logspace(*a, **k):
        r = real_logspace(*a, **k)
        return r.astype(np.int64) if r.dtype == np.uint16 else r
    np.logspace = int_logspace
    found_peaks.clear()
    try:
        dg.MainWindow.process_heuristic(win, FFT_SIZE, hop)
    finally:
        np.logspace = real_logspace
    heur = captured.pop("mono_out.wav")
    for i, p in enumerate(found_peaks):
        out[f"heur_band_peaks_{i}"] = p
    out.update(heur_params=np.array(json.dumps(HEUR)), heur_out=heur[:, 0].astype(np.float32))
    scipy.signal.find_peaks = real_find_peaks

    path = os.path.join(HERE, "dropouts_full.npz")
    np.savez_compressed(path, pcm=mono16[:, 0], sr=np.array(sr), fft_size=np.array(FFT_SIZE), hop=np.array(hop), **out)
    print(path, os.path.getsize(path) // 1024, "KiB;", len(marks), "markers healed;", len(loc_peaks), "located peaks:",
          loc_peaks[:8], "...; heuristic band peaks", [len(p) for p in found_peaks],
          "numpy-2 run unchanged:", bool(out["heur_np2_unchanged"]),
          "heuristic changed samples:", int(np.sum(heur[:, 0] != x)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
