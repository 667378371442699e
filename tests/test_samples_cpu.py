"""CPU tests around the hot path's data formats and the dropout tools' oracle (no GPU):
FLAC decode (MD5-checked on the reference's samples when they are present, and on synthetic
streams anywhere), and the oracle restatements pinned by the golden output of the UNMODIFIED
reference (tests/golden/dropouts.npz, flutter.npz <- tests/golden/make_golden_dropouts.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle_np as onp
from pyaudiorestoration_b200.util import flac, io_ops

from flac_writer import write_flac

REF_SAMPLES = "/root/reference/samples"


@pytest.mark.parametrize("bps,channels,mid_side", [(16, 1, False), (16, 2, False), (16, 2, True), (24, 2, True), (8, 3, False)])
def test_flac_decoder_on_synthetic_streams(tmp_path, bps, channels, mid_side):
    rng = np.random.default_rng(bps + channels)
    lim = 1 << (bps - 2)
    pcm = [rng.integers(-lim, lim, 2500).tolist() for _ in range(channels)]
    pcm[0][1000:2000] = [pcm[0][1000]] * 1000                    # a CONSTANT subframe
    data = write_flac(pcm, 48000, bps=bps, blocksize=1000, mid_side=mid_side)
    got, sr, b = flac.decode_flac(data, verify_md5=True)
    assert sr == 48000 and b == bps and got.shape == (2500, channels)
    assert np.array_equal(got.T, np.array(pcm))
    p = tmp_path / "x.flac"
    p.write_bytes(data)
    sig, sr2, ch = io_ops.read_file(str(p))                      # util/io_ops.py:7-16 contract
    assert sig.dtype == np.float32 and sig.shape == (2500, channels) and ch == channels and sr2 == 48000
    assert np.array_equal(sig, (np.array(pcm).T / float(1 << (bps - 1))).astype(np.float32))
    corrupted = bytearray(data)
    corrupted[-10] ^= 0x40
    with pytest.raises(ValueError):
        flac.decode_flac(bytes(corrupted))


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="reference samples only exist in the authoring container")
def test_flac_decoder_on_reference_samples():
    files = sorted(glob.glob(os.path.join(REF_SAMPLES, "*.flac")))
    assert len(files) >= 3
    for f in files[:3]:                                          # Rice + fixed/LPC subframes; MD5 pins the decode
        sig, sr, ch = flac.read_flac(f, verify_md5=True)
        assert ch == 1 and sr in (44100, 192000) and len(sig) > 100000
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "flutter.npz"))
    pcm, _, _ = flac.decode_flac(open(os.path.join(REF_SAMPLES, "flutter.flac"), "rb").read())
    assert np.array_equal(pcm[:, 0], z["pcm"])


def test_cfg1_oracle_matches_reference_on_flutter(golden_dir):
    """BASELINE config 1: util.fourier.stft(n_fft=4096, hop=1024) on samples/flutter.flac."""
    z = np.load(os.path.join(golden_dir, "flutter.npz"))
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    s = onp.stft_ref(x, 4096, 1024)
    assert tuple(z["shape"]) == s.shape == (2049, len(x) // 1024 + 1)
    assert np.array_equal(s[:, z["frames"]].astype(np.complex64), z["S"])
    assert np.allclose(np.sum(np.abs(s) ** 2, axis=0), z["frame_energy"], rtol=1e-12)


def test_cfg4_oracle_heal_matches_reference_bitwise(golden_dir):
    """BASELINE config 4: the oracle's heal reproduces the unmodified dropout_healer_gui body."""
    z = np.load(os.path.join(golden_dir, "dropouts.npz"))
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    sr, fft_size, hop = int(z["sr"]), int(z["fft_size"]), int(z["hop"])
    assert np.array_equal(onp.heal_regions(z["markers"], sr, fft_size, hop), z["regions"])
    assert z["regions"].shape == (25, 5)
    y = onp.heal_ref(x, sr, z["markers"], fft_size, hop)
    assert y.dtype == np.float32 and np.array_equal(y, z["healed"])
    assert np.max(np.abs(y - x)) > 0.01                           # the markers did change the audio
    # product-side integer rules agree with the reference's (no GPU needed for these)
    from pyaudiorestoration_b200 import dropouts
    regs = [dropouts.marker_region(dropouts.Dropout(*m), sr, fft_size, hop) for m in z["markers"].tolist()]
    assert np.array_equal(np.array(regs), z["regions"])


def test_tracker_oracle_matches_reference_classes_bitwise(golden_dir):
    """the oracle's restatement of PeakTracker / PeakTrackTracker / CenterOfGravity
    against tests/golden/trackers.npz (the unmodified reference classes, make_golden_trackers.py)."""
    z = np.load(os.path.join(golden_dir, "trackers.npz"))
    rng = np.random.default_rng(5)
    sr, dur = 44100, 3.0
    t = np.arange(int(sr * dur)) / sr
    phase = 2 * np.pi * np.cumsum(3150.0 * (1 + 0.006 * np.sin(2 * np.pi * 0.8 * t))) / sr
    x = (0.3 * np.sin(phase) + 0.05 * np.sin(2.31 * phase) + 0.01 * rng.standard_normal(len(t))).astype(np.float32)
    fft_size, hop, sr2, zp = (int(v) for v in z["params"])
    spec = onp.to_mag(onp.stft_ref(x, fft_size, hop)).astype(np.float32)
    assert float(spec.astype(np.float64).sum()) == z["spec_checksum"][0]
    trail = [tuple(r) for r in z["trail"]]
    for mode in ("peak", "peak_track", "cog"):
        times, freqs = onp.track_ref(mode, spec, trail, fft_size, hop, sr)
        assert np.array_equal(times, z[mode + "__times"]) and np.array_equal(freqs, z[mode + "__freqs"]), mode


# ---- BASELINE config 4 on the WHOLE sample: fixtures from tests/golden/make_golden_dropouts_full.py --------------
@pytest.fixture(scope="module")
def full(golden_dir):
    z = np.load(os.path.join(golden_dir, "dropouts_full.npz"))
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    return z, x, int(z["sr"]), int(z["fft_size"]), int(z["hop"])


def test_cfg4_full_oracle_heal_all_markers_bitwise(full):
    """D1: the oracle's heal equals the unmodified Canvas.resample_files on the whole file, all 32 markers."""
    z, x, sr, fft_size, hop = full
    assert len(z["markers"]) == 32
    assert np.array_equal(onp.heal_regions(z["markers"], sr, fft_size, hop), z["regions"])
    y = onp.heal_ref(x, sr, z["markers"], fft_size, hop)
    assert y.dtype == np.float32 and np.array_equal(y, z["healed"])


def test_cfg4_full_oracle_locator_peaks_are_the_references(full):
    """D2: the integer frame indices find_peaks returned INSIDE the reference's Alt-drag handler
    (dropout_healer_gui.py:185-204) equal the oracle's, so the oracle is pinned for the bit-exact index check."""
    z, x, sr, fft_size, hop = full
    (t0, f0), (t1, f1) = z["loc_corners"]
    f_lo, f_hi = max(min(f0, f1), 1), min(max(f0, f1), sr // 2 - 1)            # Spectrum.get_times_freqs, util/spectrum.py:173-178
    mag = onp.to_mag(onp.stft_ref(x, fft_size, hop))
    peaks = onp.locate_peaks_ref(mag, sr, fft_size, hop, min(t0, t1), max(t0, t1), f_lo, f_hi, float(z["loc_sensitivity"]))
    assert len(z["loc_peaks"]) >= 30 and np.array_equal(peaks, z["loc_peaks"])


def test_cfg4_full_oracle_batch_tool(full):
    """D3: max/min-mono outputs and the per-band peaks of process_heuristic (dropouts_gui.py:137-163, :241-323)."""
    z, x, sr, fft_size, hop = full
    right = np.zeros_like(x)
    d = int(z["mm_right_delay"])
    pcm_r = np.zeros_like(z["pcm"])
    pcm_r[d:] = (z["pcm"][:-d].astype(np.int32) * 4 // 5).astype(np.int16)
    right = (pcm_r.astype(np.float64) / 32768.0).astype(np.float32)
    out = onp.max_mono_ref(np.stack([x, right], axis=1), fft_size, hop)
    # the numpy back-end hands istft a complex128 matrix, so the reference's result is float64; the fixture keeps float32
    assert np.array_equal(out["max"].astype(np.float32), z["mm_max"]) and np.array_equal(out["min"].astype(np.float32), z["mm_min"])
    # heuristic: under this numpy the unmodified reference finds nothing (uint16 overflow) ...
    assert bool(z["heur_np2_unchanged"]) and not z["heur_np2_band_peak_counts"].any()
    # ... with integer band edges (numpy 1.x arithmetic) it finds these valleys, and so does the oracle
    cpu_h = onp.to_mag(onp.stft_ref(x, fft_size, hop, "hann"))
    got = onp.heuristic_peaks_ref(cpu_h, sr, fft_size, 100, 15000, 5)
    assert len(got) == 4
    for i, g in enumerate(got):
        assert len(g) > 0 and np.array_equal(g, z[f"heur_band_peaks_{i}"])
    assert np.sum(z["heur_out"] != x) > 1000


def test_noise_gate_oracle_matches_reference_run(golden_dir):
    """the oracle's spectral gate equals the unmodified renoiser_gui.Canvas.run_resample output
    (tests/golden/make_golden_gate.py)."""
    z = np.load(os.path.join(golden_dir, "gate.npz"))
    x = (z["pcm"].astype(np.float64) / 32768.0).astype(np.float32)
    y = onp.noise_gate_ref(x, z["profile_db"], float(z["gain_db"]), int(z["fft_size"]), int(z["hop"]))
    assert np.array_equal(np.asarray(y).astype(np.float32), z["gated"])
