import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from test_gpu_parity import synth, wow_curve
from pyaudiorestoration_b200.util import resampling
sr = 96000
sig = np.stack([synth(sr * 4, 61), synth(sr * 4, 62)], axis=1)
curve = wow_curve(4.0, sr, 1024, depth=0.05, freq=0.9)
ref = resampling.varispeed(sig, sr, curve, None, "Sinc", 50).copy()
for chunk in ("4096", "70000", "1000000"):
    os.environ["PAR_B200_CHUNK_BYTES"] = chunk
    v = resampling.varispeed(sig, sr, curve, None, "Sinc", 50)
    d = np.abs(v - ref).max(axis=1)
    bad = np.nonzero(d > 0)[0]
    print(chunk, len(bad), d.max(), bad[:20], bad[-5:] if len(bad) else None)
