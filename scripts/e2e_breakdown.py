"""Where does the end-to-end time of the reference-facing calls go?  (development aid)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyaudiorestoration_b200 import _lib
from pyaudiorestoration_b200.util import fourier, resampling
import bench

sr, dur, C = 96000, 600.0, 2
n = int(sr * dur)
sig = _lib.pinned_empty((n, C), np.float32)
tmp = np.empty(n, np.float32)
for c in range(C):
    bench.synth_channel(n, sr, 1234 + c, out=tmp); sig[:, c] = tmp
curve = bench.wow_curve(dur, sr)

def t(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0); del r
    return best * 1e3

# raw PCIe
d = torch.empty(n * C, dtype=torch.float32, device="cuda")
h = torch.from_numpy(sig.reshape(-1))
print("H2D pinned 461 MB: %.1f ms" % t(lambda: d.copy_(h, non_blocking=True)))
big = torch.empty(922 * 1000 * 1000 // 4, dtype=torch.float32, device="cuda")
hb = torch.empty(big.numel(), dtype=torch.float32).pin_memory()
ms = t(lambda: hb.copy_(big, non_blocking=True)); print("D2H pinned 922 MB: %.1f ms (%.1f GB/s)" % (ms, 0.922 / ms * 1e3))
pg = np.empty(n * C, np.float32); hp = torch.from_numpy(pg)
print("H2D pageable 461 MB: %.1f ms" % t(lambda: d.copy_(hp)))
print("stft(sig[:,0])          %.1f ms" % t(lambda: fourier.stft(sig[:, 0], 4096, 1024)))
print("get_mag(sig[:,0])       %.1f ms" % t(lambda: fourier.get_mag(sig[:, 0], 4096, 1024)))
print("stft_multi(sig)         %.1f ms" % t(lambda: fourier.stft_multi(sig, 4096, 1024)))
print("varispeed sinc 128      %.1f ms" % t(lambda: resampling.varispeed(sig, sr, curve, None, "Sinc", 128)))
print("varispeed sinc 50       %.1f ms" % t(lambda: resampling.varispeed(sig, sr, curve, None, "Sinc", 50)))
print("varispeed linear        %.1f ms" % t(lambda: resampling.varispeed(sig, sr, curve, None, "Linear", 50)))
print("speed_to_pos            %.1f ms" % t(lambda: resampling.speed_to_pos(curve[:, 0] * sr, curve[:, 1], n)))
print("pinned_empty 922MB (cached) %.2f ms" % t(lambda: _lib.pinned_empty((56251, 2049), np.complex64)))
