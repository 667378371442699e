"""Where the end-to-end time of the cfg2 step goes: wall clock of each public call on pinned host arrays next to the
PCIe floor measured with plain pinned copies of the same byte counts (alone and in both directions at once)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyaudiorestoration_b200 import _lib  # noqa: E402
from pyaudiorestoration_b200.util import fourier, resampling  # noqa: E402

sr, dur, C = 96000, 600.0, 2
n = int(sr * dur)
dev = torch.device("cuda", 0)
sig = _lib.pinned_empty((n, C), np.float32)
for c in range(C):
    sig[:, c] = bench.synth_channel(n, sr, 1234 + c)
curve = bench.wow_curve(dur, sr)


def wall(f, reps=4):
    f()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        del r
    return float(np.median(ts)), float(np.min(ts))


for name, f in (("stft(sig[:, 0])", lambda: fourier.stft(sig[:, 0], 4096, 1024)),
                ("stft(sig[:, 1])", lambda: fourier.stft(sig[:, 1], 4096, 1024)),
                ("get_mag(sig[:, 0])", lambda: fourier.get_mag(sig[:, 0], 4096, 1024)),
                ("varispeed Sinc 128", lambda: resampling.varispeed(sig, sr, curve, range(C), "Sinc", 128)),
                ("varispeed Sinc 50", lambda: resampling.varispeed(sig, sr, curve, range(C), "Sinc", 50)),
                ("varispeed Linear", lambda: resampling.varispeed(sig, sr, curve, range(C), "Linear", 50))):
    med, best = wall(f)
    print(f"{name:24s} median {med:8.2f} ms  best {best:8.2f} ms", flush=True)

# PCIe floor with the same byte counts
hb_in = torch.from_numpy(sig.reshape(-1))
T, F = n // 1024 + 1, 2049
out_bytes = T * F * 8
hb_out = torch.empty(out_bytes // 4, dtype=torch.float32).pin_memory()
d_in = torch.empty(n * C, dtype=torch.float32, device=dev)
d_out = torch.empty(out_bytes // 4, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def copies(up, down):
    def f():
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(hb_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                hb_out.copy_(d_out, non_blocking=True)
    return f


for name, f, nbytes in (("H2D 461 MB alone", copies(True, False), n * C * 4), ("D2H 922 MB alone", copies(False, True), out_bytes),
                        ("H2D 461 MB + D2H 922 MB together", copies(True, True), out_bytes)):
    med, best = wall(f)
    print(f"{name:36s} median {med:8.2f} ms  ({nbytes / med / 1e6:.1f} GB/s on the longer leg)", flush=True)

if os.environ.get("PAR_B200_TRACE") == "1":
    import ctypes
    L = _lib.lib()
    for nm in ("par_stft_f32", "par_varispeed_f32"):
        fn = getattr(L, nm)

        def wrap(fn=fn, nm=nm):
            def g(*a):
                t0 = time.perf_counter()
                r = fn(*a)
                print(f"[py trace] {nm} took {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr, flush=True)
                return r
            return g
        setattr(L, nm, wrap())
    for name, f in (("stft", lambda: fourier.stft(sig[:, 0], 4096, 1024)),
                    ("varispeed", lambda: resampling.varispeed(sig, sr, curve, range(C), "Sinc", 128))):
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f()
            print(f"[py trace] {name} python call total {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr, flush=True)
            del r
