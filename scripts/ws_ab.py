"""A/B of the two sinc kernels: PAR_B200_SINC_WS=0|1 python scripts/ws_ab.py -> one line per case with time and sha1.
The two runs must print identical hashes (same records, same units, same arithmetic)."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyaudiorestoration_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
sr = 96000


def case(name, n, C, nt, speed_fn, reps=5, flags=0):
    x = torch.from_numpy(np.stack([bench.synth_channel(n, sr, 77 + c) for c in range(C)])).to(dev)
    t = np.arange(0, n, 1024, dtype=np.float64)
    t = np.append(t, float(n))
    sp = speed_fn(t / n)
    st = np.ascontiguousarray(t)
    cap = int(n * max(1.0 / max(sp.min(), 0.2), sp.max()) * 1.05) + 4096
    pos = torch.empty(cap, dtype=torch.float64, device=dev)
    out = torch.zeros((C, cap), dtype=torch.float32, device=dev)
    mbox = np.zeros(1, np.int64)
    _lib.check(L.par_speed_to_pos_f64(st.ctypes.data, np.ascontiguousarray(sp).ctypes.data, len(st), float(n), pos.data_ptr(), cap,
                                      mbox.ctypes.data, _lib.PAR_DEVICE_PTRS, 0, stream), "pos")
    m = int(mbox[0])

    def run():
        _lib.check(L.par_sinc_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, C, n, nt, out.data_ptr(), 1, cap,
                                           _lib.PAR_DEVICE_PTRS | flags, 0, stream), "sinc")
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    h = hashlib.sha1(out[:, :m].cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"{name:34s} C={C} nt={nt:3d} m={m:9d}  {np.median(ts):8.3f} ms  {h}", flush=True)


wow = lambda u: 1.0 + 0.003 * np.sin(2 * np.pi * 40 * u)
n10 = sr * 600
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    case("cfg2 wow 600 s", n10, 2, 128, wow)
    case("wow 600 s 100 taps", n10, 2, 50, wow)
    case("fast 1.3x (lowpass)", sr * 120, 2, 128, lambda u: 1.3 + 0.05 * np.sin(2 * np.pi * 9 * u))
    case("wow 150 s 4 ch", sr * 150, 4, 128, wow)
    sys.exit(0)
case("cfg2 wow 600 s", n10, 2, 128, wow)
case("cfg2 wow 600 s", n10, 1, 128, wow)
case("wow 600 s 100 taps", n10, 2, 50, wow)
case("wow 150 s 4 ch", sr * 150, 4, 128, wow)
case("wow 150 s 3 ch", sr * 150, 3, 128, wow)
case("fast 1.3x (lowpass)", sr * 120, 2, 128, lambda u: 1.3 + 0.05 * np.sin(2 * np.pi * 9 * u))
case("slow 0.7x", sr * 120, 2, 128, lambda u: 0.7 + 0.05 * np.sin(2 * np.pi * 9 * u))
case("1.03x + wow", sr * 300, 2, 128, lambda u: 1.03 + 0.003 * np.sin(2 * np.pi * 40 * u))
case("0.95x + wow", sr * 300, 2, 128, lambda u: 0.95 + 0.003 * np.sin(2 * np.pi * 40 * u))
case("mixed 0.5..2.6x", sr * 60, 2, 128, lambda u: 1.55 + 1.05 * np.sin(2 * np.pi * 5 * u))
case("3.5x (span over cap)", sr * 60, 2, 128, lambda u: 3.5 + 0 * u)
case("short 3000 samples", 3000, 2, 128, wow)
case("short 300 samples", 300, 1, 128, wow)
case("nt 8", sr * 60, 2, 8, wow)
case("nt 300 (large table)", sr * 30, 2, 300, wow)
case("aligned edges", sr * 30, 2, 128, wow, flags=_lib.PAR_SINC_ALIGNED_EDGES)
