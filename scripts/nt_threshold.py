import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from pyaudiorestoration_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev).cuda_stream
sr = 96000; n = sr * 300; C = 2
x = torch.from_numpy(np.stack([bench.synth_channel(n, sr, 77 + c) for c in range(C)])).to(dev)
curve = bench.wow_curve(300.0, sr)
st, sp = np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1])
cap = int(n * 1.02) + 4096
pos = torch.empty(cap, dtype=torch.float64, device=dev); out = torch.empty((C, cap), dtype=torch.float32, device=dev)
mbox = np.zeros(1, np.int64)
_lib.check(L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos.data_ptr(), cap, mbox.ctypes.data, _lib.PAR_DEVICE_PTRS, 0, stream), "pos")
m = int(mbox[0])
for nt in (16, 24, 32, 40, 50, 64):
    res = []
    for flag in (_lib.PAR_SINC_KERNEL_TILED, _lib.PAR_SINC_KERNEL_WS):
        def run():
            _lib.check(L.par_sinc_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, C, n, nt, out.data_ptr(), 1, cap, _lib.PAR_DEVICE_PTRS | flag, 0, stream), "sinc")
        run(); torch.cuda.synchronize(); ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res.append(np.median(ts))
    print(f"nt={nt:3d}  tiled {res[0]:.3f} ms   ws {res[1]:.3f} ms", flush=True)
