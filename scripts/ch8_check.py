"""development aid: time the 8-channel sinc resampler with channel groups of 4 vs 8 (PAR_B200_SINC_CH8)"""
import os, sys, subprocess
if len(sys.argv) == 1:
    for v in ("0", "1"):
        env = dict(os.environ, PAR_B200_SINC_CH8=v)
        print("PAR_B200_SINC_CH8=" + v, subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip())
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from pyaudiorestoration_b200 import _lib
L = _lib.lib(); sr, dur, C = 192000, 60.0, 8
n = int(sr * dur); dev = torch.device("cuda", 0)
x = torch.randn((C, n), device=dev) * 0.1
curve = bench.wow_curve(dur, sr); st, sp = np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1])
cap = int(n * 1.02) + 4096
pos = torch.empty(cap, dtype=torch.float64, device=dev); out = torch.empty((C, cap), device=dev); mb = np.zeros(1, np.int64)
s = torch.cuda.current_stream().cuda_stream
_lib.check(L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos.data_ptr(), cap, mb.ctypes.data, 1, 0, s), "pos")
m = int(mb[0])
def f(): _lib.check(L.par_sinc_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, C, n, 128, out.data_ptr(), 1, cap, 1, 0, s), "sinc")
for _ in range(3): f()
torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); [f() for _ in range(5)]; b.record(); torch.cuda.synchronize()
print("8 ch x %d samples, NT 128: %.3f ms" % (n, a.elapsed_time(b) / 5))
