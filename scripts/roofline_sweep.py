"""BASELINE config 5: roofline sweep -- n_fft in {1024, 4096, 16384, 65536}, hop n_fft/4,
STFT (complex and magnitude) + 256-tap and 100-tap sinc resample on the cfg2 waveform, device resident,
CUDA-event timed; achieved algorithmic GB/s against MEASURED_PEAKS.json; torch.stft (cuFFT, the
reference's own GPU back-end, util/fourier.py:92-121) timed beside it as the competitor.
Usage: python scripts/roofline_sweep.py [seconds]   (prints a markdown table)
       torchrun --nproc-per-node N scripts/roofline_sweep.py [seconds]: every rank sweeps its own two channels on its own
       GPU (the path shards by channel with nothing to exchange); times are the MAX over ranks, GB/s the whole job's."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyaudiorestoration_b200 import _lib  # noqa: E402
from scipy import signal as dsp  # noqa: E402

sr, dur, C = 96000, float(sys.argv[1]) if len(sys.argv) > 1 else 600.0, 2
n = int(sr * dur)
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
os.environ["PAR_B200_DEVICE"] = str(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
L = _lib.lib()
peak, src = bench.measured_peaks()
x_host = np.stack([bench.synth_channel(n, sr, 1234 + C * rank + c) for c in range(C)])
x = torch.from_numpy(x_host).to(dev)
stream = torch.cuda.current_stream(dev).cuda_stream
_print = print


def print(*a, **k):                                  # noqa: A001 -- rank 0 reports
    if rank == 0:
        _print(*a, **k)


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor(ts, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)         # per repetition: the slowest rank
    ts = t.cpu().numpy()
    return float(np.median(ts)), float(np.min(ts))


print(f"peak = {peak} GB/s per GPU ({src}); {world} GPU(s) x {C} ch x {n} samples; GB/s and frac are per GPU (whole job = x{world})\n")
print("| n_fft | hop | stage | median ms | best ms | algorithmic GB | GB/s (median) | frac of peak | torch.stft ms |")
print("|---|---|---|---|---|---|---|---|---|")
for n_fft in (1024, 4096, 16384, 65536):
    hop = n_fft // 4
    T = int(L.par_stft_num_frames(n, n_fft, hop))
    F = n_fft // 2 + 1
    win = np.ascontiguousarray(dsp.get_window("blackmanharris", n_fft), dtype=np.float32)
    wt = torch.from_numpy(win).to(dev)
    S = torch.empty((C, T, F), dtype=torch.complex64, device=dev)
    M = torch.empty((C, T, F), dtype=torch.float32, device=dev)

    def ours(mag=False):
        out = M if mag else S
        _lib.check(L.par_stft_f32(x.data_ptr(), n, 1, C, n, n_fft, hop, 1, win.ctypes.data, out.data_ptr(), F, T * F,
                                  _lib.PAR_DEVICE_PTRS | (_lib.PAR_OUT_MAGNITUDE if mag else 0), local, stream), "stft")

    def cufft():
        s = torch.stft(x, n_fft, hop_length=hop, window=wt, win_length=n_fft, center=True, pad_mode="reflect",
                       normalized=False, onesided=True, return_complex=True)
        s /= np.sqrt(n_fft)
        return s
    tc_med, _ = timed(cufft, 5)
    y = torch.empty((C, n), dtype=torch.float32, device=dev)

    def inverse():
        _lib.check(L.par_istft_f32(S.data_ptr(), n_fft, T, F, C, T * F, hop, win.ctypes.data, n_fft // 2, n, y.data_ptr(), 1,
                                   n, _lib.PAR_DEVICE_PTRS, local, stream), "istft")
    stages = [("stft complex", lambda: ours(False), C * (n * 4 + T * F * 8)),
              ("stft magnitude", lambda: ours(True), C * (n * 4 + T * F * 4))]
    if n_fft <= 32768:
        stages.append(("istft", inverse, C * (n * 4 + T * F * 8)))
    for name, fn, nbytes in stages:
        med, best = timed(fn)
        gbs = nbytes / med / 1e6
        print(f"| {n_fft} | {hop} | {name} | {med:.3f} | {best:.3f} | {nbytes / 1e9:.3f} | {gbs:.0f} | {gbs / peak:.3f} | "
              f"{tc_med if name.startswith('stft') else float('nan'):.3f} |")
    del S, M, y
curve = bench.wow_curve(dur, sr)
st, sp = np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1])
cap = int(n * 1.02) + 4096
pos = torch.empty(cap, dtype=torch.float64, device=dev)
out = torch.empty((C, cap), dtype=torch.float32, device=dev)
mbox = np.zeros(1, np.int64)


def positions():
    _lib.check(L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos.data_ptr(), cap,
                                      mbox.ctypes.data, _lib.PAR_DEVICE_PTRS, local, stream), "pos")
med, best = timed(positions)
m = int(mbox[0])
print(f"| - | 1024 | positions (curve -> {m} x f64) | {med:.3f} | {best:.3f} | {m * 8 / 1e9:.3f} | {m * 8 / med / 1e6:.0f} | "
      f"{m * 8 / med / 1e6 / peak:.3f} | - |")
for nt in (128, 50):
    for ch in (2, 1):
        def sinc():
            _lib.check(L.par_sinc_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, ch, n, nt, out.data_ptr(), 1, cap,
                                               _lib.PAR_DEVICE_PTRS, local, stream), "sinc")
        med, best = timed(sinc, 5)
        nbytes = ch * (n * 4 + m * 4) + m * 8
        print(f"| - | - | sinc NT={nt} ({2 * nt} taps), {ch} ch | {med:.3f} | {best:.3f} | {nbytes / 1e9:.3f} | "
              f"{nbytes / med / 1e6:.0f} | {nbytes / med / 1e6 / peak:.3f} | - |")
        print(f"|  |  | -> {ch * m * 2 * nt / med / 1e9:.2f} T taps/s |  |  |  |  |  |  |")


def linear():
    _lib.check(L.par_linear_resample_f32(pos.data_ptr(), m, x.data_ptr(), n, 1, C, n, out.data_ptr(), 1, cap,
                                         _lib.PAR_DEVICE_PTRS, local, stream), "linear")
med, best = timed(linear)
nbytes = C * (n * 4 + m * 4) + m * 8
print(f"| - | - | linear resample, {C} ch | {med:.3f} | {best:.3f} | {nbytes / 1e9:.3f} | {nbytes / med / 1e6:.0f} | "
      f"{nbytes / med / 1e6 / peak:.3f} | - |")

if world > 1:
    dist.destroy_process_group()
