#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric for the STFT + varispeed-resample hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|small]

A "step" is one pass of the hot path over one batch of synthetic audio: STFT (n_fft 4096, hop 1024,
blackmanharris) of every channel, speed curve -> read positions, windowed-sinc resample of every
channel (NT = 128, i.e. the 256-tap sinc of the roofline sweep).  The metric is per-channel input
samples processed per second over all GPUs.

  value      device-resident: inputs already in HBM, outputs left in HBM, timed with CUDA events on
             the launch stream, barrier + synchronize on both sides, max over ranks.
  e2e        the same step through the reference-facing Python API: util.fourier.stft(sig[:, c]) per
             channel (as the GUIs call it) + util.resampling.varispeed (= run() minus the WAV write) on
             an interleaved pinned (frames, channels) float32 array, numpy arrays out: every step copies
             its inputs host->device and its results device->host.
  roofline   the dominant kernel (the sinc interpolator), algorithmic bytes / CUDA-event time,
             against MEASURED_PEAKS.json; roofline_stft / roofline_positions give the other two.
  cpu_baseline  the CPU oracle (port of the reference's numpy/numba path) on a bounded sample.

Under torchrun (N > 1) every rank runs the same per-GPU workload on its own channels (weak
scaling; --workload cfg3 shards the 8 channels of the 60-min config over the ranks instead); the
only exchange is one NCCL broadcast of the speed curve and one gather of the output lengths.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_FFT, HOP, NT = 4096, 1024, 128
WORKLOADS = {
    # name: (sample rate, seconds, channels, description)
    "cfg2": (96000, 600.0, 2, "synthetic 10-min stereo 96 kHz (BASELINE configs[1])"),
    "cfg3": (192000, 3600.0, 8, "synthetic 60-min 8-ch 192 kHz (BASELINE configs[2]), channels sharded over ranks"),
    "small": (96000, 20.0, 2, "synthetic 20-s stereo 96 kHz (plumbing check, not a bench line)"),
}
METRIC = "audio samples/sec (STFT+varispeed resample)"
UNIT = "samples/s"


# ------------------------------------------------------------------------------------ synthetic data
def synth_channel(n, sr, seed, out=None):
    """0.25 sin(2 pi 1000 t) + 0.1 sin(2 pi sr/4.3 t) + 0.05 N(0,1), float32."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty(n, np.float32)
    blk = 1 << 22
    for s in range(0, n, blk):
        e = min(n, s + blk)
        t = np.arange(s, e, dtype=np.float64) / sr
        out[s:e] = (0.25 * np.sin(2 * np.pi * 1000.0 * t) + 0.1 * np.sin(2 * np.pi * (sr / 4.3) * t)
                    + 0.05 * rng.standard_normal(e - s)).astype(np.float32)
    return out


def wow_curve(duration, sr, hop=HOP, depth=0.01, freq=0.5556):
    """The speed curve as the GUI builds it (util/markers.py:585-599): K = int(duration*sr/hop)
    points on linspace(0, duration, K), +-1 % sinusoidal wow."""
    k = int(duration * sr / hop)
    times = np.linspace(0, duration, k)
    return np.stack((times, 1 + depth * np.sin(2 * np.pi * freq * times)), -1)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 20 ms on a thread (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz, self._stop, self._thr = None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture summary, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ------------------------------------------------------------------------------------ CPU arm
def cpu_pass(seconds, sr, seed, cores):
    """One pass of the hot path on the host: the oracle's port of the reference's numpy STFT
    (single-threaded per-frame rfft, util/fourier.py:136-157), speed_to_pos (:93-137) and
    sinc_wrapper_mt (all cores, util/resampling.py:30-46) on `seconds` of one channel."""
    import oracle
    from oracle import oracle_np as onp
    n = int(seconds * sr)
    x = synth_channel(n, sr, seed)
    curve = wow_curve(seconds, sr)
    t0 = time.perf_counter()
    s = onp.stft_ref(x, N_FFT, HOP)
    t1 = time.perf_counter()
    pos = oracle.speed_to_pos_c(curve[:, 0] * sr, curve[:, 1], n)
    t2 = time.perf_counter()
    y = oracle.sinc_c(pos, x, NT, nthreads=cores)
    t3 = time.perf_counter()
    del s, y
    return n, (t1 - t0, t2 - t1, t3 - t2)


def cpu_baseline(sr, cores, budget_s=15.0):
    n, (a, b, c) = cpu_pass(1.0, sr, 1234, cores)         # calibration (also warms caches / threads)
    rate = n / (a + b + c)
    seconds = float(np.clip(budget_s * rate / sr, 2.0, 120.0))
    n, (a, b, c) = cpu_pass(seconds, sr, 1234, cores)
    return {"value": n / (a + b + c), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {seconds:.1f} s of channel 0 at {sr} Hz ({n} samples): numpy per-frame rfft STFT "
                      f"{a:.2f} s (1 thread) + speed_to_pos {b:.2f} s + float64 sinc NT={NT} {c:.2f} s ({cores} threads)",
            "stft_samples_per_s": n / a, "sinc_samples_per_s": n / c}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port; the reference itself is
    Python + numba and its sources may not travel to the GPU box) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sr, dur, ch, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n, (a, b, c) = cpu_pass(1.0, sr, 1234, cores)
    rate = n / (a + b + c)
    seconds = float(np.clip(8.0 * rate / sr, 2.0, 60.0))   # ~8 s of CPU work per step
    for _ in range(args.warmup):
        cpu_pass(min(seconds, 2.0), sr, 1234, cores)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        n, _ = cpu_pass(seconds, sr, 1234, cores)
        total += n
    dt = time.perf_counter() - t0
    # synthesis of the sample is outside the reference's path; re-time it and subtract
    t1 = time.perf_counter()
    for _ in range(args.steps):
        synth_channel(int(seconds * sr), sr, 1234)
        wow_curve(seconds, sr)
    dt -= time.perf_counter() - t1
    value = total / dt
    sample = f"{seconds:.1f} s of one channel per step ({int(seconds * sr)} samples), oracle port, {cores} threads for the sinc stage"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 STFT / f64 sinc",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "n_fft": N_FFT, "hop": HOP, "sinc_quality": NT},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm, time shards
def run_time_sharded(args, torch, dist, rank, world, local, dev):
    """ONE job (all channels of the workload) cut into `world` time chunks: per step
    one broadcast of the speed curve, one all-gather of the chunks' edge samples, then every rank
    transforms the frames and resamples the outputs that fall into its chunk.  Strong scaling."""
    from pyaudiorestoration_b200 import _lib, dist as pdist
    L = _lib.lib()
    sr, dur, C, desc = WORKLOADS[args.workload]
    n = int(sr * dur)
    sh = pdist.TimeShard(n, N_FFT, HOP, NT, rank, world)
    ln = sh.s1 - sh.s0
    host_chunk = _lib.pinned_empty((C, ln), np.float32)
    for c in range(C):
        synth_channel(ln, sr, 1234 + c + 100 * rank, out=host_chunk[c])
    buf = sh.local_buffer(C, dev)
    sh.chunk_view(buf).copy_(torch.from_numpy(host_chunk))
    curve = wow_curve(dur, sr) if rank == 0 else None
    window = np.ascontiguousarray(__import__("scipy.signal").signal.get_window("blackmanharris", N_FFT), dtype=np.float32)
    nfr, F = sh.frame1 - sh.frame0, N_FFT // 2 + 1
    S_out = torch.empty((C, nfr, F), dtype=torch.complex64, device=dev)
    pos_buf = torch.empty(int(ln * 1.1) + 8 * HOP, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(upload=False):
        if upload:                                   # e2e: this step's audio comes from (pinned) host memory
            sh.chunk_view(buf).copy_(torch.from_numpy(host_chunk), non_blocking=True)
        cv = pdist.broadcast_curve(curve, src=0, device=dev)
        sh.exchange_halos(buf)
        sh.stft(buf, window, out=S_out)
        ps, p0, m = sh.positions(cv[:, 0] * sr, cv[:, 1], dev, out=pos_buf)
        y = sh.resample(buf, ps, "Sinc", pos_origin=p0, m=m)
        return y, m

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = L.par_kernel_launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    for _ in range(args.steps):
        y, m = step()
    t1.record(stream)
    barrier()
    tm = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    clk = clocks.stop()
    launches = L.par_kernel_launch_count() - launches0
    # e2e: host chunk in, spectrogram + resampled audio back in pinned host memory
    S_host = torch.empty(S_out.shape, dtype=S_out.dtype).pin_memory()
    y_host = torch.empty((C, int(ln * 1.1) + 8 * HOP), dtype=torch.float32).pin_memory()
    e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        y, m = step(upload=True)
        S_host.copy_(S_out, non_blocking=True)
        y_host[:, :y.shape[1]].copy_(y, non_blocking=True)
        torch.cuda.synchronize(dev)
        return y.shape[1]
    e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e_steps):
        mine = e2e_step()
    dt = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {
            "metric": METRIC, "value": n * C * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 (positions f64)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "sample_rate": sr, "seconds": dur, "channels": C,
                       "samples_per_channel": n, "n_fft": N_FFT, "hop": HOP, "sinc_quality": NT,
                       "parallelism": f"time chunks x{world}, halo {sh.H} samples, 1 all-gather of edge blocks + 1 curve "
                                      f"broadcast per step",
                       "l2": "per-rank inputs and outputs exceed the 126 MB L2; no flush"},
            "e2e": {"value": n * C * e_steps / float(dt.item()), "unit": UNIT,
                    "h2d_bytes_per_step": int(C * ln * 4), "d2h_bytes_per_step": int(C * nfr * F * 8 + C * mine * 4),
                    "steps": e_steps, "ms_per_step": float(dt.item()) / e_steps * 1e3,
                    "api": "per rank: pinned host chunk -> TimeShard.exchange_halos/stft/positions/resample -> pinned host"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": None, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="channels", choices=["channels", "time"],
                    help="N > 1: every rank its own channels (weak scaling, default) or one job cut into time chunks "
                         "with a boundary all-gather (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    os.environ["PAR_B200_DEVICE"] = str(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from pyaudiorestoration_b200 import _lib
    from pyaudiorestoration_b200.util import fourier, resampling
    L = _lib.lib()
    _lib.require_device()

    if args.shard == "time" and world > 1:
        run_time_sharded(args, torch, dist, rank, world, local, dev)
        dist.destroy_process_group()
        return

    sr, dur, ch_total, desc = WORKLOADS[args.workload]
    if args.workload == "cfg3":
        if ch_total % world:
            raise SystemExit("cfg3 needs a rank count dividing 8")
        my_ch = list(range(rank * (ch_total // world), (rank + 1) * (ch_total // world)))
        scaling = "strong"
    else:
        my_ch = [rank * ch_total + c for c in range(ch_total)]       # every rank: its own channels
        scaling = "weak"
    C = len(my_ch)
    n = int(sr * dur)
    T = int(L.par_stft_num_frames(n, N_FFT, HOP))
    F = N_FFT // 2 + 1

    # ---- inputs: host (pinned, interleaved (frames, channels) like the arrays soundfile hands the
    #      reference, util/io_ops.py:10) + device planar copy for the device-resident arm
    host_sig = _lib.pinned_empty((n, C), np.float32)
    tmp_ch = np.empty(n, np.float32)
    for i, c in enumerate(my_ch):
        synth_channel(n, sr, 1234 + c, out=tmp_ch)
        host_sig[:, i] = tmp_ch
    del tmp_ch
    curve = wow_curve(dur, sr)
    curve_t = torch.from_numpy(curve.copy()).to(dev)
    window = np.ascontiguousarray(__import__("scipy.signal").signal.get_window("blackmanharris", N_FFT), dtype=np.float32)
    x_dev = torch.from_numpy(host_sig).to(dev).t().contiguous()
    S_dev = torch.empty((C, T, F), dtype=torch.complex64, device=dev)
    cap = int(n * 1.02) + 4096
    pos_dev = torch.empty(cap, dtype=torch.float64, device=dev)
    out_dev = torch.empty((C, cap), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    sh = stream.cuda_stream
    m_box = np.zeros(1, np.int64)
    lens = torch.zeros(world, dtype=torch.int64, device=dev)
    ev = {k: [] for k in ("stft", "pos", "sinc")}

    def step(timed):
        """One device-resident pass.  Returns the number of output samples."""
        if world > 1:
            dist.broadcast(curve_t, 0)            # the speed curve comes from rank 0
        cv = curve_t.cpu().numpy() if world > 1 else curve
        st = np.ascontiguousarray(cv[:, 0] * sr)
        sp = np.ascontiguousarray(cv[:, 1])
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
        if timed:
            e[0].record(stream)
        _lib.check(L.par_stft_f32(x_dev.data_ptr(), n, 1, C, n, N_FFT, HOP, 1, window.ctypes.data,
                                  S_dev.data_ptr(), F, T * F, _lib.PAR_DEVICE_PTRS, local, sh), "par_stft_f32")
        if timed:
            e[1].record(stream)
        _lib.check(L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos_dev.data_ptr(), cap,
                                          m_box.ctypes.data, _lib.PAR_DEVICE_PTRS, local, sh), "par_speed_to_pos_f64")
        m = int(m_box[0])
        if timed:
            e[2].record(stream)
        _lib.check(L.par_sinc_resample_f32(pos_dev.data_ptr(), m, x_dev.data_ptr(), n, 1, C, n, NT,
                                           out_dev.data_ptr(), 1, cap, _lib.PAR_DEVICE_PTRS, local, sh),
                   "par_sinc_resample_f32")
        if timed:
            e[3].record(stream)
            ev["stft"].append((e[0], e[1]))
            ev["pos"].append((e[1], e[2]))
            ev["sinc"].append((e[2], e[3]))
        if world > 1:
            mine = torch.tensor([m], dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(lens, mine)
        return m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        m = step(False)
    barrier()
    launches0 = L.par_kernel_launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for _ in range(args.steps):
        m = step(True)
    t_end.record(stream)
    barrier()
    ms = t_start.elapsed_time(t_end)
    clk = clocks.stop()
    launches = L.par_kernel_launch_count() - launches0
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    samples_per_step = n * C * world if scaling == "weak" else n * ch_total
    value = samples_per_step * args.steps / (ms * 1e-3)

    def avg_ms(pairs):
        return float(np.mean([a.elapsed_time(b) for a, b in pairs]))
    k_stft, k_pos, k_sinc = avg_ms(ev["stft"]), avg_ms(ev["pos"]), avg_ms(ev["sinc"])
    peak, peak_src = measured_peaks()
    bytes_stft = C * (n * 4 + T * F * 8) + N_FFT * 4
    bytes_pos = m * 8 + len(curve) * 16
    bytes_sinc = C * (n * 4 + m * 4) + m * 8

    def roof(nbytes, ms_, kernel, extra=None):
        a = nbytes / (ms_ * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
             "traffic": ncu_traffic(kernel), "kernel": kernel, "ms_per_launch": ms_,
             "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src}
        if extra:
            r.update(extra)
        return r

    # ---- e2e through the reference-facing API (host buffers in, host arrays out)
    e2e = None
    if not args.no_e2e:
        sig = host_sig
        speed_curve = curve

        def e2e_step():
            res = [fourier.stft(sig[:, c], N_FFT, HOP) for c in range(C)]        # as the GUIs call it, per channel
            out = resampling.varispeed(sig, sr, speed_curve, range(C), "Sinc", NT)   # run() minus the WAV write
            return res, out
        for _ in range(2):
            r = e2e_step()
        del r
        barrier()
        t0 = time.perf_counter()
        e_steps = max(2, min(args.steps, 5))
        for _ in range(e_steps):
            res, out = e2e_step()
            m_e = out.shape[0]
            del res, out
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            td = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
        e2e = {"value": samples_per_step * e_steps / dt, "unit": UNIT,
               # stft(sig[:, c]) uploads the interleaved span of the column view (C*n floats) per call
               "h2d_bytes_per_step": int(C * (C * n * 4) + C * n * 4 + len(curve) * 40 + N_FFT * 4),
               "d2h_bytes_per_step": int(C * T * F * 8 + C * m_e * 4 + len(curve) * 8),
               "steps": e_steps, "ms_per_step": dt / e_steps * 1e3,
               "api": "util.fourier.stft(sig[:, c]) per channel + util.resampling.varispeed (= run() minus the WAV write), "
                      "interleaved pinned float32 (frames, channels) in, pinned numpy arrays out"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(sr, os.cpu_count() or 1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32 (positions f64)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "sample_rate": sr, "seconds": dur,
                       "channels_per_gpu": C, "samples_per_channel": n, "n_fft": N_FFT, "hop": HOP, "zeropad": 1,
                       "window": "blackmanharris", "sinc_quality": NT, "output_samples_per_channel": m,
                       "speed_curve": "1 + 0.01 sin(2 pi 0.5556 t), one point per hop",
                       "l2": "inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2; no flush" % (
                           C * n * 4 / 1e6, (C * T * F * 8 + C * m * 4 + m * 8) / 1e6),
                       "parallelism": f"channels x{world}" if world > 1 else "1 GPU"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof(bytes_sinc, k_sinc, "sinc_kernel",
                             {"note": "sinc stage is FP32-issue/MUFU bound (2*NT reciprocals per output sample), "
                                      "not HBM bound; see DESIGN.md",
                              "taps_per_s": C * m * 2 * NT / (k_sinc * 1e-3)}),
            "roofline_stft": roof(bytes_stft, k_stft, "stft_kernel",
                                  {"note": "stft_tma_kernel<11,0>: TMA-staged frames, HBM-bound by design"}),
            "roofline_positions": roof(bytes_pos, k_pos, "expand_positions_kernel",
                                       {"note": "stage time includes the serial host chain of speed_to_pos (2 stream "
                                                "synchronisations); kernels: expand_positions + add_offsets"}),
            "stage_ms": {"stft": k_stft, "positions": k_pos, "sinc": k_sinc},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
