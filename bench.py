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
  parity     a sampled check of THIS run's device results against the CPU oracle: read positions bit
             for bit, ~2*10^4 resampled outputs and a set of STFT frames to 1e-6.
  competitor_torch_stft   the reference's own GPU back-end (util/fourier.py:92-121: torch.stft + /sqrt(N)
             [+ .cpu()]) on the same input, device-resident and end to end.
  strong_cfg3   BASELINE configs[2] (60 min, 8 ch, 192 kHz) as ONE job cut into N time chunks (strong
             scaling; N = 1: the whole job on one GPU), device-generated input.
  cpu_baseline  the CPU oracle (port of the reference's numpy/numba path) on a bounded sample.

Under torchrun (N > 1) every rank runs the same per-GPU cfg2 workload on its own channels (weak scaling):
rank 0 broadcasts the speed curve ONCE before the steps and the output lengths are gathered once after
them -- the steps themselves need no communication.  `--impl reference` runs the UNMODIFIED reference
(util/fourier.np_rfft_pick, util/resampling.speed_to_pos + sinc_wrapper_mt, copied to git-ignored
baseline/_ref by scripts/install_reference.py) on the host cores; if that copy is missing it falls back
to the oracle port and says so.
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
TOL = 1e-6


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


def device_synth(torch, out, sr, seed, start=0):
    """The same kind of signal generated on the device (cfg3 is 22 GB: host synthesis would take minutes):
    `out` is a float32 (n,) device view holding samples start .. start + n of a channel."""
    n = out.numel()
    g = torch.Generator(device=out.device)
    g.manual_seed(int(seed))
    blk = 1 << 24
    for s in range(0, n, blk):
        e = min(n, s + blk)
        t = (torch.arange(s, e, dtype=torch.float64, device=out.device) + float(start)) / float(sr)
        v = 0.25 * torch.sin(2 * np.pi * 1000.0 * t) + 0.1 * torch.sin(2 * np.pi * (sr / 4.3) * t)
        out[s:e] = v.to(torch.float32) + 0.05 * torch.randn(e - s, generator=g, device=out.device, dtype=torch.float32)
    return out


def wow_curve(duration, sr, hop=HOP, depth=0.01, freq=0.5556):
    """The speed curve as the GUI builds it (util/markers.py:585-599): K = int(duration*sr/hop)
    points on linspace(0, duration, K), +-1 % sinusoidal wow."""
    k = int(duration * sr / hop)
    times = np.linspace(0, duration, k)
    return np.stack((times, 1 + depth * np.sin(2 * np.pi * freq * times)), -1)


def get_window():
    import scipy.signal
    return np.ascontiguousarray(scipy.signal.get_window("blackmanharris", N_FFT), dtype=np.float32)


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


def workload_config(workload, world=1, channels_per_gpu=None, m=None):
    """The `config` object of a line -- identical for the GPU arm and the reference arm of one workload."""
    sr, dur, ch, desc = WORKLOADS[workload]
    C = ch if channels_per_gpu is None else channels_per_gpu
    n = int(sr * dur)
    cfg = {"workload": f"{workload}: {desc}", "sample_rate": sr, "seconds": dur,
           "channels_per_gpu": C, "samples_per_channel": n,
           "n_fft": N_FFT, "hop": HOP, "zeropad": 1, "window": "blackmanharris", "sinc_quality": NT,
           "speed_curve": "1 + 0.01 sin(2 pi 0.5556 t), one point per hop"}
    if m is None:                       # the reference arm: length of the reference's speed_to_pos on this curve (CPU)
        import oracle
        curve = wow_curve(dur, sr)
        m = len(oracle.speed_to_pos_c(np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1]), n))
    T, F = n // HOP + 1, N_FFT // 2 + 1
    cfg.update({"output_samples_per_channel": int(m),
                "l2": "inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2; no flush" % (
                    C * n * 4 / 1e6, (C * T * F * 8 + C * m * 4 + m * 8) / 1e6),
                "parallelism": (f"channels x{world}: every rank its own {C} channels; curve broadcast once before, output "
                                f"lengths gathered once after the steps") if world > 1 else "1 GPU"})
    return cfg


# ------------------------------------------------------------------------------------ CPU arms
def cpu_pass_port(seconds, sr, seed, cores):
    """One pass of the hot path on the host with the ORACLE PORT: numpy per-frame rfft STFT (single thread,
    util/fourier.py:136-157), speed_to_pos (:93-137) and the float64 sinc with the reference's thread fan-out
    (util/resampling.py:30-46) on `seconds` of one channel."""
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


_REF = None


def reference_modules():
    """The UNMODIFIED reference modules from baseline/_ref (scripts/install_reference.py), or None."""
    global _REF
    if _REF is None:
        _REF = False
        base = os.path.join(ROOT, "baseline", "_ref")
        if os.path.exists(os.path.join(base, "util", "resampling.py")):
            import importlib
            import logging
            import warnings
            sys.path.insert(0, base)              # stays: numba resolves the jitted functions' globals through sys.modules
            try:
                logging.disable(logging.CRITICAL)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    f = importlib.import_module("util.fourier")
                    r = importlib.import_module("util.resampling")
                if os.path.realpath(f.__file__).startswith(os.path.realpath(base)):
                    _REF = (f, r)
            except Exception as e:                                   # noqa: BLE001
                print(f"reference import failed: {e!r}", file=sys.stderr)
            finally:
                logging.disable(logging.NOTSET)
    return _REF or None


def cpu_pass_reference(seconds, sr, seed, cores):
    """The same pass with the reference's own functions: util.fourier.np_rfft_pick (its CPU back-end on this
    image: pyfftw is not installed), util.resampling.speed_to_pos and sinc_wrapper_mt (numba, all cores)."""
    import warnings
    f, r = reference_modules()
    n = int(seconds * sr)
    x = synth_channel(n, sr, seed)
    curve = wow_curve(seconds, sr)
    win = get_window()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        s = f.np_rfft_pick(N_FFT, HOP, win, x, 1)
        t1 = time.perf_counter()
        pos = r.speed_to_pos(curve[:, 0] * sr, curve[:, 1], n)
        t2 = time.perf_counter()
        out = np.empty((len(pos), 1), np.float32)
        r.sinc_wrapper_mt(out[:, 0], pos, x, 0, NT)
        t3 = time.perf_counter()
    del s, out
    return n, (t1 - t0, t2 - t1, t3 - t2)


def cpu_baseline(sr, cores, budget_s=15.0):
    n, (a, b, c) = cpu_pass_port(1.0, sr, 1234, cores)         # calibration (also warms caches / threads)
    rate = n / (a + b + c)
    seconds = float(np.clip(budget_s * rate / sr, 2.0, 120.0))
    n, (a, b, c) = cpu_pass_port(seconds, sr, 1234, cores)
    return {"value": n / (a + b + c), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {seconds:.1f} s of channel 0 at {sr} Hz ({n} samples): numpy per-frame rfft STFT "
                      f"{a:.2f} s (1 thread) + speed_to_pos {b:.2f} s + float64 sinc NT={NT} {c:.2f} s ({cores} threads)",
            "stft_samples_per_s": n / a, "sinc_samples_per_s": n / c}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sr, dur, ch, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    real = reference_modules() is not None
    cpu_pass = cpu_pass_reference if real else cpu_pass_port
    kind = "reference" if real else "port"
    cpu_pass(1.0, sr, 1234, cores)                         # numba JIT / thread start-up outside every measurement
    n, (a, b, c) = cpu_pass(2.0, sr, 1234, cores)
    rate = n / (a + b + c)
    seconds = float(np.clip(8.0 * rate / sr, 2.0, 60.0))   # ~8 s of CPU work per step
    for _ in range(args.warmup):
        cpu_pass(min(seconds, 2.0), sr, 1234, cores)
    total, dt, parts = 0, 0.0, np.zeros(3)
    for _ in range(args.steps):
        n, t = cpu_pass(seconds, sr, 1234, cores)          # synthesis of the sample is outside the timed parts
        total += n
        dt += sum(t)
        parts += np.array(t)
    value = total / dt
    what = ("unmodified reference (baseline/_ref): util.fourier.np_rfft_pick + util.resampling.speed_to_pos + "
            "sinc_wrapper_mt (numba)") if real else "oracle port (baseline/_ref missing: run scripts/install_reference.py)"
    sample = (f"{seconds:.1f} s of one channel per step ({int(seconds * sr)} samples): STFT {parts[0] / args.steps:.2f} s (1 thread) "
              f"+ speed_to_pos {parts[1] / args.steps:.2f} s + sinc NT={NT} {parts[2] / args.steps:.2f} s ({cores} threads); {what}; "
              "throughput is linear in samples, so it stands for the full workload")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 STFT / f64 sinc",
            "data": "synthetic", "config": workload_config(args.workload, max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ sampled parity
def sampled_parity(torch, x_dev, pos_dev, m, out_dev, S_dev, curve, sr, n, n_points=20000, n_frames=24, seed=0):
    """Checks device results of one step against the CPU oracle WITHOUT computing the oracle for the whole job:
    * read positions at `n_points` random output indices (+ their successors, + both ends) and the output count,
      bit for bit against the serial float64 recurrence (oracle.speed_to_pos_at_c);
    * the resampled value of every sampled interior output of every channel against the float64 sinc of the
      oracle evaluated on the same 2*NT input samples (oracle.sinc_windows_c);
    * `n_frames` STFT frames per channel against the float64 oracle transform of the same samples.
    x_dev (C, n) float32, pos_dev (>= m) float64, out_dev (C, >= m) float32, S_dev (C, T, F) complex64 or None."""
    import oracle
    from oracle import oracle_np as onp
    rng = np.random.default_rng(seed)
    C = x_dev.shape[0]
    st, sp = np.ascontiguousarray(curve[:, 0] * sr), np.ascontiguousarray(curve[:, 1])
    pick = np.unique(np.concatenate([rng.integers(0, max(m - 1, 1), n_points), [0, 1, max(m - 2, 0)]]))
    pick = pick[pick + 1 < m]
    idx = np.unique(np.concatenate([pick, pick + 1, [m - 1]]))
    want, m_ref = oracle.speed_to_pos_at_c(st, sp, n, idx)
    dev = x_dev.device
    got = pos_dev[torch.as_tensor(idx, device=dev)].cpu().numpy()
    res = {"points": int(len(pick)), "tolerance": TOL, "output_count_equal": bool(m_ref == m),
           "positions_bit_exact": bool(m_ref == m and np.array_equal(got, want))}
    # sinc: interior outputs (all 2*NT taps inside the signal)
    p = got[np.searchsorted(idx, pick)]
    pn = got[np.searchsorted(idx, pick + 1)]
    ind = np.rint(p).astype(np.int64)
    base = ind - NT - 2
    W = 2 * NT + 4
    keep = (base >= 0) & (base + W <= n)
    pick, p, pn, base = pick[keep], p[keep], pn[keep], base[keep]
    gidx = torch.as_tensor(base[:, None] + np.arange(W)[None, :], device=dev)
    pick_t = torch.as_tensor(pick, device=dev)
    rel_l2, rel_max = 0.0, 0.0
    for c in range(C):
        wins = x_dev[c][gidx].cpu().numpy()
        y = out_dev[c][pick_t].cpu().numpy().astype(np.float64)
        ref = oracle.sinc_windows_c(p - base, pn - base, wins, NT).astype(np.float64)
        rel_l2 = max(rel_l2, float(np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-300)))
        rel_max = max(rel_max, float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-300)))
    res.update(sinc_points=int(len(pick)) * C, sinc_rel_l2=rel_l2, sinc_rel_max=rel_max)
    ok = res["positions_bit_exact"] and rel_l2 <= TOL and rel_max <= TOL
    if S_dev is not None:
        T = S_dev.shape[1]
        j = N_FFT // (2 * HOP)                              # frame j of a lone n_fft-sample segment is centred on it
        lo_t, hi_t = -(-(N_FFT // 2) // HOP), (n - N_FFT // 2) // HOP
        frames = np.unique(rng.integers(lo_t, hi_t + 1, n_frames))
        seg_idx = torch.as_tensor((frames * HOP - N_FFT // 2)[:, None] + np.arange(N_FFT)[None, :], device=dev)
        worst = 0.0
        for c in range(C):
            segs = x_dev[c][seg_idx].cpu().numpy()
            got_f = S_dev[c][torch.as_tensor(frames, device=dev)].cpu().numpy().astype(np.complex128)
            for k in range(len(frames)):
                ref_f = onp.stft_f64(segs[k], N_FFT, HOP)[j]
                worst = max(worst, float(np.linalg.norm(got_f[k] - ref_f) / max(np.linalg.norm(ref_f), 1e-300)))
        res.update(stft_frames=int(len(frames)) * C, stft_rel_l2=worst, stft_frame_count=int(T))
        ok = ok and worst <= TOL
    res["pass"] = bool(ok)
    return res


# ------------------------------------------------------------------------------------ cfg3, strong scaling
def run_cfg3_strong(torch, dist, rank, world, dev, steps, workload="cfg3"):
    """ONE job (all channels of the workload) cut into `world` time chunks: the speed curve is
    broadcast once, then per step one all-gather of the chunks' edge samples, one all-gather of the per-segment
    totals of the curve, and every rank transforms the frames and resamples the outputs that fall into its chunk.
    Input is generated on the device.  Returns the record on rank 0 (None elsewhere)."""
    from pyaudiorestoration_b200 import _lib, dist as pdist
    L = _lib.lib()
    sr, dur, C, desc = WORKLOADS[workload]
    n = int(sr * dur)
    sh = pdist.TimeShard(n, N_FFT, HOP, NT, rank, world)
    buf = sh.local_buffer(C, dev)
    chunk = sh.chunk_view(buf)
    for c in range(C):
        device_synth(torch, chunk[c], sr, 1234 + c, start=sh.s0)
    curve = wow_curve(dur, sr) if rank == 0 else None
    cv = pdist.broadcast_curve(curve, src=0, device=dev)           # once per job: the curve is an input like the audio
    st, sp = np.ascontiguousarray(cv[:, 0] * sr), np.ascontiguousarray(cv[:, 1])
    window = get_window()
    nfr, F = sh.frame1 - sh.frame0, N_FFT // 2 + 1
    S_out = torch.empty((C, nfr, F), dtype=torch.complex64, device=dev)
    pos_buf = torch.empty(int((sh.s1 - sh.s0) * 1.1) + 8 * HOP, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)

    side = torch.cuda.Stream(dev)

    def step():
        # order matters: the transform is a persistent kernel that fills every SM, so the two small collectives
        # go first; the serial host chain of speed_to_pos and its copies then run on a side stream BEHIND the
        # transform (its expansion kernel starts as the transform's CTAs retire)
        sh.exchange_halos(buf)
        sums = sh.position_sums(st, sp, dev) if world > 1 else None
        side.wait_stream(stream)
        sh.stft(buf, window, out=S_out)
        ps, p0, m = sh.positions(st, sp, dev, out=pos_buf, sums=sums, stream=side)
        stream.wait_stream(side)
        y = sh.resample(buf, ps, "Sinc", pos_origin=p0, m=m)
        side.wait_stream(stream)                     # the next step's positions overwrite pos_buf
        return y, m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(3):
        y, m = step()
        del y
    barrier()
    launches0 = L.par_kernel_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    for _ in range(steps):
        y, m = step()
        del y
    t1.record(stream)
    barrier()
    ms = t0.elapsed_time(t1)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    launches = L.par_kernel_launch_count() - launches0
    rec = None
    if rank == 0:
        rec = {"workload": f"{workload}: {desc}", "scaling": "strong", "n_gpus": world, "steps": steps,
               "ms_per_step": ms / steps, "value": n * C * steps / (ms * 1e-3), "unit": UNIT,
               "channels": C, "samples_per_channel": n, "output_samples_per_channel": int(m),
               "gpu_launches": int(launches),
               "parallelism": (f"time chunks x{world}, halo {sh.H} samples: per step 1 all-gather of edge blocks + 1 all-gather of "
                               f"{len(st) - 1} segment totals; curve broadcast once per job") if world > 1 else "1 GPU, whole job",
               "note": "speed-up vs 1 GPU = this ms_per_step against strong_cfg3.ms_per_step of the N=1 line"}
    del buf, S_out, pos_buf
    torch.cuda.empty_cache()
    L.par_release_cached_memory(dev.index)
    return rec


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-competitor", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the cfg3 strong-scaling record")
    ap.add_argument("--strong-steps", type=int, default=5)
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
    # the speed curve comes from rank 0: ONE broadcast per job, before the steps
    curve = wow_curve(dur, sr) if rank == 0 else None
    if world > 1:
        from pyaudiorestoration_b200 import dist as pdist
        curve = pdist.broadcast_curve(curve, src=0, device=dev)
    st = np.ascontiguousarray(curve[:, 0] * sr)
    sp = np.ascontiguousarray(curve[:, 1])
    window = get_window()
    x_dev = torch.from_numpy(host_sig).to(dev).t().contiguous()
    S_dev = torch.empty((C, T, F), dtype=torch.complex64, device=dev)
    cap = int(n * 1.02) + 4096
    pos_dev = torch.empty(cap, dtype=torch.float64, device=dev)
    out_dev = torch.empty((C, cap), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    sh = stream.cuda_stream
    m_box = np.zeros(1, np.int64)
    ev = {k: [] for k in ("stft", "pos", "sinc")}
    side = torch.cuda.Stream(dev)                      # the positions run beside the transform (they share no data)
    sinc_done = torch.cuda.Event()

    def step(timed, overlap=True):
        """One device-resident pass.  Returns the number of output samples.
        overlap: the STFT is enqueued on the main stream and speed_to_pos runs on a side stream meanwhile -- its
        kernels and its serial host chain (two synchronisations of ITS stream) hide behind the transform; the
        resampler then waits for both.  Without overlap the three calls run back to back on one stream: that is how
        the per-stage times (stage_ms, the rooflines) are taken."""
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
        if timed:
            e[0].record(stream)
        _lib.check(L.par_stft_f32(x_dev.data_ptr(), n, 1, C, n, N_FFT, HOP, 1, window.ctypes.data,
                                  S_dev.data_ptr(), F, T * F, _lib.PAR_DEVICE_PTRS, local, sh), "par_stft_f32")
        if timed:
            e[1].record(stream)
        pos_stream = side if overlap else stream
        if overlap:
            side.wait_event(sinc_done)                # the previous step's resampler still reads pos_dev
        _lib.check(L.par_speed_to_pos_f64(st.ctypes.data, sp.ctypes.data, len(st), float(n), pos_dev.data_ptr(), cap,
                                          m_box.ctypes.data, _lib.PAR_DEVICE_PTRS, local, pos_stream.cuda_stream),
                   "par_speed_to_pos_f64")
        m = int(m_box[0])
        if overlap:
            stream.wait_stream(side)
        if timed:
            e[2].record(stream)
        _lib.check(L.par_sinc_resample_f32(pos_dev.data_ptr(), m, x_dev.data_ptr(), n, 1, C, n, NT,
                                           out_dev.data_ptr(), 1, cap, _lib.PAR_DEVICE_PTRS, local, sh),
                   "par_sinc_resample_f32")
        sinc_done.record(stream)
        if timed:
            e[3].record(stream)
            ev["stft"].append((e[0], e[1]))
            ev["pos"].append((e[1], e[2]))
            ev["sinc"].append((e[2], e[3]))
        return m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sinc_done.record(stream)
    for _ in range(args.warmup):
        m = step(False)
    barrier()
    for _ in range(3):                                 # per-stage times: serial steps, outside the timed region
        m = step(True, overlap=False)
    barrier()
    launches0 = L.par_kernel_launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for _ in range(args.steps):
        m = step(False)
    t_end.record(stream)
    barrier()
    ms = t_start.elapsed_time(t_end)
    clk = clocks.stop()
    launches = L.par_kernel_launch_count() - launches0
    lens = None
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
        lens = torch.zeros(world, dtype=torch.int64, device=dev)      # one gather of the output lengths per job
        dist.all_gather_into_tensor(lens, torch.tensor([m], dtype=torch.int64, device=dev))
        lens = [int(v) for v in lens.tolist()]
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

    # ---- parity of THIS run's device results (sampled), rank 0
    parity = None
    if rank == 0 and not args.no_parity:
        parity = sampled_parity(torch, x_dev, pos_dev, m, out_dev, S_dev, curve, sr, n)

    # ---- the reference's own GPU back-end on the same input (util/fourier.py:92-121), rank 0
    competitor = None
    if rank == 0 and not args.no_competitor:
        win_t = torch.from_numpy(window).to(dev)

        def torch_stft_dev():
            out = []
            for c in range(C):
                s_ = torch.stft(x_dev[c], N_FFT, hop_length=HOP, window=win_t, win_length=N_FFT, center=True,
                                pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
                s_ /= np.sqrt(N_FFT)
                out.append(s_)
            return out

        def timeit(fn, reps):
            fn()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                r_ = fn()
                del r_
            b.record(stream)
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / reps
        dev_ms = timeit(torch_stft_dev, 5)

        def torch_stft_e2e():                       # torch_rfft2 as written: host array in, .cpu() tensor out, per channel
            out = []
            for c in range(C):
                xs_ = torch.from_numpy(np.ascontiguousarray(host_sig[:, c])).to(dev)
                s_ = torch.stft(xs_, N_FFT, hop_length=HOP, window=win_t, win_length=N_FFT, center=True, pad_mode="reflect",
                                normalized=False, onesided=True, return_complex=True)
                s_ /= np.sqrt(N_FFT)
                out.append(s_.cpu())
            return out
        torch_stft_e2e()
        t0 = time.perf_counter()
        r_ = torch_stft_e2e()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        del r_

        def ours_e2e():
            return [fourier.stft(host_sig[:, c], N_FFT, HOP) for c in range(C)]
        ours_e2e()
        t0 = time.perf_counter()
        r_ = ours_e2e()
        ours_ms = (time.perf_counter() - t0) * 1e3
        del r_
        competitor = {"what": "torch.stft (cuFFT) with the reference's arguments + /sqrt(n_fft), util/fourier.py:92-121",
                      "device_ms": dev_ms, "ours_device_ms": k_stft, "device_speedup": dev_ms / k_stft,
                      "e2e_ms": e2e_ms, "ours_e2e_ms": ours_ms, "e2e_speedup": e2e_ms / ours_ms,
                      "e2e_note": "host float32 column in, host complex64 (F, T) out, per channel, pageable .cpu() result vs "
                                  "this repo's util.fourier.stft"}
        torch.cuda.empty_cache()

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

    # ---- BASELINE configs[2] as one strong-scaling job on the same ranks
    strong = None
    if not args.no_strong and args.workload == "cfg2":
        del x_dev, S_dev, pos_dev, out_dev
        torch.cuda.empty_cache()
        L.par_release_cached_memory(local)
        try:
            strong = run_cfg3_strong(torch, dist, rank, world, dev, args.strong_steps)
        except Exception as e:                                                   # noqa: BLE001
            if world > 1:
                raise
            strong = {"error": repr(e)}

    if rank == 0:
        cfg = workload_config(args.workload, world, C, m)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32 (positions f64)", "data": "synthetic",
            "config": cfg,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof(bytes_sinc, k_sinc, "sinc_kernel",
                             {"note": "sinc_kernel_ws<2,64> (warp-specialised: 12 interpolating + 4 set-up warps per SM); the stage is "
                                      "FP32-pipe bound (4.5 / 7.5 FMA-pipe operations per tap at fc = 1 / fc < 1 for 2 channels), "
                                      "not HBM bound; see DESIGN.md",
                              "taps_per_s": C * m * 2 * NT / (k_sinc * 1e-3)}),
            "roofline_stft": roof(bytes_stft, k_stft, "stft_kernel",
                                  {"note": "stft_tma_kernel<11,0>: TMA-staged frames, HBM-bound by design"}),
            "roofline_positions": roof(bytes_pos, k_pos, "expand_positions_kernel",
                                       {"note": "stage time includes the serial host chain of speed_to_pos (2 stream "
                                                "synchronisations); kernels: segment_sums (totals) + expand_positions "
                                                "(every position written once, offset included)"}),
            "stage_ms": {"stft": k_stft, "positions": k_pos, "sinc": k_sinc,
                         "note": "measured in 3 serial steps before the timed region; in the timed steps speed_to_pos runs on a "
                                 "side stream beside the STFT"},
            "parity": parity,
            "competitor_torch_stft": competitor,
            "strong_cfg3": strong,
            "output_lengths": lens,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
