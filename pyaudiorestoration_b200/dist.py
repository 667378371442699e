"""Sharding the hot path over the GPUs of one box.

The reference is single-process; its only parallelism is the thread fan-out of ``sinc_wrapper_mt``
over contiguous output ranges (util/resampling.py:30-46) and FFTW threads (util/fourier.py:131).
Channels are independent given the speed curve (util/resampling.py:225-231), and both the STFT and
the resampler are local in time, so a job shards two ways with no data-path collective beyond one
exchange of boundary samples:

* by channel (``shard_channels``): nothing to exchange but the speed curve;
* by time chunk (``TimeShard``): rank r owns samples ``[s0, s1)`` of every channel (chunk
  boundaries on multiples of the hop), the STFT frames centred in that range and the output
  samples whose read positions fall into it.  Frames and taps that straddle a boundary read a halo
  of ``H = max(n_fft/2, NT + 2)`` samples from each neighbour, moved by ONE all-gather of the
  ranks' edge blocks (``exchange_halos``; NCCL on GPUs, gloo in the CPU tests).

One process per GPU; ``torch.distributed`` is plumbing only -- the compute goes through the range
entry points of the C ABI (``par_stft_range_f32``, ``par_resample_range_f32``).
"""
import numpy as np

from . import _lib


def shard_channels(n_channels, rank, world):
    """Contiguous, balanced block of channel indices for ``rank``."""
    base, extra = divmod(n_channels, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def time_chunks(n, world, hop):
    """``[(s0, s1)] * world`` partitioning ``[0, n)`` with boundaries on multiples of ``hop``."""
    hops = -(-n // hop)
    per = -(-hops // world)
    out = []
    for r in range(world):
        s0 = min(r * per * hop, n)
        s1 = n if r == world - 1 else min((r + 1) * per * hop, n)
        out.append((s0, s1))
    return out


def halo_width(n_fft, nt):
    """Samples a rank needs from each neighbour: half a window for the STFT (util/fourier.py:80-81
    centres frames), NT taps (+ rounding slack) for the resampler; a multiple of 4 so that slices
    stay 16-byte aligned for the bulk-copy staged transform."""
    h = max(int(n_fft) // 2, int(nt) + 2)
    return (h + 3) & ~3


class TimeShard:
    """Index plan of one rank of a time-sharded job over a signal of ``n`` samples per channel."""

    def __init__(self, n, n_fft, hop, nt, rank, world):
        if hop % 4 or (n_fft // 2) % 4:
            raise ValueError("time sharding needs hop and n_fft/2 to be multiples of 4")
        self.n, self.n_fft, self.hop, self.nt, self.rank, self.world = int(n), int(n_fft), int(hop), int(nt), rank, world
        self.chunks = time_chunks(self.n, world, self.hop)
        self.s0, self.s1 = self.chunks[rank]
        self.H = halo_width(n_fft, nt)
        if world > 1 and min(s1 - s0 for s0, s1 in self.chunks) < self.H:
            raise ValueError("chunks shorter than the halo: use fewer ranks for this signal")
        self.halo_left = self.H if rank > 0 else 0
        self.halo_right = self.H if rank < world - 1 else 0
        self.origin = self.s0 - self.halo_left                 # global index of local sample 0
        self.local_len = self.halo_left + (self.s1 - self.s0) + self.halo_right
        # STFT frames centred in [s0, s1); the last rank also owns the frames centred at >= n
        total_frames = self.n // self.hop + 1
        self.frame0 = -(-self.s0 // self.hop)
        self.frame1 = total_frames if rank == world - 1 else -(-self.s1 // self.hop)
        self.total_frames = total_frames

    # -------------------------------------------------------------------------------- buffers
    def local_buffer(self, channels, device):
        import torch
        return torch.zeros((channels, self.local_len), dtype=torch.float32, device=device)

    def chunk_view(self, buf):
        return buf[:, self.halo_left:self.halo_left + (self.s1 - self.s0)]

    def exchange_halos(self, buf, group=None):
        """Fill the halo columns of ``buf`` from the neighbours' chunks: one all-gather of every
        rank's first and last ``H`` samples."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return buf
        chunk = self.chunk_view(buf)
        edges = torch.stack((chunk[:, :self.H], chunk[:, -self.H:])).contiguous()       # (2, C, H)
        flat = torch.empty((self.world * 2,) + tuple(edges.shape[1:]), dtype=edges.dtype, device=edges.device)
        dist.all_gather_into_tensor(flat, edges, group=group)
        gathered = flat.view((self.world, 2) + tuple(edges.shape[1:]))
        if self.rank > 0:
            buf[:, :self.H] = gathered[self.rank - 1, 1]
        if self.rank < self.world - 1:
            buf[:, -self.H:] = gathered[self.rank + 1, 0]
        return buf

    # -------------------------------------------------------------------------------- compute
    def stft(self, buf, window, zeropad=1, magnitude=False, out=None):
        """Frames ``[frame0, frame1)`` of every channel of the global transform:
        ``(channels, frame1 - frame0, n_freqs)`` complex64 (or float32 magnitudes) on ``buf``'s device."""
        import torch
        L = _lib.lib()
        ch = buf.shape[0]
        nfr = self.frame1 - self.frame0
        F = self.n_fft * zeropad // 2 + 1
        if out is None:
            out = torch.empty((ch, nfr, F), dtype=torch.float32 if magnitude else torch.complex64, device=buf.device)
        window = np.ascontiguousarray(window, dtype=np.float32)
        flags = _lib.PAR_DEVICE_PTRS | (_lib.PAR_OUT_MAGNITUDE if magnitude else 0)
        stream = torch.cuda.current_stream(buf.device).cuda_stream
        rc = L.par_stft_range_f32(buf.data_ptr(), buf.shape[1], self.origin, self.n, ch, buf.stride(0), self.n_fft,
                                  self.hop, zeropad, window.ctypes.data, self.frame0, nfr, out.data_ptr(), F, nfr * F,
                                  flags, buf.device.index, stream)
        _lib.check(rc, "par_stft_range_f32")
        return out

    def segment_slice(self, n_seg):
        """Balanced block ``(per, a, b)`` of curve segments whose totals this rank computes."""
        per = -(-n_seg // self.world)
        a = min(self.rank * per, n_seg)
        return per, a, min(a + per, n_seg)

    def position_sums(self, sampletimes, speeds, device, group=None):
        """First half of the shared-sums positions: this rank's block of per-segment totals
        (``par_segment_sums_f64``) and ONE all-gather of ``8 * n_segments`` bytes.  Returns the handle
        ``positions(..., sums=...)`` takes.  Splitting the call lets a caller enqueue other work (the STFT of the
        step) between the collective and the serial host chain of the second half, which then hides behind it."""
        import torch
        import torch.distributed as dist
        L = _lib.lib()
        st = np.ascontiguousarray(sampletimes, dtype=np.float64)
        sp = np.ascontiguousarray(speeds, dtype=np.float64)
        n_seg = len(st) - 1
        per, a, b = self.segment_slice(n_seg)
        dev = torch.device(device)
        mine = torch.zeros(per, dtype=torch.float64, device=dev)
        seg_n = np.empty(n_seg, dtype=np.int64)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = L.par_segment_sums_f64(st.ctypes.data, sp.ctypes.data, len(st), a, b, mine.data_ptr(), seg_n.ctypes.data,
                                    _lib.PAR_DEVICE_PTRS, dev.index, stream)
        _lib.check(rc, "par_segment_sums_f64")
        sums = torch.empty(per * self.world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(sums, mine, group=group)
        return {"sums": sums, "seg_n": seg_n, "st": st, "sp": sp}

    def positions(self, sampletimes, speeds, device, out=None, group=None, share_sums=True, sums=None, stream=None):
        """This rank's slice of the read positions of a speed curve: the positions that fall into
        ``[s0 - 1, s1 + 1]`` plus the segment after them.  Returns ``(pos_slice, pos_origin, m_global)``;
        ``pos_slice[i]`` is position ``pos_origin + i``.

        With several ranks (and ``share_sums``) the per-segment totals the serial offset chain needs are
        computed once per JOB instead of once per rank: every rank sums its block of segments, one all-gather
        shares them (``position_sums``; pass its result as ``sums`` to split the call), and
        ``par_speed_to_pos_range_sums_f64`` expands the rank's window.  Otherwise ``par_speed_to_pos_range_f64``
        does everything locally.  Both give identical bits.  ``stream``: a torch stream for the second half (its
        copies, its two synchronisations and the expansion kernel); default the current stream."""
        import torch
        import torch.distributed as dist
        L = _lib.lib()
        if sums is None and share_sums and self.world > 1 and dist.is_initialized():
            sums = self.position_sums(sampletimes, speeds, device, group=group)
        if sums is not None:
            st, sp = sums["st"], sums["sp"]
        else:
            st = np.ascontiguousarray(sampletimes, dtype=np.float64)
            sp = np.ascontiguousarray(speeds, dtype=np.float64)
        if out is None:
            # a chunk read at >= 0.5x speed yields at most 2x its length in outputs (+ 2 segments of slack)
            out = torch.empty(2 * (self.s1 - self.s0) + 4 * int(np.max(np.diff(st)) * 2 + 16), dtype=torch.float64,
                              device=device)
        box = np.zeros(3, dtype=np.int64)
        lo = -np.inf if self.rank == 0 else float(self.s0) - 1.0
        hi = np.inf if self.rank == self.world - 1 else float(self.s1) + 1.0
        cu = (stream if stream is not None else torch.cuda.current_stream(out.device)).cuda_stream
        if sums is not None:
            rc = L.par_speed_to_pos_range_sums_f64(st.ctypes.data, sp.ctypes.data, len(st), float(self.n), lo, hi,
                                                   sums["sums"].data_ptr(), sums["seg_n"].ctypes.data, out.data_ptr(),
                                                   out.numel(), box[0:].ctypes.data, box[1:].ctypes.data,
                                                   box[2:].ctypes.data, _lib.PAR_DEVICE_PTRS, out.device.index, cu)
            _lib.check(rc, "par_speed_to_pos_range_sums_f64")
        else:
            rc = L.par_speed_to_pos_range_f64(st.ctypes.data, sp.ctypes.data, len(st), float(self.n), lo, hi, out.data_ptr(),
                                              out.numel(), box[0:].ctypes.data, box[1:].ctypes.data, box[2:].ctypes.data,
                                              _lib.PAR_DEVICE_PTRS, out.device.index, cu)
            _lib.check(rc, "par_speed_to_pos_range_f64")
        return out[:int(box[1])], int(box[0]), int(box[2])

    def output_range(self, pos, pos_origin=0, m=None):
        """Outputs whose read position falls into this rank's chunk ``[s0, s1)``: ``(o0, o1)``.
        ``pos`` is a monotone float64 device tensor holding positions ``pos_origin ...``."""
        import torch
        m = pos_origin + pos.numel() if m is None else m
        bounds = torch.tensor([float(self.s0), float(self.s1)], dtype=torch.float64, device=pos.device)
        o = torch.searchsorted(pos, bounds).tolist()
        o0 = 0 if self.rank == 0 else pos_origin + o[0]
        o1 = m if self.rank == self.world - 1 else pos_origin + o[1]
        return o0, max(o1, o0)

    def resample(self, buf, pos, mode="Sinc", out_range=None, pos_origin=0, m=None):
        """Resampled outputs ``[o0, o1)`` of every channel: ``(channels, o1 - o0)`` float32.
        ``pos`` holds positions ``[pos_origin, pos_origin + len(pos))`` of ``m`` in total."""
        import torch
        L = _lib.lib()
        m = pos_origin + pos.numel() if m is None else m
        o0, o1 = out_range if out_range is not None else self.output_range(pos, pos_origin, m)
        ch = buf.shape[0]
        out = torch.empty((ch, max(o1 - o0, 1)), dtype=torch.float32, device=buf.device)
        stream = torch.cuda.current_stream(buf.device).cuda_stream
        rc = L.par_resample_range_f32(pos.data_ptr(), pos_origin, pos.numel(), m, o0, o1, buf.data_ptr(), self.origin,
                                      buf.shape[1], self.n, ch, buf.stride(0),
                                      _lib.PAR_MODE_SINC if mode == "Sinc" else _lib.PAR_MODE_LINEAR, self.nt,
                                      out.data_ptr(), 1, out.stride(0), _lib.PAR_DEVICE_PTRS, buf.device.index, stream)
        _lib.check(rc, "par_resample_range_f32")
        return out[:, :o1 - o0]


def broadcast_curve(curve, src=0, device=None, group=None):
    """The speed curve lives on rank ``src`` (the GUI thread of the reference builds it,
    util/markers.py:595-599): one broadcast of its length and one of its ``(K, 2)`` float64 points."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(curve, dtype=np.float64)
    rank = dist.get_rank(group)
    k = torch.tensor([len(curve) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(k, src, group=group)
    t = torch.empty((int(k.item()), 2), dtype=torch.float64, device=device)
    if rank == src:
        t.copy_(torch.as_tensor(np.asarray(curve, dtype=np.float64)))
    dist.broadcast(t, src, group=group)
    return t.cpu().numpy()
