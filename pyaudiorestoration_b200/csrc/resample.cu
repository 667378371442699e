// resample.cu -- varispeed resampler for sm_100a: speed curve -> read positions -> windowed sinc.
//
// Replaces util/resampling.py:93-137 (speed_to_pos: the per-segment cumsum expansion; the serial
// segment chain stays on the host, see api.cu), :51-90 (sinc_core) with :21-46 (its wrappers)
// and the np.interp call of the "Linear" mode (:228-229).
//
// sinc kernel: one thread per output sample, one tile of SINC_TILE consecutive outputs per block
// iteration.  The input samples a tile touches ([min lower, max upper) over the tile) are staged
// once in shared memory with coalesced loads, so every input sample is read from HBM once per
// tile; the 2*NT tap weights of a sample are computed once and applied to all CH channels of the
// group.  Arithmetic:
//   weight_k = h[k] * fc * sinc((d_k - s) * fc) = h[k]/pi * sin(pi*fc*(d_k - s)) / (d_k - s)
//   fc == 1 (speed >= 1):  sin(pi*(d - s)) = (-1)^(d+1) * sin(pi*s)   -> one sinpi per sample
//   fc <  1 (speed <  1):  sin(theta_block + j*pi*fc) by angle addition from per-sample tables
//     (16 anchors + 16 steps), every table angle reduced EXACTLY modulo one turn in 64-bit
//     fixed point (fc needs more than float32 precision: its error is multiplied by up to NT).
// Positions, the rounding to the nearest input sample, the fractional shift and fc are
// float64 like the reference; the tap loop is float32 (parity bound 1e-6, see tests).
#include <limits.h>
#include <stdlib.h>

#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "par_internal.h"
#include "sinc_core.cuh"
#include "../../include/par_b200.h"

namespace par {

// ------------------------------------------------------------------------------------------
// positions
// ------------------------------------------------------------------------------------------

// v_j and the running sum exactly as np.arange(n)/(n-1)*(s1-s0)+s0 and np.cumsum(1/v) evaluate
// them (util/resampling.py:120,125): IEEE double ops, no FMA contraction.
// The quotient j/(n-1) is needed for every element but its divisor is fixed per segment: with
// rcp = RN(1/(n-1)), q0 = RN(j*rcp), the exact remainder r = j - q0*(n-1) (one FMA) and
// q = RN(q0 + r*rcp) give the correctly rounded quotient (Markstein's final-correction step, the
// same one the hardware division sequence ends with) at the cost of three FMA-class operations.
struct SegDiv {
	double nm1, rcp;
	bool exact;       // n - 1 >= 1: the correction step applies; otherwise fall back to a true division
	__device__ __forceinline__ explicit SegDiv(int64_t n) {
		nm1 = (double)(n - 1);
		exact = n >= 2;
		rcp = exact ? __ddiv_rn(1.0, nm1) : 0.0;
	}
	__device__ __forceinline__ double quot(double j) const {
		if (!exact) return __ddiv_rn(j, nm1);
		const double q0 = __dmul_rn(j, rcp);
		const double r = __fma_rn(-q0, nm1, j);
		return __fma_rn(r, rcp, q0);
	}
};

__device__ __forceinline__ double seg_speed(int64_t j, const SegDiv &sd, double ds, double s0) {
	return __dadd_rn(__dmul_rn(sd.quot((double)j), ds), s0);
}

__global__ void __launch_bounds__(128)
segment_sums_kernel(const double *__restrict__ speeds, const int64_t *__restrict__ seg_n,
                    int64_t n_seg, double *__restrict__ sums) {
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n_seg) return;
	const int64_t n = seg_n[i];
	const double s0 = speeds[i];
	const double ds = __dsub_rn(speeds[i + 1], s0);
	const SegDiv nm1(n);
	double acc = 0.0;
	int64_t j = 0;
	for (; j + 8 <= n; j += 8) {      // the divisions are independent: let them pipeline
		double r[8];
#pragma unroll
		for (int u = 0; u < 8; u++) r[u] = __ddiv_rn(1.0, seg_speed(j + u, nm1, ds, s0));
#pragma unroll
		for (int u = 0; u < 8; u++) acc = __dadd_rn(acc, r[u]);
	}
	for (; j < n; j++) acc = __dadd_rn(acc, __ddiv_rn(1.0, seg_speed(j, nm1, ds, s0)));
	sums[i] = acc;
}

// One thread per segment (the cumsum is serial in float64 by contract), but the stores go through
// a per-warp shared-memory tile so that every segment's 32-position run leaves as one 256-byte
// coalesced write instead of 32 scattered 8-byte ones.
constexpr int EXP_WARPS = 4;
__global__ void __launch_bounds__(32 * EXP_WARPS)
expand_positions_kernel(const double *__restrict__ speeds, const int64_t *__restrict__ seg_n,
                        const int64_t *__restrict__ seg_start, const double *__restrict__ seg_off,
                        int64_t n_seg, double *__restrict__ pos, int64_t m, double *__restrict__ sums_out) {
	// seg_off == nullptr: write the bare per-segment cumsum (offset 0) and its total to sums_out;
	// add_offsets_kernel finishes the job once the host has chained the offsets
	__shared__ double tile[EXP_WARPS][32][33];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	int64_t start = 0, n = 0;
	double s0 = 1.0, ds = 0.0, off = 0.0;
	int64_t n_full = 2;
	if (i < n_seg) {
		start = seg_start[i];
		n = seg_n[i];
		if (n < 0 || start >= m) n = 0;
		if (start + n > m) n = m - start;
		s0 = speeds[i];
		ds = __dsub_rn(speeds[i + 1], s0);
		n_full = seg_n[i];
		off = seg_off ? seg_off[i] : 0.0;
		if (sums_out) n = n_full > 0 ? n_full : 0;      // the total needs the whole segment
	}
	const SegDiv nm1(n_full);
	int64_t nmax = n;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
	double acc = 0.0;
	for (int64_t j0 = 0; j0 < nmax; j0 += 32) {
		if (j0 < n) {
			double r[32];
#pragma unroll
			for (int jj = 0; jj < 32; jj++) r[jj] = __ddiv_rn(1.0, seg_speed(j0 + jj, nm1, ds, s0));
#pragma unroll
			for (int jj = 0; jj < 32; jj++) {
				if (j0 + jj < n) acc = __dadd_rn(acc, r[jj]);
				tile[warp][lane][jj] = seg_off ? __dadd_rn(acc, off) : acc;
			}
		}
		__syncwarp();
		for (int row = 0; row < 32; row++) {
			const int64_t rn = __shfl_sync(0xffffffffu, n, row);
			const int64_t rs = __shfl_sync(0xffffffffu, start, row);
			if (j0 + lane < rn && rs + j0 + lane < m) pos[rs + j0 + lane] = tile[warp][row][lane];
		}
		__syncwarp();
	}
	if (sums_out && i < n_seg) sums_out[i] = acc;
}

// pos[start_i + j] = cumsum_j + off_i for every segment i (np.cumsum(...) + offset,
// util/resampling.py:125): one warp per segment, lanes streaming it with independent accesses.
__global__ void __launch_bounds__(256)
add_offsets_kernel(const int64_t *__restrict__ seg_n, const int64_t *__restrict__ seg_start,
                   const double *__restrict__ seg_off, int64_t n_seg, double *__restrict__ pos, int64_t m) {
	const int lane = threadIdx.x & 31;
	for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < n_seg;
	     i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
		const int64_t start = seg_start[i];
		int64_t n = seg_n[i];
		if (start + n > m) n = m - start;
		const double off = seg_off[i];
		double *p = pos + start;
		int64_t j = lane;
		for (; j + 96 < n; j += 128) {
			const double a = p[j], b = p[j + 32], c = p[j + 64], d = p[j + 96];
			p[j] = __dadd_rn(a, off);
			p[j + 32] = __dadd_rn(b, off);
			p[j + 64] = __dadd_rn(c, off);
			p[j + 96] = __dadd_rn(d, off);
		}
		for (; j < n; j += 32) p[j] = __dadd_rn(p[j], off);
	}
}

// Self-test of SegDiv::quot against the IEEE division for every j < n, n = 2 .. max_n.
__global__ void quotient_selftest_kernel(int64_t max_n, unsigned long long *mismatches) {
	unsigned long long bad = 0;
	for (int64_t n = 2 + blockIdx.x; n <= max_n; n += gridDim.x) {
		const SegDiv sd(n);
		for (int64_t j = threadIdx.x; j < n; j += blockDim.x)
			if (sd.quot((double)j) != __ddiv_rn((double)j, sd.nm1)) bad++;
	}
	if (bad) atomicAdd(mismatches, bad);
}

int launch_quotient_selftest(int64_t max_n, unsigned long long *mismatches_dev, cudaStream_t st) {
	quotient_selftest_kernel<<<1184, 256, 0, st>>>(max_n, mismatches_dev);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

int launch_segment_sums(const double *speeds_dev, const int64_t *seg_n_dev, int64_t n_seg,
                        double *sums_dev, cudaStream_t st) {
	if (n_seg <= 0) return PAR_OK;
	segment_sums_kernel<<<(unsigned)((n_seg + 127) / 128), 128, 0, st>>>(speeds_dev, seg_n_dev, n_seg,
	                                                                      sums_dev);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

int launch_add_offsets(const int64_t *seg_n_dev, const int64_t *seg_start_dev, const double *seg_offset_dev,
                       int64_t n_seg, double *pos_dev, int64_t m, cudaStream_t st) {
	if (n_seg <= 0 || m <= 0) return PAR_OK;
	int64_t blocks = (n_seg + 7) / 8;                 // 8 warps per block, one warp per segment
	if (blocks > 148 * 32) blocks = 148 * 32;
	add_offsets_kernel<<<(unsigned)blocks, 256, 0, st>>>(seg_n_dev, seg_start_dev, seg_offset_dev, n_seg, pos_dev, m);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

int launch_expand_positions(const double *speeds_dev, const int64_t *seg_n_dev,
                            const int64_t *seg_start_dev, const double *seg_offset_dev,
                            int64_t n_seg, double *pos_dev, int64_t m, cudaStream_t st, double *sums_out_dev) {
	if (n_seg <= 0 || (m <= 0 && !sums_out_dev)) return PAR_OK;
	const int tpb = 32 * EXP_WARPS;
	expand_positions_kernel<<<(unsigned)((n_seg + tpb - 1) / tpb), tpb, 0, st>>>(
	    speeds_dev, seg_n_dev, seg_start_dev, seg_offset_dev, n_seg, pos_dev, m, sums_out_dev);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

// ------------------------------------------------------------------------------------------
// windowed-sinc interpolation (arithmetic: sinc_core.cuh)
// ------------------------------------------------------------------------------------------
//
// One tile = SINC_TILE consecutive outputs of one channel group, handled by a CTA of 256 threads:
//   A  every output's float64 set-up (position -> nearest sample, fractional shift, fc, fixed-point
//      phases) by one thread per output; the tile's input span by a 32-bit redux over the tile;
//   B  the set-ups go to shared memory; the span is staged PLANAR in shared memory (coalesced loads,
//      every input sample leaves HBM once per tile); a block scan forms the work UNITS: an output
//      whose tap run starts at an even input index (E) followed by one starting at the next index (O)
//      become one unit, anything else is a unit of its own;
//   C  one thread per unit runs all 2*NT taps of its one or two outputs from the same 8-byte sample
//      pairs (sinc_unit), all channels of the group sharing the weights;
//   D  outputs at the edges of the signal (fewer than 2*NT taps, the reference's start-edge
//      misalignment) and tiles whose span does not fit take the scalar path that reads global memory.
// Which path an output takes and the arithmetic applied to it depend only on the positions, never on
// the tile or the chunk of a host call it falls into.

#ifndef SINC_THREADS_N
#define SINC_THREADS_N 256
#endif
constexpr int SINC_THREADS = SINC_THREADS_N;
constexpr int SINC_TILE = 2 * SINC_THREADS - 16;   // outputs per tile: ~T/2 units + room for 16 unpaired outputs in one round
constexpr int SINC_XPAD = 16;        // zeroed floats behind the staged span (block padding with zero coefficients reads them)
constexpr int SINC_XFRONT = 8;       // ... and in front of it
#ifndef SINC_MIN_BLOCKS
#define SINC_MIN_BLOCKS 2
#endif

constexpr unsigned SO_LIVE = 1u, SO_FAST = 2u, SO_LOWPASS = 4u;

// Weight of weight-index widx, any sample (edge / fallback path): h[k] fc sinc(fc (d - s)).
__device__ __forceinline__ float weight_single(const SincSetup &su, int widx, int nt,
                                               const float *__restrict__ ctab, const float *__restrict__ hptab) {
	const int d = widx - nt;
	const float q = (float)d - su.slot.s;
	if (!su.lowpass) return __ldg(ctab + widx) * sc_rcp(q);
	float sn, cs;
	if (d == 0) sn = sinpif(su.slot.fc * q);
	else sincos_fx(su.f_fx * (uint64_t)(int64_t)d - (uint64_t)su.slot.s_fx, &sn, &cs);
	return __ldg(hptab + widx) * sn * sc_rcp(q);
}

// Edge / fallback path: any tap count, samples read from global memory; each half of the tap run is
// accumulated from its far end towards the centre.
__device__ __forceinline__ float taps_slow(const SincSetup &su, int nt, const float *__restrict__ ctab,
                                           const float *__restrict__ hptab, const float *__restrict__ x,
                                           int64_t stride) {
	float accl = 0.f, accr = 0.f;
	const int mid = min(max(nt - su.koff, 0), su.cnt);       // taps [0, mid) lie left of the centre
	for (int k = 0; k < mid; k++)
		accl = fmaf(__ldg(x + (su.lower + k) * stride), weight_single(su, k + su.koff, nt, ctab, hptab), accl);
	for (int k = su.cnt - 1; k >= mid; k--)
		accr = fmaf(__ldg(x + (su.lower + k) * stride), weight_single(su, k + su.koff, nt, ctab, hptab), accr);
	float y = accl + accr;
	if (!su.lowpass) y *= sinpif(su.slot.s);
	return y;
}

__device__ __forceinline__ SincSetup sinc_setup_at(const SincArgs &a, int64_t i) {
	const double *pos = a.pos - a.pos_origin;
	const double p = pos[i];
	double per;
	if (i + 1 < a.m) per = fmax(1e-12, pos[i + 1] - p);
	else per = a.m >= 2 ? fmax(1e-12, pos[a.m - 1] - pos[a.m - 2]) : 0.0;
	return sinc_setup(p, per, a.nt, a.n_in, a.aligned_edges != 0);
}

// cp.async (LDGSTS): global -> shared copies that land while the CTA interpolates the previous tile
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Per-tile shared-memory state (double-buffered: tile w+1 is prepared while tile w is interpolated).
struct SincTileBuf {
	double pos[SINC_TILE + 2];     // read positions of the tile's outputs (+ the one after), prefetched
	unsigned long long g[SINC_TILE + 2];   // entry SINC_TILE of g, sfx, s, fc: a harmless stand-in for the empty half of a unit
	long long sfx[SINC_TILE + 2];
	int lo[SINC_TILE + 2];         // centre tap index relative to the tile's staged span (-1: not on the fast path)
	float s[SINC_TILE + 2];
	float fc[SINC_TILE + 2];
	unsigned flags[SINC_TILE + 2];
	int unit[SINC_TILE];           // first output of the unit | paired << 16
};
struct SincSmem {
	SincTileBuf tb[2];
	int red[2][SINC_THREADS / 32];
	int wsum[SINC_THREADS / 32];
};

// Uniform (per-CTA) description of a prepared tile, kept in registers by every thread.
struct SincTileInfo {
	int64_t i0;        // first output
	long long tlo;     // absolute input index of the staged span's first sample (multiple of 4)
	int span;          // staged samples
	int n_units;
	int ch0;
	bool staged;
};

template <int CH, int CAP>
__global__ void __launch_bounds__(SINC_THREADS, SINC_MIN_BLOCKS)
sinc_kernel(const __grid_constant__ SincArgs a, const __grid_constant__ SincTab<CAP> tab,
            const float *__restrict__ ctab, const float *__restrict__ hptab, const int span_cap) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SincSmem &sm = *reinterpret_cast<SincSmem *>(smem_raw);
	float *xs_all = reinterpret_cast<float *>(smem_raw + ((sizeof(SincSmem) + 15) & ~(size_t)15));
	const int xpitch = SINC_XFRONT + span_cap + SINC_XPAD;            // multiple of 4
	const int nt = a.nt;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int64_t tiles = (a.out_end - a.out_begin + SINC_TILE - 1) / SINC_TILE;
	const int groups = (a.n_ch + CH - 1) / CH;
	const int64_t work = tiles * groups;
	// work items are dealt round-robin (item j of this CTA is work item blockIdx.x + j * gridDim.x): all CTAs sit on the
	// same phase of the speed curve at any time, so stretches of low-passed outputs do not pile up on a few of them
	const int64_t w0 = 0;
	const int64_t w1 = (int64_t)blockIdx.x < work ? (work - 1 - blockIdx.x) / gridDim.x + 1 : 0;
	if (w0 >= w1) return;
	const double *posg = a.pos - a.pos_origin;
	for (int e = tid; e < 2 * CH * SINC_XFRONT; e += SINC_THREADS) {          // front padding of both buffers (both parity planes)
		const int bufi = e / (CH * SINC_XFRONT), r = e % (CH * SINC_XFRONT), par = r / (CH * SINC_XFRONT / 2), q = r % (CH * SINC_XFRONT / 2);
		xs_all[bufi * CH * xpitch + par * (xpitch / 2) * CH + q] = 0.f;
	}

	// positions of work item w -> tb[buf].pos (asynchronous)
	auto prefetch_pos = [&](int64_t tile, int buf) {
		const int64_t i0 = a.out_begin + tile * SINC_TILE;
		// a shard's positions reach one past its last output (par_resample_range_f32), no further
		int64_t cnt = (a.out_end + 1 < a.m ? a.out_end + 1 : a.m) - i0;
		if (cnt > SINC_TILE + 1) cnt = SINC_TILE + 1;
		for (int e = tid; e < cnt; e += SINC_THREADS) cp_async8(&sm.tb[buf].pos[e], posg + i0 + e);
	};

	// set-up of work item w from tb[buf].pos: per-output records, span, units; returns the tile description
	auto prepare = [&](int grp, int64_t tile, int buf) -> SincTileInfo {
		SincTileBuf &tb = sm.tb[buf];
		SincTileInfo ti;
		ti.ch0 = grp * CH;
		ti.i0 = a.out_begin + tile * SINC_TILE;
		double rf = rint(tb.pos[0]);
		if (!(rf > -9.0e15)) rf = -9.0e15;
		if (rf > 9.0e15) rf = 9.0e15;
		const long long ref = (long long)rf - nt;
		SincSetup su[2];
		bool live[2], fast[2];
		int lo_min = INT_MAX, hi_max = INT_MIN;
#pragma unroll
		for (int r = 0; r < 2; r++) {
			const int o = tid + r * SINC_THREADS;
			const int64_t i = ti.i0 + o;
			live[r] = o < SINC_TILE && i < a.out_end;
			fast[r] = false;
			if (live[r]) {
				const double p = tb.pos[o];
				double per;
				if (i + 1 < a.m) per = fmax(1e-12, tb.pos[o + 1] - p);
				else per = a.m >= 2 ? fmax(1e-12, posg[a.m - 1] - posg[a.m - 2]) : 0.0;
				su[r] = sinc_setup(p, per, nt, a.n_in, a.aligned_edges != 0);
				fast[r] = su[r].cnt == 2 * nt && su[r].koff == 0;
				if (fast[r]) {
					long long rel = su[r].lower - ref;
					rel = rel < -(1ll << 30) ? -(1ll << 30) : (rel > (1ll << 30) ? (1ll << 30) : rel);
					lo_min = min(lo_min, (int)rel);
					hi_max = max(hi_max, (int)rel + 2 * nt);
				}
			}
		}
		lo_min = __reduce_min_sync(0xffffffffu, lo_min);
		hi_max = __reduce_max_sync(0xffffffffu, hi_max);
		if (lane == 0) { sm.red[0][warp] = lo_min; sm.red[1][warp] = hi_max; }
		__syncthreads();
#pragma unroll
		for (int wv = 0; wv < SINC_THREADS / 32; wv++) { lo_min = min(lo_min, sm.red[0][wv]); hi_max = max(hi_max, sm.red[1][wv]); }
		ti.tlo = (ref + lo_min) & ~3ll;               // multiple of 4 (lower >= 0 on the fast path)
		const bool any = hi_max > lo_min;
		ti.span = any ? (int)(ref + hi_max - ti.tlo) : 0;
		ti.staged = any && lo_min > -(1 << 30) && hi_max < (1 << 30) && ti.span <= span_cap;
#pragma unroll
		for (int r = 0; r < 2; r++) {
			const int o = tid + r * SINC_THREADS;
			if (o < SINC_TILE) {
				const bool f = fast[r] && ti.staged;
				tb.flags[o] = (live[r] ? SO_LIVE : 0u) | (f ? SO_FAST : 0u) | (live[r] && su[r].lowpass ? SO_LOWPASS : 0u);
				if (f) {
					tb.lo[o] = (int)(su[r].lower - ti.tlo) + nt;
					tb.s[o] = su[r].slot.s;
					tb.fc[o] = su[r].slot.fc;
					tb.g[o] = su[r].slot.g_fx;
					tb.sfx[o] = su[r].slot.s_fx;
				} else {
					tb.lo[o] = -1;
				}
			}
		}
		if (tid < 2) {
			tb.flags[SINC_TILE + tid] = 0u; tb.lo[SINC_TILE + tid] = -1;
			tb.s[SINC_TILE + tid] = 0.5f; tb.fc[SINC_TILE + tid] = 1.f; tb.g[SINC_TILE + tid] = 0; tb.sfx[SINC_TILE + tid] = 0;
		}
		__syncthreads();
		// units: outputs 2 tid and 2 tid + 1; a unit starts at every fast output that is not the O half of a pair
		const int o = 2 * tid;
		bool st0 = false, st1 = false, hd0 = false, hd1 = false;
		if (o < SINC_TILE) {
			const unsigned fm = tb.flags[o > 0 ? o - 1 : SINC_TILE], f0 = tb.flags[o], f1 = tb.flags[o + 1], f2 = tb.flags[o + 2];
			const int lm = tb.lo[o > 0 ? o - 1 : SINC_TILE], l0 = tb.lo[o], l1 = tb.lo[o + 1], l2 = tb.lo[o + 2];
			const bool hdm = o > 0 && (fm & f0 & SO_FAST) && !(lm & 1) && l0 == lm + 1 && !((fm ^ f0) & SO_LOWPASS);
			hd0 = (f0 & f1 & SO_FAST) && !(l0 & 1) && l1 == l0 + 1 && !((f0 ^ f1) & SO_LOWPASS);
			hd1 = (f1 & f2 & SO_FAST) && !(l1 & 1) && l2 == l1 + 1 && !((f1 ^ f2) & SO_LOWPASS);
			st0 = (f0 & SO_FAST) && !hdm;
			st1 = (f1 & SO_FAST) && !hd0;
		}
		const int cnt = (int)st0 + (int)st1;
		int inc = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int v = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += v;
		}
		if (lane == 31) sm.wsum[warp] = inc;
		__syncthreads();
		int base = 0;
		ti.n_units = 0;
#pragma unroll
		for (int wv = 0; wv < SINC_THREADS / 32; wv++) {
			if (wv < warp) base += sm.wsum[wv];
			ti.n_units += sm.wsum[wv];
		}
		int idx = base + inc - cnt;
		if (st0) tb.unit[idx++] = o | (hd0 ? 0x10000 : 0);
		if (st1) tb.unit[idx] = (o + 1) | (hd1 ? 0x10000 : 0);
		return ti;
	};

	// the tile's input span -> xs buffer `buf` (asynchronous, channel-interleaved), zero padding behind it
	auto stage_span = [&](const SincTileInfo &ti, int buf) {
		if (!ti.staged) return;
		float *xs = xs_all + buf * CH * xpitch;
		const int plane = (xpitch / 2) * CH;
		const int len = ti.span + SINC_XPAD;
#pragma unroll
		for (int c = 0; c < CH; c++) {
			const bool have = ti.ch0 + c < a.n_ch;
			const float *src = a.signal + (int64_t)(ti.ch0 + c) * a.sig_ch_stride + (ti.tlo - a.sig_origin) * a.sig_stride;
			// element e -> parity plane (e & 1), slot (e + XFRONT) >> 1; two elements per thread and round
			float *d0 = xs + ((SINC_XFRONT >> 1) + tid) * CH + c, *d1 = d0 + plane;
			const int span = have ? ti.span : 0;
			if (a.sig_stride == 1) {
				const float *sp = src + 2 * tid;
				for (int e = 2 * tid; e < len; e += 2 * SINC_THREADS, sp += 2 * SINC_THREADS, d0 += SINC_THREADS * CH, d1 += SINC_THREADS * CH) {
					if (e < span) cp_async4(d0, sp); else *d0 = 0.f;
					if (e + 1 < span) cp_async4(d1, sp + 1); else *d1 = 0.f;
				}
			} else {
				for (int e = 2 * tid; e < len; e += 2 * SINC_THREADS, d0 += SINC_THREADS * CH, d1 += SINC_THREADS * CH) {
					if (e < span) cp_async4(d0, src + (int64_t)e * a.sig_stride); else *d0 = 0.f;
					if (e + 1 < span) cp_async4(d1, src + (int64_t)(e + 1) * a.sig_stride); else *d1 = 0.f;
				}
			}
		}
	};

	// (group, tile) of the work items w, w + 1, w + 2, advanced without 64-bit divisions
	int grp0 = (int)((int64_t)blockIdx.x / tiles);
	int64_t tile0 = (int64_t)blockIdx.x - (int64_t)grp0 * tiles;
	auto advance = [&](int &g, int64_t &t) { t += gridDim.x; while (t >= tiles) { t -= tiles; g++; } };
	int grp1 = grp0, grp2;
	int64_t tile1 = tile0, tile2;
	advance(grp1, tile1);
	grp2 = grp1; tile2 = tile1;
	advance(grp2, tile2);

	prefetch_pos(tile0, 0);
	cp_async_commit();
	cp_async_wait_all();
	__syncthreads();
	SincTileInfo cur = prepare(grp0, tile0, 0);
	stage_span(cur, 0);
	if (w0 + 1 < w1) prefetch_pos(tile1, 1);
	cp_async_commit();

	for (int64_t w = w0; w < w1; w++) {
		const int buf = (int)((w - w0) & 1);
		cp_async_wait_all();            // this tile's samples and the next tile's positions have landed
		__syncthreads();
		SincTileInfo nxt = cur;
		if (w + 1 < w1) {
			nxt = prepare(grp1, tile1, buf ^ 1);
			stage_span(nxt, buf ^ 1);
			if (w + 2 < w1) prefetch_pos(tile2, buf);     // tb[buf].pos is no longer needed
			cp_async_commit();
		}
		grp1 = grp2; tile1 = tile2;
		advance(grp2, tile2);
		__syncthreads();                // unit table of the next tile / of the first tile complete
		const SincTileBuf &tb = sm.tb[buf];
		const float *xs = xs_all + buf * CH * xpitch;

		// ---- one thread per unit ----
		for (int u = tid; u < cur.n_units; u += SINC_THREADS) {
			const int code = tb.unit[u];
			const int o = code & 0xffff;
			const bool paired = (code >> 16) != 0;
			const int lo0 = tb.lo[o];
			const int oE = (paired || !(lo0 & 1)) ? o : -1;
			const int oO = paired ? o + 1 : ((lo0 & 1) ? o : -1);
			const int j0 = lo0 & ~1;
			const bool lowpass = (tb.flags[o] & SO_LOWPASS) != 0;
			const int iE = oE >= 0 ? oE : SINC_TILE, iO = oO >= 0 ? oO : SINC_TILE;
			const SincSlotRef sl[2] = {{&tb.s[iE], &tb.fc[iE], &tb.g[iE], &tb.sfx[iE]}, {&tb.s[iO], &tb.fc[iO], &tb.g[iO], &tb.sfx[iO]}};
			float y[2][CH];
			const SincWin<CH> xu{xs + ((SINC_XFRONT + j0) >> 1) * CH, (xpitch / 2) * CH};
#if defined(SINC_EXPERIMENT_SKIP_TAPS)            // development aid: everything but the tap loop
#pragma unroll
			for (int c = 0; c < CH; c++) { y[0][c] = xu.at(0)[c] * *sl[0].s; y[1][c] = xu.at(1)[c] * *sl[1].s + (float)*sl[1].g_fx; }
#elif defined(SINC_EXPERIMENT_ALL_FC1)
			sinc_unit<CH, false, 2, CAP>(nt, tab, xu, sl, y);
#elif defined(SINC_EXPERIMENT_ALL_LOWPASS)
			sinc_unit<CH, true, 2, CAP>(nt, tab, xu, sl, y);
#else
			if (lowpass) sinc_unit<CH, true, 2, CAP>(nt, tab, xu, sl, y);
			else sinc_unit<CH, false, 2, CAP>(nt, tab, xu, sl, y);
#endif
#pragma unroll
			for (int c = 0; c < CH; c++) {
				if (cur.ch0 + c < a.n_ch) {
					float *dst = a.out + (int64_t)(cur.ch0 + c) * a.out_ch_stride - a.out_origin * a.out_stride;
					if (oE >= 0) dst[(cur.i0 + oE) * a.out_stride] = y[0][c];
					if (oO >= 0) dst[(cur.i0 + oO) * a.out_stride] = y[1][c];
				}
			}
		}

		// ---- edge outputs / unstaged tiles: scalar path from global memory ----
#pragma unroll
		for (int r = 0; r < 2; r++) {
			const int o = tid + r * SINC_THREADS;
			if (o < SINC_TILE && (tb.flags[o] & SO_LIVE) && !(tb.flags[o] & SO_FAST)) {
				const int64_t i = cur.i0 + o;
				const SincSetup su = sinc_setup_at(a, i);
				for (int c = 0; c < CH; c++) {
					if (cur.ch0 + c >= a.n_ch) break;
					float y = 0.f;
					if (su.cnt > 0)
						y = taps_slow(su, nt, ctab, hptab,
						              a.signal + (int64_t)(cur.ch0 + c) * a.sig_ch_stride - a.sig_origin * a.sig_stride, a.sig_stride);
					a.out[(int64_t)(cur.ch0 + c) * a.out_ch_stride + (i - a.out_origin) * a.out_stride] = y;
				}
			}
		}
		cur = nxt;
	}
}


// ------------------------------------------------------------------------------------------
// Warp-specialised variant: one CTA per SM, WS_CWARPS interpolating warps + WS_PWARPS set-up warps.
// The set-up warps run stages A and B of tile w + 1 (positions -> records, span, unit table, cp.async of
// the span) while the interpolating warps run stage C/D of tile w; the two sides meet only at four
// named barriers (full / empty per buffer).  Same records, same units, same sinc_unit arithmetic as
// sinc_kernel: the outputs are bit-identical, only who computes the records differs.
// ------------------------------------------------------------------------------------------
#ifndef SINC_WS_CWARPS
#define SINC_WS_CWARPS 12
#endif
#ifndef SINC_WS_PWARPS
#define SINC_WS_PWARPS 4
#endif
#ifndef SINC_WS_UNROLL
#define SINC_WS_UNROLL 2
#endif
// Registers after the role split (setmaxnreg; per scheduler 3 * C + P <= 512): the interpolating warps run the block
// loop unrolled twice with 152, the set-up warps need 56.  Measured: 5.80 vs 5.94 ms against 128 / 128, unroll 1.
#ifndef SINC_WS_REGS_C
#define SINC_WS_REGS_C 152
#endif
#ifndef SINC_WS_REGS_P
#define SINC_WS_REGS_P 56
#endif
#ifndef SINC_WS_BLOCK_UNROLL
#define SINC_WS_BLOCK_UNROLL 2
#endif
constexpr int WS_CT = 32 * SINC_WS_CWARPS, WS_PT = 32 * SINC_WS_PWARPS, WS_THREADS = WS_CT + WS_PT;
constexpr int WS_TILE = 2 * WS_CT - 16;
constexpr int WS_PER = (WS_TILE + WS_PT - 1) / WS_PT;       // outputs per set-up thread
constexpr int kWsUnroll = SINC_WS_UNROLL;
constexpr int WS_BAR_FULL = 1, WS_BAR_EMPTY = 3, WS_BAR_PROD = 5;   // named barriers (0 is __syncthreads)

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// One word per output: flags in bits 0..2, (tap-run start relative to the tile's reference sample + WS_BIAS) above.
constexpr int WS_BIAS = 1 << 27;
struct WsTileBuf {
	unsigned long long g[WS_TILE + 2];     // entry WS_TILE of g, sfx, s, fc: stand-in for the empty half of a unit
	long long sfx[WS_TILE + 2];
	float s[WS_TILE + 2];
	float fc[WS_TILE + 2];
	unsigned rec[WS_TILE + 2];      // entries WS_TILE, WS_TILE + 1 stay 0 (not fast)
	int unit[WS_TILE];              // first output of the unit | paired << 16
	long long i0;
	int n_units, ch0, adj, staged;  // centre index in the staged span = (rec >> 3) + adj
	int rot, pad_;                  // unit u belongs to interpolating thread (u + rot) % WS_CT
};
struct WsSmem {
	WsTileBuf tb[2];
	double pos[2][WS_TILE + 2];
	int red[2][SINC_WS_PWARPS];
	int wsum[SINC_WS_PWARPS];
	unsigned long long full_bar[2]; // mbarriers: tile buffer b is ready (the set-up threads arrive, the others only wait)
};

__device__ __forceinline__ void ws_mbar_init(unsigned long long *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(unsigned long long *bar) {
	asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool ws_mbar_test(unsigned long long *bar, unsigned parity) {
	unsigned ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
	             : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
	return ok != 0;
}

template <int CH, int CAP>
__global__ void __launch_bounds__(WS_THREADS, 1)
sinc_kernel_ws(const __grid_constant__ SincArgs a, const __grid_constant__ SincTab<CAP> tab,
               const float *__restrict__ ctab, const float *__restrict__ hptab, const int span_cap, const int TL) {
	// TL <= WS_TILE: outputs per tile of this launch, chosen by the host so that a tile's units fill one round of the
	// interpolating threads at the expected read period (launch_sinc_ws)
	extern __shared__ __align__(16) unsigned char smem_raw[];
	WsSmem &sm = *reinterpret_cast<WsSmem *>(smem_raw);
	float *xs_all = reinterpret_cast<float *>(smem_raw + ((sizeof(WsSmem) + 15) & ~(size_t)15));
	const int xpitch = SINC_XFRONT + span_cap + SINC_XPAD;
	const int nt = a.nt;
	const int tid = threadIdx.x;
	const int64_t tiles = (a.out_end - a.out_begin + TL - 1) / TL;
	const int groups = (a.n_ch + CH - 1) / CH;
	const int64_t work = tiles * groups;
	// Work items (tile, channel group) are dealt round-robin: at any moment the CTAs work on neighbouring tiles, i.e. on
	// the same phase of the speed curve, so stretches of low-passed outputs (which cost ~1.6x) do not pile up on a few
	// CTAs the way they do with one contiguous range per CTA; neighbours also find each other's tap halo in L2.
	// w0 / w1 count this CTA's items: item j is work item blockIdx.x + j * gridDim.x.
	const int64_t w0 = 0;
	const int64_t w1 = (int64_t)blockIdx.x < work ? (work - 1 - blockIdx.x) / gridDim.x + 1 : 0;
	if (w0 >= w1) return;
	const double *posg = a.pos - a.pos_origin;
	for (int e = tid; e < 2 * CH * SINC_XFRONT; e += WS_THREADS) {
		const int bufi = e / (CH * SINC_XFRONT), r = e % (CH * SINC_XFRONT), par = r / (CH * SINC_XFRONT / 2), q = r % (CH * SINC_XFRONT / 2);
		xs_all[bufi * CH * xpitch + par * (xpitch / 2) * CH + q] = 0.f;
	}
	if (tid == 0) {
		ws_mbar_init(&sm.full_bar[0], WS_PT);
		ws_mbar_init(&sm.full_bar[1], WS_PT);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (tid < 4) {
		WsTileBuf &t = sm.tb[tid >> 1];
		const int e = WS_TILE + (tid & 1);
		t.rec[TL + (tid & 1)] = 0u;          // the unit scan looks one output past the tile
		t.s[e] = 0.5f; t.fc[e] = 1.f; t.g[e] = 0; t.sfx[e] = 0;
	}
	__syncthreads();
	int grp = (int)((int64_t)blockIdx.x / tiles);
	int64_t tile = (int64_t)blockIdx.x - (int64_t)grp * tiles;

	if (tid >= WS_CT) {
		// =============================== set-up warps ===============================
#if SINC_WS_REGS_P > 0
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SINC_WS_REGS_P));
#endif
		const int pt = tid - WS_CT, lane = pt & 31, pw = pt >> 5;
		auto prefetch_pos = [&](int64_t tl, int pb) {
			const int64_t i0 = a.out_begin + tl * TL;
			// a shard's positions reach one past its last output (par_resample_range_f32), no further
			int64_t cnt = (a.out_end + 1 < a.m ? a.out_end + 1 : a.m) - i0;
			if (cnt > TL + 1) cnt = TL + 1;
			for (int e = pt; e < cnt; e += WS_PT) cp_async8(&sm.pos[pb][e], posg + i0 + e);
		};
		prefetch_pos(tile, 0);
		cp_async_commit();
		// read period of the last output (the reference repeats the previous one); only the launch that holds the last
		// output has those two positions in its slice
		const double per_tail = (a.out_end >= a.m && a.m >= 2) ? fmax(1e-12, posg[a.m - 1] - posg[a.m - 2]) : 0.0;
		const bool aligned = a.aligned_edges != 0;
		int rot = 0;
		for (int64_t w = w0; w < w1; w++) {
			const int buf = (int)((w - w0) & 1);
			WsTileBuf &tb = sm.tb[buf];
			const double *tpos = sm.pos[buf];
			int grp_n = grp;
			int64_t tile_n = tile + gridDim.x;
			while (tile_n >= tiles) { tile_n -= tiles; grp_n++; }
			cp_async_wait_all();                                   // positions of this tile (issued a tile ago)
			named_sync(WS_BAR_PROD, WS_PT);
			if (w + 1 < w1) prefetch_pos(tile_n, buf ^ 1);         // pos[buf ^ 1] was last read two barriers ago
			cp_async_commit();
			if (w >= w0 + 2) named_sync(WS_BAR_EMPTY + buf, WS_THREADS);   // tile w - 2 has left this buffer
			const int64_t i0 = a.out_begin + tile * TL;
			double rf = rint(tpos[0]);
			if (!(rf > -9.0e15)) rf = -9.0e15;
			if (rf > 9.0e15) rf = 9.0e15;
			const long long ref = (long long)rf - nt;
			int lo_min = INT_MAX, hi_max = INT_MIN;
			// pass 1: the float64 set-up of every output, straight-line so that the unrolled iterations interleave
			// (outputs behind the end of the range are set up from whatever the buffer holds and marked dead)
#pragma unroll kWsUnroll
			for (int o = pt; o < TL; o += WS_PT) {
				const int64_t i = i0 + o;
				const double p = tpos[o];
				const double per = i + 1 < a.m ? fmax(1e-12, tpos[o + 1] - p) : per_tail;
				const SincSetup su = sinc_setup(p, per, nt, a.n_in, aligned);
				const bool live = i < a.out_end;
				const bool fast = live && su.cnt == 2 * nt && su.koff == 0;
				long long rel = su.lower - ref;
				rel = rel < -(1ll << 26) ? -(1ll << 26) : (rel > (1ll << 26) ? (1ll << 26) : rel);
				const int rel32 = (int)rel;
				lo_min = fast ? min(lo_min, rel32) : lo_min;
				hi_max = fast ? max(hi_max, rel32 + 2 * nt) : hi_max;
				tb.s[o] = su.slot.s;
				tb.fc[o] = su.slot.fc;
				tb.g[o] = su.slot.g_fx;
				tb.sfx[o] = su.slot.s_fx;
				tb.rec[o] = (live ? SO_LIVE : 0u) | (live && su.lowpass ? SO_LOWPASS : 0u) |
				            (fast ? SO_FAST | ((unsigned)(rel32 + WS_BIAS) << 3) : 0u);
			}
			lo_min = __reduce_min_sync(0xffffffffu, lo_min);
			hi_max = __reduce_max_sync(0xffffffffu, hi_max);
			if (lane == 0) { sm.red[0][pw] = lo_min; sm.red[1][pw] = hi_max; }
			named_sync(WS_BAR_PROD, WS_PT);
#pragma unroll
			for (int wv = 0; wv < SINC_WS_PWARPS; wv++) { lo_min = min(lo_min, sm.red[0][wv]); hi_max = max(hi_max, sm.red[1][wv]); }
			const long long tlo = (ref + lo_min) & ~3ll;
			const bool any = hi_max > lo_min;
			const int span = any ? (int)(ref + hi_max - tlo) : 0;
			const bool staged = any && lo_min > -(1 << 26) && hi_max < (1 << 26) && span <= span_cap;
			// the span's copies go out first: they fly while the units are formed
			if (staged) {
				float *xs = xs_all + buf * CH * xpitch;
				const int plane = (xpitch / 2) * CH;
				const int len = span + SINC_XPAD;
#pragma unroll
				for (int c = 0; c < CH; c++) {
					const bool have = grp * CH + c < a.n_ch;
					const float *src = a.signal + (int64_t)(grp * CH + c) * a.sig_ch_stride + (tlo - a.sig_origin) * a.sig_stride;
					float *d0 = xs + ((SINC_XFRONT >> 1) + pt) * CH + c, *d1 = d0 + plane;
					const int sp_n = have ? span : 0;
					if (a.sig_stride == 1) {
						const float *sp = src + 2 * pt;
						for (int e = 2 * pt; e < len; e += 2 * WS_PT, sp += 2 * WS_PT, d0 += WS_PT * CH, d1 += WS_PT * CH) {
							if (e < sp_n) cp_async4(d0, sp); else *d0 = 0.f;
							if (e + 1 < sp_n) cp_async4(d1, sp + 1); else *d1 = 0.f;
						}
					} else {
						for (int e = 2 * pt; e < len; e += 2 * WS_PT, d0 += WS_PT * CH, d1 += WS_PT * CH) {
							if (e < sp_n) cp_async4(d0, src + (int64_t)e * a.sig_stride); else *d0 = 0.f;
							if (e + 1 < sp_n) cp_async4(d1, src + (int64_t)(e + 1) * a.sig_stride); else *d1 = 0.f;
						}
					}
				}
			}
			cp_async_commit();
			// pass 2: units over this thread's run of consecutive outputs; the centre index of an output in
			// the staged span is (rec >> 3) + adj, so its parity is that of (rec >> 3) + adj
			const int adj = (int)(ref - tlo) + nt - WS_BIAS;
			const int c0 = pt * WS_PER;
			unsigned wd[WS_PER + 2];
#pragma unroll
			for (int j = 0; j < WS_PER + 2; j++) {
				const int o = c0 - 1 + j;
				wd[j] = (o >= 0 && o <= TL) ? tb.rec[o] : 0u;
			}
			unsigned starts = 0u, heads = 0u;
			if (staged) {
				bool hprev;
				{
					const unsigned k0 = wd[0] >> 3, k1 = wd[1] >> 3;
					hprev = (wd[0] & wd[1] & SO_FAST) && !((k0 + (unsigned)adj) & 1u) && k1 == k0 + 1 && !((wd[0] ^ wd[1]) & SO_LOWPASS);
				}
#pragma unroll
				for (int j = 1; j <= WS_PER; j++) {
					const unsigned k0 = wd[j] >> 3, k1 = wd[j + 1] >> 3;
					const bool h = (wd[j] & wd[j + 1] & SO_FAST) && !((k0 + (unsigned)adj) & 1u) && k1 == k0 + 1 &&
					               !((wd[j] ^ wd[j + 1]) & SO_LOWPASS);
					if (c0 + j - 1 < TL) {
						if ((wd[j] & SO_FAST) && !hprev) starts |= 1u << (j - 1);
						if (h) heads |= 1u << (j - 1);
					}
					hprev = h;
				}
			}
			const int cnt = __popc(starts);
			int inc = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int v = __shfl_up_sync(0xffffffffu, inc, d);
				if (lane >= d) inc += v;
			}
			if (lane == 31) sm.wsum[pw] = inc;
			named_sync(WS_BAR_PROD, WS_PT);
			int base = 0, total = 0;
#pragma unroll
			for (int wv = 0; wv < SINC_WS_PWARPS; wv++) {
				if (wv < pw) base += sm.wsum[wv];
				total += sm.wsum[wv];
			}
			int idx = base + inc - cnt;
			while (starts) {
				const int j = __ffs(starts) - 1;
				starts &= starts - 1;
				tb.unit[idx++] = (c0 + j) | (((heads >> j) & 1u) ? 0x10000 : 0);
			}
			if (pt == 0) {
				tb.i0 = i0; tb.n_units = total; tb.ch0 = grp * CH; tb.adj = adj; tb.staged = staged ? 1 : 0;
				// the units beyond a full round go to a different set of threads every tile
				tb.rot = rot;
				rot = (rot + total) % WS_CT;
			}
			cp_async_wait_all();                                    // the span (and the next positions) have landed
			__threadfence_block();
			ws_mbar_arrive(&sm.full_bar[buf]);
			grp = grp_n; tile = tile_n;
		}
		return;
	}

	// =============================== interpolating warps ===============================
#if SINC_WS_REGS_C > 0
	asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SINC_WS_REGS_C));
#endif
	for (int64_t w = w0; w < w1; w++) {
		const int buf = (int)((w - w0) & 1);
		// no barrier among the interpolating warps: a warp moves on as soon as ITS units of the tile are done
		while (!ws_mbar_test(&sm.full_bar[buf], (unsigned)(((w - w0) >> 1) & 1))) {}
		const WsTileBuf &tb = sm.tb[buf];
		const float *xs = xs_all + buf * CH * xpitch;
		const int64_t i0 = tb.i0;
		const int n_units = tb.n_units, ch0 = tb.ch0, adj = tb.adj;
		int u_first = tid - tb.rot;
		if (u_first < 0) u_first += WS_CT;
		const bool staged = tb.staged != 0;
		for (int u = u_first; u < n_units; u += WS_CT) {
			const int code = tb.unit[u];
			const int o = code & 0xffff;
			const bool paired = (code >> 16) != 0;
			const unsigned word = tb.rec[o];
			const int lo0 = (int)(word >> 3) + adj;
			const int oE = (paired || !(lo0 & 1)) ? o : -1;
			const int oO = paired ? o + 1 : ((lo0 & 1) ? o : -1);
			const int j0 = lo0 & ~1;
			const bool lowpass = (word & SO_LOWPASS) != 0;
			const int iE = oE >= 0 ? oE : WS_TILE, iO = oO >= 0 ? oO : WS_TILE;
			const SincSlotRef sl[2] = {{&tb.s[iE], &tb.fc[iE], &tb.g[iE], &tb.sfx[iE]}, {&tb.s[iO], &tb.fc[iO], &tb.g[iO], &tb.sfx[iO]}};
			float y[2][CH];
			const SincWin<CH> xu{xs + ((SINC_XFRONT + j0) >> 1) * CH, (xpitch / 2) * CH};
			if (lowpass) sinc_unit<CH, true, 2, CAP, SINC_WS_BLOCK_UNROLL>(nt, tab, xu, sl, y);
			else sinc_unit<CH, false, 2, CAP, SINC_WS_BLOCK_UNROLL>(nt, tab, xu, sl, y);
#pragma unroll
			for (int c = 0; c < CH; c++) {
				if (ch0 + c < a.n_ch) {
					float *dst = a.out + (int64_t)(ch0 + c) * a.out_ch_stride - a.out_origin * a.out_stride;
					if (oE >= 0) dst[(i0 + oE) * a.out_stride] = y[0][c];
					if (oO >= 0) dst[(i0 + oO) * a.out_stride] = y[1][c];
				}
			}
		}
		for (int o = tid; o < TL; o += WS_CT) {
			const unsigned fl = tb.rec[o];
			if ((fl & SO_LIVE) && !((fl & SO_FAST) && staged)) {
				const int64_t i = i0 + o;
				const SincSetup su = sinc_setup_at(a, i);
				for (int c = 0; c < CH; c++) {
					if (ch0 + c >= a.n_ch) break;
					float y = 0.f;
					if (su.cnt > 0)
						y = taps_slow(su, nt, ctab, hptab,
						              a.signal + (int64_t)(ch0 + c) * a.sig_ch_stride - a.sig_origin * a.sig_stride, a.sig_stride);
					a.out[(int64_t)(ch0 + c) * a.out_ch_stride + (i - a.out_origin) * a.out_stride] = y;
				}
			}
		}
		if (w + 2 < w1) named_arrive(WS_BAR_EMPTY + buf, WS_THREADS);
	}
}

// distance table of sinc_core.cuh per NT (host cache, passed to the kernel by value)
template <int CAP>
static const SincTab<CAP> *sinc_param_table(int nt) {
	static std::mutex mu;
	static std::map<int, std::unique_ptr<SincTab<CAP>>> cache;
	std::lock_guard<std::mutex> lk(mu);
	auto it = cache.find(nt);
	if (it == cache.end()) {
		std::unique_ptr<SincTab<CAP>> t(new SincTab<CAP>());
		sinc_fill_table<CAP>(nt, t.get());
		it = cache.emplace(nt, std::move(t)).first;
	}
	return it->second.get();
}

// Which kernel: the warp-specialised one from 64 taps (NT 32) up.  Measured on B200, 300 s x 2 ch at 96 kHz, two-CTA vs
// warp-specialised: NT 16: 0.99 / 1.08 ms, 24: 1.11 / 1.13, 32: 1.22 / 1.19, 40: 1.33 / 1.25, 50: 1.58 / 1.39, 64: 1.71 / 1.49
// (with few taps the set-up warps cannot keep up).  PAR_B200_SINC_WS=0/1 or the PAR_SINC_KERNEL_* flags force one.
static bool sinc_use_ws(int nt, int kernel) {
	if (kernel) return kernel == 2;
	static const int forced = [] { const char *e = getenv("PAR_B200_SINC_WS"); return e && (e[0] == '0' || e[0] == '1') ? e[0] - '0' : -1; }();
	if (forced >= 0) return forced == 1;
	return nt >= 32;
}

template <int CH, int CAP>
static int launch_sinc_ws(const SincArgs &a, int device, cudaStream_t st, const SincTables &tb) {
	// widest span staged in shared memory: a tile read at up to ~4x speed (one CTA per SM: room for it)
	int span_cap = 4 * WS_TILE + 2 * a.nt + 8;
	span_cap = (span_cap + 3) & ~3;
	const int smem = (int)((sizeof(WsSmem) + 15) & ~(size_t)15) + 2 * CH * (SINC_XFRONT + span_cap + SINC_XPAD) * (int)sizeof(float);
	auto kern = sinc_kernel_ws<CH, CAP>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	// Outputs per tile.  Consecutive outputs pair into one unit only while the read position advances by about one
	// sample: a tile of T outputs holds ~T/2 * (1 + |period - 1|) units (T at twice the speed).  A few units more than
	// interpolating threads would cost a whole extra round, so the tile is sized for the largest period deviation the
	// caller expects (the speed curve's, when there is one; else the average over the job plus a margin).
	double dev = a.period_dev >= 0.0 ? a.period_dev : fabs((double)a.n_in / (double)(a.m > 0 ? a.m : 1) - 1.0) + 0.01;
	if (!(dev < 1.0)) dev = 1.0;
	int tile_len = (int)(2.0 * WS_CT / (1.0 + dev)) - 8;
	tile_len = tile_len < 64 ? 64 : (tile_len > WS_TILE ? WS_TILE : tile_len);
	const int64_t tiles = (a.out_end - a.out_begin + tile_len - 1) / tile_len;
	const int64_t work = tiles * ((a.n_ch + CH - 1) / CH);
	int64_t grid = sm_count(device);
	if (grid > work) grid = work;
	if (grid < 1) return PAR_OK;
	const SincTab<CAP> *pt = sinc_param_table<CAP>(a.nt);
	kern<<<(unsigned)grid, WS_THREADS, smem, st>>>(a, *pt, tb.c, tb.hp, span_cap, tile_len);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

template <int CH, int CAP>
static int launch_sinc_ch(const SincArgs &a, int device, cudaStream_t st, const SincTables &tb) {
	if (sinc_use_ws(a.nt, a.kernel)) return launch_sinc_ws<CH, CAP>(a, device, st, tb);
	// widest span staged in shared memory: a tile read at up to ~2.5x speed
	int span_cap = 2 * SINC_TILE + 2 * a.nt + 8;
	span_cap = (span_cap + 3) & ~3;
	const int smem = (int)((sizeof(SincSmem) + 15) & ~(size_t)15) + 2 * CH * (SINC_XFRONT + span_cap + SINC_XPAD) * (int)sizeof(float);
	auto kern = sinc_kernel<CH, CAP>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SINC_THREADS, smem));
	if (occ < 1) occ = 1;
	const int64_t tiles = (a.out_end - a.out_begin + SINC_TILE - 1) / SINC_TILE;
	const int64_t work = tiles * ((a.n_ch + CH - 1) / CH);
	int64_t grid = (int64_t)occ * sm_count(device);
	if (grid > work) grid = work;
	if (grid < 1) return PAR_OK;
	const SincTab<CAP> *pt = sinc_param_table<CAP>(a.nt);
	kern<<<(unsigned)grid, SINC_THREADS, smem, st>>>(a, *pt, tb.c, tb.hp, span_cap);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

template <int CH>
static int launch_sinc_cap(const SincArgs &a, int device, cudaStream_t st, const SincTables &tb) {
	if (sinc_num_blocks(a.nt) * (SINC_BLOCK / 2) <= SINC_TAB_SMALL) return launch_sinc_ch<CH, SINC_TAB_SMALL>(a, device, st, tb);
	return launch_sinc_ch<CH, SINC_TAB_LARGE>(a, device, st, tb);
}

int launch_sinc(const SincArgs &a, int device, cudaStream_t st) {
	if (a.m <= 0 || a.n_ch <= 0 || a.out_end <= a.out_begin) return PAR_OK;
	SincTables tb;
	int rc = sinc_tables(device, a.nt, st, &tb);
	if (rc != PAR_OK) return rc;
	// The tap weights of an output are computed once per channel group of up to 4 channels.
	if (a.n_ch >= 4) return launch_sinc_cap<4>(a, device, st, tb);
	if (a.n_ch >= 2) return launch_sinc_cap<2>(a, device, st, tb);
	return launch_sinc_cap<1>(a, device, st, tb);
}

// ------------------------------------------------------------------------------------------
// linear interpolation (np.interp(sample_at, arange(L), signal, left=0, right=0))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_kernel(SincArgs a) {
	const int64_t span = a.out_end - a.out_begin;
	const int64_t total = span * a.n_ch;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
	     t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t ch = t / span, i = a.out_begin + (t - ch * span);
		const double p = a.pos[i - a.pos_origin];
		const float *x = a.signal + ch * a.sig_ch_stride - a.sig_origin * a.sig_stride;
		float y = 0.f;
		if (p >= 0.0 && p <= (double)(a.n_in - 1)) {
			long long j = (long long)floor(p);
			if (j >= a.n_in - 1) {
				y = __ldg(x + (a.n_in - 1) * a.sig_stride);
			} else {
				// numpy: slope * (x - xp[j]) + fp[j], float64
				const double f0 = (double)__ldg(x + j * a.sig_stride);
				const double f1 = (double)__ldg(x + (j + 1) * a.sig_stride);
				y = (float)__dadd_rn(__dmul_rn(__dsub_rn(f1, f0), __dsub_rn(p, (double)j)), f0);
			}
		}
		a.out[ch * a.out_ch_stride + (i - a.out_origin) * a.out_stride] = y;
	}
}

int launch_linear(const SincArgs &a, int device, cudaStream_t st) {
	const int64_t total = (a.out_end - a.out_begin) * a.n_ch;
	if (total <= 0) return PAR_OK;
	int64_t grid = (total + 255) / 256;
	const int64_t cap = (int64_t)sm_count(device) * 16;
	if (grid > cap) grid = cap;
	linear_kernel<<<(unsigned)grid, 256, 0, st>>>(a);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

}  // namespace par
