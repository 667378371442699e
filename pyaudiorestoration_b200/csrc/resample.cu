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
#include <stdlib.h>

#include "par_internal.h"
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
// windowed-sinc interpolation
// ------------------------------------------------------------------------------------------

constexpr int SINC_TILE = 256;       // outputs per block iteration == threads per block
#ifndef SINC_MIN_BLOCKS
#define SINC_MIN_BLOCKS 3
#endif
constexpr int SINC_XPAD = 32;        // zero padding behind the staged span (the last tap block may overhang)

__device__ __forceinline__ float rcp_approx(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// sin/cos of the angle  pi * phase / 2^63  (phase wraps at one full turn = 2^64).
// The top 32 bits are split into the nearest quarter turn and a residual in [-pi/4, pi/4) that is
// converted to float32 (absolute error <= 5e-8 rad) and fed to Taylor polynomials whose truncation
// error is < 2e-9 on that interval.
__device__ __forceinline__ void sincos_fx(uint64_t phase, float *s, float *c) {
	const uint32_t t = (uint32_t)(phase >> 32);
	const uint32_t quad = (t + 0x20000000u) >> 30;
	const int32_t res = (int32_t)(t - (quad << 30));
	const float x = (float)res * 1.4629180792671596e-9f;      // pi / 2^31
	const float x2 = x * x;
	float ps = fmaf(x2, 2.7557319e-6f, -1.9841270e-4f);
	ps = fmaf(x2, ps, 8.3333333e-3f);
	ps = fmaf(x2, ps, -1.6666667e-1f);
	const float sn = fmaf(x * x2, ps, x);
	float pc = fmaf(x2, -2.7557319e-7f, 2.4801587e-5f);
	pc = fmaf(x2, pc, -1.3888889e-3f);
	pc = fmaf(x2, pc, 4.1666667e-2f);
	pc = fmaf(x2, pc, -0.5f);
	const float cs = fmaf(x2, pc, 1.0f);
	const float a = (quad & 1) ? cs : sn;
	const float b = (quad & 1) ? sn : cs;
	*s = (quad & 2) ? -a : a;
	*c = ((quad + 1) & 2) ? -b : b;
}

struct SampleSetup {
	int64_t lower;     // first input sample of the tap run
	int cnt;           // number of taps (0 .. 2NT)
	int koff;          // weight index of tap 0 (0 unless PAR_SINC_ALIGNED_EDGES at the start edge)
	float s;           // fractional shift p - round(p), never exactly 0
	float fc;          // fc rounded to float32 (centre tap only)
	bool lowpass;      // fc < 1
	uint64_t f_fx;     // fc in units of 2^-63 half-turns
	int64_t s_fx;      // fc * s in the same units
};

// float64 part of util/resampling.py:67-84 for output i
__device__ __forceinline__ SampleSetup sample_setup(const SincArgs &a, int64_t i) {
	SampleSetup su;
	const int nt = a.nt;
	const double *pos = a.pos - a.pos_origin;
	const double p = pos[i];
	double per;
	if (i + 1 < a.m) per = fmax(1e-12, pos[i + 1] - p);
	else per = a.m >= 2 ? fmax(1e-12, pos[a.m - 1] - pos[a.m - 2]) : 0.0;
	double fc = 1.0 / per;
	if (!(fc < 1.0)) fc = 1.0;
	double pr = rint(p);                        // half to even, like Python's round()
	if (!(pr > -9.0e15)) pr = -9.0e15;          // NaN / -inf guard (garbage in, zeros out)
	if (pr > 9.0e15) pr = 9.0e15;
	const long long ind = (long long)pr;
	const double sd = p - pr;
	long long lower = ind - nt, upper = ind + nt;
	if (lower < 0) lower = 0;
	if (upper > a.n_in) upper = a.n_in;
	su.lower = lower;
	su.cnt = upper > lower ? (int)(upper - lower) : 0;
	su.koff = a.aligned_edges ? (int)(lower - (ind - nt)) : 0;
	float s = (float)sd;
	if (s == 0.f) s = 1e-30f;
	su.s = s;
	su.lowpass = fc < 1.0;
	su.fc = (float)fc;
	su.f_fx = 0;
	su.s_fx = 0;
	if (su.lowpass) {
		su.f_fx = __double2ull_rn(fc * 9223372036854775808.0);
		su.s_fx = __double2ll_rn(fc * sd * 9223372036854775808.0);
	}
	return su;
}

// Per-sample rotation table of the fc < 1 path: (cos, sin)(j * pi * fc), j = 1 .. 15.
// j = 1,2,3,4,8,12 are evaluated exactly from the fixed-point angle, the rest are one complex
// product of two exact entries (absolute error ~1e-7).
struct RotTable {
	float c[16], s[16];
	__device__ __forceinline__ void build(uint64_t f_fx) {
		c[0] = 1.f; s[0] = 0.f;
#pragma unroll
		for (int j = 1; j <= 4; j++) sincos_fx(f_fx * (uint64_t)j, &s[j], &c[j]);
		sincos_fx(f_fx * 8ull, &s[8], &c[8]);
		sincos_fx(f_fx * 12ull, &s[12], &c[12]);
#pragma unroll
		for (int hi = 4; hi <= 12; hi += 4) {
#pragma unroll
			for (int lo = 1; lo <= 3; lo++) {
				c[hi + lo] = fmaf(c[hi], c[lo], -s[hi] * s[lo]);
				s[hi + lo] = fmaf(s[hi], c[lo], c[hi] * s[lo]);
			}
		}
	}
};

// One block of 16 taps (weight indices 16b .. 16b+15) of one output sample, all CH channels.
// xrow points at the shared-memory sample that pairs with weight index 0 (channel-interleaved).
//   fc == 1: w = c[widx] / (d - s)                      (sinpi(s) is applied once at the end)
//   fc <  1: w = hp[widx] * sin(theta_b + j*pi*fc) / (d - s), theta_b exact per block
template <int CH, bool LOWPASS, bool DESC>
__device__ __forceinline__ void tap_block(const SampleSetup &su, int b, int nt, const float *tab_s,
                                          const RotTable &rot, float sa, float ca, float centre_sn, const float *xrow,
                                          float (&acc)[CH]) {
	const int d0 = 16 * b - nt;
	const bool near = d0 < 16 && d0 > -31;       // block holds a tap with |d| < 16
	const float base = (float)d0 - su.s;         // far blocks: |q| >= 15.5, one rounding is harmless
	const float4 *t4 = reinterpret_cast<const float4 *>(tab_s + 16 * b);
#pragma unroll
	for (int h = 0; h < 2; h++) {
		const int hh = DESC ? 1 - h : h;
		const float4 ta = t4[2 * hh], tb = t4[2 * hh + 1];
		const float coef[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
		float w[8];
		if (!LOWPASS && !near) {
			// far taps of the fc == 1 path: one reciprocal serves two neighbouring taps,
			//   1/q = (q+1) * t,  1/(q+1) = q * t,  t = 1/(q (q+1)),
			// which halves the load on the MUFU pipe (the limiter of this path); |q| >= 15.5 here,
			// so the extra roundings cost ~2e-7 relative on weights below 0.02
#pragma unroll
			for (int u = 0; u < 8; u += 2) {
				const float q = base + (float)(8 * hh + u);
				const float t = rcp_approx(fmaf(q, q, q));
				w[u] = fmaf(coef[u], q, coef[u]) * t;
				w[u + 1] = (coef[u + 1] * q) * t;
			}
		} else
#pragma unroll
		for (int u = 0; u < 8; u++) {
			const int j = 8 * hh + u;
			const float q = near ? (float)(d0 + j) - su.s : base + (float)j;
			float num = coef[u];
			if (LOWPASS) {
				float sn = fmaf(sa, rot.c[j], ca * rot.s[j]);
				// centre tap (|q| <= 0.5): the fixed-point angle has ABSOLUTE accuracy only, but
				// sin(pi fc q) / q needs RELATIVE accuracy as q -> 0
				if (near && d0 == -j) sn = centre_sn;
				num *= sn;
			}
			w[u] = num * rcp_approx(q);
		}
#pragma unroll
		for (int u = 0; u < 8; u++) {
			const int uu = DESC ? 7 - u : u;
			const float *xp = xrow + (16 * b + 8 * hh + uu) * CH;
			if (CH == 1) {
				acc[0] = fmaf(xp[0], w[uu], acc[0]);
			} else if (CH == 2) {
				const float2 v = *reinterpret_cast<const float2 *>(xp);
				acc[0] = fmaf(v.x, w[uu], acc[0]);
				acc[1] = fmaf(v.y, w[uu], acc[1]);
			} else {
#pragma unroll
				for (int c4 = 0; c4 < CH; c4 += 4) {
					const float4 v = *reinterpret_cast<const float4 *>(xp + c4);
					acc[c4] = fmaf(v.x, w[uu], acc[c4]);
					acc[c4 + 1] = fmaf(v.y, w[uu], acc[c4 + 1]);
					acc[c4 + 2] = fmaf(v.z, w[uu], acc[c4 + 2]);
					acc[c4 + 3] = fmaf(v.w, w[uu], acc[c4 + 3]);
				}
			}
		}
	}
}

// All 2*NT taps of an interior output sample whose inputs are staged in shared memory.
// Summation order: the weights decay like 1/|d| away from the centre tap, so each half of the tap
// run is accumulated from its far end towards the centre (small terms first) in its own
// accumulator; a plain left-to-right float32 sum costs ~1e-6 relative at NT >= 128.
template <int CH, bool LOWPASS>
__device__ __forceinline__ void taps_fast(const SampleSetup &su, int nt, int nblk, const float *tab_s,
                                          const float *xrow, float (&out)[CH]) {
	RotTable rot;
	float centre_sn = 0.f, s16 = 0.f, c16 = 1.f;
	if (LOWPASS) {
		rot.build(su.f_fx);
		sincos_fx(su.f_fx * 16ull, &s16, &c16);
		centre_sn = sinpif(su.fc * (0.f - su.s));     // numerator of the d = 0 tap, q = -s
	}
	float accl[CH], accr[CH];
#pragma unroll
	for (int c = 0; c < CH; c++) accl[c] = accr[c] = 0.f;
	const int half = nblk >> 1;
	// Block anchors sin/cos(pi fc (d0 - s)): exact from the fixed-point phase at the far end of each
	// side and at every 4th block counted from the centre (always the block next to the centre);
	// in between, one rotation by 16*pi*fc per block (error <= ~1e-7 per step, at |d| >= 16 only).
	float sal = 0.f, cal = 1.f, sar = 0.f, car = 1.f;
	for (int it = 0; it < half; it++) {
		const int bl = it, br = nblk - 1 - it;
		if (LOWPASS) {
			if (it == 0 || ((half - 1 - it) & 3) == 0) {
				sincos_fx(su.f_fx * (uint64_t)(int64_t)(16 * bl - nt) - (uint64_t)su.s_fx, &sal, &cal);
				sincos_fx(su.f_fx * (uint64_t)(int64_t)(16 * br - nt) - (uint64_t)su.s_fx, &sar, &car);
			} else {
				const float nsl = fmaf(sal, c16, cal * s16), ncl = fmaf(cal, c16, -sal * s16);
				const float nsr = fmaf(sar, c16, -car * s16), ncr = fmaf(car, c16, sar * s16);
				sal = nsl; cal = ncl; sar = nsr; car = ncr;
			}
		}
		tap_block<CH, LOWPASS, false>(su, bl, nt, tab_s, rot, sal, cal, centre_sn, xrow, accl);
		tap_block<CH, LOWPASS, true>(su, br, nt, tab_s, rot, sar, car, centre_sn, xrow, accr);
	}
	if (nblk & 1) {
		if (LOWPASS) sincos_fx(su.f_fx * (uint64_t)(int64_t)(16 * half - nt) - (uint64_t)su.s_fx, &sal, &cal);
		tap_block<CH, LOWPASS, false>(su, half, nt, tab_s, rot, sal, cal, centre_sn, xrow, accl);
	}
#pragma unroll
	for (int c = 0; c < CH; c++) out[c] = accl[c] + accr[c];
}

// Weight of weight-index widx, any sample (edge / fallback path).
__device__ __forceinline__ float weight_single(const SampleSetup &su, int widx, int nt,
                                               const float *__restrict__ ctab, const float *__restrict__ hptab) {
	const int d = widx - nt;
	const float q = (float)d - su.s;
	if (!su.lowpass) return __ldg(ctab + widx) * rcp_approx(q);
	float sn, cs;
	if (d == 0) sn = sinpif(su.fc * q);
	else sincos_fx(su.f_fx * (uint64_t)(int64_t)d - (uint64_t)su.s_fx, &sn, &cs);
	return __ldg(hptab + widx) * sn * rcp_approx(q);
}

// Edge / fallback path: any tap count, samples read from global memory, same summation order.
__device__ __forceinline__ float taps_slow(const SampleSetup &su, int nt, const float *__restrict__ ctab,
                                           const float *__restrict__ hptab, const float *__restrict__ x,
                                           int64_t stride) {
	float accl = 0.f, accr = 0.f;
	const int mid = min(max(nt - su.koff, 0), su.cnt);       // taps [0, mid) lie left of the centre
	for (int k = 0; k < mid; k++)
		accl = fmaf(__ldg(x + (su.lower + k) * stride), weight_single(su, k + su.koff, nt, ctab, hptab), accl);
	for (int k = su.cnt - 1; k >= mid; k--)
		accr = fmaf(__ldg(x + (su.lower + k) * stride), weight_single(su, k + su.koff, nt, ctab, hptab), accr);
	return accl + accr;
}

template <int CH>
__global__ void __launch_bounds__(SINC_TILE, CH >= 8 ? 2 : SINC_MIN_BLOCKS)
sinc_kernel(SincArgs a, const float *__restrict__ ctab, const float *__restrict__ hptab, int nblk, int span_cap) {
	extern __shared__ __align__(16) float smem_f[];
	// [c table | hp table] (16 * (nblk + 1) floats each), then the staged samples:
	// (span_cap + SINC_XPAD) samples x CH, channel-interleaved
	float *ctab_s = smem_f;
	float *hptab_s = smem_f + 16 * (nblk + 1);
	float *xs = smem_f + 32 * (nblk + 1);
	__shared__ long long red_lo[SINC_TILE / 32], red_hi[SINC_TILE / 32];
	const int nt = a.nt;
	for (int i = threadIdx.x; i < 16 * (nblk + 1); i += SINC_TILE) {
		ctab_s[i] = __ldg(ctab + i);
		hptab_s[i] = __ldg(hptab + i);
	}
	const int64_t tiles = (a.out_end - a.out_begin + SINC_TILE - 1) / SINC_TILE;
	const int groups = (a.n_ch + CH - 1) / CH;
	const int64_t work = tiles * groups;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	for (int64_t wk = blockIdx.x; wk < work; wk += gridDim.x) {
		const int grp = (int)(wk / tiles);
		const int64_t tile = wk - (int64_t)grp * tiles;
		const int ch0 = grp * CH;
		const int64_t i = a.out_begin + tile * SINC_TILE + threadIdx.x;
		const bool live = i < a.out_end;

		SampleSetup su;
		su.lower = 0; su.cnt = 0; su.koff = 0; su.s = 1e-30f; su.fc = 1.f; su.lowpass = false; su.f_fx = 0; su.s_fx = 0;
		long long lo = LLONG_MAX, hi = LLONG_MIN;
		if (live) {
			su = sample_setup(a, i);
			if (su.cnt > 0) { lo = su.lower; hi = su.lower + su.cnt; }
		}
		// ---- input span of the tile ----
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
			hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
		}
		if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
		__syncthreads();
		long long tlo = red_lo[0], thi = red_hi[0];
#pragma unroll
		for (int w = 1; w < SINC_TILE / 32; w++) { tlo = min(tlo, red_lo[w]); thi = max(thi, red_hi[w]); }
		const bool any = thi > tlo;
		const bool staged = any && (thi - tlo) <= span_cap;
		if (staged) {
			const int total = ((int)(thi - tlo) + SINC_XPAD) * CH;
			const int valid = (int)(thi - tlo);
			for (int idx = threadIdx.x; idx < total; idx += SINC_TILE) {
				const int e = idx / CH, c = idx - e * CH;
				float v = 0.f;
				if (e < valid && ch0 + c < a.n_ch)
					v = __ldg(a.signal + (int64_t)(ch0 + c) * a.sig_ch_stride + (tlo + e - a.sig_origin) * a.sig_stride);
				xs[idx] = v;
			}
		}
		__syncthreads();

		float acc[CH];
#pragma unroll
		for (int c = 0; c < CH; c++) acc[c] = 0.f;
		// per-sample choice (not per warp): the result of a sample must not depend on which other
		// samples share its warp, i.e. on how a host-pointer call was cut into chunks
		const bool interior = su.cnt == 2 * nt && su.koff == 0;
		if (live && su.cnt > 0) {
			if (staged && interior) {
				const float *xrow = xs + (int)(su.lower - tlo) * CH;
#if defined(SINC_EXPERIMENT_SKIP_TAPS)          // development aid (scripts/try_variants.sh): fixed cost only
				acc[0] = xrow[0] * su.s + (float)su.f_fx;
#elif defined(SINC_EXPERIMENT_ALL_FC1)
				taps_fast<CH, false>(su, nt, nblk, ctab_s, xrow, acc);
#elif defined(SINC_EXPERIMENT_ALL_LOWPASS)
				taps_fast<CH, true>(su, nt, nblk, hptab_s, xrow, acc);
#else
				if (su.lowpass) taps_fast<CH, true>(su, nt, nblk, hptab_s, xrow, acc);
				else taps_fast<CH, false>(su, nt, nblk, ctab_s, xrow, acc);
#endif
			} else {
				// first / last NT outputs of a file, or a span too wide for shared memory
#pragma unroll
				for (int c = 0; c < CH; c++)
					if (ch0 + c < a.n_ch)
						acc[c] = taps_slow(su, nt, ctab, hptab,
						                   a.signal + (int64_t)(ch0 + c) * a.sig_ch_stride - a.sig_origin * a.sig_stride,
						                   a.sig_stride);
			}
			if (!su.lowpass) {
				const float sp = sinpif(su.s);
#pragma unroll
				for (int c = 0; c < CH; c++) acc[c] *= sp;
			}
		}
		if (live) {
#pragma unroll
			for (int c = 0; c < CH; c++)
				if (ch0 + c < a.n_ch) a.out[(int64_t)(ch0 + c) * a.out_ch_stride + (i - a.out_origin) * a.out_stride] = acc[c];
		}
		__syncthreads();
	}
}

template <int CH>
static int launch_sinc_ch(const SincArgs &a, int device, cudaStream_t st, const SincTables &tb) {
	// widest span staged in shared memory: a tile read at up to 4x speed
	const int span_cap = 4 * SINC_TILE + 2 * a.nt;
	const int smem = (CH * (span_cap + SINC_XPAD) + 2 * tb.padded) * (int)sizeof(float);
	auto kern = sinc_kernel<CH>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SINC_TILE, smem));
	if (occ < 1) occ = 1;
	const int64_t tiles = (a.out_end - a.out_begin + SINC_TILE - 1) / SINC_TILE;
	const int64_t work = tiles * ((a.n_ch + CH - 1) / CH);
	int64_t grid = (int64_t)occ * sm_count(device);
	if (grid > work) grid = work;
	if (grid < 1) return PAR_OK;
	kern<<<(unsigned)grid, SINC_TILE, smem, st>>>(a, tb.c, tb.hp, tb.padded / 16 - 1, span_cap);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

int launch_sinc(const SincArgs &a, int device, cudaStream_t st) {
	if (a.m <= 0 || a.n_ch <= 0 || a.out_end <= a.out_begin) return PAR_OK;
	SincTables tb;
	int rc = sinc_tables(device, a.nt, st, &tb);
	if (rc != PAR_OK) return rc;
	// The tap weights of an output sample are computed once per channel group.  Groups of 4 are the
	// sweet spot: a group of 8 needs 128 registers (2 CTAs/SM) and is shared-memory-bandwidth bound
	// just like two groups of 4 -- measured 25 % slower on the 8-channel config ($PAR_B200_SINC_CH8=1
	// selects it for experiments).
	static const bool ch8 = getenv("PAR_B200_SINC_CH8") && atoi(getenv("PAR_B200_SINC_CH8")) > 0;
	if (ch8 && a.n_ch >= 8) return launch_sinc_ch<8>(a, device, st, tb);
	if (a.n_ch >= 4) return launch_sinc_ch<4>(a, device, st, tb);
	if (a.n_ch >= 2) return launch_sinc_ch<2>(a, device, st, tb);
	return launch_sinc_ch<1>(a, device, st, tb);
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
