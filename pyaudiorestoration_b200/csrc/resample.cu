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
#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

// ------------------------------------------------------------------------------------------
// positions
// ------------------------------------------------------------------------------------------

// v_j and the running sum exactly as np.arange(n)/(n-1)*(s1-s0)+s0 and np.cumsum(1/v) evaluate
// them (util/resampling.py:120,125): IEEE double ops, no FMA contraction.
__device__ __forceinline__ double seg_speed(int64_t j, double nm1, double ds, double s0) {
	return __dadd_rn(__dmul_rn(__ddiv_rn((double)j, nm1), ds), s0);
}

__global__ void __launch_bounds__(128)
segment_sums_kernel(const double *__restrict__ speeds, const int64_t *__restrict__ seg_n,
                    int64_t n_seg, double *__restrict__ sums) {
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n_seg) return;
	const int64_t n = seg_n[i];
	const double s0 = speeds[i];
	const double ds = __dsub_rn(speeds[i + 1], s0);
	const double nm1 = (double)(n - 1);
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
                        int64_t n_seg, double *__restrict__ pos, int64_t m) {
	__shared__ double tile[EXP_WARPS][32][33];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	int64_t start = 0, n = 0;
	double s0 = 1.0, ds = 0.0, nm1 = 1.0, off = 0.0;
	if (i < n_seg) {
		start = seg_start[i];
		n = seg_n[i];
		if (n < 0 || start >= m) n = 0;
		if (start + n > m) n = m - start;
		s0 = speeds[i];
		ds = __dsub_rn(speeds[i + 1], s0);
		nm1 = (double)(seg_n[i] - 1);
		off = seg_off[i];
	}
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
				acc = __dadd_rn(acc, r[jj]);
				tile[warp][lane][jj] = __dadd_rn(acc, off);
			}
		}
		__syncwarp();
		for (int row = 0; row < 32; row++) {
			const int64_t rn = __shfl_sync(0xffffffffu, n, row);
			const int64_t rs = __shfl_sync(0xffffffffu, start, row);
			if (j0 + lane < rn) pos[rs + j0 + lane] = tile[warp][row][lane];
		}
		__syncwarp();
	}
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

int launch_expand_positions(const double *speeds_dev, const int64_t *seg_n_dev,
                            const int64_t *seg_start_dev, const double *seg_offset_dev,
                            int64_t n_seg, double *pos_dev, int64_t m, cudaStream_t st) {
	if (n_seg <= 0 || m <= 0) return PAR_OK;
	const int tpb = 32 * EXP_WARPS;
	expand_positions_kernel<<<(unsigned)((n_seg + tpb - 1) / tpb), tpb, 0, st>>>(
	    speeds_dev, seg_n_dev, seg_start_dev, seg_offset_dev, n_seg, pos_dev, m);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

// ------------------------------------------------------------------------------------------
// windowed-sinc interpolation
// ------------------------------------------------------------------------------------------

constexpr int SINC_TILE = 256;       // outputs per block iteration == threads per block
constexpr int SINC_XPAD = 32;        // zero padding behind the staged span

__device__ __forceinline__ float rcp_approx(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// sin/cos of pi * (phase / 2^63), phase a 64-bit fixed-point angle (wraps at one full turn)
__device__ __forceinline__ void sincos_fx(uint64_t phase, float *s, float *c) {
	const int32_t top = (int32_t)(phase >> 32);
	sincospif((float)top * 4.656612873077393e-10f /* 2^-31 */, s, c);
}

struct SampleSetup {
	int64_t lower;     // first input sample of the tap run
	int cnt;           // number of taps (0 .. 2NT)
	int koff;          // weight index of tap 0 (0 unless PAR_SINC_ALIGNED_EDGES at the start edge)
	float s;           // fractional shift p - round(p), never exactly 0
	bool lowpass;      // fc < 1
	float fc;          // fc rounded to float32 (centre tap only)
	uint64_t f_fx;     // fc in units of 2^-63 half-turns
	int64_t s_fx;      // fc * s in the same units
};

// Weights w[widx] for widx = 16*b .. 16*b+15 of one sample.
template <bool LOWPASS>
struct BlockWeights {
	// fc == 1: w = c[widx] * sinpi(s) / (d - s); sinpi(s) is applied once at the end
	// fc <  1: w = hp[widx] * sin(theta_b + j*pi*fc) / (d - s)
	__device__ __forceinline__ static void run(const SampleSetup &su, int b, int nt,
	                                           const float *__restrict__ tab, const float *cj,
	                                           const float *sj, float (&w)[16]) {
		const int d0 = 16 * b - nt;
		float sa = 0.f, ca = 0.f;
		if (LOWPASS) sincos_fx(su.f_fx * (uint64_t)(int64_t)d0 - (uint64_t)su.s_fx, &sa, &ca);
		const float4 *t4 = reinterpret_cast<const float4 *>(tab + 16 * b);
		float coef[16];
#pragma unroll
		for (int v = 0; v < 4; v++) {
			const float4 t = t4[v];
			coef[4 * v] = t.x; coef[4 * v + 1] = t.y; coef[4 * v + 2] = t.z; coef[4 * v + 3] = t.w;
		}
		const bool far = d0 >= 16 || d0 + 15 <= -16;
		const float base = (float)d0 - su.s;      // only used when every |q| of the block is >= 15.5
#pragma unroll
		for (int j = 0; j < 16; j++) {
			const float q = far ? base + (float)j : (float)(d0 + j) - su.s;
			float num = coef[j];
			if (LOWPASS) {
				// centre tap (|q| <= 0.5): the fixed-point angle has ABSOLUTE accuracy only, but the
				// weight sin(pi fc q) / q needs RELATIVE accuracy as q -> 0
				if (d0 == -j) num *= sinpif(su.fc * q);
				else num *= fmaf(sa, cj[j], ca * sj[j]);
			}
			w[j] = num * rcp_approx(q);
		}
	}
};

template <int CH, bool LOWPASS, bool FAST, class XLoad>
__device__ __forceinline__ void sinc_taps(const SampleSetup &su, int nt, int nblk,
                                          const float *__restrict__ tab, XLoad xload,
                                          float (&acc)[CH]) {
	float cj[16], sj[16];
	if (LOWPASS) {
#pragma unroll
		for (int j = 0; j < 16; j++) sincos_fx(su.f_fx * (uint64_t)j, &sj[j], &cj[j]);
	}
	// Summation order: the weights decay like 1/|d| away from the centre tap, so each half of the
	// tap run is accumulated from its far end towards the centre (small terms first) in its own
	// accumulator; a plain left-to-right float32 sum costs ~1e-6 relative at NT >= 128.
	float accl[CH], accr[CH];
#pragma unroll
	for (int c = 0; c < CH; c++) accl[c] = accr[c] = 0.f;
	const int half = nblk >> 1;
	for (int it = 0; it < nblk; it++) {
		const bool right = it & 1;
		const int b = right ? nblk - 1 - (it >> 1) : (it >> 1);
		// odd block count: the middle block is visited last, on the left accumulator
		float w[16];
		BlockWeights<LOWPASS>::run(su, b, nt, tab, cj, sj, w);
		if (right && b >= half) {
#pragma unroll
			for (int j = 15; j >= 0; j--) {
				const int k = 16 * b + j - su.koff;
				if (FAST) {
#pragma unroll
					for (int c = 0; c < CH; c++) accr[c] = fmaf(xload(c, k), w[j], accr[c]);
				} else {
					const bool on = k >= 0 && k < su.cnt;
#pragma unroll
					for (int c = 0; c < CH; c++) accr[c] = fmaf(on ? xload(c, k) : 0.f, on ? w[j] : 0.f, accr[c]);
				}
			}
		} else {
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const int k = 16 * b + j - su.koff;      // tap number; its sample is lower + k
				if (FAST) {
#pragma unroll
					for (int c = 0; c < CH; c++) accl[c] = fmaf(xload(c, k), w[j], accl[c]);
				} else {
					const bool on = k >= 0 && k < su.cnt;
#pragma unroll
					for (int c = 0; c < CH; c++) accl[c] = fmaf(on ? xload(c, k) : 0.f, on ? w[j] : 0.f, accl[c]);
				}
			}
		}
	}
#pragma unroll
	for (int c = 0; c < CH; c++) acc[c] = accl[c] + accr[c];
}

struct SmemX {
	const float *xs;
	int plane, off;   // plane stride, offset of tap 0's sample
	__device__ __forceinline__ float operator()(int c, int k) const { return xs[c * plane + off + k]; }
};
struct GlobalX {
	const float *x;   // first channel of the group
	int64_t ch_stride, stride, lower;
	__device__ __forceinline__ float operator()(int c, int k) const {
		return __ldg(x + c * ch_stride + (lower + k) * stride);
	}
};

template <int CH>
__global__ void __launch_bounds__(SINC_TILE)
sinc_kernel(SincArgs a, const float *__restrict__ ctab, const float *__restrict__ hptab, int nblk,
            int span_cap) {
	extern __shared__ float xs[];        // CH planes of span_cap + SINC_XPAD floats
	__shared__ long long red_lo[SINC_TILE / 32], red_hi[SINC_TILE / 32];
	__shared__ long long tile_lo, tile_hi;
	const int plane = span_cap + SINC_XPAD;
	const int nt = a.nt;
	const int64_t tiles = (a.m + SINC_TILE - 1) / SINC_TILE;
	const int groups = (a.n_ch + CH - 1) / CH;
	const int64_t work = tiles * groups;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	for (int64_t wk = blockIdx.x; wk < work; wk += gridDim.x) {
		const int grp = (int)(wk / tiles);
		const int64_t tile = wk - (int64_t)grp * tiles;
		const int ch0 = grp * CH;
		const int64_t i = tile * SINC_TILE + threadIdx.x;
		const bool live = i < a.m;

		// ---- per-sample setup in float64 (util/resampling.py:67-84) ----
		SampleSetup su;
		su.lower = 0; su.cnt = 0; su.koff = 0; su.s = 1e-30f; su.lowpass = false; su.fc = 1.f; su.f_fx = 0; su.s_fx = 0;
		long long lo = LLONG_MAX, hi = LLONG_MIN;
		if (live) {
			const double p = a.pos[i];
			double per;
			if (i + 1 < a.m) per = fmax(1e-12, a.pos[i + 1] - p);
			else per = a.m >= 2 ? fmax(1e-12, a.pos[a.m - 1] - a.pos[a.m - 2]) : 0.0;
			double fc = 1.0 / per;
			if (!(fc < 1.0)) fc = 1.0;
			double pr = rint(p);
			if (!(pr > -9.0e15)) pr = -9.0e15;      // NaN / -inf guard (garbage in, zeros out)
			if (pr > 9.0e15) pr = 9.0e15;
			const long long ind = (long long)pr;
			const double sd = p - pr;
			long long lower = ind - nt, upper = ind + nt;
			if (lower < 0) lower = 0;
			if (upper > a.n_in) upper = a.n_in;
			su.lower = lower;
			su.cnt = upper > lower ? (int)(upper - lower) : 0;
			if (a.aligned_edges) su.koff = (int)(lower - (ind - nt));
			float s = (float)sd;
			if (s == 0.f) s = 1e-30f;
			su.s = s;
			su.lowpass = fc < 1.0;
			if (su.lowpass) {
				su.fc = (float)fc;
				su.f_fx = __double2ull_rn(fc * 9223372036854775808.0);
				su.s_fx = __double2ll_rn(fc * sd * 9223372036854775808.0);
			}
			if (su.cnt > 0) { lo = lower; hi = upper; }
		}
		// ---- tile span ----
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
			hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
		}
		if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
		__syncthreads();
		if (threadIdx.x == 0) {
			long long l = red_lo[0], h = red_hi[0];
			for (int w = 1; w < SINC_TILE / 32; w++) { l = min(l, red_lo[w]); h = max(h, red_hi[w]); }
			tile_lo = l; tile_hi = h;
		}
		__syncthreads();
		const long long tlo = tile_lo, thi = tile_hi;
		const bool any = thi > tlo;
		const bool staged = any && (thi - tlo) <= span_cap;
		if (staged) {
			const int span = (int)(thi - tlo);
			for (int c = 0; c < CH; c++) {
				const bool chv = ch0 + c < a.n_ch;
				const float *src = a.signal + (int64_t)(ch0 + c) * a.sig_ch_stride;
				float *dst = xs + c * plane;
				for (int e = threadIdx.x; e < span + SINC_XPAD; e += SINC_TILE)
					dst[e] = (chv && e < span) ? __ldg(src + (tlo + e) * a.sig_stride) : 0.f;
			}
		}
		__syncthreads();

		float acc[CH];
#pragma unroll
		for (int c = 0; c < CH; c++) acc[c] = 0.f;
		if (any) {
			const bool interior = su.cnt == 2 * nt && su.koff == 0;
			const bool warp_fast = staged && __all_sync(0xffffffffu, interior || !live);
			const float *tab = su.lowpass ? hptab : ctab;
			if (staged) {
				SmemX xl{xs, plane, (int)(su.lower - tlo)};
				if (!live || su.cnt == 0) {
					// nothing
				} else if (warp_fast) {
					if (su.lowpass) sinc_taps<CH, true, true>(su, nt, nblk, tab, xl, acc);
					else sinc_taps<CH, false, true>(su, nt, nblk, tab, xl, acc);
				} else {
					if (su.lowpass) sinc_taps<CH, true, false>(su, nt, nblk, tab, xl, acc);
					else sinc_taps<CH, false, false>(su, nt, nblk, tab, xl, acc);
				}
			} else if (live && su.cnt > 0) {
				// span too wide for shared memory (wildly non-monotone positions): read through L1/L2
				for (int c = 0; c < CH; c++) {
					if (ch0 + c >= a.n_ch) break;
					GlobalX xl{a.signal + (int64_t)(ch0 + c) * a.sig_ch_stride, 0, a.sig_stride, su.lower};
					float one[1];
					if (su.lowpass) sinc_taps<1, true, false>(su, nt, nblk, tab, xl, one);
					else sinc_taps<1, false, false>(su, nt, nblk, tab, xl, one);
					acc[c] = one[0];
				}
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
				if (ch0 + c < a.n_ch) a.out[(int64_t)(ch0 + c) * a.out_ch_stride + i * a.out_stride] = acc[c];
		}
		__syncthreads();
	}
}

template <int CH>
static int launch_sinc_ch(const SincArgs &a, int device, cudaStream_t st, const SincTables &tb) {
	// widest span staged in shared memory: a tile read at up to 4x speed
	const int span_cap = 4 * SINC_TILE + 2 * a.nt;
	const int smem = CH * (span_cap + SINC_XPAD) * (int)sizeof(float);
	auto kern = sinc_kernel<CH>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SINC_TILE, smem));
	if (occ < 1) occ = 1;
	const int64_t tiles = (a.m + SINC_TILE - 1) / SINC_TILE;
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
	if (a.m <= 0 || a.n_ch <= 0) return PAR_OK;
	SincTables tb;
	int rc = sinc_tables(device, a.nt, st, &tb);
	if (rc != PAR_OK) return rc;
	if (a.n_ch >= 2) return launch_sinc_ch<2>(a, device, st, tb);
	return launch_sinc_ch<1>(a, device, st, tb);
}

// ------------------------------------------------------------------------------------------
// linear interpolation (np.interp(sample_at, arange(L), signal, left=0, right=0))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_kernel(SincArgs a) {
	const int64_t total = a.m * a.n_ch;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
	     t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t ch = t / a.m, i = t - ch * a.m;
		const double p = a.pos[i];
		const float *x = a.signal + ch * a.sig_ch_stride;
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
		a.out[ch * a.out_ch_stride + i * a.out_stride] = y;
	}
}

int launch_linear(const SincArgs &a, int device, cudaStream_t st) {
	const int64_t total = a.m * a.n_ch;
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
