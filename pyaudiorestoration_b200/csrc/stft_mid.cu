// stft_mid.cu -- STFT for N*Z = 16384 and 32768: too large for the table-in-shared-memory design of
// stft_tma_kernel (the flat twiddle table alone is 98 KB), small enough to keep one whole frame in
// shared memory.  Same contract as stft.cu (util/fourier.py:37-166 of the reference).
//
// The packed complex transform of M = N*Z/2 points is split M = M1 * M2 inside ONE CTA ("four-step"
// without leaving the SM), n = M2*n1 + n2, k = k1 + M1*k2:
//   1. columns: M1-point Stockham transform over n1 for every n2 -- thread t owns column t % M2, so a
//      warp's global loads are 256 contiguous bytes and its shared accesses one contiguous row piece;
//   2. rows: inter-step twiddle W_M^(k1*n2) on load, M2-point transform over n2 for every k1; the
//      last pass writes TRANSPOSED (z'[k2][k1]) so that
//   3. the real-FFT split walks k = k1 + M1*k2 with consecutive lanes on consecutive bins: contiguous
//      shared reads of X[k] and X[M-k], 256-byte coalesced stores.
// Only tiny twiddle tables are needed (two sub-transforms + two-level tables), they stay in L1.
// Shared memory: M1 rows of pad16(M2) float2 (row stride = 64 bytes mod 128: conflict-free for both
// phases) = 69.6 KB (16384) / 139 KB (32768).
#include "fft_core.cuh"
#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

__device__ __forceinline__ int64_t reflect_index_m(int64_t i, int64_t n) {
	if (n == 1) return 0;
	const int64_t period = 2 * (n - 1);
	i %= period;
	if (i < 0) i += period;
	return i < n ? i : period - i;
}

__device__ __forceinline__ float2 tw2m(const float2 *__restrict__ lo, const float2 *__restrict__ hi, int p) {
	return cmul(__ldg(lo + (p & 1023)), __ldg(hi + (p >> 10)));
}

// One in-place Stockham pass of a length-2^LOG2L transform whose element n lives at base[n * STR]
// (STR = row stride for the column phase) or base[pad16(n)] (row phase).  `Load` supplies the pass
// input (global memory for the very first pass), `Store` takes the pass output.
template <int LOG2L, int PASS, class Load, class Store>
__device__ __forceinline__ void mid_pass(int tid, Load load, Store store, const float2 *__restrict__ tw) {
	using S = FftSched<LOG2L>;
	constexpr int RB = S::bits(PASS);
	constexpr int R = 1 << RB;
	constexpr int NSL = S::ns_log2(PASS);
	constexpr int NS = 1 << NSL;
	constexpr int NB = S::P / R;
	constexpr int STRIDE = S::M / R;
	float2 v[NB][R];
#pragma unroll
	for (int b = 0; b < NB; b++) {
		const int j = tid + b * S::TPF;
#pragma unroll
		for (int t = 0; t < R; t++) v[b][t] = load(j + t * STRIDE);
	}
	__syncthreads();
#pragma unroll
	for (int b = 0; b < NB; b++) {
		const int j = tid + b * S::TPF;
		const int k = j & (NS - 1);
		if (PASS > 0) {
			const float2 *twp = tw + S::tw_offset(PASS) + k;
#pragma unroll
			for (int t = 1; t < R; t++) v[b][t] = ctw<false>(v[b][t], __ldg(twp + (t - 1) * NS));
		}
		Dft<R, false>::run(v[b]);
		const int j0 = ((j >> NSL) << (NSL + RB)) + k;
#pragma unroll
		for (int t = 0; t < R; t++) store(j0 + t * NS, v[b][t]);
	}
}

template <int LOG2M1, int LOG2M2>
struct MidCfg {
	static constexpr int M1 = 1 << LOG2M1, M2 = 1 << LOG2M2, M = M1 * M2;
	static constexpr int TPF1 = FftSched<LOG2M1>::TPF, TPF2 = FftSched<LOG2M2>::TPF;
	static constexpr int THREADS = M2 * TPF1;
	static_assert(M2 * TPF1 == M1 * TPF2, "both phases must use every thread");
	static constexpr int RS = M2 + M2 / 16;           // row stride of z[n1][pad16(n2)], float2
	static constexpr int RST = M1 + 2;                // row stride of the transposed result z'[k2][k1]
	static_assert(M2 * RST <= M1 * RS, "the transposed result must fit the same buffer");
	static constexpr int SMEM = M1 * RS * (int)sizeof(float2);
	static constexpr int MIN_BLOCKS = SMEM > 100 * 1024 ? 1 : 2;
	static constexpr int NP1 = FftSched<LOG2M1>::NP, NP2 = FftSched<LOG2M2>::NP;
};

// column-phase passes 1 .. NP1-1 (pass 0 is issued by the kernel: it loads from global memory)
template <int LOG2M1, int PASS, int RS>
struct ColPasses {
	__device__ __forceinline__ static void run(int tid, float2 *col, const float2 *__restrict__ tw) {
		if constexpr (PASS < FftSched<LOG2M1>::NP) {
			__syncthreads();
			mid_pass<LOG2M1, PASS>(tid, [&](int n) { return col[n * RS]; }, [&](int n, float2 v) { col[n * RS] = v; }, tw);
			ColPasses<LOG2M1, PASS + 1, RS>::run(tid, col, tw);
		}
	}
};

// row-phase passes 1 .. NP2-2 in place; the last pass is issued by the kernel (transposed store)
template <int LOG2M2, int PASS>
struct RowPasses {
	__device__ __forceinline__ static void run(int tid, float2 *row, const float2 *__restrict__ tw) {
		if constexpr (PASS < FftSched<LOG2M2>::NP - 1) {
			__syncthreads();
			mid_pass<LOG2M2, PASS>(tid, [&](int n) { return row[pad16(n)]; }, [&](int n, float2 v) { row[pad16(n)] = v; },
			                       tw);
			RowPasses<LOG2M2, PASS + 1>::run(tid, row, tw);
		}
	}
};

template <int LOG2M1, int LOG2M2, bool MAG>
__global__ void __launch_bounds__(MidCfg<LOG2M1, LOG2M2>::THREADS, MidCfg<LOG2M1, LOG2M2>::MIN_BLOCKS)
stft_mid_kernel(StftArgs a, const float2 *__restrict__ tw1, const float2 *__restrict__ tw2t,
                const float2 *__restrict__ t_lo, const float2 *__restrict__ t_hi, const float2 *__restrict__ s_lo,
                const float2 *__restrict__ s_hi, float half_scale) {
	using C = MidCfg<LOG2M1, LOG2M2>;
	using S2 = FftSched<LOG2M2>;
	static_assert(S2::NP >= 2, "row transform needs at least two passes");
	extern __shared__ __align__(16) float2 z[];
	const int t = threadIdx.x;
	// column phase: column n2 = t % M2, sub-thread t / M2
	const int n2 = t & (C::M2 - 1), ctid = t >> LOG2M2;
	float2 *col = z + pad16(n2);
	// row phase: row k1 = t / TPF2, sub-thread t % TPF2
	const int k1 = t / C::TPF2, rtid = t % C::TPF2;
	float2 *row = z + k1 * C::RS;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int half_n = a.n_fft >> 1;

	for (int64_t f = blockIdx.x; f < total; f += gridDim.x) {
		const int64_t ch = f / a.n_frames;
		const int64_t fr = a.frame0 + (f - ch * a.n_frames);
		const float *x = a.x + ch * a.x_ch_stride;
		const int64_t base = fr * a.hop - half_n;
		const bool interior = a.x_stride == 1 && base >= 0 && base + a.n_fft <= a.n &&
		                      ((reinterpret_cast<uintptr_t>(x + (base - a.x_origin)) & 7) == 0);
		// ---- 1. columns (pass 0 reads global memory: window, reflect, pack, zero tail)
		auto gload = [&](int n1) -> float2 {
			const int64_t e2 = 2 * ((int64_t)n1 * C::M2 + n2);
			if (e2 >= a.n_fft) return make_float2(0.f, 0.f);
			const float2 w = __ldg(reinterpret_cast<const float2 *>(a.window + e2));
			float s0, s1;
			if (interior) {
				const float2 v = __ldg(reinterpret_cast<const float2 *>(x + (base + e2 - a.x_origin)));
				s0 = v.x;
				s1 = v.y;
			} else {
				s0 = __ldg(x + (reflect_index_m(base + e2, a.n) - a.x_origin) * a.x_stride);
				s1 = __ldg(x + (reflect_index_m(base + e2 + 1, a.n) - a.x_origin) * a.x_stride);
			}
			return make_float2(s0 * w.x, s1 * w.y);
		};
		// the barrier inside the pass separates the previous frame's epilogue reads from these stores
		mid_pass<LOG2M1, 0>(ctid, gload, [&](int n, float2 v) { col[n * C::RS] = v; }, tw1);
		ColPasses<LOG2M1, 1, C::RS>::run(ctid, col, tw1);
		__syncthreads();
		// ---- 2. rows: inter-step twiddle on the first load, last pass stored transposed
		mid_pass<LOG2M2, 0>(rtid, [&](int n) { return cmul(row[pad16(n)], tw2m(t_lo, t_hi, k1 * n)); },
		                    [&](int n, float2 v) { row[pad16(n)] = v; }, tw2t);
		RowPasses<LOG2M2, 1>::run(rtid, row, tw2t);
		__syncthreads();
		mid_pass<LOG2M2, S2::NP - 1>(rtid, [&](int n) { return row[pad16(n)]; },
		                             [&](int k2, float2 v) { z[k2 * C::RST + k1] = v; }, tw2t);
		__syncthreads();
		// ---- 3. real-FFT split over k = k1 + M1*k2 (z'[k2][k1] is exactly index k with the row padding)
		const int64_t orow = ch * a.out_ch_stride + fr * a.out_pitch;
		for (int k = t; k <= C::M / 2; k += C::THREADS) {
			const int km = (C::M - k) & (C::M - 1);
			const float2 zk = z[(k >> LOG2M1) * C::RST + (k & (C::M1 - 1))];
			const float2 zm = z[(km >> LOG2M1) * C::RST + (km & (C::M1 - 1))];
			const float2 w = tw2m(s_lo, s_hi, k);
			// 2E = zk + conj(zm); 2O = -i (zk - conj(zm)); X[k] = E + W^k O; X[M-k] = conj(E - W^k O)
			const float ex = zk.x + zm.x, ey = zk.y - zm.y;
			const float ox = zk.y + zm.y, oy = zm.x - zk.x;
			const float2 wo = cmul(make_float2(ox, oy), w);
			const float2 xa = make_float2((ex + wo.x) * half_scale, (ey + wo.y) * half_scale);
			const float2 xb = make_float2((ex - wo.x) * half_scale, (wo.y - ey) * half_scale);
			if (MAG) {
				float *o = reinterpret_cast<float *>(a.out) + orow;
				o[k] = cmag(xa);
				if (k != C::M - k) o[C::M - k] = cmag(xb);
			} else {
				float2 *o = reinterpret_cast<float2 *>(a.out) + orow;
				o[k] = xa;
				if (k != C::M - k) o[C::M - k] = xb;
			}
		}
		// next iteration: the barrier inside its first pass orders these reads before the new stores
	}
}

template <int LOG2M1, int LOG2M2, bool MAG>
static int launch_mid_t(const StftArgs &a, int device, cudaStream_t st) {
	using C = MidCfg<LOG2M1, LOG2M2>;
	const float2 *tw1 = fft_twiddles(device, LOG2M1, st);
	const float2 *tw2t = fft_twiddles(device, LOG2M2, st);
	const float2 *big = large_fft_tables(device, LOG2M1 + LOG2M2, st);
	if (!tw1 || !tw2t || !big) return PAR_ECUDA;
	const int64_t n_hi = (int64_t)C::M >> 10;
	auto kern = stft_mid_kernel<LOG2M1, LOG2M2, MAG>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SMEM));
	if (occ < 1) occ = 1;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	int64_t grid = (int64_t)occ * sm_count(device);
	if (grid > total) grid = total;
	if (grid < 1) return PAR_OK;
	const float half_scale = (float)(0.5 / sqrt((double)a.n_fft));
	kern<<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(a, tw1, tw2t, big, big + 1024, big + 1024 + n_hi,
	                                                 big + 2048 + n_hi, half_scale);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

// log2m = log2(N*Z/2) in {13, 14}
int launch_stft_mid(const StftArgs &a, int log2m, int device, cudaStream_t st) {
	if (log2m == 13)
		return a.magnitude ? launch_mid_t<6, 7, true>(a, device, st) : launch_mid_t<6, 7, false>(a, device, st);
	return a.magnitude ? launch_mid_t<7, 7, true>(a, device, st) : launch_mid_t<7, 7, false>(a, device, st);
}

}  // namespace par
