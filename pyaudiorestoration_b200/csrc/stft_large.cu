// stft_large.cu -- STFT for transforms too large for one SM's shared memory (65536 <= N*Z <= 1048576).
//
// Same contract as stft.cu (util/fourier.py:37-166 of the reference; the GUIs offer FFT sizes up to
// 1048576, util/widgets.py:333-335, and humspeed_gui.py:18-22 uses 524288).  The packed complex
// transform of M = N*Z/2 points is split M = M1 * M2 ("four-step"), n = M2*n1 + n2, k = k1 + M1*k2:
//
//   kernel A  (columns): window + reflect + pack on load, M1-point transform over n1 for a tile of 16
//             adjacent n2, inter-step twiddle W_M^(k1*n2), store to scratch[k1][n2];
//   kernel B  (rows):    M2-point transform over n2 for 16 adjacent rows k1 AND their 16 mirror rows
//             M1-k1, so that X[k] and X[M-k] meet in shared memory; real-FFT split, 1/sqrt(N) scale,
//             optional magnitude, store -- 128-byte runs of adjacent bins.
//
// Frames are processed in batches whose scratch (8*M bytes per frame) stays inside the 126 MB L2,
// so DRAM sees each sample once and each output bin once, like the single-kernel path.
#include "fft_core.cuh"
#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

constexpr int LG_CB = 16;          // adjacent columns (kernel A) / rows (kernel B) per CTA
// slots are S::BUF + 1 float2 apart: lanes that read the same index of 16 different slots (the
// coalesced epilogues below) then fall into 16 different bank pairs
template <int LOG2M> constexpr int slot_stride() { return FftSched<LOG2M>::BUF + 1; }

struct BlockSync {
	__device__ __forceinline__ void operator()() const { __syncthreads(); }
};

__device__ __forceinline__ int64_t reflect_index_l(int64_t i, int64_t n) {
	if (n == 1) return 0;
	const int64_t period = 2 * (n - 1);
	i %= period;
	if (i < 0) i += period;
	return i < n ? i : period - i;
}

// W^p from a two-level table: tab_lo[p & 1023] * tab_hi[p >> 10]
__device__ __forceinline__ float2 tw2(const float2 *__restrict__ lo, const float2 *__restrict__ hi, int p) {
	return cmul(__ldg(lo + (p & 1023)), __ldg(hi + (p >> 10)));
}

struct LargeTables {
	const float2 *t_lo, *t_hi;     // W_M^p
	const float2 *s_lo, *s_hi;     // W_2M^k
};

struct ColumnLoad {
	const float *x;        // channel base = sample `origin`
	const float *win;
	int64_t n, base, origin;
	int n_fft, m2, n2;
	bool interior;
	__device__ __forceinline__ float2 operator()(int n1) const {
		const int64_t e = (int64_t)n1 * m2 + n2;          // packed element
		const int64_t k = 2 * e;
		if (k >= n_fft) return make_float2(0.f, 0.f);
		const float2 w = __ldg(reinterpret_cast<const float2 *>(win + k));
		float a, b;
		if (interior) {
			const float2 v = __ldg(reinterpret_cast<const float2 *>(x + (base + k - origin)));
			a = v.x;
			b = v.y;
		} else {
			a = __ldg(x + (reflect_index_l(base + k, n) - origin));
			b = __ldg(x + (reflect_index_l(base + k + 1, n) - origin));
		}
		return make_float2(a * w.x, b * w.y);
	}
};

template <int LOG2M1>
__global__ void __launch_bounds__(LG_CB * FftSched<LOG2M1>::TPF)
stft_large_cols_kernel(StftArgs a, int log2m2, int64_t batch0, const float2 *__restrict__ tw1, LargeTables lt,
                       float2 *__restrict__ scratch) {
	using S = FftSched<LOG2M1>;
	extern __shared__ __align__(16) float2 smem[];
	const int slot = threadIdx.x / S::TPF, tid = threadIdx.x % S::TPF;
	float2 *buf = smem + slot * slot_stride<LOG2M1>();
	const int m2 = 1 << log2m2;
	const int tiles = m2 / LG_CB;
	const int64_t fb = blockIdx.x / tiles;                 // frame within the batch (all channels)
	const int n2 = (int)(blockIdx.x - fb * tiles) * LG_CB + slot;
	const int64_t f = batch0 + fb;
	const int64_t ch = f / a.n_frames;
	const int64_t t = a.frame0 + (f - ch * a.n_frames);
	ColumnLoad ld;
	ld.x = a.x + ch * a.x_ch_stride;
	ld.win = a.window;
	ld.n = a.n;
	ld.base = t * a.hop - (a.n_fft >> 1);
	ld.origin = a.x_origin;
	ld.n_fft = a.n_fft;
	ld.m2 = m2;
	ld.n2 = n2;
	ld.interior = a.x_stride == 1 && ld.base >= 0 && ld.base + a.n_fft <= a.n &&
	              ((reinterpret_cast<uintptr_t>(ld.x + (ld.base - ld.origin)) & 7) == 0);
	const BlockSync sync;
	stockham_pass_slot<LOG2M1, 0, false>(tid, ld, buf, tw1, sync);
	RunPassesSlot<LOG2M1, false, 1, BlockSync>::run(tid, buf, tw1, sync);
	__syncthreads();
	// store with 16 consecutive lanes on 16 adjacent columns: 128-byte runs of scratch[k1][n2..n2+15]
	const int col = threadIdx.x & (LG_CB - 1);
	const int n2c = n2 - slot + col;
	float2 *dst = scratch + ((size_t)fb << (LOG2M1 + log2m2)) + n2c;
	const float2 *src = smem + col * slot_stride<LOG2M1>();
	for (int k1 = threadIdx.x / LG_CB; k1 < S::M; k1 += S::TPF) {
		const float2 w = tw2(lt.t_lo, lt.t_hi, k1 * n2c);
		dst[(size_t)k1 << log2m2] = cmul(src[pad16(k1)], w);
	}
}

struct RowLoad {
	const float2 *row;
	__device__ __forceinline__ float2 operator()(int n2) const { return __ldg(row + n2); }
};

template <int LOG2M2, bool MAG>
__global__ void __launch_bounds__(2 * LG_CB * FftSched<LOG2M2>::TPF)
stft_large_rows_kernel(StftArgs a, int log2m1, int64_t batch0, const float2 *__restrict__ tw2t, LargeTables lt,
                       const float2 *__restrict__ scratch, float half_scale) {
	using S = FftSched<LOG2M2>;
	extern __shared__ __align__(16) float2 smem[];
	const int slot = threadIdx.x / S::TPF, tid = threadIdx.x % S::TPF;     // slots 0..15: rows, 16..31: mirrors
	float2 *buf = smem + slot * slot_stride<LOG2M2>();
	const int m1 = 1 << log2m1;
	const int tiles = m1 / (2 * LG_CB) + 1;                // k1 tiles 0, 16, ..., m1/2
	const int64_t fb = blockIdx.x / tiles;
	const int k1_0 = (int)(blockIdx.x - fb * tiles) * LG_CB;
	const int64_t f = batch0 + fb;
	const int64_t ch = f / a.n_frames;
	const int64_t t = a.frame0 + (f - ch * a.n_frames);
	const int k1s = k1_0 + (slot & (LG_CB - 1));           // row of the slot's pair
	const int k1 = slot < LG_CB ? (k1s & (m1 - 1)) : ((m1 - k1s) & (m1 - 1));
	RowLoad ld{scratch + ((size_t)fb << (LOG2M2 + log2m1)) + ((size_t)k1 << LOG2M2)};
	const BlockSync sync;
	stockham_pass_slot<LOG2M2, 0, false>(tid, ld, buf, tw2t, sync);
	RunPassesSlot<LOG2M2, false, 1, BlockSync>::run(tid, buf, tw2t, sync);
	__syncthreads();
	// Epilogue with 16 consecutive lanes on 16 adjacent rows k1: bins k = k1 + m1*k2 of one k2 are
	// contiguous in the output, so every 16-lane group writes a 128-byte run (and its mirror run).
	// X[k] pairs with X[M-k] = mirror row, column m2-1-k2 (row 0: m2-k2, wrapping).
	const int r = threadIdx.x & (LG_CB - 1);
	const int k1e = k1_0 + r;
	if (k1e > m1 / 2) return;
	const int64_t M = (int64_t)1 << (LOG2M2 + log2m1);
	const float2 *own = smem + r * slot_stride<LOG2M2>();
	const float2 *mir = smem + (r + LG_CB) * slot_stride<LOG2M2>();
	const int64_t row = ch * a.out_ch_stride + t * a.out_pitch;
	for (int k2 = threadIdx.x / LG_CB; k2 < S::M; k2 += 2 * S::TPF) {
		const int64_t k = k1e + ((int64_t)k2 << log2m1);
		const int k2m = k1e == 0 ? ((S::M - k2) & (S::M - 1)) : (S::M - 1 - k2);
		const float2 zk = own[pad16(k2)];
		const float2 zm = mir[pad16(k2m)];
		const float2 w = tw2(lt.s_lo, lt.s_hi, (int)k);
		const float ex = zk.x + zm.x, ey = zk.y - zm.y;
		const float ox = zk.y + zm.y, oy = zm.x - zk.x;
		const float2 wo = cmul(make_float2(ox, oy), w);
		const float2 xa = make_float2((ex + wo.x) * half_scale, (ey + wo.y) * half_scale);
		const float2 xb = make_float2((ex - wo.x) * half_scale, (wo.y - ey) * half_scale);
		if (MAG) {
			float *o = reinterpret_cast<float *>(a.out) + row;
			o[k] = cmag(xa);
			if (k != M - k) o[M - k] = cmag(xb);
		} else {
			float2 *o = reinterpret_cast<float2 *>(a.out) + row;
			o[k] = xa;
			if (k != M - k) o[M - k] = xb;
		}
	}
}

template <int LOG2M1>
static int launch_cols(const StftArgs &a, int log2m2, int64_t batch0, int64_t nb, const float2 *tw1,
                       const LargeTables &lt, float2 *scratch, cudaStream_t st) {
	using S = FftSched<LOG2M1>;
	auto kern = stft_large_cols_kernel<LOG2M1>;
	const int smem = LG_CB * slot_stride<LOG2M1>() * (int)sizeof(float2);
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	const int64_t grid = nb * ((1 << log2m2) / LG_CB);
	kern<<<(unsigned)grid, LG_CB * S::TPF, smem, st>>>(a, log2m2, batch0, tw1, lt, scratch);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

template <int LOG2M2, bool MAG>
static int launch_rows(const StftArgs &a, int log2m1, int64_t batch0, int64_t nb, const float2 *tw2t,
                       const LargeTables &lt, const float2 *scratch, cudaStream_t st) {
	using S = FftSched<LOG2M2>;
	auto kern = stft_large_rows_kernel<LOG2M2, MAG>;
	const int smem = 2 * LG_CB * slot_stride<LOG2M2>() * (int)sizeof(float2);
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	const int64_t grid = nb * ((1 << log2m1) / (2 * LG_CB) + 1);
	const float half_scale = (float)(0.5 / sqrt((double)a.n_fft));
	kern<<<(unsigned)grid, 2 * LG_CB * S::TPF, smem, st>>>(a, log2m1, batch0, tw2t, lt, scratch, half_scale);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

// log2m = log2(N*Z/2) in [15, 19]
int launch_stft_large(const StftArgs &a, int log2m, int device, cudaStream_t st) {
	const int log2m2 = log2m >= 18 ? 9 : 8;
	const int log2m1 = log2m - log2m2;                     // 7 .. 10
	const float2 *tw1 = fft_twiddles(device, log2m1, st);
	const float2 *tw2t = fft_twiddles(device, log2m2, st);
	const float2 *big = large_fft_tables(device, log2m, st);
	if (!tw1 || !tw2t || !big) return PAR_ECUDA;
	const int64_t M = (int64_t)1 << log2m;
	LargeTables lt;
	lt.t_lo = big;
	lt.t_hi = big + 1024;
	lt.s_lo = lt.t_hi + (M >> 10);
	lt.s_hi = lt.s_lo + 1024;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	int64_t batch = (48ll << 20) / (M * (int64_t)sizeof(float2));
	if (batch < 1) batch = 1;
	if (batch > total) batch = total;
	float2 *scratch = nullptr;
	PAR_CUDA(cudaMallocAsync((void **)&scratch, (size_t)batch * M * sizeof(float2), st));
	int rc = PAR_OK;
	for (int64_t b0 = 0; b0 < total && rc == PAR_OK; b0 += batch) {
		const int64_t nb = b0 + batch < total ? batch : total - b0;
		switch (log2m1) {
		case 7: rc = launch_cols<7>(a, log2m2, b0, nb, tw1, lt, scratch, st); break;
		case 8: rc = launch_cols<8>(a, log2m2, b0, nb, tw1, lt, scratch, st); break;
		case 9: rc = launch_cols<9>(a, log2m2, b0, nb, tw1, lt, scratch, st); break;
		case 10: rc = launch_cols<10>(a, log2m2, b0, nb, tw1, lt, scratch, st); break;
		default: set_error("stft: unsupported large size"); rc = PAR_EUNSUPPORTED;
		}
		if (rc != PAR_OK) break;
		if (log2m2 == 8)
			rc = a.magnitude ? launch_rows<8, true>(a, log2m1, b0, nb, tw2t, lt, scratch, st)
			                 : launch_rows<8, false>(a, log2m1, b0, nb, tw2t, lt, scratch, st);
		else
			rc = a.magnitude ? launch_rows<9, true>(a, log2m1, b0, nb, tw2t, lt, scratch, st)
			                 : launch_rows<9, false>(a, log2m1, b0, nb, tw2t, lt, scratch, st);
	}
	cudaFreeAsync(scratch, st);
	return rc;
}

}  // namespace par
