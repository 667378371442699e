// istft.cu -- inverse short-time Fourier transform for sm_100a.
//
// Replaces util/fourier.py:314-437 (istft): de-normalisation by sqrt(n_fft) (:359), per-frame
// irfft * window (:401), overlap-add (__overlap_add, :677-687), division by the window
// sum-square envelope (window_sumsquare, :481-546, :409-417) and the centre trim (:419-435).
//
// Kernel 1: one transform per frame.  The one-sided spectrum is folded into the M = N/2 point
// complex spectrum of z[n] = x[2n] + i x[2n+1], inverse Stockham passes (fft_core.cuh), window,
// store the windowed frame to scratch.
// Kernel 2: gather-form overlap-add: every output sample sums the <= ceil(N/hop) frame samples
// that cover it in ascending frame order (the reference's accumulation order), accumulates the
// window sum-square in the same loop and divides.
#include "fft_core.cuh"
#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

template <int LOG2M>
struct IstftCfg {
	using S = FftSched<LOG2M>;
	static constexpr bool INPLACE = LOG2M >= 14;
	static constexpr int BLOCK = S::TPF < 128 ? 128 : S::TPF;
	static constexpr int FPB = BLOCK / S::TPF;
	static constexpr int SMEM = FPB * (INPLACE ? 1 : 2) * S::BUF * (int)sizeof(float2);
};

template <int LOG2M>
__global__ void __launch_bounds__(IstftCfg<LOG2M>::BLOCK)
istft_frames_kernel(IstftArgs a, const float2 *__restrict__ tw, float scale) {
	using S = FftSched<LOG2M>;
	using C = IstftCfg<LOG2M>;
	extern __shared__ float2 smem[];
	const int slot = threadIdx.x / S::TPF;
	const int tid = threadIdx.x % S::TPF;
	float2 *buf0 = smem + slot * (C::INPLACE ? 1 : 2) * S::BUF;
	float2 *buf1 = C::INPLACE ? buf0 : buf0 + S::BUF;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int64_t groups = (total + C::FPB - 1) / C::FPB;
	const float2 *tws = tw + S::TW_SPLIT_OFFSET;
	const float2 *win2 = reinterpret_cast<const float2 *>(a.window);

	for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
		const int64_t f = g * C::FPB + slot;
		const bool valid = f < total;
		const int64_t ch = valid ? f / a.n_frames : 0;
		const int64_t t = valid ? f - ch * a.n_frames : 0;
		const float2 *row = a.S + ch * a.s_ch_stride + t * a.s_pitch;
		for (int k = tid; k <= S::M / 2; k += S::TPF) {
			float2 xk = make_float2(0.f, 0.f), xm = xk;
			if (valid) {
				xk = __ldg(row + k);
				xm = __ldg(row + S::M - k);
			}
			if (k == 0) {   // irfft ignores the imaginary parts of the DC and Nyquist bins
				xk.y = 0.f;
				xm.y = 0.f;
			}
			const float2 w = __ldg(tws + k);
			const float2 e = make_float2(xk.x + xm.x, xk.y - xm.y);
			const float2 d = make_float2(xk.x - xm.x, xk.y + xm.y);
			const float2 o = ctw<true>(d, w);   // d * conj(W^k)
			buf0[pad16(k)] = make_float2(e.x - o.y, e.y + o.x);
			if (k != 0 && k != S::M - k) buf0[pad16(S::M - k)] = make_float2(e.x + o.y, o.x - e.y);
		}
		float2 *res = RunPasses<LOG2M, true, C::INPLACE, 0>::run(tid, buf0, buf1, tw);
		__syncthreads();
		if (valid) {
			float2 *dst = reinterpret_cast<float2 *>(a.frames + f * (int64_t)(2 * S::M));
			for (int n = tid; n < S::M; n += S::TPF) {
				const float2 z = res[pad16(n)];
				const float2 w = __ldg(win2 + n);
				dst[n] = make_float2(z.x * scale * w.x, z.y * scale * w.y);
			}
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256)
istft_ola_kernel(IstftArgs a) {
	const int64_t per_ch = a.length;
	const int64_t total = per_ch * a.n_ch;
	const int64_t n = a.n_fft;
	const int64_t timeline = n + (int64_t)a.hop * (a.n_frames - 1);
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
	     i += (int64_t)gridDim.x * blockDim.x) {
		const int64_t ch = i / per_ch;
		const int64_t o = i - ch * per_ch;
		const int64_t j = a.start + o;
		float y = 0.f;
		if (j < timeline) {
			int64_t t_hi = j / a.hop;
			if (t_hi > a.n_frames - 1) t_hi = a.n_frames - 1;
			const int64_t t_lo = j >= n ? (j - n) / a.hop + 1 : 0;
			const float *fr = a.frames + ch * a.n_frames * n;
			float acc = 0.f, wss = 0.f;
			for (int64_t t = t_lo; t <= t_hi; t++) {
				const int64_t r = j - t * a.hop;
				acc += __ldg(fr + t * n + r);
				const float w = __ldg(a.window + r);
				wss = fmaf(w, w, wss);
			}
			y = wss > 1.17549435e-38f ? acc / wss : acc;
		}
		a.y[ch * a.y_ch_stride + o * a.y_stride] = y;
	}
}

template <int LOG2M>
static int launch_frames(const IstftArgs &a, int device, cudaStream_t st) {
	using C = IstftCfg<LOG2M>;
	const float2 *tw = fft_twiddles(device, LOG2M, st);
	if (!tw) return PAR_ECUDA;
	auto kern = istft_frames_kernel<LOG2M>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, C::SMEM));
	if (occ < 1) occ = 1;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int64_t groups = (total + C::FPB - 1) / C::FPB;
	int64_t grid = (int64_t)occ * sm_count(device);
	if (grid > groups) grid = groups;
	if (grid < 1) return PAR_OK;
	// S * sqrt(N) (util/fourier.py:359), 1/M of the inverse transform, 1/2 of the fold
	const float scale = (float)(1.0 / sqrt((double)a.n_fft));
	kern<<<(unsigned)grid, C::BLOCK, C::SMEM, st>>>(a, tw, scale);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

int launch_istft(const IstftArgs &a, int device, cudaStream_t st) {
	int log2m = -1;
	for (int b = 4; b <= 14; b++)
		if (a.n_fft == (2 << b)) log2m = b;
	int rc;
	switch (log2m) {
	case 4: rc = launch_frames<4>(a, device, st); break;
	case 5: rc = launch_frames<5>(a, device, st); break;
	case 6: rc = launch_frames<6>(a, device, st); break;
	case 7: rc = launch_frames<7>(a, device, st); break;
	case 8: rc = launch_frames<8>(a, device, st); break;
	case 9: rc = launch_frames<9>(a, device, st); break;
	case 10: rc = launch_frames<10>(a, device, st); break;
	case 11: rc = launch_frames<11>(a, device, st); break;
	case 12: rc = launch_frames<12>(a, device, st); break;
	case 13: rc = launch_frames<13>(a, device, st); break;
	case 14: rc = launch_frames<14>(a, device, st); break;
	default:
		set_error("istft: n_fft must be a power of two in [32, 32768]");
		return PAR_EUNSUPPORTED;
	}
	if (rc != PAR_OK) return rc;
	const int64_t total = a.length * a.n_ch;
	if (total > 0) {
		int64_t grid = (total + 255) / 256;
		const int64_t cap = (int64_t)sm_count(device) * 16;
		if (grid > cap) grid = cap;
		istft_ola_kernel<<<(unsigned)grid, 256, 0, st>>>(a);
		count_launch();
		PAR_CUDA(cudaGetLastError());
	}
	return PAR_OK;
}

}  // namespace par
