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
#include <stdlib.h>

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

// ---- fused variant: inverse transform + overlap-add in shared memory, no scratch round trip -----------------
// A slot (TPF threads) walks a run of consecutive frames of one channel in ascending order and adds each
// windowed frame into a circular buffer of N floats; after frame t has been added the samples
// [t*hop, (t+1)*hop) are complete (no later frame covers them), so they are divided by the window sum-square
// and stored, and their slots are cleared.  A run starts with ceil(N/hop) - 1 warm-up frames whose output belongs
// to the previous run (their contributions are needed, their samples are not stored).  Every output sample is
// accumulated in ascending frame order from 0, exactly like istft_ola_kernel, so both paths give identical bits;
// HBM traffic is the spectrogram once (+ warm-up frames) and the audio once.
template <int LOG2M>
struct IstftFusedCfg {
	using S = FftSched<LOG2M>;
	static constexpr int BLOCK = S::TPF < 128 ? 128 : S::TPF;
	static constexpr int SLOTS = BLOCK / S::TPF;
	static constexpr int SLOT_FLOATS = 2 * S::BUF + 2 * S::M;          // FFT buffer + circular overlap-add buffer
	static constexpr int SMEM = SLOTS * SLOT_FLOATS * (int)sizeof(float);
};

template <int LOG2M>
__global__ void __launch_bounds__(IstftFusedCfg<LOG2M>::BLOCK)
istft_fused_kernel(IstftArgs a, const float2 *__restrict__ tw, float scale, int64_t run_len, int64_t runs_per_ch) {
	using S = FftSched<LOG2M>;
	using C = IstftFusedCfg<LOG2M>;
	constexpr int N = 2 * S::M;
	extern __shared__ __align__(16) float smem_f[];
	const int slot = threadIdx.x / S::TPF;
	const int tid = threadIdx.x % S::TPF;
	float2 *buf = reinterpret_cast<float2 *>(smem_f + slot * C::SLOT_FLOATS);
	float *ola = smem_f + slot * C::SLOT_FLOATS + 2 * S::BUF;
	const SlotSync<S::TPF> sync{slot + 1};
	const float2 *tws = tw + S::TW_SPLIT_OFFSET;
	const float2 *win2 = reinterpret_cast<const float2 *>(a.window);
	const int64_t total = (int64_t)a.n_ch * runs_per_ch;
	const int64_t ov = (N + a.hop - 1) / a.hop;
	const int64_t timeline = N + (int64_t)a.hop * (a.n_frames - 1);

	for (int64_t unit = (int64_t)blockIdx.x * C::SLOTS + slot; unit < total; unit += (int64_t)gridDim.x * C::SLOTS) {
		const int64_t ch = unit / runs_per_ch;
		const int64_t own0 = (unit - ch * runs_per_ch) * run_len;
		const int64_t own1 = own0 + run_len < a.n_frames ? own0 + run_len : a.n_frames;
		const int64_t first = own0 - (ov - 1) > 0 ? own0 - (ov - 1) : 0;
		for (int r = tid; r < N; r += S::TPF) ola[r] = 0.f;
		float *ych = a.y + ch * a.y_ch_stride;
		for (int64_t t = first; t < own1; t++) {
			// fold the one-sided spectrum into the M-point complex spectrum of z[n] = x[2n] + i x[2n+1]
			const float2 *row = a.S + ch * a.s_ch_stride + t * a.s_pitch;
			for (int k = tid; k <= S::M / 2; k += S::TPF) {
				float2 xk = __ldg(row + k), xm = __ldg(row + S::M - k);
				if (k == 0) {   // irfft ignores the imaginary parts of the DC and Nyquist bins
					xk.y = 0.f;
					xm.y = 0.f;
				}
				const float2 w = __ldg(tws + k);
				const float2 e = make_float2(xk.x + xm.x, xk.y - xm.y);
				const float2 d = make_float2(xk.x - xm.x, xk.y + xm.y);
				const float2 o = ctw<true>(d, w);   // d * conj(W^k)
				buf[pad16(k)] = make_float2(e.x - o.y, e.y + o.x);
				if (k != 0 && k != S::M - k) buf[pad16(S::M - k)] = make_float2(e.x + o.y, o.x - e.y);
			}
			RunPassesSlot<LOG2M, true, 0, SlotSync<S::TPF>>::run(tid, buf, tw, sync);
			sync();
			// overlap-add: sample r of frame t lands in slot (t*hop + r) mod N
			const int base = (int)((t * a.hop) & (N - 1));
			for (int n = tid; n < S::M; n += S::TPF) {
				const float2 z = buf[pad16(n)];
				const float2 w = __ldg(win2 + n);
				const int i0 = (base + 2 * n) & (N - 1), i1 = (base + 2 * n + 1) & (N - 1);
				ola[i0] += z.x * scale * w.x;
				ola[i1] += z.y * scale * w.y;
			}
			sync();
			// the samples this frame completes: [t*hop, (t+1)*hop), or up to the end of the timeline after the last frame
			const int64_t j0 = t * a.hop;
			const int64_t j1 = t == a.n_frames - 1 ? timeline : j0 + a.hop;
			const bool emit = t >= own0;
			for (int64_t j = j0 + tid; j < j1; j += S::TPF) {
				const int idx = (int)(j & (N - 1));
				const float acc = ola[idx];
				ola[idx] = 0.f;
				const int64_t o = j - a.start;
				if (emit && o >= 0 && o < a.length) {
					int64_t t_hi = j / a.hop;
					if (t_hi > a.n_frames - 1) t_hi = a.n_frames - 1;
					const int64_t t_lo = j >= N ? (j - N) / a.hop + 1 : 0;
					float wss = 0.f;
					for (int64_t q = t_lo; q <= t_hi; q++) {
						const float w = __ldg(a.window + (j - q * a.hop));
						wss = fmaf(w, w, wss);
					}
					ych[o * a.y_stride] = wss > 1.17549435e-38f ? acc / wss : acc;
				}
			}
			sync();
		}
	}
}

// zero fill of the outputs past the end of the timeline (length longer than the frames cover)
__global__ void __launch_bounds__(256)
istft_tail_kernel(IstftArgs a, int64_t o_begin) {
	const int64_t span = a.length - o_begin;
	const int64_t total = span * a.n_ch;
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
		const int64_t ch = i / span, o = o_begin + (i - ch * span);
		a.y[ch * a.y_ch_stride + o * a.y_stride] = 0.f;
	}
}

template <int LOG2M>
static int launch_fused(const IstftArgs &a, int device, cudaStream_t st) {
	using C = IstftFusedCfg<LOG2M>;
	using S = FftSched<LOG2M>;
	const float2 *tw = fft_twiddles(device, LOG2M, st);
	if (!tw) return PAR_ECUDA;
	auto kern = istft_fused_kernel<LOG2M>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, C::SMEM));
	if (occ < 1) occ = 1;
	// runs: long enough that the warm-up frames stay below ~12 % of the work, short enough to fill the machine
	const int64_t ov = (2 * S::M + a.hop - 1) / a.hop;
	const int64_t slots = (int64_t)occ * sm_count(device) * C::SLOTS;
	int64_t run_len = ((int64_t)a.n_ch * a.n_frames + slots - 1) / slots;
	if (run_len < 8 * (ov - 1)) run_len = 8 * (ov - 1);
	if (run_len < 1) run_len = 1;
	const int64_t runs_per_ch = (a.n_frames + run_len - 1) / run_len;
	const int64_t units = runs_per_ch * a.n_ch;
	int64_t grid = (units + C::SLOTS - 1) / C::SLOTS;
	if (grid > (int64_t)occ * sm_count(device)) grid = (int64_t)occ * sm_count(device);
	if (grid < 1) return PAR_OK;
	const float scale = (float)(1.0 / sqrt((double)a.n_fft));
	kern<<<(unsigned)grid, C::BLOCK, C::SMEM, st>>>(a, tw, scale, run_len, runs_per_ch);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	const int64_t timeline = (int64_t)a.n_fft + (int64_t)a.hop * (a.n_frames - 1);
	const int64_t covered = timeline - a.start > 0 ? timeline - a.start : 0;
	if (a.length > covered) {
		const int64_t total = (a.length - covered) * a.n_ch;
		int64_t g = (total + 255) / 256;
		if (g > (int64_t)sm_count(device) * 16) g = (int64_t)sm_count(device) * 16;
		istft_tail_kernel<<<(unsigned)g, 256, 0, st>>>(a, covered);
		count_launch();
		PAR_CUDA(cudaGetLastError());
	}
	return PAR_OK;
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

static bool istft_two_pass_forced() {
	static const bool v = getenv("PAR_B200_ISTFT_TWO_PASS") && atoi(getenv("PAR_B200_ISTFT_TWO_PASS")) > 0;
	return v;
}

// does launch_istft need IstftArgs::frames (n_ch * n_frames * n_fft floats of scratch)?
bool istft_needs_scratch(int n_fft) { return istft_two_pass_forced() || n_fft > 8192; }

int launch_istft(const IstftArgs &a, int device, cudaStream_t st) {
	int log2m = -1;
	for (int b = 4; b <= 14; b++)
		if (a.n_fft == (2 << b)) log2m = b;
	int rc;
	// fused inverse transform + overlap-add (no scratch) up to 8192 points (above, a slot's serial frame chain is slower
	// than the two-pass pair: measured 3.0 vs 2.6 ms at 16384); $PAR_B200_ISTFT_TWO_PASS=1 selects the
	// frames-to-scratch + gather pair for every size (tests compare the two: identical bits)
	const bool two_pass = istft_two_pass_forced();
	if (!two_pass && log2m >= 4 && log2m <= 12 && a.n_frames > 0 && a.length > 0) {
		switch (log2m) {
		case 4: return launch_fused<4>(a, device, st);
		case 5: return launch_fused<5>(a, device, st);
		case 6: return launch_fused<6>(a, device, st);
		case 7: return launch_fused<7>(a, device, st);
		case 8: return launch_fused<8>(a, device, st);
		case 9: return launch_fused<9>(a, device, st);
		case 10: return launch_fused<10>(a, device, st);
		case 11: return launch_fused<11>(a, device, st);
		case 12: return launch_fused<12>(a, device, st);
		}
	}
	if (!a.frames) { set_error("istft: internal error (no scratch for the two-pass path)"); return PAR_EINVAL; }
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
