// stft.cu -- fused short-time Fourier transform for sm_100a.
//
// Replaces, in one kernel, the reference's estimate_and_center (np.pad reflect,
// util/fourier.py:78-82), segment_array (frame gather * window, zero tail, :160-166), the
// batched real FFT of pyfftw_rfft2 / np_rfft_pick / torch_rfft2 (:92-157) with its
// 1/sqrt(n_fft) scaling (:105,:157) and, optionally, to_mag (:23-24).
//
// One transform = one frame.  The real frame of length N*Z is packed as M = N*Z/2 complex
// points z[n] = (w[2n] x[2n], w[2n+1] x[2n+1]) (zeros beyond N), transformed with the Stockham
// passes of fft_core.cuh, and split into the N*Z/2+1 one-sided bins on the way out.
// Algorithmic HBM traffic per frame: hop*4 B of new input (overlap re-reads are L1/L2 hits)
// + (N*Z/2+1)*8 B of output (4 B in magnitude mode).
#include "fft_core.cuh"
#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t n) {
	// np.pad(mode='reflect') index map, any pad length
	if (n == 1) return 0;
	const int64_t period = 2 * (n - 1);
	i %= period;
	if (i < 0) i += period;
	return i < n ? i : period - i;
}

struct FrameLoad {
	const float *x;        // channel base = sample `origin`
	int64_t n, stride, base, origin;
	const float2 *win2;    // window as float2 pairs
	int half_n;            // N/2: packed elements that carry samples
	bool fast, valid;
	__device__ __forceinline__ float2 operator()(int e) const {
		if (e >= half_n || !valid) return make_float2(0.f, 0.f);
		const float2 w = __ldg(win2 + e);
		const int64_t s = base + 2 * (int64_t)e;
		float a, b;
		if (fast) {
			const float2 v = __ldg(reinterpret_cast<const float2 *>(x + (s - origin)));
			a = v.x;
			b = v.y;
		} else {
			a = __ldg(x + (reflect_index(s, n) - origin) * stride);
			b = __ldg(x + (reflect_index(s + 1, n) - origin) * stride);
		}
		return make_float2(a * w.x, b * w.y);
	}
};

// Planar copy of a strided channel set (host-pointer calls upload interleaved audio as it lies in
// host memory; the transform kernel's vector loads want unit stride).
__global__ void __launch_bounds__(256)
deinterleave_kernel(const float *__restrict__ src, int64_t n, int64_t stride, int n_ch, int64_t ch_stride,
                    float *__restrict__ dst, int64_t dst_ch_stride) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const float *s = src + i * stride;
		for (int c = 0; c < n_ch; c++) dst[c * dst_ch_stride + i] = __ldg(s + c * ch_stride);
	}
}

int launch_deinterleave(const float *src, int64_t n, int64_t stride, int n_ch, int64_t ch_stride, float *dst,
                        int64_t dst_ch_stride, int device, cudaStream_t st) {
	if (n <= 0 || n_ch <= 0) return PAR_OK;
	int64_t grid = (n + 255) / 256;
	const int64_t cap = (int64_t)sm_count(device) * 16;
	if (grid > cap) grid = cap;
	deinterleave_kernel<<<(unsigned)grid, 256, 0, st>>>(src, n, stride, n_ch, ch_stride, dst, dst_ch_stride);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

template <int LOG2M>
struct StftCfg {
	using S = FftSched<LOG2M>;
	static constexpr bool INPLACE = LOG2M >= 14;
	static constexpr int BLOCK = S::TPF < 128 ? 128 : S::TPF;
	static constexpr int FPB = BLOCK / S::TPF;                      // frames per block pass
	static constexpr int SMEM = FPB * (INPLACE ? 1 : 2) * S::BUF * (int)sizeof(float2);
};

template <int LOG2M, bool MAG>
__global__ void __launch_bounds__(StftCfg<LOG2M>::BLOCK)
stft_kernel(StftArgs a, const float2 *__restrict__ tw, float half_scale) {
	using S = FftSched<LOG2M>;
	using C = StftCfg<LOG2M>;
	extern __shared__ float2 smem[];
	const int slot = threadIdx.x / S::TPF;
	const int tid = threadIdx.x % S::TPF;
	float2 *buf0 = smem + slot * (C::INPLACE ? 1 : 2) * S::BUF;
	float2 *buf1 = C::INPLACE ? buf0 : buf0 + S::BUF;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int64_t groups = (total + C::FPB - 1) / C::FPB;
	const int half_n = a.n_fft >> 1;

	for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
		const int64_t f = g * C::FPB + slot;
		const bool valid = f < total;
		const int64_t ch = valid ? f / a.n_frames : 0;
		const int64_t t = a.frame0 + (valid ? f - ch * a.n_frames : 0);
		FrameLoad ld;
		ld.x = a.x + ch * a.x_ch_stride;
		ld.n = a.n;
		ld.stride = a.x_stride;
		ld.base = t * a.hop - half_n;
		ld.origin = a.x_origin;
		ld.win2 = reinterpret_cast<const float2 *>(a.window);
		ld.half_n = half_n;
		ld.valid = valid;
		ld.fast = a.x_stride == 1 && ld.base >= 0 && ld.base + a.n_fft <= a.n &&
		          ((reinterpret_cast<uintptr_t>(ld.x + (ld.base - ld.origin)) & 7) == 0);

		if (C::INPLACE)
			stockham_pass_inplace<LOG2M, 0, false>(tid, ld, buf0, tw);
		else
			stockham_pass<LOG2M, 0, false>(tid, ld, buf0, tw);
		float2 *res = RunPasses<LOG2M, false, C::INPLACE, 1>::run(tid, buf0, buf1, tw);
		__syncthreads();

		if (valid) {
			const int64_t row = ch * a.out_ch_stride + t * a.out_pitch;
			const float2 *tws = tw + S::TW_SPLIT_OFFSET;
			for (int k = tid; k <= S::M / 2; k += S::TPF) {
				const float2 zk = res[pad16(k)];
				const float2 zm = res[pad16((S::M - k) & (S::M - 1))];
				const float2 w = __ldg(tws + k);
				// 2E = zk + conj(zm); 2O = -i (zk - conj(zm)); X[k] = E + W^k O; X[M-k] = conj(E - W^k O)
				const float ex = zk.x + zm.x, ey = zk.y - zm.y;
				const float ox = zk.y + zm.y, oy = zm.x - zk.x;
				const float2 wo = cmul(make_float2(ox, oy), w);
				const float2 xa = make_float2((ex + wo.x) * half_scale, (ey + wo.y) * half_scale);
				const float2 xb = make_float2((ex - wo.x) * half_scale, (wo.y - ey) * half_scale);
				if (MAG) {
					float *o = reinterpret_cast<float *>(a.out) + row;
					o[k] = cmag(xa);
					if (k != S::M - k) o[S::M - k] = cmag(xb);
				} else {
					float2 *o = reinterpret_cast<float2 *>(a.out) + row;
					o[k] = xa;
					if (k != S::M - k) o[S::M - k] = xb;
				}
			}
		}
		__syncthreads();
	}
}

// ---- TMA-staged variant ---------------------------------------------------------------------
// For 1024 <= N*Z <= 8192 with unit-stride, 16-byte aligned channels and hop % 4 == 0 the raw
// samples of the block's NEXT frame(s) are fetched by one elected thread with a bulk async copy
// (cp.async.bulk, the TMA engine; completion on an mbarrier) while the current frame is being
// transformed, so no warp ever waits on a global load of samples.  The transform runs in place in
// a single padded buffer, which keeps enough of the SM's unified L1 free for the twiddle and
// window tables.  Frames that touch the reflected edges are loaded the slow way, in place.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "LAB_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra LAB_DONE;\n"
	    "bra LAB_WAIT;\n"
	    "LAB_DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

template <bool FULL>     // FULL: zeropad == 1, every packed element carries samples
struct StagedLoad {
	const float2 *raw2;    // staged raw samples of this slot, as float2 pairs (shared memory)
	const float2 *win2;    // window (shared memory)
	int half_n;
	__device__ __forceinline__ float2 operator()(int e) const {
		if (!FULL && e >= half_n) return make_float2(0.f, 0.f);
		const float2 w = win2[e];
		const float2 v = raw2[e];
		return make_float2(v.x * w.x, v.y * w.y);
	}
};

// CTA = SLOTS independent transform slots of TPF threads each, sharing one shared-memory copy of the
// twiddle tables and the window.  Each slot owns an in-place FFT buffer, a staging buffer for the raw
// samples of its next frame, an mbarrier and a named barrier.
template <int LOG2M>
struct StftTmaCfg {
	using S = FftSched<LOG2M>;
	static constexpr int SLOTS = S::TPF >= 128 ? 2 : 256 / S::TPF;
	static constexpr int BLOCK = SLOTS * S::TPF;
	static constexpr int STAGE_FLOATS = 2 * S::M;                         // up to n_fft = N*Z samples
	static constexpr int TW_F2 = (S::TW_TOTAL + 1) & ~1;                  // keep 16-byte alignment behind it
	static constexpr int SMEM = TW_F2 * (int)sizeof(float2) + STAGE_FLOATS * (int)sizeof(float) +
	                            SLOTS * (S::BUF * (int)sizeof(float2) + STAGE_FLOATS * (int)sizeof(float));
	static constexpr int MIN_BLOCKS = SMEM > 110 * 1024 ? 1 : 2;
};

template <int LOG2M, bool MAG>
__global__ void __launch_bounds__(StftTmaCfg<LOG2M>::BLOCK, StftTmaCfg<LOG2M>::MIN_BLOCKS)
stft_tma_kernel(StftArgs a, const float2 *__restrict__ tw, float half_scale) {
	using S = FftSched<LOG2M>;
	using C = StftTmaCfg<LOG2M>;
	extern __shared__ __align__(16) float2 smem[];
	__shared__ __align__(8) uint64_t bars[C::SLOTS];
	const int slot = threadIdx.x / S::TPF;
	const int tid = threadIdx.x % S::TPF;
	float2 *tw_s = smem;
	float *win_s = reinterpret_cast<float *>(smem + C::TW_F2);
	float *slot_base = win_s + C::STAGE_FLOATS + slot * (2 * S::BUF + C::STAGE_FLOATS);
	float *stage = slot_base;                                             // 16-byte aligned
	float2 *buf = reinterpret_cast<float2 *>(slot_base + C::STAGE_FLOATS);
	uint64_t *bar = &bars[slot];
	const SlotSync<S::TPF> sync{slot + 1};
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int64_t stride_f = (int64_t)gridDim.x * C::SLOTS;
	const int half_n = a.n_fft >> 1;
	const uint32_t frame_bytes = (uint32_t)a.n_fft * 4u;

	// ---- one-time: tables to shared memory, barriers
	for (int i = threadIdx.x; i < S::TW_TOTAL; i += C::BLOCK) tw_s[i] = __ldg(tw + i);
	for (int i = threadIdx.x; i < a.n_fft; i += C::BLOCK) win_s[i] = __ldg(a.window + i);
	if (tid == 0) mbar_init(bar, 1);
	if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();

	// can frame f be staged with one bulk copy?  (deterministic: producer and consumers agree)
	auto stageable = [&](int64_t f, const float **src) -> bool {
		if (f >= total) return false;
		const int64_t ch = f / a.n_frames;
		const int64_t t = a.frame0 + (f - ch * a.n_frames);
		const int64_t base = t * a.hop - half_n;
		if (base < 0 || base + a.n_fft > a.n) return false;
		*src = a.x + ch * a.x_ch_stride + (base - a.x_origin);
		return true;
	};
	auto issue = [&](int64_t f) {        // the slot's thread 0: prefetch the raw samples of frame f
		const float *src;
		if (!stageable(f, &src)) return;
		mbar_expect_tx(bar, frame_bytes);
		tma_load_1d(stage, src, frame_bytes, bar);
	};

	int64_t f = (int64_t)blockIdx.x * C::SLOTS + slot;
	if (tid == 0) issue(f);
	uint32_t parity = 0;
	const float2 *tws = tw_s + S::TW_SPLIT_OFFSET;
	for (; f < total; f += stride_f) {
		const int64_t ch = f / a.n_frames;
		const int64_t t = a.frame0 + (f - ch * a.n_frames);
		const float *src_unused;
		const bool staged = stageable(f, &src_unused);
		if (staged) {
			mbar_wait(bar, parity);
			parity ^= 1;
		}
		// pass 0: its slot barrier sits between the loads (staging buffer, and the previous frame's
		// epilogue reads of `buf`) and the stores into `buf`.  `staged` is uniform over the slot.
		if (staged) {
			if (a.zeropad == 1) {
				StagedLoad<true> ld{reinterpret_cast<const float2 *>(stage), reinterpret_cast<const float2 *>(win_s), half_n};
				stockham_pass_slot<LOG2M, 0, false>(tid, ld, buf, tw_s, sync);
			} else {
				StagedLoad<false> ld{reinterpret_cast<const float2 *>(stage), reinterpret_cast<const float2 *>(win_s), half_n};
				stockham_pass_slot<LOG2M, 0, false>(tid, ld, buf, tw_s, sync);
			}
		} else {
			FrameLoad ld;
			ld.x = a.x + ch * a.x_ch_stride;
			ld.n = a.n;
			ld.stride = a.x_stride;
			ld.base = t * a.hop - half_n;
			ld.origin = a.x_origin;
			ld.win2 = reinterpret_cast<const float2 *>(a.window);
			ld.half_n = half_n;
			ld.valid = true;
			ld.fast = false;
			stockham_pass_slot<LOG2M, 0, false>(tid, ld, buf, tw_s, sync);
		}
		// every thread of the slot has consumed its staged samples: prefetch the next frame
		if (tid == 0) issue(f + stride_f);
		RunPassesSlot<LOG2M, false, 1, SlotSync<S::TPF>>::run(tid, buf, tw_s, sync);
		sync();

		const int64_t row = ch * a.out_ch_stride + t * a.out_pitch;
		// bins k and M-k come from z[k], z[M-k]: 2E = zk + conj(zm); 2O = -i (zk - conj(zm));
		// X[k] = E + W^k O; X[M-k] = conj(E - W^k O).  k = 0 pairs with itself (bins 0 and M).
		auto split_store = [&](int k, bool both) {
			const float2 zk = buf[pad16(k)];
			const float2 zm = buf[pad16((S::M - k) & (S::M - 1))];
			const float2 w = tws[k];
			const float ex = zk.x + zm.x, ey = zk.y - zm.y;
			const float ox = zk.y + zm.y, oy = zm.x - zk.x;
			const float2 wo = cmul(make_float2(ox, oy), w);
			const float2 xa = make_float2((ex + wo.x) * half_scale, (ey + wo.y) * half_scale);
			const float2 xb = make_float2((ex - wo.x) * half_scale, (wo.y - ey) * half_scale);
			if (MAG) {
				float *o = reinterpret_cast<float *>(a.out) + row;
				o[k] = cmag(xa);
				if (both) o[S::M - k] = cmag(xb);
			} else {
				float2 *o = reinterpret_cast<float2 *>(a.out) + row;
				o[k] = xa;
				if (both) o[S::M - k] = xb;
			}
		};
		if constexpr ((S::M / 2) % S::TPF == 0) {
#pragma unroll
			for (int i = 0; i < (S::M / 2) / S::TPF; i++) split_store(tid + i * S::TPF, true);
			if (tid == 0) split_store(S::M / 2, false);
		} else {
			for (int k = tid; k <= S::M / 2; k += S::TPF) split_store(k, k != S::M - k);
		}
	}
}

template <int LOG2M, bool MAG>
static int launch_tma(const StftArgs &a, int device, cudaStream_t st) {
	using C = StftTmaCfg<LOG2M>;
	const float2 *tw = fft_twiddles(device, LOG2M, st);
	if (!tw) return PAR_ECUDA;
	auto kern = stft_tma_kernel<LOG2M, MAG>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, C::SMEM));
	if (occ < 1) occ = 1;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int64_t groups = (total + C::SLOTS - 1) / C::SLOTS;
	int64_t grid = (int64_t)occ * sm_count(device);
	if (grid > groups) grid = groups;
	if (grid < 1) return PAR_OK;
	const float half_scale = (float)(0.5 / sqrt((double)a.n_fft));
	kern<<<(unsigned)grid, C::BLOCK, C::SMEM, st>>>(a, tw, half_scale);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

static bool tma_eligible(const StftArgs &a) {
	return a.x_stride == 1 && (a.hop & 3) == 0 && ((a.n_fft >> 1) & 3) == 0 && (a.x_ch_stride & 3) == 0 &&
	       (a.x_origin & 3) == 0 &&
	       (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
}

template <int LOG2M, bool MAG>
static int launch_one(const StftArgs &a, int device, cudaStream_t st) {
	using C = StftCfg<LOG2M>;
	const float2 *tw = fft_twiddles(device, LOG2M, st);
	if (!tw) return PAR_ECUDA;
	auto kern = stft_kernel<LOG2M, MAG>;
	PAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
	int occ = 0;
	PAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, C::SMEM));
	if (occ < 1) occ = 1;
	const int64_t total = (int64_t)a.n_ch * a.n_frames;
	const int64_t groups = (total + C::FPB - 1) / C::FPB;
	int64_t grid = (int64_t)occ * sm_count(device);
	if (grid > groups) grid = groups;
	if (grid < 1) return PAR_OK;
	const float half_scale = (float)(0.5 / sqrt((double)a.n_fft));
	kern<<<(unsigned)grid, C::BLOCK, C::SMEM, st>>>(a, tw, half_scale);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

template <bool MAG>
static int dispatch(const StftArgs &a, int log2m, int device, cudaStream_t st) {
	switch (log2m) {
	case 4: return launch_one<4, MAG>(a, device, st);
	case 5: return launch_one<5, MAG>(a, device, st);
	case 6: return launch_one<6, MAG>(a, device, st);
	case 7: return launch_one<7, MAG>(a, device, st);
	case 8: return launch_one<8, MAG>(a, device, st);
	case 9: return tma_eligible(a) ? launch_tma<9, MAG>(a, device, st) : launch_one<9, MAG>(a, device, st);
	case 10: return tma_eligible(a) ? launch_tma<10, MAG>(a, device, st) : launch_one<10, MAG>(a, device, st);
	case 11: return tma_eligible(a) ? launch_tma<11, MAG>(a, device, st) : launch_one<11, MAG>(a, device, st);
	case 12: return tma_eligible(a) ? launch_tma<12, MAG>(a, device, st) : launch_one<12, MAG>(a, device, st);
	case 13:
	case 14: return launch_stft_mid(a, log2m, device, st);     // whole frame in shared memory, two-level
	}
	set_error("stft: n_fft*zeropad must be a power of two in [32, 32768]");
	return PAR_EUNSUPPORTED;
}

int launch_stft(const StftArgs &a, int device, cudaStream_t st) {
	const int64_t nz = (int64_t)a.n_fft * a.zeropad;
	int log2m = -1;
	for (int b = 4; b <= 19; b++)
		if (nz == (int64_t)2 << b) log2m = b;
	if (log2m < 0 || (a.n_fft & 1)) {
		set_error("stft: n_fft*zeropad must be a power of two in [32, 1048576] (n_fft even)");
		return PAR_EUNSUPPORTED;
	}
	if (log2m >= 15) {
		if (a.x_stride != 1) {
			set_error("stft: transforms above 32768 points need unit-stride channels");
			return PAR_EUNSUPPORTED;
		}
		return launch_stft_large(a, log2m, device, st);
	}
	return a.magnitude ? dispatch<true>(a, log2m, device, st) : dispatch<false>(a, log2m, device, st);
}

}  // namespace par
