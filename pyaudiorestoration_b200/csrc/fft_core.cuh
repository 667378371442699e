// fft_core.cuh -- register/shared-memory Stockham FFT building blocks for sm_100a.
//
// A length-M complex transform (M = 2^LOG2M, 16 <= M <= 16384) is run by M/16 threads, each
// holding 16 points in registers per pass.  Passes are radix 2^b with b chosen so that the
// LOG2M bits are split as evenly as possible over ceil(LOG2M/4) passes (e.g. 2048 = 16*16*8).
// Between passes the points go through a padded shared-memory buffer (Stockham autosort: reads
// at stride M/R are contiguous across lanes, the scatter on the write side is made
// conflict-free by one float2 of padding per 16).
//
// Twiddles come from a table generated on the host in float64 and rounded once to float32
// (no in-kernel recurrences: the 1e-6 parity bound of the path leaves no room for them).
#pragma once
#include <cuda_runtime.h>

namespace par {

__host__ __device__ constexpr int imin(int a, int b) { return a < b ? a : b; }

template <int LOG2M>
struct FftSched {
	static constexpr int M = 1 << LOG2M;
	static constexpr int NP = (LOG2M + 3) / 4;            // number of passes
	static constexpr int P = imin(16, M);                 // points per thread
	static constexpr int TPF = M / P;                     // threads per transform
	__host__ __device__ static constexpr int bits(int p) {
		return LOG2M / NP + (p < LOG2M % NP ? 1 : 0);
	}
	__host__ __device__ static constexpr int ns_log2(int p) {   // log2 of product of earlier radices
		int s = 0;
		for (int q = 0; q < p; q++) s += bits(q);
		return s;
	}
	// twiddle table: pass p >= 1 holds (R-1)*Ns entries laid out [t-1][k]
	__host__ __device__ static constexpr int tw_offset(int p) {
		int s = 0;
		for (int q = 1; q < p; q++) s += ((1 << bits(q)) - 1) << ns_log2(q);
		return s;
	}
	static constexpr int TW_PASS_TOTAL = tw_offset(NP);
	// followed by the real-FFT split table W_{2M}^k, k = 0 .. M/2
	static constexpr int TW_SPLIT_OFFSET = TW_PASS_TOTAL;
	static constexpr int TW_TOTAL = TW_PASS_TOTAL + M / 2 + 1;
	static constexpr int BUF = M + M / 16;                // padded buffer length (float2)
};

__device__ __forceinline__ int pad16(int a) { return a + (a >> 4); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
	return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// to_mag (util/fourier.py:23-24): |z| + 1e-7.  One MUFU square root (relative error 2^-23) instead of the
// IEEE-rounded sequence: the magnitude epilogue must not cost more than the complex one it halves.
__device__ __forceinline__ float cmag(float2 z) {
	float r;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(z.x, z.x, z.y * z.y)));
	return r + 1e-7f;
}
// a * w for forward transforms, a * conj(w) for inverse ones (tables hold forward twiddles)
template <bool INV>
__device__ __forceinline__ float2 ctw(float2 a, float2 w) {
	if (INV) return make_float2(fmaf(a.x, w.x, a.y * w.y), fmaf(a.y, w.x, -a.x * w.y));
	return cmul(a, w);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
	return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
	float2 t = a;
	a = cadd(t, b);
	b = csub(t, b);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
	float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
	float2 t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
	a0 = cadd(t0, t2);
	a1 = cadd(t1, t3);
	a2 = csub(t0, t2);
	a3 = csub(t1, t3);
}

// multiply by W8^1 / W8^3 (forward) or their conjugates (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_w8_1(float2 a) {
	const float h = 0.70710678118654752440f;
	return INV ? make_float2((a.x - a.y) * h, (a.x + a.y) * h)
	           : make_float2((a.x + a.y) * h, (a.y - a.x) * h);
}
template <bool INV>
__device__ __forceinline__ float2 mul_w8_3(float2 a) {
	const float h = 0.70710678118654752440f;
	return INV ? make_float2((-a.x - a.y) * h, (a.x - a.y) * h)
	           : make_float2((a.y - a.x) * h, (-a.x - a.y) * h);
}

template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<2, INV> {
	__device__ __forceinline__ static void run(float2 *v) { dft2<INV>(v[0], v[1]); }
};
template <bool INV>
struct Dft<4, INV> {
	__device__ __forceinline__ static void run(float2 *v) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV>
struct Dft<8, INV> {
	__device__ __forceinline__ static void run(float2 *v) {
		float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
		float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
		dft4<INV>(e0, e1, e2, e3);
		dft4<INV>(o0, o1, o2, o3);
		o1 = mul_w8_1<INV>(o1);
		o2 = mul_mi<INV>(o2);
		o3 = mul_w8_3<INV>(o3);
		v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
		v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
		v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
		v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
	}
};
template <bool INV>
struct Dft<16, INV> {
	// n = n1 + 4*n2, k = 4*k1 + k2:  X[4k1+k2] = sum_n1 W4^(n1 k1) W16^(n1 k2) DFT4_n2(x[n1+4n2])[k2]
	__device__ __forceinline__ static void run(float2 *v) {
		const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin(pi/8)
		const float h = 0.70710678118654752440f;
		float2 y[4][4];
#pragma unroll
		for (int n1 = 0; n1 < 4; n1++) {
			y[n1][0] = v[n1]; y[n1][1] = v[n1 + 4]; y[n1][2] = v[n1 + 8]; y[n1][3] = v[n1 + 12];
			dft4<INV>(y[n1][0], y[n1][1], y[n1][2], y[n1][3]);
		}
		// W16^m = (cos(pi m/8), -sin(pi m/8)); table for m = 1,2,3,4,6,9
		y[1][1] = ctw<INV>(y[1][1], make_float2(c1, -s1));        // m=1
		y[1][2] = mul_w8_1<INV>(y[1][2]);                          // m=2
		y[1][3] = ctw<INV>(y[1][3], make_float2(s1, -c1));        // m=3
		y[2][1] = mul_w8_1<INV>(y[2][1]);                          // m=2
		y[2][2] = mul_mi<INV>(y[2][2]);                            // m=4
		y[2][3] = mul_w8_3<INV>(y[2][3]);                          // m=6
		y[3][1] = ctw<INV>(y[3][1], make_float2(s1, -c1));        // m=3
		y[3][2] = mul_w8_3<INV>(y[3][2]);                          // m=6
		y[3][3] = ctw<INV>(y[3][3], make_float2(-c1, s1));        // m=9
		(void)h;
#pragma unroll
		for (int k2 = 0; k2 < 4; k2++) {
			dft4<INV>(y[0][k2], y[1][k2], y[2][k2], y[3][k2]);
			v[k2] = y[0][k2]; v[4 + k2] = y[1][k2]; v[8 + k2] = y[2][k2]; v[12 + k2] = y[3][k2];
		}
	}
};

// One Stockham pass for the calling thread (lane id `tid` in [0, TPF) of its transform).
//   Load(n)  -> float2 element n of the pass input (natural index, 0 <= n < M)
//   dst      -> padded shared buffer of the pass output
template <int LOG2M, int PASS, bool INV, class Load>
__device__ __forceinline__ void stockham_pass(int tid, Load load, float2 *dst,
                                              const float2 *__restrict__ tw) {
	using S = FftSched<LOG2M>;
	constexpr int RB = S::bits(PASS);
	constexpr int R = 1 << RB;
	constexpr int NSL = S::ns_log2(PASS);
	constexpr int NS = 1 << NSL;
	constexpr int NB = S::P / R;           // butterflies per thread
	constexpr int STRIDE = S::M / R;
	float2 v[NB][R];
#pragma unroll
	for (int b = 0; b < NB; b++) {
		const int j = tid + b * S::TPF;
#pragma unroll
		for (int t = 0; t < R; t++) v[b][t] = load(j + t * STRIDE);
	}
#pragma unroll
	for (int b = 0; b < NB; b++) {
		const int j = tid + b * S::TPF;
		const int k = j & (NS - 1);
		if (PASS > 0) {
			const float2 *twp = tw + S::tw_offset(PASS) + k;
#pragma unroll
			for (int t = 1; t < R; t++) v[b][t] = ctw<INV>(v[b][t], __ldg(twp + (t - 1) * NS));
		}
		Dft<R, INV>::run(v[b]);
		const int j0 = ((j >> NSL) << (NSL + RB)) + k;
#pragma unroll
		for (int t = 0; t < R; t++) dst[pad16(j0 + t * NS)] = v[b][t];
	}
}

// Same pass, but reading and writing the SAME buffer (single-buffer mode for the largest
// sizes): all loads, barrier, all stores.
template <int LOG2M, int PASS, bool INV, class Load>
__device__ __forceinline__ void stockham_pass_inplace(int tid, Load load, float2 *buf,
                                                      const float2 *__restrict__ tw) {
	using S = FftSched<LOG2M>;
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
			for (int t = 1; t < R; t++) v[b][t] = ctw<INV>(v[b][t], __ldg(twp + (t - 1) * NS));
		}
		Dft<R, INV>::run(v[b]);
		const int j0 = ((j >> NSL) << (NSL + RB)) + k;
#pragma unroll
		for (int t = 0; t < R; t++) buf[pad16(j0 + t * NS)] = v[b][t];
	}
}

struct SmemLoad {
	const float2 *src;
	__device__ __forceinline__ float2 operator()(int n) const { return src[pad16(n)]; }
};

// Sub-block ("slot") barrier: the TPF threads that share one transform synchronise among
// themselves only (named barrier `id`, or a warp barrier when a transform fits one warp), so several
// transforms in a CTA progress independently.
template <int TPF>
struct SlotSync {
	int id;
	__device__ __forceinline__ void operator()() const {
		if (TPF == 32) {
			__syncwarp();
		} else if (TPF < 32) {
			// several slots share a warp and may run different numbers of frames: synchronise this slot's lanes only
			const unsigned lane = threadIdx.x & 31u;
			const unsigned mask = ((TPF >= 32 ? 0u : (1u << (TPF & 31))) - 1u) << (lane & ~(unsigned)(TPF - 1));
			__syncwarp(mask);
		} else {
			asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TPF) : "memory");
		}
	}
};

// In-place Stockham pass with the twiddle table in SHARED memory and a slot barrier.
template <int LOG2M, int PASS, bool INV, class Load, class Sync>
__device__ __forceinline__ void stockham_pass_slot(int tid, Load load, float2 *buf, const float2 *tw_s, Sync sync) {
	using S = FftSched<LOG2M>;
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
	sync();
#pragma unroll
	for (int b = 0; b < NB; b++) {
		const int j = tid + b * S::TPF;
		const int k = j & (NS - 1);
		if (PASS > 0) {
			const float2 *twp = tw_s + S::tw_offset(PASS) + k;
#pragma unroll
			for (int t = 1; t < R; t++) v[b][t] = ctw<INV>(v[b][t], twp[(t - 1) * NS]);
		}
		Dft<R, INV>::run(v[b]);
		const int j0 = ((j >> NSL) << (NSL + RB)) + k;
#pragma unroll
		for (int t = 0; t < R; t++) buf[pad16(j0 + t * NS)] = v[b][t];
	}
}

template <int LOG2M, bool INV, int PASS, class Sync>
struct RunPassesSlot {
	__device__ __forceinline__ static void run(int tid, float2 *buf, const float2 *tw_s, Sync sync) {
		using S = FftSched<LOG2M>;
		if constexpr (PASS < S::NP) {
			sync();   // stores of the previous pass -> loads of this one
			stockham_pass_slot<LOG2M, PASS, INV>(tid, SmemLoad{buf}, buf, tw_s, sync);
			RunPassesSlot<LOG2M, INV, PASS + 1, Sync>::run(tid, buf, tw_s, sync);
		}
	}
};


// Runs passes FIRST..NP-1 of a transform whose pass-(FIRST-1) output (or input, FIRST == 0)
// sits in `cur`; ping-pongs between cur and alt (or works in place).  Returns the buffer that
// holds the natural-order result.  Ends WITHOUT a trailing barrier after the last pass's stores.
template <int LOG2M, bool INV, bool INPLACE, int PASS>
struct RunPasses {
	__device__ __forceinline__ static float2 *run(int tid, float2 *cur, float2 *alt,
	                                              const float2 *__restrict__ tw) {
		using S = FftSched<LOG2M>;
		if constexpr (PASS >= S::NP) {
			return cur;
		} else {
			__syncthreads();   // producer of `cur` -> this pass's loads
			if constexpr (INPLACE) {
				stockham_pass_inplace<LOG2M, PASS, INV>(tid, SmemLoad{cur}, cur, tw);
				return RunPasses<LOG2M, INV, INPLACE, PASS + 1>::run(tid, cur, alt, tw);
			} else {
				stockham_pass<LOG2M, PASS, INV>(tid, SmemLoad{cur}, alt, tw);
				return RunPasses<LOG2M, INV, INPLACE, PASS + 1>::run(tid, alt, cur, tw);
			}
		}
	}
};

}  // namespace par
