// sinc_core.cuh -- per-thread arithmetic of the windowed-sinc interpolator (util/resampling.py:51-90),
// shared by the sm_100a kernel (resample.cu) and by a host build of the same source
// (tests/sinc_host_emulation.cu) that checks the numerics on the CPU before GPU time is spent.
//
// Formulation.  For output i with read position p: ind = rint(p), s = p - ind, fc = min(1/period, 1),
// g = 1 - fc, taps k = 0 .. 2NT-1 at d = k - NT, q = d - s:
//     w_k = h[k] fc sinc(fc q) = C_k sin(theta_d) / q,   C_k = (-1)^(d+1) h[k] / pi,
//     theta_d = pi (g d + fc s)            (sin(pi fc (d - s)) = (-1)^(d+1) sin(pi (g d + fc s)))
// fc == 1:  theta = pi s for every tap, so sin(pi s) is applied once at the end.
// fc <  1:  theta advances by pi g per tap (a SMALL angle for tape-speed curves): block anchors
//           (sin, cos)(theta_d0) exact from a 64-bit fixed-point phase every 8th block, rotated by
//           8 pi g in between, and sin(theta_d0 + j pi g) by angle addition from a per-output table
//           of (cos, sin)(j pi g), j < 8.  The centre tap (q -> 0 needs RELATIVE accuracy of the sine)
//           is taken out of the table (coefficient zeroed) and added separately.
//
// Pairing.  Taps are processed as PAIRS aligned to EVEN ABSOLUTE input indices, so that
//   * one 8-byte shared-memory load fetches both samples of a pair (planar staging),
//   * one reciprocal serves both taps:  t = 1/(q (q+1)),  1/q = (q+1) t,  1/(q+1) = q t,
//   * all per-pair arithmetic is packed (fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2, IEEE per lane),
//   * two outputs whose tap runs start at j0 (even, "E slot") and j0 + 1 ("O slot") read the SAME
//     sample pairs: a thread interpolates both and every staged sample is fetched once per two outputs.
// An output's arithmetic depends only on the parity of its own first tap index, never on its
// neighbour, the tile or the chunking of a call, so results are bit-identical however the work is cut.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define SC_HD __host__ __device__ __forceinline__
#else
#error "compile with nvcc (host emulation builds use nvcc's host pass)"
#endif

#ifndef SINC_BLOCK_UNROLL
#define SINC_BLOCK_UNROLL 1
#endif
#ifndef SINC_EXACT_MASK
#define SINC_EXACT_MASK 7           // fc < 1: exact block anchors at the far end and at every block b with (b & mask) == 0
#endif

namespace par {

// ---- packed float32 pairs -------------------------------------------------------------------------
SC_HD float2 ffma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
	float2 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;"
	    : "=l"(*reinterpret_cast<uint64_t *>(&d))
	    : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)),
	      "l"(*reinterpret_cast<const uint64_t *>(&c)));
	return d;
#else
	return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
SC_HD float2 fmul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	float2 d;
	asm("mul.rn.f32x2 %0, %1, %2;"
	    : "=l"(*reinterpret_cast<uint64_t *>(&d))
	    : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
	return d;
#else
	return make_float2(a.x * b.x, a.y * b.y);
#endif
}
SC_HD float2 fadd2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	float2 d;
	asm("add.rn.f32x2 %0, %1, %2;"
	    : "=l"(*reinterpret_cast<uint64_t *>(&d))
	    : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
	return d;
#else
	return make_float2(a.x + b.x, a.y + b.y);
#endif
}
SC_HD float2 fsub2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	float2 d;
	asm("sub.rn.f32x2 %0, %1, %2;"
	    : "=l"(*reinterpret_cast<uint64_t *>(&d))
	    : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
	return d;
#else
	return make_float2(a.x - b.x, a.y - b.y);
#endif
}
SC_HD float2 bcast2(float v) { return make_float2(v, v); }

SC_HD float sc_rcp(float x) {
#if defined(__CUDA_ARCH__)
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
#else
	return 1.0f / x;
#endif
}

SC_HD float sc_sinpi(float x) {
#if defined(__CUDA_ARCH__)
	return sinpif(x);
#else
	return (float)sin(M_PI * (double)x);
#endif
}

// sin/cos of the angle  pi * phase / 2^63  (phase wraps at one full turn = 2^64).
// The top 32 bits are split into the nearest quarter turn and a residual in [-pi/4, pi/4) that is
// converted to float32 (absolute error <= 5e-8 rad) and fed to Taylor polynomials whose truncation
// error is < 2e-9 on that interval.
SC_HD void sincos_fx(uint64_t phase, float *s, float *c) {
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

// ---- coefficient table (a kernel parameter: read through the constant bank with uniform loads) ----
// The Hann window is symmetric, so the two taps at distance d on either side of the centre tap share
//   C_d = (-1)^(d+1) hanning(2NT+1)[NT +- d] / pi
// and for an output with fractional shift s (|s| <= 1/2) their weights (times sin(theta), see above) are
//   w(+d) = C_d / (d - s),   w(-d) = C_d / (-d - s).
// NEAR distances (d <= 16):   w(+d) = C_d (s + d) t,  w(-d) = C_d (s - d) t,  t = 1 / (d^2 - s^2)
//     -- one reciprocal and five FMA-pipe operations per tap pair;
// FAR distances (d > 16):     1 / (d -+ s) as a power series in s / d (|s/d| < 1/33, four terms: the
//     truncation is < 1e-6 of a weight that is itself < 0.02):
//     w(+d) = O + E,  w(-d) = E - O,  O = a0 + a2 s^2,  E = s (a1 + a3 s^2),  a_k = C_d / d^(k+1)
//     -- five FMA-pipe operations per tap pair and NO reciprocal.
// Entry p holds the distances d = 2p+1 and d' = 2p+2 so that two consecutive distances run as packed pairs:
//     near:  a = (d^2, d'^2, C_d, C_d'),      b = (C_d d, C_d' d', -C_d d, -C_d' d')
//     far:   a = (a0_d, a0_d', a2_d, a2_d'),  b = (a1_d, a1_d', a3_d, a3_d')
// Distances past NT-1 (block padding) have all-zero far entries.  c0 is the centre coefficient C_0.
template <int CAP>
struct SincTab {
	float4 a[CAP];
	float4 b[CAP];
	float c0;
};
constexpr int SINC_TAB_SMALL = 64;      // NT <= 128: (NT - 1) distances in pairs, padded to whole blocks of 8
constexpr int SINC_TAB_LARGE = 256;     // NT <= 512
constexpr int SINC_BLOCK = 8;           // distances per block
constexpr int SINC_NEAR_BLOCKS = 2;     // blocks (from the centre) that use the exact reciprocal form

SC_HD int sinc_num_blocks(int nt) { return (nt - 1 + SINC_BLOCK - 1) / SINC_BLOCK; }

// One interior output sample (all 2NT taps inside the signal, no start-edge shift).
struct SincSlot {
	float s;          // fractional shift p - rint(p), never exactly 0
	float fc;         // fc rounded to float32 (centre tap of the fc < 1 path)
	uint64_t g_fx;    // (1 - fc) in units of 2^-63 half-turns
	int64_t s_fx;     // fc * s in the same units
};

// fc < 1: rotation table (cos, sin)(j pi g), j < 8, as pairs (j, j+1); the step (cos, sin)(8 pi g).
// j = 1, 2, 4 and 8 come exactly from the fixed-point angle, 3, 5, 6, 7 are one or two complex products.
struct SincRot {
	float2 c[4], s[4];
	float c8, s8;
	SC_HD void build(uint64_t g_fx) {
		float cj[8], sj[8];
		cj[0] = 1.f; sj[0] = 0.f;
		sincos_fx(g_fx, &sj[1], &cj[1]);
		sincos_fx(g_fx * 2ull, &sj[2], &cj[2]);
		sincos_fx(g_fx * 4ull, &sj[4], &cj[4]);
		sincos_fx(g_fx * 8ull, &s8, &c8);
		cj[3] = fmaf(cj[2], cj[1], -sj[2] * sj[1]); sj[3] = fmaf(sj[2], cj[1], cj[2] * sj[1]);
		cj[5] = fmaf(cj[4], cj[1], -sj[4] * sj[1]); sj[5] = fmaf(sj[4], cj[1], cj[4] * sj[1]);
		cj[6] = fmaf(cj[4], cj[2], -sj[4] * sj[2]); sj[6] = fmaf(sj[4], cj[2], cj[4] * sj[2]);
		cj[7] = fmaf(cj[4], cj[3], -sj[4] * sj[3]); sj[7] = fmaf(sj[4], cj[3], cj[4] * sj[3]);
#pragma unroll
		for (int u = 0; u < 4; u++) {
			c[u] = make_float2(cj[2 * u], cj[2 * u + 1]);
			s[u] = make_float2(sj[2 * u], sj[2 * u + 1]);
		}
	}
};

// Where the set-up of an output lives (shared memory in the kernel): the tap loop re-reads the fixed-point
// phases when it needs an exact anchor instead of holding them in registers.
struct SincSlotRef {
	const float *s, *fc;
	const unsigned long long *g_fx;
	const long long *s_fx;
};

// exact anchor: (sin, cos)(theta_d), theta_d = pi (g d + fc s)
SC_HD void sinc_anchor(const SincSlotRef &sl, int d, float *sa, float *ca) {
	sincos_fx((uint64_t)*sl.g_fx * (uint64_t)(int64_t)d + (uint64_t)*sl.s_fx, sa, ca);
}

// Channel vectors: the staging buffer is channel-interleaved, x[sample * CH + ch], so one 4/8/16-byte load
// fetches all channels of a sample and one packed FMA with the broadcast weight updates two channels.
template <int CH> struct SincVec;
template <> struct SincVec<1> {
	float v;
	SC_HD void zero() { v = 0.f; }
	SC_HD void load(const float *p) { v = *p; }
	SC_HD void fma(const SincVec &x, float w) { v = fmaf(x.v, w, v); }
	SC_HD void add(const SincVec &o) { v += o.v; }
	SC_HD float get(int) const { return v; }
};
template <> struct SincVec<2> {
	float2 v;
	SC_HD void zero() { v = make_float2(0.f, 0.f); }
	SC_HD void load(const float *p) { v = *reinterpret_cast<const float2 *>(p); }
	SC_HD void fma(const SincVec &x, float w) { v = ffma2(x.v, bcast2(w), v); }
	SC_HD void add(const SincVec &o) { v = fadd2(v, o.v); }
	SC_HD float get(int c) const { return c ? v.y : v.x; }
};
template <> struct SincVec<4> {
	float2 a, b;
	SC_HD void zero() { a = b = make_float2(0.f, 0.f); }
	SC_HD void load(const float *p) {
		const float4 t = *reinterpret_cast<const float4 *>(p);
		a = make_float2(t.x, t.y); b = make_float2(t.z, t.w);
	}
	SC_HD void fma(const SincVec &x, float w) { a = ffma2(x.a, bcast2(w), a); b = ffma2(x.b, bcast2(w), b); }
	SC_HD void add(const SincVec &o) { a = fadd2(a, o.a); b = fadd2(b, o.b); }
	SC_HD float get(int c) const { return c == 0 ? a.x : c == 1 ? a.y : c == 2 ? b.x : b.y; }
};

// Per-output running state of the tap loop.
template <int CH, bool LOWPASS>
struct SincAcc {
	float s, s2;
	SincVec<CH> accl, accr;
	float sar, car, sal, ncal;      // fc < 1: (sin, cos) of theta at +d_a and (sin, -cos) at -d_a, d_a = first distance of the block
};

// Weights of the tap pairs at distances (d, d+1) = (8 b + 2 u + 1, 8 b + 2 u + 2) of one output.
template <int CH, bool LOWPASS, bool NEAR>
SC_HD void sinc_pair_weights(const float4 ta, const float4 tb, int u, const SincRot &rot,
                             const SincAcc<CH, LOWPASS> &A, float2 *wp, float2 *wm) {
	float2 np, nm;
	if (NEAR) {
		const float2 D = fadd2(make_float2(ta.x, ta.y), bcast2(-A.s2));           // d^2 - s^2
		const float2 t = make_float2(sc_rcp(D.x), sc_rcp(D.y));
		np = fmul2(ffma2(bcast2(A.s), make_float2(ta.z, ta.w), make_float2(tb.x, tb.y)), t);   // C (s + d) t
		nm = fmul2(ffma2(bcast2(A.s), make_float2(ta.z, ta.w), make_float2(tb.z, tb.w)), t);   // C (s - d) t
	} else {
		const float2 O = ffma2(make_float2(ta.z, ta.w), bcast2(A.s2), make_float2(ta.x, ta.y));
		const float2 E = fmul2(ffma2(make_float2(tb.z, tb.w), bcast2(A.s2), make_float2(tb.x, tb.y)), bcast2(A.s));
		np = fadd2(O, E);
		nm = fsub2(E, O);
	}
	if (LOWPASS) {
		// sin(theta(+d_a) + j pi g) and sin(theta(-d_a) - j pi g), j = 2u, 2u + 1
		np = fmul2(np, ffma2(bcast2(A.sar), rot.c[u], fmul2(bcast2(A.car), rot.s[u])));
		nm = fmul2(nm, ffma2(bcast2(A.sal), rot.c[u], fmul2(bcast2(A.ncal), rot.s[u])));
	}
	*wp = np;
	*wm = nm;
}

// Staging layout.  A thread interpolates the output E whose centre tap is the EVEN staged sample c and the
// output O centred at c + 1; consecutive threads have centres 2 samples apart.  To keep every warp-wide load
// free of bank conflicts the staged samples are split by parity: sample P lives at
//     x[(P & 1) * plane + (P >> 1) * CH + ch]
// so that the lanes of a load (same parity, centres 2 apart) read consecutive CH-vectors.
template <int CH>
struct SincWin {
	const float *even;     // the thread's centre sample c (even plane)
	int plane;             // floats between the parity planes
	// sample c + off (off known at compile time after unrolling)
	SC_HD const float *at(int off) const {
		// floor division by 2 for negative offsets too
		const int m = (off - (off & 1)) / 2;
		return even + (off & 1) * plane + m * CH;
	}
};

// One block of 8 distances d = 8 b + 1 .. 8 b + 8 (visited from far to near) for the R outputs of a thread
// (slot r is centred at sample c + r).
template <int CH, bool LOWPASS, bool NEAR, int R, int CAP>
SC_HD void sinc_block(int b, const SincTab<CAP> &tab, const SincWin<CH> &x,
                      const SincRot (&rot)[LOWPASS ? R : 1], SincAcc<CH, LOWPASS> (&A)[R]) {
	// 8 + R - 1 samples on either side: right[k] = sample c + 8 b + 1 + k, left[k] = sample c - 8 b - 8 + k
	SincVec<CH> xr[7 + R], xl[7 + R];
	SincWin<CH> xb = x;
	xb.even = x.even + 4 * b * CH;
	SincWin<CH> xa = x;
	xa.even = x.even - 4 * b * CH;
#pragma unroll
	for (int k = 0; k < 7 + R; k++) {
		xr[k].load(xb.at(1 + k));
		xl[k].load(xa.at(k - 8));
	}
#pragma unroll
	for (int u = 3; u >= 0; u--) {
		const float4 ta = tab.a[4 * b + u], tb = tab.b[4 * b + u];
		const int j1 = 2 * u + 1, j2 = 2 * u + 2;              // d - 8 b of the two distances
#pragma unroll
		for (int r = 0; r < R; r++) {
			float2 wp, wm;
			sinc_pair_weights<CH, LOWPASS, NEAR>(ta, tb, u, rot[LOWPASS ? r : 0], A[r], &wp, &wm);
			// +d is right[d - 8b - 1 + r], -d is left[8 - (d - 8b) + r]; the farther distance first
			A[r].accr.fma(xr[j2 - 1 + r], wp.y);
			A[r].accl.fma(xl[8 - j2 + r], wm.y);
			A[r].accr.fma(xr[j1 - 1 + r], wp.x);
			A[r].accl.fma(xl[8 - j1 + r], wm.x);
		}
	}
}

// All taps of the R outputs of one thread.  Summation order: the weights decay like 1/|d| away from the
// centre tap, so each half of the tap run is accumulated from its far end towards the centre (small terms
// first) in its own accumulator; the centre tap comes last.
template <int CH, bool LOWPASS, int R, int CAP, int UNROLL = SINC_BLOCK_UNROLL>
SC_HD void sinc_unit(int nt, const SincTab<CAP> &tab, const SincWin<CH> &x,
                     const SincSlotRef (&sl)[R], float (&out)[R][CH]) {
	const int nblk = sinc_num_blocks(nt);
	SincRot rot[LOWPASS ? R : 1];
	SincAcc<CH, LOWPASS> A[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		A[r].s = *sl[r].s; A[r].s2 = A[r].s * A[r].s;
		A[r].accl.zero(); A[r].accr.zero();
		A[r].sar = A[r].sal = 0.f; A[r].car = 1.f; A[r].ncal = -1.f;
		if (LOWPASS) rot[r].build(*sl[r].g_fx);
	}
	constexpr int kUnroll = UNROLL;
#pragma unroll kUnroll
	for (int b = nblk - 1; b >= 0; b--) {
		if (LOWPASS) {
			const int da = 8 * b + 1;
			if (b == nblk - 1 || (b & SINC_EXACT_MASK) == 0) {
#pragma unroll
				for (int r = 0; r < R; r++) {
					float cl;
					sinc_anchor(sl[r], da, &A[r].sar, &A[r].car);
					sinc_anchor(sl[r], -da, &A[r].sal, &cl);
					A[r].ncal = -cl;
				}
			} else {
				// moving towards the centre: the right anchor turns by -8 pi g, the left one by +8 pi g
#pragma unroll
				for (int r = 0; r < R; r++) {
					const SincRot &q = rot[LOWPASS ? r : 0];
					const float a = fmaf(A[r].sar, q.c8, -A[r].car * q.s8), b2 = fmaf(A[r].car, q.c8, A[r].sar * q.s8);
					const float c = fmaf(A[r].sal, q.c8, -A[r].ncal * q.s8), d = fmaf(A[r].ncal, q.c8, A[r].sal * q.s8);
					A[r].sar = a; A[r].car = b2; A[r].sal = c; A[r].ncal = d;
				}
			}
		}
		if (b < SINC_NEAR_BLOCKS) sinc_block<CH, LOWPASS, true, R, CAP>(b, tab, x, rot, A);
		else sinc_block<CH, LOWPASS, false, R, CAP>(b, tab, x, rot, A);
	}
	// centre tap: q = -s; fc < 1: sin(theta_0) = sin(pi fc s) with relative accuracy as s -> 0
	const float c0 = tab.c0;
#pragma unroll
	for (int r = 0; r < R; r++) {
		float w = c0 * sc_rcp(-A[r].s);
		if (LOWPASS) w *= sc_sinpi(*sl[r].fc * A[r].s);
		const float sp = LOWPASS ? 1.f : sc_sinpi(A[r].s);
		SincVec<CH> xc;
		xc.load(x.at(r));
		A[r].accl.add(A[r].accr);
		A[r].accl.fma(xc, w);
#pragma unroll
		for (int ch = 0; ch < CH; ch++) out[r][ch] = LOWPASS ? A[r].accl.get(ch) : A[r].accl.get(ch) * sp;
	}
}

// float64 part of util/resampling.py:67-84 for one output (shared by the kernel and the emulation)
struct SincSetup {
	int64_t lower;     // first input sample of the tap run
	int cnt;           // number of taps (0 .. 2NT)
	int koff;          // weight index of tap 0 (0 unless PAR_SINC_ALIGNED_EDGES at the start edge)
	bool lowpass;      // fc < 1
	SincSlot slot;
	uint64_t f_fx;     // fc in units of 2^-63 half-turns (edge path)
};

// fc = min(1 / per, 1): a single-precision reciprocal refined by two Newton steps in float64 (relative error
// < 1e-15: fc only enters the weights through pi * (1 - fc) * d, d < 512) instead of the IEEE division sequence
SC_HD double sinc_fc(double per) {
	// straight-line on purpose (several outputs are set up interleaved): the Newton value is dropped for per <= 1
	const double pc = per > 1.0 ? per : 2.0;
	const double r0 = (double)sc_rcp((float)pc);
	const double r1 = fma(fma(-pc, r0, 1.0), r0, r0);
	const double r2 = fma(fma(-pc, r1, 1.0), r1, r1);
	return (per > 1.0 && r2 < 1.0) ? r2 : 1.0;
}

// Index part of the set-up: which taps an output reads and whether it is low-passed.
struct SincIndex {
	int64_t lower;
	int cnt, koff;
	bool lowpass;
};
SC_HD SincIndex sinc_index(double p, double per, int nt, int64_t n_in, bool aligned_edges, double *fc_out = nullptr,
                           double *pr_out = nullptr) {
	SincIndex ix;
	const double fc = sinc_fc(per);
	double pr = rint(p);                        // half to even, like Python's round()
	if (!(pr > -9.0e15)) pr = -9.0e15;          // NaN / -inf guard (garbage in, zeros out)
	if (pr > 9.0e15) pr = 9.0e15;
	const long long ind = (long long)pr;
	long long lower = ind - nt, upper = ind + nt;
	if (lower < 0) lower = 0;
	if (upper > n_in) upper = n_in;
	ix.lower = lower;
	ix.cnt = upper > lower ? (int)(upper - lower) : 0;
	ix.koff = aligned_edges ? (int)(lower - (ind - nt)) : 0;
	ix.lowpass = fc < 1.0;
	if (fc_out) *fc_out = fc;
	if (pr_out) *pr_out = pr;
	return ix;
}

SC_HD SincSetup sinc_setup(double p, double per, int nt, int64_t n_in, bool aligned_edges) {
	SincSetup su;
	double fc, pr;
	const SincIndex ix = sinc_index(p, per, nt, n_in, aligned_edges, &fc, &pr);
	const double sd = p - pr;
	su.lower = ix.lower;
	su.cnt = ix.cnt;
	su.koff = ix.koff;
	float s = (float)sd;
	if (s == 0.f) s = 1e-30f;
	su.slot.s = s;
	su.lowpass = ix.lowpass;
	su.slot.fc = (float)fc;
	// fc in (0, 1): fc * 2^63 < 2^63; g = 1 - fc is exact in float64 for fc >= 0.5.  Straight-line: fc = 1 converts 0.
	const double f63 = su.lowpass ? fc * 9223372036854775808.0 : 0.0;
	const double s63 = su.lowpass ? fc * sd * 9223372036854775808.0 : 0.0;
#if defined(__CUDA_ARCH__)
	su.f_fx = __double2ull_rn(f63);
	su.slot.s_fx = __double2ll_rn(s63);
#else
	su.f_fx = (uint64_t)llrint(f63);
	su.slot.s_fx = (int64_t)llrint(s63);
#endif
	su.slot.g_fx = su.lowpass ? 9223372036854775808ull - su.f_fx : 0ull;
	return su;
}

// Host: the distance table of one NT (float64 arithmetic, rounded once).
template <int CAP>
inline void sinc_fill_table(int nt, SincTab<CAP> *t) {
	const int mm = 2 * nt + 1;
	auto coef = [&](int d) -> double {
		if (d < 0 || d >= nt) return 0.0;
		// np.hanning(2nt+1)[nt + d] rounded to float32 (util/resampling.py:24,36), then the sign of
		// sin(pi (d - s)) and 1/pi folded in
		const double nn = (double)(1 - mm + 2 * (nt + d));
		const float h = (float)(0.5 + 0.5 * cos(M_PI * nn / (double)(mm - 1)));
		return (double)(float)((((d + 1) & 1) ? -1.0 : 1.0) * (double)h / M_PI);
	};
	t->c0 = (float)coef(0);
	for (int p = 0; p < CAP; p++) {
		const int d1 = 2 * p + 1, d2 = 2 * p + 2;
		const double c1 = coef(d1), c2 = coef(d2);
		if (p < SINC_NEAR_BLOCKS * (SINC_BLOCK / 2)) {
			t->a[p] = make_float4(d1 < nt ? (float)((double)d1 * d1) : 1.f, d2 < nt ? (float)((double)d2 * d2) : 1.f,
			                      (float)c1, (float)c2);
			t->b[p] = make_float4((float)(c1 * d1), (float)(c2 * d2), (float)(-c1 * d1), (float)(-c2 * d2));
		} else {
			const double e1 = d1, e2 = d2;
			t->a[p] = make_float4((float)(c1 / e1), (float)(c2 / e2), (float)(c1 / (e1 * e1 * e1)), (float)(c2 / (e2 * e2 * e2)));
			t->b[p] = make_float4((float)(c1 / (e1 * e1)), (float)(c2 / (e2 * e2)), (float)(c1 / (e1 * e1 * e1 * e1)),
			                      (float)(c2 / (e2 * e2 * e2 * e2)));
		}
	}
}

}  // namespace par
