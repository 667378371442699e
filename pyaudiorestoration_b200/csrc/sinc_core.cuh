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
// Entry m (tap pair m of a thread, absolute input indices j0 + 2m, j0 + 2m + 1):
//   .x, .y = C[2m], C[2m+1]      E slot (first tap at j0:     weight indices 2m,   2m+1)
//   .z, .w = C[2m-1], C[2m]      O slot (first tap at j0 + 1: weight indices 2m-1, 2m)
// with C[k] = 0 outside 0 .. 2NT-1.  `lp` is the same table with the centre coefficient C[NT] zeroed.
template <int CAP>
struct SincTab {
	float4 full[CAP];
	float4 lp[CAP];
};
constexpr int SINC_TAB_SMALL = 132;     // NT <= 128 (+ padding to whole blocks of 4 pairs)
constexpr int SINC_TAB_LARGE = 516;     // NT <= 512
constexpr int SINC_PAIRS_PER_BLOCK = 4;

SC_HD int sinc_num_blocks(int nt) { return (nt + 1 + SINC_PAIRS_PER_BLOCK - 1) / SINC_PAIRS_PER_BLOCK; }

// One interior output sample (all 2NT taps inside the signal, no start-edge shift).
struct SincSlot {
	float s;          // fractional shift p - rint(p), never exactly 0
	float fc;         // fc rounded to float32 (centre tap of the fc < 1 path)
	uint64_t g_fx;    // (1 - fc) in units of 2^-63 half-turns
	int64_t s_fx;     // fc * s in the same units
};

// fc < 1: rotation table (cos, sin)(j pi g), j < 8, as pairs; the step (cos, sin)(8 pi g).
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

// exact block anchor: (sin, cos)(theta_d0), theta_d0 = pi (g d0 + fc s)
SC_HD void sinc_anchor(const SincSlot &sl, int d0, float *sa, float *ca) {
	sincos_fx(sl.g_fx * (uint64_t)(int64_t)d0 + (uint64_t)sl.s_fx, sa, ca);
}

// One block of 4 tap pairs (B = block index) for the E and the O output of a thread, all CH channels.
//   xs    : this thread's pair 0 of channel 0 in the planar staging buffer (8-byte aligned)
//   xpitch: floats between channels
//   NEAR  : the block holds a tap with |d| < 16: q = d - s is formed from the exact integer d
//   DESC  : pairs are visited from high to low (right half of the tap run, accumulated towards the centre)
template <int CH, bool LOWPASS, bool DESC, bool NEAR>
SC_HD void sinc_block(int B, int nt, const float4 *tab, const float *xs, int xpitch,
                      const SincSlot &E, const SincSlot &O, const SincRot &rotE, const SincRot &rotO,
                      float saE, float caE, float saO, float caO, float2 (&accE)[CH], float2 (&accO)[CH]) {
	const float d0f = (float)(8 * B - nt);                 // d of the E slot's first tap in this block
	const float baseE = d0f - E.s, baseO = (d0f - 1.f) - O.s;
#pragma unroll
	for (int uu = 0; uu < 4; uu++) {
		const int u = DESC ? 3 - uu : uu;
		const int m = 4 * B + u;
		const float4 c = tab[m];
		const float2 J = make_float2((float)(2 * u + 1), (float)(2 * u));
		float2 QE, QO;           // (q of the pair's second tap, q of its first tap) = (q + 1, q)
		if (NEAR) {
			const float2 DE = fadd2(J, bcast2(d0f));       // exact small integers
			QE = fadd2(DE, bcast2(-E.s));
			QO = fadd2(fadd2(DE, bcast2(-1.f)), bcast2(-O.s));
		} else {
			QE = fadd2(J, bcast2(baseE));                  // |q| >= 15: one more rounding is harmless
			QO = fadd2(J, bcast2(baseO));
		}
		const float tE = sc_rcp(QE.x * QE.y), tO = sc_rcp(QO.x * QO.y);
		float2 nE = fmul2(QE, make_float2(c.x, c.y));      // (C_k0 (q+1), C_k1 q)
		float2 nO = fmul2(QO, make_float2(c.z, c.w));
		if (LOWPASS) {
			const float2 snE = ffma2(bcast2(saE), rotE.c[u], fmul2(bcast2(caE), rotE.s[u]));
			const float2 snO = ffma2(bcast2(saO), rotO.c[u], fmul2(bcast2(caO), rotO.s[u]));
			nE = fmul2(nE, snE);
			nO = fmul2(nO, snO);
		}
		const float2 wE = fmul2(nE, bcast2(tE)), wO = fmul2(nO, bcast2(tO));
#pragma unroll
		for (int ch = 0; ch < CH; ch++) {
			const float2 X = *reinterpret_cast<const float2 *>(xs + ch * xpitch + 2 * m);
			accE[ch] = ffma2(X, wE, accE[ch]);
			accO[ch] = ffma2(X, wO, accO[ch]);
		}
	}
}

// All taps of the E and O outputs of one thread.  Summation order: the weights decay like 1/|d| away
// from the centre tap, so each half of the tap run is accumulated from its far end towards the centre
// (small terms first), even and odd taps in separate lanes: four partial sums per output and channel.
template <int CH, bool LOWPASS>
SC_HD void sinc_unit(int nt, const float4 *tab, float centre_c, const float *xs, int xpitch,
                     const SincSlot &E, const SincSlot &O, float (&outE)[CH], float (&outO)[CH]) {
	const int nblk = sinc_num_blocks(nt);
	const int half = nblk >> 1;
	SincRot rotE, rotO;
	if (LOWPASS) {
		rotE.build(E.g_fx);
		rotO.build(O.g_fx);
	}
	float2 aEl[CH], aEr[CH], aOl[CH], aOr[CH];
#pragma unroll
	for (int ch = 0; ch < CH; ch++) aEl[ch] = aEr[ch] = aOl[ch] = aOr[ch] = make_float2(0.f, 0.f);
	float saEl = 0.f, caEl = 1.f, saEr = 0.f, caEr = 1.f, saOl = 0.f, caOl = 1.f, saOr = 0.f, caOr = 1.f;
	for (int it = 0; it < half; it++) {
		const int bl = it, br = nblk - 1 - it;
		if (LOWPASS) {
			if (it == 0 || ((half - 1 - it) & 7) == 0) {
				sinc_anchor(E, 8 * bl - nt, &saEl, &caEl);
				sinc_anchor(E, 8 * br - nt, &saEr, &caEr);
				sinc_anchor(O, 8 * bl - 1 - nt, &saOl, &caOl);
				sinc_anchor(O, 8 * br - 1 - nt, &saOr, &caOr);
			} else {
				// left anchors advance by +8 pi g, right anchors by -8 pi g
				const float a = fmaf(saEl, rotE.c8, caEl * rotE.s8), b = fmaf(caEl, rotE.c8, -saEl * rotE.s8);
				const float c = fmaf(saEr, rotE.c8, -caEr * rotE.s8), d = fmaf(caEr, rotE.c8, saEr * rotE.s8);
				saEl = a; caEl = b; saEr = c; caEr = d;
				const float e = fmaf(saOl, rotO.c8, caOl * rotO.s8), f = fmaf(caOl, rotO.c8, -saOl * rotO.s8);
				const float g = fmaf(saOr, rotO.c8, -caOr * rotO.s8), h = fmaf(caOr, rotO.c8, saOr * rotO.s8);
				saOl = e; caOl = f; saOr = g; caOr = h;
			}
		}
		// |d| < 16 somewhere in the block (E and O slots together cover d0 - 1 .. d0 + 7)
		const bool nearl = 8 * bl + 7 - nt > -16 && 8 * bl - 1 - nt < 16;
		const bool nearr = 8 * br + 7 - nt > -16 && 8 * br - 1 - nt < 16;
		if (nearl) sinc_block<CH, LOWPASS, false, true>(bl, nt, tab, xs, xpitch, E, O, rotE, rotO, saEl, caEl, saOl, caOl, aEl, aOl);
		else sinc_block<CH, LOWPASS, false, false>(bl, nt, tab, xs, xpitch, E, O, rotE, rotO, saEl, caEl, saOl, caOl, aEl, aOl);
		if (nearr) sinc_block<CH, LOWPASS, true, true>(br, nt, tab, xs, xpitch, E, O, rotE, rotO, saEr, caEr, saOr, caOr, aEr, aOr);
		else sinc_block<CH, LOWPASS, true, false>(br, nt, tab, xs, xpitch, E, O, rotE, rotO, saEr, caEr, saOr, caOr, aEr, aOr);
	}
	if (nblk & 1) {
		if (LOWPASS) {
			sinc_anchor(E, 8 * half - nt, &saEl, &caEl);
			sinc_anchor(O, 8 * half - 1 - nt, &saOl, &caOl);
		}
		sinc_block<CH, LOWPASS, false, true>(half, nt, tab, xs, xpitch, E, O, rotE, rotO, saEl, caEl, saOl, caOl, aEl, aOl);
	}
	if (LOWPASS) {
		// centre tap: q = -s, sin(theta_0) = sin(pi fc s) with relative accuracy as s -> 0
		const float wE = centre_c * sc_sinpi(E.fc * E.s) * sc_rcp(-E.s);
		const float wO = centre_c * sc_sinpi(O.fc * O.s) * sc_rcp(-O.s);
#pragma unroll
		for (int ch = 0; ch < CH; ch++) {
			const float l = aEl[ch].x + aEl[ch].y, r = aEr[ch].x + aEr[ch].y;
			outE[ch] = fmaf(xs[ch * xpitch + nt], wE, l + r);
			const float lo = aOl[ch].x + aOl[ch].y, ro = aOr[ch].x + aOr[ch].y;
			outO[ch] = fmaf(xs[ch * xpitch + nt + 1], wO, lo + ro);
		}
	} else {
		const float spE = sc_sinpi(E.s), spO = sc_sinpi(O.s);
#pragma unroll
		for (int ch = 0; ch < CH; ch++) {
			outE[ch] = ((aEl[ch].x + aEl[ch].y) + (aEr[ch].x + aEr[ch].y)) * spE;
			outO[ch] = ((aOl[ch].x + aOl[ch].y) + (aOr[ch].x + aOr[ch].y)) * spO;
		}
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

SC_HD SincSetup sinc_setup(double p, double per, int nt, int64_t n_in, bool aligned_edges) {
	SincSetup su;
	double fc = 1.0 / per;
	if (!(fc < 1.0)) fc = 1.0;
	double pr = rint(p);                        // half to even, like Python's round()
	if (!(pr > -9.0e15)) pr = -9.0e15;          // NaN / -inf guard (garbage in, zeros out)
	if (pr > 9.0e15) pr = 9.0e15;
	const long long ind = (long long)pr;
	const double sd = p - pr;
	long long lower = ind - nt, upper = ind + nt;
	if (lower < 0) lower = 0;
	if (upper > n_in) upper = n_in;
	su.lower = lower;
	su.cnt = upper > lower ? (int)(upper - lower) : 0;
	su.koff = aligned_edges ? (int)(lower - (ind - nt)) : 0;
	float s = (float)sd;
	if (s == 0.f) s = 1e-30f;
	su.slot.s = s;
	su.lowpass = fc < 1.0;
	su.slot.fc = (float)fc;
	su.slot.g_fx = 0;
	su.slot.s_fx = 0;
	su.f_fx = 0;
	if (su.lowpass) {
		// fc in (0, 1): fc * 2^63 < 2^63; g = 1 - fc is exact in float64 for fc >= 0.5
		const double f63 = fc * 9223372036854775808.0;
#if defined(__CUDA_ARCH__)
		su.f_fx = __double2ull_rn(f63);
		su.slot.s_fx = __double2ll_rn(fc * sd * 9223372036854775808.0);
#else
		su.f_fx = (uint64_t)llrint(f63);
		su.slot.s_fx = (int64_t)llrint(fc * sd * 9223372036854775808.0);
#endif
		su.slot.g_fx = 9223372036854775808ull - su.f_fx;
	}
	return su;
}

// Host: C[k] per NT, packed for the E / O slots.  Returns the centre coefficient C[NT].
template <int CAP>
inline float sinc_fill_table(int nt, SincTab<CAP> *t) {
	float c[2 * 512 + 2];
	const int mm = 2 * nt + 1;
	for (int k = 0; k < 2 * nt; k++) {
		// np.hanning(2nt+1)[k] rounded to float32 (util/resampling.py:24,36), then the sign of
		// sin(pi (d - s)) and 1/pi folded in, in float64, rounded once
		const double nn = (double)(1 - mm + 2 * k);
		const float h = (float)(0.5 + 0.5 * cos(M_PI * nn / (double)(mm - 1)));
		const int d = k - nt;
		c[k] = (float)((((d + 1) & 1) ? -1.0 : 1.0) * (double)h / M_PI);
	}
	auto at = [&](int k) { return (k >= 0 && k < 2 * nt) ? c[k] : 0.f; };
	auto lp = [&](int k) { return k == nt ? 0.f : at(k); };
	for (int m = 0; m < CAP; m++) {
		t->full[m] = make_float4(at(2 * m), at(2 * m + 1), at(2 * m - 1), at(2 * m));
		t->lp[m] = make_float4(lp(2 * m), lp(2 * m + 1), lp(2 * m - 1), lp(2 * m));
	}
	return c[nt];
}

}  // namespace par
