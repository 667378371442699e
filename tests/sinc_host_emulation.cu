// Host build of csrc/sinc_core.cuh: the per-output arithmetic of the sm_100a sinc kernel, compiled for
// the CPU (packed ops lane by lane, rcp.approx as an IEEE division) so that tests/test_sinc_emulation_cpu.py
// can check the numerics against the float64 oracle without a GPU.  Test infrastructure only.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../pyaudiorestoration_b200/csrc/sinc_core.cuh"

using namespace par;

// Interior outputs only (all 2*NT taps inside the signal); edge outputs are written as NaN.
extern "C" int sinc_emulate(const double *pos, long long m, const float *x, long long n_in, int nt, float *out) {
	static SincTab<SINC_TAB_LARGE> tab;
	sinc_fill_table<SINC_TAB_LARGE>(nt, &tab);
	// parity planes of the zero-padded signal (sinc_core.cuh, "Staging layout"): staged sample P = input sample P - 16
	const long long half = (n_in + 64) / 2 + 2;
	std::vector<float> xp(2 * (size_t)half, 0.f);
	for (long long j = 0; j < n_in; j++) { const long long P = j + 16; xp[(P & 1) * half + (P >> 1)] = x[j]; }
	for (long long i = 0; i < m; i++) {
		const double p = pos[i];
		double per;
		if (i + 1 < m) per = fmax(1e-12, pos[i + 1] - p);
		else per = m >= 2 ? fmax(1e-12, pos[m - 1] - pos[m - 2]) : 0.0;
		const SincSetup su = sinc_setup(p, per, nt, n_in, false);
		if (!(su.cnt == 2 * nt && su.koff == 0)) { out[i] = NAN; continue; }
		SincSlot dummy;
		dummy.s = 0.5f; dummy.fc = 1.f; dummy.g_fx = 0; dummy.s_fx = 0;
		const long long cen = su.lower + nt + 16;            // staged index of the centre tap
		const bool odd = cen & 1;
		const long long c0 = cen & ~1ll;
		const SincSlot &sE = odd ? dummy : su.slot, &sO = odd ? su.slot : dummy;
		const unsigned long long gE = sE.g_fx, gO = sO.g_fx;
		const long long fE = sE.s_fx, fO = sO.s_fx;
		const SincSlotRef sl[2] = {{&sE.s, &sE.fc, &gE, &fE}, {&sO.s, &sO.fc, &gO, &fO}};
		float y[2][1];
		const SincWin<1> xs{xp.data() + (c0 >> 1), (int)half};
		if (su.lowpass) sinc_unit<1, true, 2, SINC_TAB_LARGE>(nt, tab, xs, sl, y);
		else sinc_unit<1, false, 2, SINC_TAB_LARGE>(nt, tab, xs, sl, y);
		out[i] = y[odd ? 1 : 0][0];
	}
	return 0;
}
