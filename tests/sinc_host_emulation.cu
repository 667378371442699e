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
	const float centre = sinc_fill_table<SINC_TAB_LARGE>(nt, &tab);
	std::vector<float> xp((size_t)n_in + 64, 0.f);
	memcpy(xp.data() + 2, x, (size_t)n_in * sizeof(float));       // 2 floats of slack in front (O slot reads j0 = lo - 1)
	for (long long i = 0; i < m; i++) {
		const double p = pos[i];
		double per;
		if (i + 1 < m) per = fmax(1e-12, pos[i + 1] - p);
		else per = m >= 2 ? fmax(1e-12, pos[m - 1] - pos[m - 2]) : 0.0;
		const SincSetup su = sinc_setup(p, per, nt, n_in, false);
		if (!(su.cnt == 2 * nt && su.koff == 0)) { out[i] = NAN; continue; }
		SincSlot dummy;
		dummy.s = 0.5f; dummy.fc = 1.f; dummy.g_fx = 0; dummy.s_fx = 0;
		const bool odd = su.lower & 1;
		const long long j0 = su.lower & ~1ll;
		const SincSlot &E = odd ? dummy : su.slot, &O = odd ? su.slot : dummy;
		float yE[1], yO[1];
		const float *xs = xp.data() + 2 + j0;
		if (su.lowpass) sinc_unit<1, true>(nt, tab.lp, centre, xs, 0, E, O, yE, yO);
		else sinc_unit<1, false>(nt, tab.full, centre, xs, 0, E, O, yE, yO);
		out[i] = odd ? yO[0] : yE[0];
	}
	return 0;
}
