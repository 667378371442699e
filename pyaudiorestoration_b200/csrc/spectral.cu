// spectral.cu -- STFT-domain mask operators that run between the analysis and the synthesis transform
// WITHOUT the spectrogram leaving the device:
//   gate    noise gate of renoiser_gui.py:273-278, :314-317   S *= to_fac(gain) where to_dB(|S| + 1e-7) <= profile[bin]
//   select  max / min mono of dropouts_gui.py:153-161         D = where(|L| > |R| (or <), L, R)
//   heal    dropout gain of dropout_healer_gui.py:134-162     per marker: mean dB before / after the gap, bilinear blend,
//           boost clipped to [earlier boost, 255] dB, S *= to_fac(boost)
// The mask arithmetic (decibels, means, powers) is float64 like the reference's numpy path; the spectrogram stays
// complex64.  These kernels touch each cell once (gate / select) or only the few thousand cells of the marked
// regions (heal): they are bandwidth-trivial next to the two transforms around them.
#include <math.h>

#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

__device__ __forceinline__ double cell_db(float2 z) {
	// to_dB(to_mag(S)) = 20 log10(|S| + 1e-7)   (util/fourier.py:23-24, util/units.py:27-28)
	return 20.0 * log10(hypot((double)z.x, (double)z.y) + 1.0e-7);
}

__global__ void __launch_bounds__(256)
spec_gate_kernel(float2 *__restrict__ S, int64_t cells, int F, const double *__restrict__ thr_db, float fac) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
		const float2 z = S[i];
		if (!(cell_db(z) > thr_db[i % F])) S[i] = make_float2(z.x * fac, z.y * fac);
	}
}

// out_max / out_min may alias L / R (every cell is read before it is written by the same thread)
__global__ void __launch_bounds__(256)
spec_select_kernel(const float2 *L, const float2 *R, int64_t cells, float2 *out_max, float2 *out_min) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
		const float2 l = L[i], r = R[i];
		const double ml = (double)l.x * l.x + (double)l.y * l.y, mr = (double)r.x * r.x + (double)r.y * r.y;
		if (out_max) out_max[i] = ml > mr ? l : r;
		if (out_min) out_min[i] = ml < mr ? l : r;
	}
}

struct HealRegion { long long frame_b, frame_a, around, bin_l, bin_u; };

// Python slice semantics of spectrum_db[:, lo:hi] along an axis of length T
__device__ __forceinline__ void py_slice(long long &lo, long long &hi, long long T) {
	if (lo < 0) lo += T;
	if (hi < 0) hi += T;
	lo = lo < 0 ? 0 : (lo > T ? T : lo);
	hi = hi < 0 ? 0 : (hi > T ? T : hi);
}

// one block per region: mean dB of the `around` frames before and after the gap, per bin (:144-145)
__global__ void __launch_bounds__(256)
heal_means_kernel(const float2 *__restrict__ S, int64_t pitch, long long T, HealRegion rg, double *__restrict__ before,
                  double *__restrict__ after) {
	long long b0 = rg.frame_b - rg.around, b1 = rg.frame_b, a0 = rg.frame_a, a1 = rg.frame_a + rg.around;
	py_slice(b0, b1, T);
	py_slice(a0, a1, T);
	for (long long b = rg.bin_l + threadIdx.x; b < rg.bin_u; b += blockDim.x) {
		double sb = 0.0, sa = 0.0;
		for (long long t = b0; t < b1; t++) sb += cell_db(S[t * pitch + b]);
		for (long long t = a0; t < a1; t++) sa += cell_db(S[t * pitch + b]);
		before[b] = b1 > b0 ? sb / (double)(b1 - b0) : NAN;      // np.mean of an empty slice
		after[b] = a1 > a0 ? sa / (double)(a1 - a0) : NAN;
	}
}

// boost of one region into the dense gain map G (frames [g0, ...), bins contiguous), :148-160
__global__ void __launch_bounds__(256)
heal_boost_kernel(const float2 *__restrict__ S, int64_t pitch, HealRegion rg, const double *__restrict__ before,
                  const double *__restrict__ after, double *__restrict__ G, long long g0, int F) {
	const long long nb = rg.bin_u - rg.bin_l, nf = rg.frame_a - rg.frame_b;
	const long long cells = nb * nf;
	// frames = np.linspace(frame_b, frame_a, num=nf): the interpolation abscissa of gap frame j
	const double step = nf > 1 ? (double)(rg.frame_a - rg.frame_b) / (double)(nf - 1) : 0.0;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
		const long long j = i / nb, b = rg.bin_l + (i - j * nb);
		const long long t = rg.frame_b + j;
		const double xj = j == nf - 1 && nf > 1 ? (double)rg.frame_a : (double)rg.frame_b + (double)j * step;
		const double y = (xj - (double)rg.frame_b) / (double)(rg.frame_a - rg.frame_b);
		const double target = before[b] * (1.0 - y) + after[b] * y;
		double boost = target - cell_db(S[t * pitch + b]);
		const double lo = G[(t - g0) * F + b];
		boost = fmin(fmax(boost, lo), 255.0);             // np.clip(boost, earlier boost, 255)
		G[(t - g0) * F + b] = boost;
	}
}

__global__ void __launch_bounds__(256)
heal_apply_kernel(float2 *__restrict__ S, int64_t pitch, const double *__restrict__ G, long long g0, long long g1, int F) {
	const long long cells = (g1 - g0) * F;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
		const double g = G[i];
		if (g != 0.0) {
			const long long t = g0 + i / F, b = i % F;
			const double fac = pow(10.0, g / 20.0);                // to_fac, util/units.py:31-32
			const float2 z = S[t * pitch + b];
			S[t * pitch + b] = make_float2((float)((double)z.x * fac), (float)((double)z.y * fac));
		}
	}
}

static unsigned grid_for(int64_t n, int device) {
	int64_t g = (n + 255) / 256;
	const int64_t cap = (int64_t)sm_count(device) * 16;
	if (g > cap) g = cap;
	return (unsigned)(g < 1 ? 1 : g);
}

int launch_spec_gate(float2 *S, int64_t cells, int F, const double *thr_db_dev, double gain_db, int device, cudaStream_t st) {
	if (cells <= 0) return PAR_OK;
	const float fac = (float)pow(10.0, gain_db / 20.0);             // to_fac(gain).astype(float32)
	spec_gate_kernel<<<grid_for(cells, device), 256, 0, st>>>(S, cells, F, thr_db_dev, fac);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

int launch_spec_select(const float2 *L, const float2 *R, int64_t cells, float2 *out_max, float2 *out_min, int device,
                       cudaStream_t st) {
	if (cells <= 0) return PAR_OK;
	spec_select_kernel<<<grid_for(cells, device), 256, 0, st>>>(L, R, cells, out_max, out_min);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

// regions: host array of n_regions x 5 int64 (frame_b, frame_a, frame_surrounding, bin_l, bin_u), applied in order.
// scratch_dev: 2 * F doubles (means) + (g1 - g0) * F doubles (gain map, zeroed here).
int launch_spec_heal(float2 *S, int64_t pitch, int64_t T, int F, const int64_t *regions, int64_t n_regions, int64_t g0,
                     int64_t g1, double *scratch_dev, int device, cudaStream_t st) {
	if (n_regions <= 0 || g1 <= g0) return PAR_OK;
	double *before = scratch_dev, *after = scratch_dev + F, *G = scratch_dev + 2 * F;
	PAR_CUDA(cudaMemsetAsync(G, 0, (size_t)(g1 - g0) * F * sizeof(double), st));
	for (int64_t r = 0; r < n_regions; r++) {
		HealRegion rg{regions[5 * r], regions[5 * r + 1], regions[5 * r + 2], regions[5 * r + 3], regions[5 * r + 4]};
		heal_means_kernel<<<1, 256, 0, st>>>(S, pitch, T, rg, before, after);
		const int64_t cells = (rg.bin_u - rg.bin_l) * (rg.frame_a - rg.frame_b);
		heal_boost_kernel<<<grid_for(cells, device), 256, 0, st>>>(S, pitch, rg, before, after, G, g0, F);
		count_launch(2);
	}
	heal_apply_kernel<<<grid_for((g1 - g0) * F, device), 256, 0, st>>>(S, pitch, G, g0, g1, F);
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

}  // namespace par
