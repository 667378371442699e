// track.cu -- frequency trackers on device-resident magnitude spectrograms.
//
// Replaces the per-frame loops of util/wow_detection.py: Track.get_peak / is_peak (:119-139) with
// util/correlation.py:42-46 (parabolic), PeakTracker.trace (:294-302), PeakTrackTracker.trace
// (:305-327) and CenterOfGravity.COG / trace (:256-291).  The spectrogram never leaves the GPU: the
// trackers read the band [NL, NU) of each frame in place (layout of stft.cu: frame t at t*pitch,
// bins contiguous) and return one frequency per frame.
//
// Peak and Peak Track are frame-parallel (the band of a frame depends only on the drawn trail, resp.
// on its first point): one warp per frame, lanes stride the band for the first maximum, lane 0
// refines it.  Center of Gravity is serial in time (the band follows the previous result): one warp
// walks the frames.  Arithmetic follows the reference's dtypes: band limits in float64, the parabolic
// refinement and bin -> Hz conversion of a float32 spectrogram in float32, the centre of gravity in
// float64.
#include "par_internal.h"
#include "../../include/par_b200.h"

namespace par {

struct Band {
	int nl, nu;
};

// util/wow_detection.py:79-80 freq_2_bin, :91-104 set_bin_limits, :106-116 freq_plus_tolerance
__device__ __forceinline__ int freq_2_bin(double f, int num_bins, int fft_size, double sr) {
	const double b = rint(__ddiv_rn(__dmul_rn(f, (double)fft_size), sr));       // Python round(): half to even
	int v = b > 2.0e9 ? 2000000000 : (b < -2.0e9 ? -2000000000 : (int)b);
	v = v < num_bins - 1 ? v : num_bins - 1;
	return v > 1 ? v : 1;
}

__device__ __forceinline__ Band band_around(double freq, double tolerance, int num_bins, int fft_size, double sr,
                                            int min_bins) {
	const double lf = log2(freq);
	double fl = exp2(lf - tolerance), fu = exp2(lf + tolerance);
	fl = fl > 1.0 ? fl : 1.0;
	fu = fu < sr / 2 ? fu : sr / 2;
	Band b;
	b.nl = freq_2_bin(fl, num_bins, fft_size, sr);
	b.nu = freq_2_bin(fu, num_bins, fft_size, sr);
	while (b.nu - b.nl < min_bins) { b.nl -= 1; b.nu += 1; }
	// the reference would slice with a negative start here (wrapping around); keep the band inside the frame
	if (b.nl < 0) b.nl = 0;
	if (b.nu > num_bins) b.nu = num_bins;
	return b;
}

// first maximum of frame[nl:nu] over the warp (np.argmax), NaN-free input assumed
__device__ __forceinline__ int warp_argmax(const float *__restrict__ frame, int nl, int nu, int lane) {
	float best = -INFINITY;
	int arg = 0x7fffffff;
	for (int b = nl + lane; b < nu; b += 32) {
		const float v = __ldg(frame + b);
		if (v > best) { best = v; arg = b; }
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const float ov = __shfl_xor_sync(0xffffffffu, best, o);
		const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
		if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
	}
	return arg == 0x7fffffff ? nl : arg;
}

// util/wow_detection.py:119-139 get_peak (rectangular window) in the dtypes numpy uses for a float32 frame
__device__ __forceinline__ double refine_peak(const float *__restrict__ frame, int p, int num_bins, int fft_size,
                                              double sr) {
	if (p >= 1 && p + 1 < num_bins) {
		const float fm = __ldg(frame + p - 1), f0 = __ldg(frame + p), fp = __ldg(frame + p + 1);
		if (fm < f0 && f0 > fp) {
			// xv = 1/2. * (f[x-1] - f[x+1]) / (f[x-1] - 2*f[x] + f[x+1]) + x: the quotient is float32
			// arithmetic on float32 bins, "+ x" (a numpy int64) and bin_2_freq promote to float64
			const float num = __fmul_rn(0.5f, __fsub_rn(fm, fp));
			const float den = __fadd_rn(__fsub_rn(fm, __fmul_rn(2.0f, f0)), fp);
			const double xv = __dadd_rn((double)__fdiv_rn(num, den), (double)p);
			return __dmul_rn(__ddiv_rn(xv, (double)fft_size), sr);
		}
	}
	return __dmul_rn(__ddiv_rn((double)p, (double)fft_size), sr);
}

struct TraceArgs {
	const float *mag;      // frame t at mag + t * pitch
	int64_t pitch, frame0, count;
	int num_bins, fft_size, mode;
	double sr, tolerance;  // tolerance in octaves (semitones / 12)
	double *freqs;         // in: trail sampled per frame; out: traced frequency
};

__global__ void __launch_bounds__(256)
trace_parallel_kernel(TraceArgs a, double first_freq) {
	const int lane = threadIdx.x & 31;
	const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
	if (i >= a.count) return;
	double centre = a.freqs[i], tol = a.tolerance;
	if (a.mode == PAR_TRACE_PEAK_TRACK) {
		centre = first_freq;                      // the reference never advances `freq` (:316-327)
		if (i > 2) tol = a.tolerance / 2;
	}
	const Band b = band_around(centre, tol, a.num_bins, a.fft_size, a.sr, 4);
	const float *frame = a.mag + (a.frame0 + i) * a.pitch;
	const int p = warp_argmax(frame, b.nl, b.nu, lane);
	if (lane == 0) a.freqs[i] = refine_peak(frame, p, a.num_bins, a.fft_size, a.sr);
}

// util/wow_detection.py:256-291: hann-weighted centre of gravity of log2(f) over the band, band
// re-centred on every result
__global__ void __launch_bounds__(32)
trace_cog_kernel(TraceArgs a) {
	const int lane = threadIdx.x;
	Band b = band_around(a.freqs[0], a.tolerance, a.num_bins, a.fft_size, a.sr, 4);
	for (int64_t i = 0; i < a.count; i++) {
		const float *frame = a.mag + (a.frame0 + i) * a.pitch;
		const int n = b.nu - b.nl;
		double num = 0.0, den = 0.0;
		for (int k = lane; k < n; k += 32) {
			const double h = n > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)k / (double)(n - 1)) : 1.0;      // np.hanning(n)
			const double w = h * (double)__ldg(frame + b.nl + k);
			const double f = (double)(b.nl + k) / (double)a.fft_size * a.sr;                          // fft_freqs
			num += w * log2(f);
			den += w;
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			num += __shfl_xor_sync(0xffffffffu, num, o);
			den += __shfl_xor_sync(0xffffffffu, den, o);
		}
		const double freq = exp2(num / den);
		if (lane == 0) a.freqs[i] = freq;
		// an all-zero band gives NaN (the reference would raise in int(round(nan))): keep the band, the
		// host interpolates NaNs afterwards like interp_nans (:19-22)
		if (freq > 0.0 && freq < 1e12) b = band_around(freq, a.tolerance, a.num_bins, a.fft_size, a.sr, 4);
	}
}

int launch_trace(const float *mag_dev, int64_t pitch, int num_bins, int64_t frame0, int64_t count, int fft_size,
                 double sr, double tolerance_octaves, int mode, double first_freq, double *freqs_dev,
                 cudaStream_t st) {
	if (count <= 0) return PAR_OK;
	TraceArgs a;
	a.mag = mag_dev; a.pitch = pitch; a.frame0 = frame0; a.count = count; a.num_bins = num_bins;
	a.fft_size = fft_size; a.mode = mode; a.sr = sr; a.tolerance = tolerance_octaves; a.freqs = freqs_dev;
	if (mode == PAR_TRACE_COG) {
		trace_cog_kernel<<<1, 32, 0, st>>>(a);
	} else {
		const int64_t blocks = (count * 32 + 255) / 256;
		trace_parallel_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, first_freq);
	}
	count_launch();
	PAR_CUDA(cudaGetLastError());
	return PAR_OK;
}

}  // namespace par
