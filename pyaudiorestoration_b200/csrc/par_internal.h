// par_internal.h -- shared declarations of libpar_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace par {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);   // records the message, returns PAR_ECUDA
void count_launch(int n = 1);
int sm_count(int device);

#define PAR_CUDA(call)                                              \
	do {                                                            \
		cudaError_t _e = (call);                                    \
		if (_e != cudaSuccess) return par::cuda_fail(_e, #call);    \
	} while (0)

// Device-resident constant tables, cached per (device, key); built on the host in float64.
const float2 *fft_twiddles(int device, int log2m, cudaStream_t st);          // FftSched<log2m> layout
// two-level tables of the large (four-step) transform of 2^log2m complex points:
// [W_M^p, p<1024 | W_M^(1024 q), q<M/1024 | W_2M^k, k<1024 | W_2M^(1024 q), q<=M/1024]
const float2 *large_fft_tables(int device, int log2m, cudaStream_t st);
const float *device_window(int device, const float *host_window, int n, cudaStream_t st);
// sinc tap coefficients: c[k] = (-1)^(k-nt+1) * hanning(2nt+1)[k] / pi  and  hp[k] = hanning[k] / pi,
// k = 0 .. 2nt-1, each padded with zeros to a multiple of 16 (+16)
struct SincTables {
	const float *c;    // fc == 1 path
	const float *hp;   // fc < 1 path
	int padded;        // entries per table
};
int sinc_tables(int device, int nt, cudaStream_t st, SincTables *out);

// ---- kernel launchers (all asynchronous on `st`, device pointers only) -----------------------
struct StftArgs {
	const float *x;        // sample `x_origin` of every channel (shards hold a slice of the signal)
	int64_t n, x_stride, x_ch_stride;   // n = GLOBAL length (reflection happens at 0 and n)
	int64_t x_origin;
	int n_ch, n_fft, hop, zeropad;
	int64_t n_frames;      // frames handled by this launch (per channel) ...
	int64_t frame0;        // ... starting at this global frame index
	const float *window;   // device
	void *out;
	int64_t out_pitch, out_ch_stride;
	int magnitude;
};
int launch_stft(const StftArgs &a, int device, cudaStream_t st);
int launch_stft_mid(const StftArgs &a, int log2m, int device, cudaStream_t st);     // log2m in {13, 14}
int launch_stft_large(const StftArgs &a, int log2m, int device, cudaStream_t st);   // 15 <= log2m <= 19

struct IstftArgs {
	const float2 *S;
	int n_fft;
	int64_t n_frames, s_pitch, s_ch_stride;
	int n_ch, hop;
	const float *window;   // device
	int64_t start, length;
	float *y;
	int64_t y_stride, y_ch_stride;
	float *frames;         // device scratch: n_ch * n_frames * n_fft floats
};
int launch_istft(const IstftArgs &a, int device, cudaStream_t st);
bool istft_needs_scratch(int n_fft);      // only the two-pass path (n_fft > 8192) uses IstftArgs::frames

// STFT-domain mask operators (spectral.cu); S is a device spectrogram, frames `pitch` float2 apart, bins contiguous
int launch_spec_gate(float2 *S, int64_t cells, int F, const double *thr_db_dev, double gain_db, int device, cudaStream_t st);
int launch_spec_select(const float2 *L, const float2 *R, int64_t cells, float2 *out_max, float2 *out_min, int device,
                       cudaStream_t st);
int launch_spec_heal(float2 *S, int64_t pitch, int64_t T, int F, const int64_t *regions, int64_t n_regions, int64_t g0,
                     int64_t g1, double *scratch_dev, int device, cudaStream_t st);

int launch_quotient_selftest(int64_t max_n, unsigned long long *mismatches_dev, cudaStream_t st);
int launch_segment_sums(const double *speeds_dev, const int64_t *seg_n_dev, int64_t n_seg,
                        double *sums_dev, cudaStream_t st);
int launch_expand_positions(const double *speeds_dev, const int64_t *seg_n_dev,
                            const int64_t *seg_start_dev, const double *seg_offset_dev,
                            int64_t n_seg, double *pos_dev, int64_t m, cudaStream_t st,
                            double *sums_out_dev = nullptr);
int launch_add_offsets(const int64_t *seg_n_dev, const int64_t *seg_start_dev, const double *seg_offset_dev,
                       int64_t n_seg, double *pos_dev, int64_t m, cudaStream_t st);

struct SincArgs {
	const double *pos;
	int64_t m;
	const float *signal;
	int64_t n_in, sig_stride, sig_ch_stride;
	int n_ch, nt;
	float *out;
	int64_t out_stride, out_ch_stride;
	int aligned_edges;
	int kernel = 0;               // 0: by tap count, 1: two-CTA tiled kernel, 2: warp-specialised kernel
	double period_dev = -1.0;     // largest |read period - 1| the caller expects (tile sizing of the resampler); < 0: unknown
	int64_t out_begin, out_end;   // output range handled by this launch, [0, m) for all of it
	// shards: pos[0] is position `pos_origin`, signal[0] is sample `sig_origin`, out[0] is output `out_origin`
	int64_t pos_origin, sig_origin, out_origin;
};
int launch_sinc(const SincArgs &a, int device, cudaStream_t st);
int launch_linear(const SincArgs &a, int device, cudaStream_t st);
int launch_trace(const float *mag_dev, int64_t pitch, int num_bins, int64_t frame0, int64_t count, int fft_size,
                 double sr, double tolerance_octaves, int mode, double first_freq, double *freqs_dev,
                 cudaStream_t st);
// dst[c * dst_ch_stride + i] = src[i * stride + c * ch_stride]
int launch_deinterleave(const float *src, int64_t n, int64_t stride, int n_ch, int64_t ch_stride, float *dst,
                        int64_t dst_ch_stride, int device, cudaStream_t st);

}  // namespace par
