/*
 * par_b200.h -- C ABI of libpar_b200.so: the B200-native STFT + varispeed-resample hot path
 * of HENDRIX-ZT2/pyaudiorestoration (util/fourier.py, util/resampling.py).
 *
 * The reference is pure Python and has NO FFI of its own: its boundary is the
 * module surface util.fourier.* / util.resampling.run.  Every entry point below names the
 * reference function (file:line, relative to the reference root) whose work it replaces; the
 * Python mirror of that surface (pyaudiorestoration_b200/util/{fourier,resampling}.py) is a thin
 * ctypes caller of this header.  INTEGRATION.md shows the two-line stubs a maintainer drops into
 * the reference's util/ package.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - Every function returns 0 (PAR_OK) or a negative PAR_E* code; the message of the last error
 *    on the calling thread is par_last_error().  No exceptions cross the boundary.
 *  - The caller owns all data buffers.  Without PAR_DEVICE_PTRS the data pointers are HOST
 *    memory (pinned or pageable): the call stages through device memory, runs, copies back and
 *    returns when the result is in the host buffer.  With PAR_DEVICE_PTRS they are device
 *    pointers on `device`, work is enqueued on `stream` (a cudaStream_t, NULL = default stream)
 *    and the call returns without synchronising.
 *  - Small parameter arrays (window, speed curve) are always HOST pointers.
 *  - Every entry sets the CUDA device itself and is re-entrant (the reference calls this path
 *    from QThread workers, util/qt_threads.py:19-35); caches are mutex-guarded.
 *  - There is no CPU fallback: without a usable CUDA device every compute entry fails with
 *    PAR_ECUDA.
 *
 * Layouts
 *  - Audio is float32.  A channel is `n` samples `stride` elements apart; channel c starts
 *    `c * ch_stride` elements after the base pointer (planar: stride 1, ch_stride >= n;
 *    the reference's interleaved (frames, channels) arrays: stride C, ch_stride 1).
 *  - A spectrogram is the memory image of the reference's F-ordered (F, T) array: frame t of
 *    channel c starts at  c * out_ch_stride + t * out_pitch  elements (complex64 = 2 floats, or
 *    float32 magnitudes), bins contiguous; F = n_fft*zeropad/2 + 1, out_pitch >= F.
 */
#ifndef PAR_B200_H
#define PAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PAR_API __attribute__((visibility("default")))
#else
#define PAR_API
#endif

#define PAR_OK 0
#define PAR_EINVAL (-1)      /* bad argument */
#define PAR_ECUDA (-2)       /* CUDA runtime error / no device */
#define PAR_EUNSUPPORTED (-3)/* valid request this build cannot run (e.g. n_fft not a power of two) */
#define PAR_ECAPACITY (-4)   /* output buffer too small; required size returned where documented */

/* flags */
#define PAR_DEVICE_PTRS (1u << 0)   /* data pointers are device memory; async on `stream` */
#define PAR_OUT_MAGNITUDE (1u << 1) /* par_stft_f32: write float32 |S| + 1e-7 instead of complex64 */
#define PAR_SINC_ALIGNED_EDGES (1u << 2) /* par_sinc_resample_f32: do NOT reproduce the reference's
                                            start-edge tap misalignment */
#define PAR_SINC_KERNEL_TILED (1u << 3) /* resamplers: force the two-CTA-per-SM kernel (default below 64 taps) */
#define PAR_SINC_KERNEL_WS (1u << 4)    /* resamplers: force the warp-specialised kernel (default from 64 taps);
                                           both kernels produce identical bits, the flags exist for A/B tests */

/* ---- diagnostics ------------------------------------------------------------------------- */
PAR_API const char *par_last_error(void);
PAR_API const char *par_version(void);
PAR_API int par_device_count(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
PAR_API int64_t par_kernel_launch_count(void);
/* CUDA-event time of the last compute kernel sequence of a HOST-pointer call, in ms, on the
 * calling thread (0 if none); informational. */
PAR_API double par_last_kernel_ms(void);

/* Device self-test of the positions kernels' exact quotient j/(n-1) (three FMAs instead of a
 * division, csrc/resample.cu SegDiv): compares it with the IEEE division for every j < n,
 * n = 2 .. max_n.  Returns the number of mismatches (0 expected), -1 on error. */
PAR_API int64_t par_selftest_positions_quotient(int64_t max_n, int device);

/* Scratch buffers come from the device's stream-ordered memory pool and stay cached there between
 * calls (a 10-minute stereo job keeps ~3 GB); this hands the cached blocks back to the driver. */
PAR_API int par_release_cached_memory(int device);

/* Pinned host allocations (the Python layer returns ndarrays backed by these so that the
 * device->host copy of a result runs at full PCIe rate). */
PAR_API void *par_host_alloc(int64_t bytes);
PAR_API void par_host_free(void *p);

/* ---- STFT: util/fourier.py:37-75 stft, :78-82 estimate_and_center, :160-166 segment_array,
 *      :124-157 pyfftw_rfft2 / np_rfft_pick, :92-121 torch_rfft2, :23-29 to_mag / get_mag ------
 * Number of frames of the centred transform (util/fourier.py:81): n // hop + 1 for even n_fft. */
PAR_API int64_t par_stft_num_frames(int64_t n, int n_fft, int hop);

/* Fused reflect-pad + frame gather + window + real FFT of length n_fft*zeropad (frame
 * left-aligned, zeros appended) + 1/sqrt(n_fft) scaling, for n_ch channels in one launch.
 * window: HOST float32[n_fft].  out: complex64 (or float32 with PAR_OUT_MAGNITUDE).
 * n_fft*zeropad must be a power of two in [32, 1048576] (above 32768: four-step path through an
 * L2-resident scratch, unit-stride channels only); n >= 1. */
PAR_API int par_stft_f32(const float *x, int64_t n, int64_t x_stride, int n_ch, int64_t x_ch_stride,
                 int n_fft, int hop, int zeropad, const float *window,
                 void *out, int64_t out_pitch, int64_t out_ch_stride,
                 unsigned flags, int device, void *stream);

/* ---- iSTFT: util/fourier.py:314-437 istft (+ :677-687 __overlap_add, :481-546 window_sumsquare)
 * S: complex64 frames (layout above, F = n_fft/2+1 bins, n_frames frames per channel), already
 * trimmed to the frames the reference would use (:373-381).  Computes
 *   y[j] = sum_t w[j-t*hop] * irfft(S[:,t] * sqrt(n_fft))[j-t*hop] / sum_t w[j-t*hop]^2
 * over the padded timeline, skips `start` samples (n_fft/2 when centred) and writes `length`
 * samples (zero-filled past the end of the timeline).  window: HOST float32[n_fft]. */
PAR_API int par_istft_f32(const void *S, int n_fft, int64_t n_frames, int64_t s_pitch, int n_ch,
                  int64_t s_ch_stride, int hop, const float *window, int64_t start,
                  int64_t length, float *y, int64_t y_stride, int64_t y_ch_stride,
                  unsigned flags, int device, void *stream);

/* ---- stft -> STFT-domain mask -> istft, the spectrogram never leaving the device --------------------------------
 * The bodies the reference's tools run between util.fourier.stft and util.fourier.istft on a signal padded by
 * fix_length(signal, n + n_fft // 2) (dropout_healer_gui.py:129-164, dropouts_gui.py:148-159,
 * renoiser_gui.py:310-317):  y = istft(op(stft(pad(x))), length = n, hop_length = hop), one upload and one
 * download per call.  window / syn_window: HOST float32[n_fft] (analysis / synthesis; the tools use
 * blackmanharris for both).  n_fft: a power of two in [32, 32768].  Operators:
 *   PAR_SPEC_GATE         renoiser_gui.py:273-278   params = float64 profile_db[n_fft/2+1], n_params = n_fft/2+1:
 *                         S[f, t] *= 10^(gain_db/20) wherever 20 log10(|S| + 1e-7) <= profile_db[f]; n_ch outputs
 *   PAR_SPEC_SELECT_MAX / _MIN / _BOTH   dropouts_gui.py:153-161   n_ch must be 2: per cell the louder (quieter)
 *                         of the two channels; 1 output channel (BOTH: 2 -- max, then min)
 *   PAR_SPEC_HEAL         dropout_healer_gui.py:134-162   params = int64 regions[n_params][5] = (frame_b, frame_a,
 *                         frame_surrounding, bin_l, bin_u) per marker (:136-141), applied in order: mean dB level of the
 *                         surrounding frames before / after the gap, bilinear blend across it, boost clipped to
 *                         [earlier boost, 255] dB; n_ch outputs
 * y: n samples per output channel (channel c at + c * y_ch_stride, element stride y_stride). */
#define PAR_SPEC_GATE 0
#define PAR_SPEC_SELECT_MAX 1
#define PAR_SPEC_SELECT_MIN 2
#define PAR_SPEC_SELECT_BOTH 3
#define PAR_SPEC_HEAL 4
PAR_API int par_spectral_process_f32(const float *x, int64_t n, int64_t x_stride, int n_ch, int64_t x_ch_stride,
                             int n_fft, int hop, const float *window, const float *syn_window, int op,
                             const void *params, int64_t n_params, double gain_db, float *y, int64_t y_stride,
                             int64_t y_ch_stride, unsigned flags, int device, void *stream);

/* ---- positions: util/resampling.py:93-137 speed_to_pos --------------------------------------
 * Host-only, serial, bit-exact: the error-diffused integer segment lengths (:111-118).
 * seg_n: int64[k-1].  Returns sum(seg_n) in *total. */
PAR_API int par_speed_segments(const double *sampletimes, const double *speeds, int64_t k,
                       int64_t *seg_n, int64_t *total);

/* Expands the speed curve into float64 read positions with the reference's operation order
 * (per-segment sequential cumsum of 1/speed, carried offset, end test + argmin trim).
 * sampletimes/speeds: HOST float64[k].  pos: capacity `cap` doubles (host, or device with
 * PAR_DEVICE_PTRS); cap >= sum(seg_n) always suffices.  *m receives the number of valid
 * positions (the reference's filled prefix).  The expansion runs on the GPU;
 * the call synchronises `stream` internally (it needs the per-segment sums on the host). */
PAR_API int par_speed_to_pos_f64(const double *sampletimes, const double *speeds, int64_t k,
                         double num_input_samples, double *pos, int64_t cap, int64_t *m,
                         unsigned flags, int device, void *stream);

/* ---- resampler: util/resampling.py:51-90 sinc_core, :21-46 sinc_wrapper(_mt) ----------------
 * out[i] = sum_k signal[lower+k] * fc*sinc((k-NT-shift)*fc) * hanning(2NT+1)[k] with the
 * reference's index rules (half-even rounding, +NT tap dropped, fc = min(1/period, 1), the last
 * element reuses the previous period; start-edge misalignment unless PAR_SINC_ALIGNED_EDGES).
 * pos: float64[m], shared by all channels.  1 <= nt <= 512. */
PAR_API int par_sinc_resample_f32(const double *pos, int64_t m, const float *signal, int64_t n_in,
                          int64_t sig_stride, int n_ch, int64_t sig_ch_stride, int nt,
                          float *out, int64_t out_stride, int64_t out_ch_stride,
                          unsigned flags, int device, void *stream);

/* "Linear" mode of run: util/resampling.py:228-229 (np.interp, left = right = 0). */
PAR_API int par_linear_resample_f32(const double *pos, int64_t m, const float *signal, int64_t n_in,
                            int64_t sig_stride, int n_ch, int64_t sig_ch_stride,
                            float *out, int64_t out_stride, int64_t out_ch_stride,
                            unsigned flags, int device, void *stream);

/* ---- speed curve -> resampled audio in one call: the "Preparing" + "Resampling" phases of
 *      util/resampling.py:162-231 (run) with speed_curve given.  Positions are expanded on the
 *      device and never leave it.  out holds out_cap samples per channel; *m receives the number
 *      of samples written per channel (PAR_ECAPACITY, with *m set, if out_cap is too small;
 *      sum(par_speed_segments) always suffices).  mode: PAR_MODE_LINEAR | PAR_MODE_SINC. */
#define PAR_MODE_LINEAR 0
#define PAR_MODE_SINC 1
PAR_API int par_varispeed_f32(const double *sampletimes, const double *speeds, int64_t k,
                      const float *signal, int64_t n_in, int64_t sig_stride, int n_ch, int64_t sig_ch_stride,
                      int mode, int nt, float *out, int64_t out_cap, int64_t out_stride, int64_t out_ch_stride,
                      int64_t *m, unsigned flags, int device, void *stream);

/* ---- frequency trackers on the magnitude spectrogram: util/wow_detection.py:119-139 (get_peak),
 *      :294-302 (PeakTracker), :305-327 (PeakTrackTracker), :256-291 (CenterOfGravity) ------------
 * mag: float32 magnitudes as par_stft_f32 writes them (frame t at t*pitch, bins contiguous; host, or
 * device with PAR_DEVICE_PTRS).  freqs: HOST float64[count]; in: the drawn trail sampled at frames
 * frame0 .. frame0+count-1 (Track.sample_trail, :66-76), out: the traced frequency per frame.
 * fft_size is the transform length (n_fft*zeropad), tolerance_st the band half-width in semitones. */
#define PAR_TRACE_PEAK 0
#define PAR_TRACE_PEAK_TRACK 1
#define PAR_TRACE_COG 2
PAR_API int par_trace_f32(const float *mag, int num_bins, int64_t n_frames, int64_t pitch, int64_t frame0,
                  int64_t count, int fft_size, double sr, double tolerance_st, int mode, double *freqs,
                  unsigned flags, int device, void *stream);
/* Fused: the magnitudes of frames [frame0, frame0+count) of the transform of x are computed on the
 * device, traced there and discarded; only `count` doubles come back (x host or device as usual). */
PAR_API int par_stft_trace_f32(const float *x, int64_t n, int64_t x_stride, int n_fft, int hop, int zeropad,
                       const float *window, int64_t frame0, int64_t count, double sr, double tolerance_st,
                       int mode, double *freqs, unsigned flags, int device, void *stream);

/* ---- time-sharded jobs: one rank's slice of a long signal --------------------
 * Device pointers only (PAR_DEVICE_PTRS must be set).  Indices are GLOBAL sample / frame / output
 * numbers; the rank passes the slice it holds and where that slice starts.
 *
 * par_stft_range_f32: frames [frame0, frame0 + n_frames) of the transform of a signal of n_global
 * samples.  x holds samples [x_origin, x_origin + n_local) of every channel (planar, unit stride,
 * channel c at + c * x_ch_stride), which must cover [frame0*hop - n_fft/2, (frame0+n_frames-1)*hop +
 * n_fft/2) -- i.e. the rank's chunk plus a halo of up to n_fft/2 samples from each neighbour;
 * reflection (util/fourier.py:80) happens only at 0 and n_global.  Row 0 of out is frame frame0. */
PAR_API int par_stft_range_f32(const float *x, int64_t n_local, int64_t x_origin, int64_t n_global, int n_ch,
                       int64_t x_ch_stride, int n_fft, int hop, int zeropad, const float *window,
                       int64_t frame0, int64_t n_frames, void *out, int64_t out_pitch, int64_t out_ch_stride,
                       unsigned flags, int device, void *stream);

/* par_speed_to_pos_range_f64: the slice of util/resampling.py:93-137's positions a time shard needs.
 * The serial segment chain is evaluated for the whole curve (it is K additions), but only the
 * segments holding positions in [lo_pos, hi_pos] -- plus the following one, for the period of the
 * last output -- are expanded.  pos[0] is output *pos_origin, *pos_count positions are written
 * (bit-identical to the same slice of par_speed_to_pos_f64), *m is the global output count. */
PAR_API int par_speed_to_pos_range_f64(const double *sampletimes, const double *speeds, int64_t k,
                               double num_input_samples, double lo_pos, double hi_pos,
                               double *pos, int64_t cap, int64_t *pos_origin, int64_t *pos_count,
                               int64_t *m, unsigned flags, int device, void *stream);

/* par_segment_sums_f64 + par_speed_to_pos_range_sums_f64: the same slice, with the per-segment totals of the
 * cumsum of 1/speed (util/resampling.py:120-126) shared between the ranks of a time-sharded job.  Every rank
 * computes the totals of segments [seg_begin, seg_end) (sums[0] is segment seg_begin; device memory), the ranks
 * all-gather their slices (one small collective, done by the caller) and pass the totals of ALL k-1 segments as
 * seg_sums: no rank then repeats the divisions of the whole curve, and only the window's rows of the segment
 * tables are uploaded.  seg_n_out (HOST int64[k-1], may be NULL) receives the error-diffused segment lengths of the
 * whole curve (par_speed_segments); passing them back as seg_n (may be NULL) saves the second evaluation of that
 * serial recurrence.  Results are bit-identical to par_speed_to_pos_range_f64. */
PAR_API int par_segment_sums_f64(const double *sampletimes, const double *speeds, int64_t k,
                         int64_t seg_begin, int64_t seg_end, double *sums, int64_t *seg_n_out, unsigned flags,
                         int device, void *stream);
PAR_API int par_speed_to_pos_range_sums_f64(const double *sampletimes, const double *speeds, int64_t k,
                                    double num_input_samples, double lo_pos, double hi_pos, const double *seg_sums,
                                    const int64_t *seg_n, double *pos, int64_t cap, int64_t *pos_origin,
                                    int64_t *pos_count, int64_t *m, unsigned flags, int device, void *stream);

/* par_resample_range_f32: outputs [out_begin, out_end) of util/resampling.py:51-90 / :228-229 over
 * m_global read positions.  pos holds positions [pos_origin, pos_origin + pos_count) and must reach
 * out_end (the period of the last output) unless out_end == m_global; signal holds samples
 * [sig_origin, sig_origin + sig_count) of every channel (planar) and must cover every tap the
 * outputs read (round(pos) +- nt, clamped to [0, n_in_global)).  out[0] is output out_begin. */
PAR_API int par_resample_range_f32(const double *pos, int64_t pos_origin, int64_t pos_count, int64_t m_global,
                           int64_t out_begin, int64_t out_end, const float *signal, int64_t sig_origin,
                           int64_t sig_count, int64_t n_in_global, int n_ch, int64_t sig_ch_stride, int mode, int nt,
                           float *out, int64_t out_stride, int64_t out_ch_stride,
                           unsigned flags, int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PAR_B200_H */
