// api.cu -- the C ABI of libpar_b200.so (include/par_b200.h): argument checking, host<->device
// staging for HOST-pointer calls, cached constant tables, and the serial host part of
// speed_to_pos.  No CPU implementation of any kernel lives here: without a CUDA device every
// compute entry fails with PAR_ECUDA.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/par_b200.h"
#include "fft_core.cuh"
#include "par_internal.h"

namespace par {

static thread_local std::string g_err;
static thread_local double g_last_ms = 0.0;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_mu;

void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what) {
	g_err = std::string(what) + ": " + cudaGetErrorString(e);
	cudaGetLastError();   // clear sticky-less errors
	return PAR_ECUDA;
}
void count_launch(int n) { g_launches += n; }

int sm_count(int device) {
	static int cache[64] = {0};
	if (device >= 0 && device < 64 && cache[device]) return cache[device];
	int n = 0;
	if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
	if (device >= 0 && device < 64) cache[device] = n;
	return n;
}

// ---- cached device tables ---------------------------------------------------------------------
struct TableKey {
	int device, kind, a;
	uint64_t hash;
	bool operator<(const TableKey &o) const {
		if (device != o.device) return device < o.device;
		if (kind != o.kind) return kind < o.kind;
		if (a != o.a) return a < o.a;
		return hash < o.hash;
	}
};
static std::map<TableKey, void *> g_tables;

static void *upload_table(const TableKey &key, const void *host, size_t bytes, cudaStream_t st) {
	// caller holds g_mu
	void *d = nullptr;
	if (cudaMalloc(&d, bytes) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc(table)"); return nullptr; }
	// synchronous copy from pageable memory: the table is complete before any kernel can use it
	if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
		cuda_fail(cudaGetLastError(), "cudaMemcpy(table)");
		cudaFree(d);
		return nullptr;
	}
	(void)st;
	g_tables[key] = d;
	return d;
}

template <int LOG2M>
static void build_twiddles(std::vector<float2> &tw) {
	using S = FftSched<LOG2M>;
	tw.assign(S::TW_TOTAL, make_float2(0.f, 0.f));
	for (int p = 1; p < S::NP; p++) {
		const int r = 1 << S::bits(p), ns = 1 << S::ns_log2(p);
		for (int t = 1; t < r; t++)
			for (int k = 0; k < ns; k++) {
				const double ang = -2.0 * M_PI * (double)t * (double)k / ((double)ns * (double)r);
				tw[S::tw_offset(p) + (t - 1) * ns + k] = make_float2((float)cos(ang), (float)sin(ang));
			}
	}
	for (int k = 0; k <= S::M / 2; k++) {
		const double ang = -M_PI * (double)k / (double)S::M;
		tw[S::TW_SPLIT_OFFSET + k] = make_float2((float)cos(ang), (float)sin(ang));
	}
}

const float2 *fft_twiddles(int device, int log2m, cudaStream_t st) {
	std::lock_guard<std::mutex> lk(g_mu);
	TableKey key{device, 1, log2m, 0};
	auto it = g_tables.find(key);
	if (it != g_tables.end()) return (const float2 *)it->second;
	std::vector<float2> tw;
	switch (log2m) {
	case 4: build_twiddles<4>(tw); break;
	case 5: build_twiddles<5>(tw); break;
	case 6: build_twiddles<6>(tw); break;
	case 7: build_twiddles<7>(tw); break;
	case 8: build_twiddles<8>(tw); break;
	case 9: build_twiddles<9>(tw); break;
	case 10: build_twiddles<10>(tw); break;
	case 11: build_twiddles<11>(tw); break;
	case 12: build_twiddles<12>(tw); break;
	case 13: build_twiddles<13>(tw); break;
	case 14: build_twiddles<14>(tw); break;
	default: set_error("fft_twiddles: unsupported size"); return nullptr;
	}
	return (const float2 *)upload_table(key, tw.data(), tw.size() * sizeof(float2), st);
}

const float2 *large_fft_tables(int device, int log2m, cudaStream_t st) {
	std::lock_guard<std::mutex> lk(g_mu);
	TableKey key{device, 4, log2m, 0};
	auto it = g_tables.find(key);
	if (it != g_tables.end()) return (const float2 *)it->second;
	const int64_t M = (int64_t)1 << log2m;
	const int64_t n_hi = M >> 10, n_shi = (M >> 10) + 1;     // the split twiddle is used for every k < M
	std::vector<float2> tab(1024 + n_hi + 1024 + n_shi);
	for (int64_t p = 0; p < 1024; p++) {
		const double ang = -2.0 * M_PI * (double)p / (double)M;
		tab[p] = make_float2((float)cos(ang), (float)sin(ang));
		const double ang2 = -M_PI * (double)p / (double)M;
		tab[1024 + n_hi + p] = make_float2((float)cos(ang2), (float)sin(ang2));
	}
	for (int64_t q = 0; q < n_hi; q++) {
		const double ang = -2.0 * M_PI * (double)(q << 10) / (double)M;
		tab[1024 + q] = make_float2((float)cos(ang), (float)sin(ang));
	}
	for (int64_t q = 0; q < n_shi; q++) {
		const double ang = -M_PI * (double)(q << 10) / (double)M;
		tab[2048 + n_hi + q] = make_float2((float)cos(ang), (float)sin(ang));
	}
	return (const float2 *)upload_table(key, tab.data(), tab.size() * sizeof(float2), st);
}

static uint64_t fnv1a(const void *p, size_t n) {
	const unsigned char *b = (const unsigned char *)p;
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

const float *device_window(int device, const float *host_window, int n, cudaStream_t st) {
	std::lock_guard<std::mutex> lk(g_mu);
	TableKey key{device, 2, n, fnv1a(host_window, (size_t)n * sizeof(float))};
	auto it = g_tables.find(key);
	if (it != g_tables.end()) return (const float *)it->second;
	return (const float *)upload_table(key, host_window, (size_t)n * sizeof(float), st);
}

int sinc_tables(int device, int nt, cudaStream_t st, SincTables *out) {
	std::lock_guard<std::mutex> lk(g_mu);
	const int padded = ((2 * nt + 15) / 16) * 16 + 16;
	TableKey key{device, 3, nt, 0};
	auto it = g_tables.find(key);
	const float *base;
	if (it != g_tables.end()) {
		base = (const float *)it->second;
	} else {
		// np.hanning(2nt+1) rounded to float32 (util/resampling.py:24,36), then /pi and the
		// alternating sign of sin(pi (d - s)) folded in, in float64, rounded once
		std::vector<float> tab(2 * (size_t)padded, 0.f);
		const int mm = 2 * nt + 1;
		for (int k = 0; k < 2 * nt; k++) {
			const double nn = (double)(1 - mm + 2 * k);
			const float h = (float)(0.5 + 0.5 * cos(M_PI * nn / (double)(mm - 1)));
			const int d = k - nt;
			const double sign = ((d + 1) & 1) ? -1.0 : 1.0;
			tab[k] = (float)(sign * (double)h / M_PI);
			tab[padded + k] = (float)((double)h / M_PI);
		}
		base = (const float *)upload_table(key, tab.data(), tab.size() * sizeof(float), st);
		if (!base) return PAR_ECUDA;
	}
	out->c = base;
	out->hp = base + padded;
	out->padded = padded;
	return PAR_OK;
}

// ---- helpers ------------------------------------------------------------------------------------
static int use_device(int device) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		cudaGetLastError();
		set_error("no usable CUDA device (libpar_b200 has no CPU fallback)");
		return PAR_ECUDA;
	}
	if (device < 0 || device >= n) { set_error("device index out of range"); return PAR_EINVAL; }
	PAR_CUDA(cudaSetDevice(device));
	// Scratch comes from the stream-ordered pool; keep freed blocks cached instead of handing them
	// back to the driver at every synchronisation (the default release threshold is 0, which turns
	// each host-pointer call into a fresh multi-GB allocation).
	static std::atomic<uint64_t> pool_ready{0};
	if (device < 64 && !(pool_ready.load() & (1ull << device))) {
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			uint64_t keep = UINT64_MAX;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		cudaGetLastError();
		pool_ready.fetch_or(1ull << device);
	}
	return PAR_OK;
}

// RAII device buffer from the stream-ordered pool
struct DevBuf {
	void *p = nullptr;
	cudaStream_t st;
	explicit DevBuf(cudaStream_t s) : st(s) {}
	int alloc(size_t bytes) {
		if (bytes == 0) bytes = 16;
		PAR_CUDA(cudaMallocAsync(&p, bytes, st));
		return PAR_OK;
	}
	~DevBuf() { if (p) cudaFreeAsync(p, st); }
	template <class T> T *as() { return (T *)p; }
};

// $PAR_B200_TRACE=1: host-side time stamps of a call's phases on stderr (development aid)
struct CallTrace {
	bool on;
	const char *name;
	std::chrono::steady_clock::time_point t0;
	explicit CallTrace(const char *n) : name(n) {
		static const bool enabled = [] { const char *e = getenv("PAR_B200_TRACE"); return e && e[0] == '1'; }();
		on = enabled;
		if (on) t0 = std::chrono::steady_clock::now();
	}
	void mark(const char *what) const {
		if (!on) return;
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		fprintf(stderr, "[par trace] %s: %s +%.3f ms\n", name, what, ms);
	}
};

struct EventTimer {
	cudaEvent_t a = nullptr, b = nullptr;
	cudaStream_t st;
	explicit EventTimer(cudaStream_t s) : st(s) {
		cudaEventCreate(&a);
		cudaEventCreate(&b);
	}
	void start() { cudaEventRecord(a, st); }
	void stop() { cudaEventRecord(b, st); }
	void finish() {
		float ms = 0.f;
		if (cudaEventSynchronize(b) == cudaSuccess && cudaEventElapsedTime(&ms, a, b) == cudaSuccess) g_last_ms = ms;
	}
	~EventTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

// ---- host <-> device staging of audio ---------------------------------------------------------
// Helper streams of a host-pointer call: uploads, kernels and downloads of successive chunks
// overlap (H2D and D2H are separate copy engines; PCIe is full duplex).  Declared AFTER the
// DevBufs of a call so that it is destroyed (and its streams drained) BEFORE they are freed.
struct SideStreams {
	cudaStream_t up = nullptr, down = nullptr;
	std::vector<cudaEvent_t> evs;
	int init() {
		PAR_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
		PAR_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
		return PAR_OK;
	}
	// event recorded on `on`
	int mark(cudaStream_t on, cudaEvent_t *e) {
		PAR_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
		evs.push_back(*e);
		PAR_CUDA(cudaEventRecord(*e, on));
		return PAR_OK;
	}
	int after(cudaStream_t waiter, cudaStream_t on) {
		cudaEvent_t e;
		int rc = mark(on, &e);
		if (rc != PAR_OK) return rc;
		PAR_CUDA(cudaStreamWaitEvent(waiter, e, 0));
		return PAR_OK;
	}
	int drain() {
		if (up) PAR_CUDA(cudaStreamSynchronize(up));
		if (down) PAR_CUDA(cudaStreamSynchronize(down));
		return PAR_OK;
	}
	~SideStreams() {
		if (up) cudaStreamSynchronize(up);
		if (down) cudaStreamSynchronize(down);
		for (cudaEvent_t e : evs) cudaEventDestroy(e);
		if (up) cudaStreamDestroy(up);
		if (down) cudaStreamDestroy(down);
	}
};

// A host channel set (n samples, element stride `stride`, channel c at +c*ch_stride), uploaded
// progressively in sample order:
//  (a) planar, one contiguous copy per channel and piece, when stride == 1;
//  (b) as contiguous pieces of the whole interleaved span when the channels interleave inside the
//      sample stride (the reference's (frames, channels) arrays and column views of them) and the
//      span is at most 4x the useful bytes -- kernels then read it with the host's strides;
//  (c) with a strided 2-D copy per channel otherwise (slow, correct).
struct DevAudio {
	float *p = nullptr;
	int64_t stride = 1, ch_stride = 0;
};

struct AudioUploader {
	const float *src = nullptr;
	int64_t n = 0, stride = 1, ch_stride = 0, n_al = 4, span = 0, done = 0;
	int n_ch = 1;
	bool interleaved = false;
	int device = 0;
	DevBuf raw, stage;
	static constexpr int64_t STAGE_FLOATS = 8ll << 20;     // 32 MB of staging for widely strided channels
	explicit AudioUploader(cudaStream_t st) : raw(st), stage(st) {}

	int init(const float *src_, int64_t n_, int64_t stride_, int n_ch_, int64_t ch_stride_) {
		src = src_; n = n_; stride = stride_; n_ch = n_ch_; ch_stride = ch_stride_;
		n_al = ((n > 0 ? n : 1) + 3) & ~(int64_t)3;
		span = n > 0 ? (n - 1) * stride + (int64_t)(n_ch - 1) * ch_stride + 1 : 0;
		interleaved = n > 0 && stride > 1 && (n_ch == 1 || (ch_stride > 0 && ch_stride < stride)) &&
		              span <= 4 * n * (int64_t)n_ch + 64;
		cudaGetDevice(&device);
		// a channel view with a wider stride (one column of an array with more than 4 channels): the strided
		// span goes up in bounded pieces and is de-interleaved on the device -- a 2-D copy with 4-byte rows
		// would issue one DMA descriptor per sample
		if (!interleaved && stride > 1) {
			int rc = stage.alloc((size_t)STAGE_FLOATS * sizeof(float));
			if (rc != PAR_OK) return rc;
		}
		return raw.alloc((size_t)((interleaved ? span + 4 : n_al * n_ch) + 4) * sizeof(float));
	}
	DevAudio view() const {
		DevAudio v;
		v.p = (float *)raw.p;
		if (interleaved) { v.stride = stride; v.ch_stride = ch_stride; }
		else { v.stride = 1; v.ch_stride = n_al; }
		return v;
	}
	// enqueue on `up` the copy of samples [done, hi)
	int upload_to(int64_t hi, cudaStream_t up) {
		if (hi > n) hi = n;
		if (hi <= done) return PAR_OK;
		float *d = (float *)raw.p;
		if (interleaved) {
			const int64_t a = done * stride;
			const int64_t b = hi == n ? span : hi * stride;
			PAR_CUDA(cudaMemcpyAsync(d + a, src + a, (size_t)(b - a) * sizeof(float), cudaMemcpyHostToDevice, up));
		} else {
			for (int c = 0; c < n_ch; c++) {
				if (stride == 1) {
					PAR_CUDA(cudaMemcpyAsync(d + c * n_al + done, src + c * ch_stride + done,
					                         (size_t)(hi - done) * sizeof(float), cudaMemcpyHostToDevice, up));
				} else {
					const int64_t piece = STAGE_FLOATS / stride > 0 ? STAGE_FLOATS / stride : 1;
					for (int64_t a0 = done; a0 < hi; a0 += piece) {
						const int64_t cnt = hi - a0 < piece ? hi - a0 : piece;
						PAR_CUDA(cudaMemcpyAsync(stage.p, src + c * ch_stride + a0 * stride,
						                         (size_t)((cnt - 1) * stride + 1) * sizeof(float), cudaMemcpyHostToDevice, up));
						int rc = launch_deinterleave(stage.as<float>(), cnt, stride, 1, 0, d + c * n_al + a0, n_al, device, up);
						if (rc != PAR_OK) return rc;
					}
				}
			}
		}
		done = hi;
		return PAR_OK;
	}
};

// Device-side image of a host output channel set.  When the channels tile the host span exactly
// (fully interleaved: ch_stride 1, stride n_ch; or one contiguous channel) the kernel writes the
// host layout and ONE contiguous copy brings it back; otherwise planar + per-channel copies.
struct DevOut {
	float *p = nullptr;
	int64_t stride = 1, ch_stride = 0;
	bool image = false;
};

static int alloc_out(DevBuf &buf, int64_t m, int64_t stride, int n_ch, int64_t ch_stride, cudaStream_t st, DevOut *o) {
	(void)st;
	int rc;
	const int64_t mm = m > 0 ? m : 1;
	if (n_ch > 1 && ch_stride == 1 && stride == n_ch) {
		if ((rc = buf.alloc((size_t)mm * n_ch * sizeof(float))) != PAR_OK) return rc;
		o->p = buf.as<float>(); o->stride = stride; o->ch_stride = 1; o->image = true;
		return PAR_OK;
	}
	if ((rc = buf.alloc((size_t)mm * n_ch * sizeof(float))) != PAR_OK) return rc;
	o->p = buf.as<float>(); o->stride = 1; o->ch_stride = mm; o->image = false;
	return PAR_OK;
}

// copy output samples [o0, o1) of every channel back to the host
static int download_out(const DevOut &o, float *dst, int64_t o0, int64_t o1, int64_t stride, int n_ch,
                        int64_t ch_stride, cudaStream_t st) {
	if (o1 <= o0) return PAR_OK;
	const int64_t cnt = o1 - o0;
	if (o.image) {
		PAR_CUDA(cudaMemcpyAsync(dst + o0 * n_ch, o.p + o0 * n_ch, (size_t)cnt * n_ch * sizeof(float),
		                         cudaMemcpyDeviceToHost, st));
		return PAR_OK;
	}
	for (int c = 0; c < n_ch; c++) {
		if (stride == 1) {
			PAR_CUDA(cudaMemcpyAsync(dst + c * ch_stride + o0, o.p + c * o.ch_stride + o0, cnt * sizeof(float),
			                         cudaMemcpyDeviceToHost, st));
		} else {
			PAR_CUDA(cudaMemcpy2DAsync(dst + c * ch_stride + o0 * stride, stride * sizeof(float),
			                           o.p + c * o.ch_stride + o0, sizeof(float), sizeof(float), cnt,
			                           cudaMemcpyDeviceToHost, st));
		}
	}
	return PAR_OK;
}

// bytes of output per pipeline chunk of a host-pointer call ($PAR_B200_CHUNK_BYTES overrides; tests
// use it to push small inputs through many chunks)
// $PAR_B200_NO_RAMP=1: full-size first chunks and no early upload (A/B measurements of the host pipelines)
static bool ramp_enabled() {
	static const bool on = [] { const char *e = getenv("PAR_B200_NO_RAMP"); return !(e && e[0] == '1'); }();
	return on;
}

static int64_t chunk_bytes() {
	const char *e = getenv("PAR_B200_CHUNK_BYTES");
	if (e && *e) {
		const long long v = atoll(e);
		if (v >= 4096) return v;
	}
	return 48ll << 20;
}

}  // namespace par

using namespace par;

extern "C" {

PAR_API const char *par_last_error(void) { return g_err.c_str(); }
PAR_API const char *par_version(void) { return "par_b200 0.1 (sm_100a)"; }
PAR_API int par_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}
PAR_API int64_t par_kernel_launch_count(void) { return g_launches.load(); }
PAR_API double par_last_kernel_ms(void) { return g_last_ms; }

PAR_API int64_t par_selftest_positions_quotient(int64_t max_n, int device) {
	if (use_device(device) != PAR_OK) return -1;
	unsigned long long *d = nullptr, h = 0;
	if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); return -1; }
	cudaMemset(d, 0, sizeof(h));
	int rc = launch_quotient_selftest(max_n, d, nullptr);
	if (rc == PAR_OK && cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
		cuda_fail(cudaGetLastError(), "selftest");
		rc = PAR_ECUDA;
	}
	cudaFree(d);
	return rc == PAR_OK ? (int64_t)h : -1;
}

PAR_API void *par_host_alloc(int64_t bytes) {
	void *p = nullptr;
	if (bytes <= 0) bytes = 16;
	if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
		cuda_fail(cudaGetLastError(), "cudaHostAlloc");
		return nullptr;
	}
	return p;
}
PAR_API void par_host_free(void *p) { if (p) cudaFreeHost(p); }

PAR_API int par_release_cached_memory(int device) {
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	PAR_CUDA(cudaDeviceSynchronize());
	cudaMemPool_t pool;
	PAR_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
	PAR_CUDA(cudaMemPoolTrimTo(pool, 0));
	return PAR_OK;
}

PAR_API int64_t par_stft_num_frames(int64_t n, int n_fft, int hop) {
	if (n < 1 || n_fft < 2 || hop < 1) return 0;
	return (n + 2 * (int64_t)(n_fft / 2) - n_fft) / hop + 1;
}

PAR_API int par_stft_f32(const float *x, int64_t n, int64_t x_stride, int n_ch, int64_t x_ch_stride,
                 int n_fft, int hop, int zeropad, const float *window,
                 void *out, int64_t out_pitch, int64_t out_ch_stride,
                 unsigned flags, int device, void *stream) {
	if (!x || !out || !window) { set_error("stft: null pointer"); return PAR_EINVAL; }
	if (n < 1 || n_ch < 1 || n_fft < 2 || hop < 1 || zeropad < 1 || x_stride < 1) {
		set_error("stft: bad size argument");
		return PAR_EINVAL;
	}
	const int64_t F = (int64_t)n_fft * zeropad / 2 + 1;
	const int64_t T = par_stft_num_frames(n, n_fft, hop);
	if (out_pitch < F) { set_error("stft: out_pitch smaller than the number of bins"); return PAR_EINVAL; }
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const float *dwin = device_window(device, window, n_fft, st);
	if (!dwin) return PAR_ECUDA;
	const bool mag = flags & PAR_OUT_MAGNITUDE;
	StftArgs a;
	a.n = n; a.n_ch = n_ch; a.n_fft = n_fft; a.hop = hop; a.zeropad = zeropad; a.n_frames = T;
	a.window = dwin; a.magnitude = mag ? 1 : 0; a.frame0 = 0; a.x_origin = 0;
	if (flags & PAR_DEVICE_PTRS) {
		a.x = x; a.x_stride = x_stride; a.x_ch_stride = x_ch_stride;
		a.out = out; a.out_pitch = out_pitch; a.out_ch_stride = out_ch_stride;
		return launch_stft(a, device, st);
	}
	// host pointers: chunks of frames flow through upload -> (de-interleave) -> transform -> download
	// on three streams, so the PCIe copies in both directions overlap each other and the kernels
	const size_t esz = mag ? sizeof(float) : sizeof(float2);
	CallTrace tr("stft");
	DevBuf dplanar(st), dout(st);
	AudioUploader up(st);
	SideStreams ss;
	if ((rc = up.init(x, n, x_stride, n_ch, x_ch_stride)) != PAR_OK) return rc;
	if ((rc = dout.alloc((size_t)T * F * n_ch * esz)) != PAR_OK) return rc;
	const DevAudio da = up.view();
	if (up.interleaved && (rc = dplanar.alloc((size_t)up.n_al * n_ch * sizeof(float))) != PAR_OK) return rc;
	if ((rc = ss.init()) != PAR_OK) return rc;
	if ((rc = ss.after(ss.up, st)) != PAR_OK) return rc;       // allocations are stream-ordered on st
	if (up.interleaved) { a.x = dplanar.as<float>(); a.x_stride = 1; a.x_ch_stride = up.n_al; }
	else { a.x = da.p; a.x_stride = 1; a.x_ch_stride = da.ch_stride; }
	a.out = dout.p; a.out_pitch = F; a.out_ch_stride = T * F;
	int64_t per_chunk = chunk_bytes() / (int64_t)(F * esz * n_ch);
	if (per_chunk < 16) per_chunk = 16;
	const int64_t half = n_fft / 2;
	EventTimer tm(st);
	tm.start();
	tr.mark("allocated, streams ready");
	int64_t planar_done = 0;
	// the first chunks are small (1/8, 1/4, 1/2 of a chunk): the download engine, which bounds the call, starts after
	// 1/8 of a chunk's upload + transform instead of a whole one
	int64_t ramp = per_chunk >= 128 && ramp_enabled() ? per_chunk / 8 : per_chunk;
	for (int64_t t0 = 0, t1; t0 < T; t0 = t1, ramp = ramp * 2 < per_chunk ? ramp * 2 : per_chunk) {
		t1 = t0 + ramp < T ? t0 + ramp : T;
		// samples the frames [t0, t1) read: up to (t1-1)*hop - half + n_fft, everything once a frame
		// reaches the reflected tail (or the signal is short enough to reflect more than once)
		int64_t need = (t1 - 1) * hop - half + n_fft + 8;
		if (need >= n - 1 || n <= 2 * (int64_t)n_fft) need = n;
		if ((rc = up.upload_to(need, ss.up)) != PAR_OK) return rc;
		if ((rc = ss.after(st, ss.up)) != PAR_OK) return rc;
		if (up.interleaved && need > planar_done) {
			rc = launch_deinterleave(da.p + planar_done * da.stride, need - planar_done, da.stride, n_ch, da.ch_stride,
			                         dplanar.as<float>() + planar_done, up.n_al, device, st);
			if (rc != PAR_OK) return rc;
			planar_done = need;
		}
		a.frame0 = t0;
		a.n_frames = t1 - t0;
		if ((rc = launch_stft(a, device, st)) != PAR_OK) return rc;
		if ((rc = ss.after(ss.down, st)) != PAR_OK) return rc;
		for (int c = 0; c < n_ch; c++) {
			char *dst = (char *)out + ((size_t)c * out_ch_stride + (size_t)t0 * out_pitch) * esz;
			const char *src = (const char *)dout.p + ((size_t)c * T * F + (size_t)t0 * F) * esz;
			if (out_pitch == F) {
				PAR_CUDA(cudaMemcpyAsync(dst, src, (size_t)(t1 - t0) * F * esz, cudaMemcpyDeviceToHost, ss.down));
			} else {
				PAR_CUDA(cudaMemcpy2DAsync(dst, out_pitch * esz, src, F * esz, F * esz, t1 - t0,
				                           cudaMemcpyDeviceToHost, ss.down));
			}
		}
	}
	tm.stop();
	tr.mark("all chunks enqueued");
	if ((rc = ss.drain()) != PAR_OK) return rc;
	PAR_CUDA(cudaStreamSynchronize(st));
	tm.finish();
	tr.mark("drained");
	return PAR_OK;
}

PAR_API int par_istft_f32(const void *S, int n_fft, int64_t n_frames, int64_t s_pitch, int n_ch,
                  int64_t s_ch_stride, int hop, const float *window, int64_t start,
                  int64_t length, float *y, int64_t y_stride, int64_t y_ch_stride,
                  unsigned flags, int device, void *stream) {
	if (!S || !y || !window) { set_error("istft: null pointer"); return PAR_EINVAL; }
	const int64_t F = n_fft / 2 + 1;
	if (n_fft < 2 || (n_fft & 1) || n_frames < 1 || n_ch < 1 || hop < 1 || s_pitch < F || start < 0 ||
	    length < 0 || y_stride < 1) {
		set_error("istft: bad size argument");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const float *dwin = device_window(device, window, n_fft, st);
	if (!dwin) return PAR_ECUDA;
	DevBuf frames(st);
	if (istft_needs_scratch(n_fft) && (rc = frames.alloc((size_t)n_ch * n_frames * n_fft * sizeof(float))) != PAR_OK) return rc;
	IstftArgs a;
	a.n_fft = n_fft; a.n_frames = n_frames; a.n_ch = n_ch; a.hop = hop; a.window = dwin;
	a.start = start; a.length = length; a.frames = frames.as<float>();
	if (flags & PAR_DEVICE_PTRS) {
		a.S = (const float2 *)S; a.s_pitch = s_pitch; a.s_ch_stride = s_ch_stride;
		a.y = y; a.y_stride = y_stride; a.y_ch_stride = y_ch_stride;
		return launch_istft(a, device, st);
	}
	DevBuf ds(st), dy(st);
	DevOut dyo;
	if ((rc = ds.alloc((size_t)n_ch * n_frames * F * sizeof(float2))) != PAR_OK) return rc;
	if ((rc = alloc_out(dy, length, y_stride, n_ch, y_ch_stride, st, &dyo)) != PAR_OK) return rc;
	for (int c = 0; c < n_ch; c++) {
		const char *src = (const char *)S + (size_t)c * s_ch_stride * sizeof(float2);
		char *dst = (char *)ds.p + (size_t)c * n_frames * F * sizeof(float2);
		PAR_CUDA(cudaMemcpy2DAsync(dst, F * sizeof(float2), src, s_pitch * sizeof(float2), F * sizeof(float2),
		                           n_frames, cudaMemcpyHostToDevice, st));
	}
	a.S = ds.as<float2>(); a.s_pitch = F; a.s_ch_stride = n_frames * F;
	a.y = dyo.p; a.y_stride = dyo.stride; a.y_ch_stride = dyo.ch_stride;
	EventTimer tm(st);
	tm.start();
	if ((rc = launch_istft(a, device, st)) != PAR_OK) return rc;
	tm.stop();
	if ((rc = download_out(dyo, y, 0, length, y_stride, n_ch, y_ch_stride, st)) != PAR_OK) return rc;
	PAR_CUDA(cudaStreamSynchronize(st));
	tm.finish();
	return PAR_OK;
}

// ---- stft -> mask -> istft with the spectrogram resident on the device -------------------
PAR_API int par_spectral_process_f32(const float *x, int64_t n, int64_t x_stride, int n_ch, int64_t x_ch_stride,
                             int n_fft, int hop, const float *window, const float *syn_window, int op,
                             const void *params, int64_t n_params, double gain_db, float *y, int64_t y_stride,
                             int64_t y_ch_stride, unsigned flags, int device, void *stream) {
	if (!x || !y || !window || !syn_window) { set_error("spectral_process: null pointer"); return PAR_EINVAL; }
	if (n < 1 || n_ch < 1 || n_fft < 32 || (n_fft & (n_fft - 1)) || n_fft > 32768 || hop < 1 || x_stride < 1 || y_stride < 1) {
		set_error("spectral_process: bad size argument (n_fft: a power of two in [32, 32768])");
		return PAR_EINVAL;
	}
	const bool select = op == PAR_SPEC_SELECT_MAX || op == PAR_SPEC_SELECT_MIN || op == PAR_SPEC_SELECT_BOTH;
	if (!(op == PAR_SPEC_GATE || op == PAR_SPEC_HEAL || select) || (select && n_ch != 2) ||
	    ((op == PAR_SPEC_GATE || (op == PAR_SPEC_HEAL && n_params > 0)) && !params) || n_params < 0) {
		set_error("spectral_process: bad operator arguments");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const int64_t F = n_fft / 2 + 1;
	const int64_t n_pad = n + n_fft / 2;                      // fix_length(signal, n + n_fft // 2), util/fourier.py:440-478
	const int64_t n_pad_al = (n_pad + 3) & ~(int64_t)3;
	const int64_t T = par_stft_num_frames(n_pad, n_fft, hop);
	const int n_out = op == PAR_SPEC_SELECT_BOTH ? 2 : (select ? 1 : n_ch);
	if (op == PAR_SPEC_GATE && n_params != F) { set_error("spectral_process: the gate needs one threshold per bin"); return PAR_EINVAL; }
	int64_t g0 = 0, g1 = 0;
	if (op == PAR_SPEC_HEAL) {
		const int64_t *rg = (const int64_t *)params;
		g0 = T;
		for (int64_t r = 0; r < n_params; r++) {
			const int64_t fb = rg[5 * r], fa = rg[5 * r + 1], ar = rg[5 * r + 2], bl = rg[5 * r + 3], bu = rg[5 * r + 4];
			if (fb < 0 || fa <= fb || fa > T || ar < 1 || bl < 0 || bu - bl < 2 || bu > F) {
				set_error("spectral_process: heal region outside the spectrogram (or thinner than 2 bins / 1 frame)");
				return PAR_EINVAL;
			}
			if (fb < g0) g0 = fb;
			if (fa > g1) g1 = fa;
		}
	}
	const float *dwin = device_window(device, window, n_fft, st);
	const float *dsyn = device_window(device, syn_window, n_fft, st);
	if (!dwin || !dsyn) return PAR_ECUDA;

	// ---- padded planar input on the device
	DevBuf dx(st), dS(st), dframes(st), dy(st), dscratch(st);
	AudioUploader up(st);
	if ((rc = dx.alloc((size_t)n_ch * n_pad_al * sizeof(float))) != PAR_OK) return rc;
	PAR_CUDA(cudaMemsetAsync(dx.p, 0, (size_t)n_ch * n_pad_al * sizeof(float), st));
	if (flags & PAR_DEVICE_PTRS) {
		if ((rc = launch_deinterleave(x, n, x_stride, n_ch, x_ch_stride, dx.as<float>(), n_pad_al, device, st)) != PAR_OK) return rc;
	} else {
		if ((rc = up.init(x, n, x_stride, n_ch, x_ch_stride)) != PAR_OK) return rc;
		if ((rc = up.upload_to(n, st)) != PAR_OK) return rc;
		const DevAudio da = up.view();
		if ((rc = launch_deinterleave(da.p, n, da.stride, n_ch, da.ch_stride, dx.as<float>(), n_pad_al, device, st)) != PAR_OK) return rc;
	}
	// ---- analysis
	if ((rc = dS.alloc((size_t)n_ch * T * F * sizeof(float2))) != PAR_OK) return rc;
	StftArgs sa;
	sa.x = dx.as<float>(); sa.n = n_pad; sa.x_stride = 1; sa.x_ch_stride = n_pad_al; sa.x_origin = 0;
	sa.n_ch = n_ch; sa.n_fft = n_fft; sa.hop = hop; sa.zeropad = 1; sa.n_frames = T; sa.frame0 = 0;
	sa.window = dwin; sa.out = dS.p; sa.out_pitch = F; sa.out_ch_stride = T * F; sa.magnitude = 0;
	if ((rc = launch_stft(sa, device, st)) != PAR_OK) return rc;
	// ---- mask
	float2 *S = dS.as<float2>();
	if (op == PAR_SPEC_GATE) {
		if ((rc = dscratch.alloc((size_t)F * sizeof(double))) != PAR_OK) return rc;
		PAR_CUDA(cudaMemcpyAsync(dscratch.p, params, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, st));
		PAR_CUDA(cudaStreamSynchronize(st));            // `params` is the caller's (possibly pageable) memory
		if ((rc = launch_spec_gate(S, (int64_t)n_ch * T * F, (int)F, dscratch.as<double>(), gain_db, device, st)) != PAR_OK) return rc;
	} else if (select) {
		float2 *L = S, *R = S + T * F;
		rc = launch_spec_select(L, R, T * F, op == PAR_SPEC_SELECT_MIN ? nullptr : L,
		                        op == PAR_SPEC_SELECT_MIN ? L : (op == PAR_SPEC_SELECT_BOTH ? R : nullptr), device, st);
		if (rc != PAR_OK) return rc;
	} else if (n_params > 0) {
		if ((rc = dscratch.alloc((size_t)(2 * F + (g1 - g0) * F) * sizeof(double))) != PAR_OK) return rc;
		for (int c = 0; c < n_ch; c++)
			if ((rc = launch_spec_heal(S + (int64_t)c * T * F, F, T, (int)F, (const int64_t *)params, n_params, g0, g1,
			                           dscratch.as<double>(), device, st)) != PAR_OK)
				return rc;
	}
	// ---- synthesis: istft(S, length=n, hop_length=hop), util/fourier.py:373-381 frame count
	int64_t n_frames = (n + n_fft + hop - 1) / hop;
	if (n_frames > T) n_frames = T;
	if (istft_needs_scratch(n_fft) && (rc = dframes.alloc((size_t)n_out * n_frames * n_fft * sizeof(float))) != PAR_OK) return rc;
	IstftArgs ia;
	ia.S = S; ia.n_fft = n_fft; ia.n_frames = n_frames; ia.s_pitch = F; ia.s_ch_stride = T * F; ia.n_ch = n_out; ia.hop = hop;
	ia.window = dsyn; ia.start = n_fft / 2; ia.length = n; ia.frames = dframes.as<float>();
	if (flags & PAR_DEVICE_PTRS) {
		ia.y = y; ia.y_stride = y_stride; ia.y_ch_stride = y_ch_stride;
		return launch_istft(ia, device, st);
	}
	DevOut dyo;
	if ((rc = alloc_out(dy, n, y_stride, n_out, y_ch_stride, st, &dyo)) != PAR_OK) return rc;
	ia.y = dyo.p; ia.y_stride = dyo.stride; ia.y_ch_stride = dyo.ch_stride;
	if ((rc = launch_istft(ia, device, st)) != PAR_OK) return rc;
	if ((rc = download_out(dyo, y, 0, n, y_stride, n_out, y_ch_stride, st)) != PAR_OK) return rc;
	PAR_CUDA(cudaStreamSynchronize(st));
	return PAR_OK;
}

PAR_API int par_speed_segments(const double *sampletimes, const double *speeds, int64_t k,
                       int64_t *seg_n, int64_t *total) {
	if (!sampletimes || !speeds || k < 2 || !seg_n) { set_error("speed_segments: bad argument"); return PAR_EINVAL; }
	// util/resampling.py:111-118.  Host code of this file is compiled with -ffp-contract=off and x86-64
	// doubles are IEEE (SSE2), so every statement below is one correctly rounded operation, in the
	// reference's order: periods[i] * mean(speeds[i:i+2]) + err.
	double err = 0.0;
	int64_t sum = 0;
	for (int64_t i = 0; i + 1 < k; i++) {
		const double period = sampletimes[i + 1] - sampletimes[i];
		const double mean = (speeds[i] + speeds[i + 1]) / 2.0;
		const double prod = period * mean;
		const double inerr = prod + err;
		const double nr = nearbyint(inerr);       // Python round(): half to even
		if (!(fabs(nr) < 9.0e15)) { set_error("speed_segments: non-finite segment length"); return PAR_EINVAL; }
		const int64_t n = (int64_t)nr;
		err = inerr - (double)n;
		seg_n[i] = n;
		if (n > 0) sum += n;
	}
	if (total) *total = sum;
	return PAR_OK;
}

// Pinned staging memory for the small per-call arrays of speed_to_pos (segment tables up, segment sums
// down): one growing buffer per host thread, so the copies run at full PCIe rate and without the driver's
// pageable-memory bounce.  The owning call synchronises its stream before returning, so the buffer is free
// for the thread's next call.
struct PinnedScratch {
	char *p = nullptr;
	size_t cap = 0;
	~PinnedScratch() { if (p) cudaFreeHost(p); }
	char *get(size_t bytes) {
		if (bytes > cap) {
			if (p) cudaFreeHost(p);
			p = nullptr; cap = 0;
			size_t want = bytes + bytes / 4 + 4096;
			if (cudaHostAlloc((void **)&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
			cap = want;
		}
		return p;
	}
};
static thread_local PinnedScratch g_pinned;

// Speed curve -> device-resident read positions.  `pos` must hold `cap` doubles on the device;
// *m_out receives the number of valid positions.  Synchronises `st` (the per-segment sums have to reach
// the host for the serial offset chain and the end test, util/resampling.py:125-135).
struct SegChain {                 // host copy of the serial part of speed_to_pos
	std::vector<int64_t> start;   // first output index of every segment
	std::vector<double> off;      // carried offset = position just before the segment's first output
};

static int positions_device(const double *sampletimes, const double *speeds, int64_t k, double num_input_samples,
                            const std::vector<int64_t> &seg_n, int64_t total, double *pos, int64_t cap,
                            int64_t *m_out, cudaStream_t st, SegChain *chain = nullptr,
                            const double *window = nullptr, int64_t *win_origin = nullptr,
                            int64_t *win_count = nullptr, const double *ext_sums_dev = nullptr) {
	// window = {lo, hi}: expand only the segments that hold positions in [lo, hi] (plus the segment
	// after them, for the period of the last output); pos[0] is then output *win_origin.
	// ext_sums_dev: the per-segment totals of the whole curve, already on the device (ranks of a time-sharded
	// job compute a slice each and all-gather them): nothing but the window's segments is uploaded then.
	int rc;
	const int64_t n_seg = k - 1;
	// Two ways to the same bits.  Default: per-segment totals first (compute only, no position traffic), the host chain,
	// then ONE expansion that writes every position once, offset included: DRAM traffic = the positions themselves.
	// $PAR_B200_POS_ONE_PASS=1: expand the bare cumsums first and add the offsets in a second streaming pass (round 1:
	// the same time, 2.8x the traffic).
	static const bool want_one_pass = [] { const char *e = getenv("PAR_B200_POS_ONE_PASS"); return e && e[0] == '1'; }();
	const bool one_pass = want_one_pass && !window && cap >= total && !ext_sums_dev;
	// pinned layout: [speeds k][seg_n n_seg][seg_start n_seg][sums n_seg][off n_seg]
	const size_t bytes = ((size_t)k + 4 * (size_t)n_seg) * 8;
	char *pin = g_pinned.get(bytes);
	if (!pin) { set_error("speed_to_pos: pinned staging allocation failed"); return PAR_ECUDA; }
	double *h_sp = (double *)pin;
	int64_t *h_n = (int64_t *)(h_sp + k);
	int64_t *h_start = h_n + n_seg;
	double *h_sums = (double *)(h_start + n_seg);
	double *h_off = h_sums + n_seg;
	memcpy(h_sp, speeds, (size_t)k * 8);
	int64_t o = 0;
	for (int64_t i = 0; i < n_seg; i++) { h_n[i] = seg_n[i]; h_start[i] = o; if (seg_n[i] > 0) o += seg_n[i]; }

	DevBuf d_tab(st), d_sum(st), d_off(st);
	double *d_sp = nullptr;
	int64_t *d_n = nullptr, *d_start = nullptr;
	if (!ext_sums_dev) {
		if ((rc = d_tab.alloc(((size_t)k + 2 * (size_t)n_seg) * 8)) != PAR_OK) return rc;
		if ((rc = d_sum.alloc((size_t)n_seg * 8)) != PAR_OK) return rc;
		d_sp = d_tab.as<double>();
		d_n = (int64_t *)(d_sp + k);
		d_start = d_n + n_seg;
		PAR_CUDA(cudaMemcpyAsync(d_sp, h_sp, ((size_t)k + 2 * (size_t)n_seg) * 8, cudaMemcpyHostToDevice, st));
		// whole-curve call: ONE pass writes every segment's bare cumsum into pos (capacity permitting) and
		// its total; after the host chain a streaming pass adds the offsets.  Windowed call: totals only,
		// then expand just the window's segments.
		if (one_pass) rc = launch_expand_positions(d_sp, d_n, d_start, nullptr, n_seg, pos, total, st, d_sum.as<double>());
		else rc = launch_segment_sums(d_sp, d_n, n_seg, d_sum.as<double>(), st);
		if (rc != PAR_OK) return rc;
		PAR_CUDA(cudaMemcpyAsync(h_sums, d_sum.p, (size_t)n_seg * 8, cudaMemcpyDeviceToHost, st));
	} else {
		PAR_CUDA(cudaMemcpyAsync(h_sums, ext_sums_dev, (size_t)n_seg * 8, cudaMemcpyDeviceToHost, st));
	}
	PAR_CUDA(cudaStreamSynchronize(st));

	// serial offset chain + end test (util/resampling.py:125-135): one addition per segment; the
	// division for the block's first position is only needed once the end test can fire
	double offset = sampletimes[0];
	int64_t m = total, last_seg = n_seg;        // last_seg: first segment the reference never reaches
	for (int64_t i = 0; i < n_seg; i++) {
		h_off[i] = offset;
		const int64_t n = seg_n[i];
		if (n <= 0) continue;                   // the reference raises on an empty block
		const double last = h_sums[i] + offset;
		if (num_input_samples <= last) {
			double inv0 = 1.0 / speeds[i];
			if (n == 1) inv0 = NAN;               // arange(1)/0 -> nan in the reference
			const double first = inv0 + offset;
			if (first <= num_input_samples) {
				// np.argmin(|sample_at - L|) over this block, first minimum
				const double ds = speeds[i + 1] - speeds[i];
				const double nm1 = (double)(n - 1);
				double acc = 0.0, best = INFINITY;
				int64_t besti = 0;
				for (int64_t j = 0; j < n; j++) {
					const double q = (double)j / nm1;
					double v = q * ds;
					v = v + speeds[i];
					const double r = 1.0 / v;
					acc = acc + r;
					const double pj = acc + offset;
					const double dist = fabs(pj - num_input_samples);
					if (dist < best) { best = dist; besti = j; }
				}
				m = h_start[i] + besti;
				last_seg = i + 1;
				break;
			}
		}
		offset = last;
	}
	// segments behind the end of the file hold no positions: they must not look like "below the window"
	for (int64_t i = last_seg; i < n_seg; i++) h_off[i] = INFINITY;
	*m_out = m;
	int64_t seg_a = 0, seg_b = n_seg;          // segments [seg_a, seg_b) are expanded
	if (window) {
		// off[i] = position just before segment i's first output; positions grow with i for positive speeds
		while (seg_a + 1 < last_seg && h_off[seg_a + 1] < window[0]) seg_a++;
		seg_b = seg_a;
		while (seg_b < last_seg && h_off[seg_b] <= window[1]) seg_b++;
		if (seg_b < last_seg) seg_b++;
		const int64_t o_begin = h_start[seg_a] < m ? h_start[seg_a] : m;
		const int64_t o_end = seg_b < n_seg ? (h_start[seg_b] < m ? h_start[seg_b] : m) : m;
		*win_origin = o_begin;
		*win_count = o_end - o_begin;
		if (*win_count > cap) {
			set_error("speed_to_pos: output capacity too small");
			return PAR_ECAPACITY;
		}
		if (*win_count <= 0) return PAR_OK;
		pos -= o_begin;                          // the kernel indexes globally
	} else if (m > cap) {
		set_error("speed_to_pos: output capacity too small");
		return PAR_ECAPACITY;
	}
	if (m == 0) return PAR_OK;
	if (chain) {
		chain->start.assign(h_start, h_start + n_seg);
		chain->off.assign(h_off, h_off + n_seg);
	}
	const int64_t cnt = seg_b - seg_a;
	if (ext_sums_dev) {
		// upload just the window's rows of the segment tables: [speeds cnt+1][seg_n cnt][seg_start cnt][off cnt]
		if ((rc = d_tab.alloc(((size_t)cnt * 4 + 1) * 8)) != PAR_OK) return rc;
		d_sp = d_tab.as<double>();
		d_n = (int64_t *)(d_sp + cnt + 1);
		d_start = d_n + cnt;
		double *d_o = (double *)(d_start + cnt);
		PAR_CUDA(cudaMemcpyAsync(d_sp, h_sp + seg_a, (size_t)(cnt + 1) * 8, cudaMemcpyHostToDevice, st));
		PAR_CUDA(cudaMemcpyAsync(d_n, h_n + seg_a, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
		PAR_CUDA(cudaMemcpyAsync(d_start, h_start + seg_a, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
		PAR_CUDA(cudaMemcpyAsync(d_o, h_off + seg_a, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
		rc = launch_expand_positions(d_sp, d_n, d_start, d_o, cnt, pos, m, st);
	} else {
		if ((rc = d_off.alloc((size_t)cnt * 8)) != PAR_OK) return rc;
		PAR_CUDA(cudaMemcpyAsync(d_off.p, h_off + seg_a, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
		if (one_pass) rc = launch_add_offsets(d_n, d_start, d_off.as<double>(), n_seg, pos, m, st);
		else rc = launch_expand_positions(d_sp + seg_a, d_n + seg_a, d_start + seg_a, d_off.as<double>(), cnt, pos, m, st);
	}
	if (rc != PAR_OK) return rc;
	// the pinned staging buffer must outlive its async copies
	PAR_CUDA(cudaStreamSynchronize(st));
	return PAR_OK;
}

PAR_API int par_speed_to_pos_f64(const double *sampletimes, const double *speeds, int64_t k,
                         double num_input_samples, double *pos, int64_t cap, int64_t *m,
                         unsigned flags, int device, void *stream) {
	if (!sampletimes || !speeds || k < 2 || !m || (!pos && cap > 0)) {
		set_error("speed_to_pos: bad argument");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	std::vector<int64_t> seg_n(k - 1);
	int64_t total = 0;
	if ((rc = par_speed_segments(sampletimes, speeds, k, seg_n.data(), &total)) != PAR_OK) return rc;
	if (flags & PAR_DEVICE_PTRS)
		return positions_device(sampletimes, speeds, k, num_input_samples, seg_n, total, pos, cap, m, st);
	DevBuf d_pos(st);
	const int64_t dcap = total < cap ? total : cap;
	if ((rc = d_pos.alloc((size_t)(dcap > 0 ? dcap : 1) * sizeof(double))) != PAR_OK) return rc;
	rc = positions_device(sampletimes, speeds, k, num_input_samples, seg_n, total, d_pos.as<double>(), dcap, m, st);
	if (rc != PAR_OK) return rc;
	if (*m > 0) {
		PAR_CUDA(cudaMemcpyAsync(pos, d_pos.p, *m * sizeof(double), cudaMemcpyDeviceToHost, st));
		PAR_CUDA(cudaStreamSynchronize(st));
	}
	return PAR_OK;
}

extern "C" PAR_API int par_speed_to_pos_range_f64(const double *sampletimes, const double *speeds, int64_t k,
                                                  double num_input_samples, double lo_pos, double hi_pos,
                                                  double *pos, int64_t cap, int64_t *pos_origin, int64_t *pos_count,
                                                  int64_t *m, unsigned flags, int device, void *stream) {
	if (!(flags & PAR_DEVICE_PTRS)) { set_error("speed_to_pos_range: device pointers only"); return PAR_EUNSUPPORTED; }
	if (!sampletimes || !speeds || k < 2 || !m || !pos_origin || !pos_count || !pos || !(lo_pos <= hi_pos)) {
		set_error("speed_to_pos_range: bad argument");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	std::vector<int64_t> seg_n(k - 1);
	int64_t total = 0;
	if ((rc = par_speed_segments(sampletimes, speeds, k, seg_n.data(), &total)) != PAR_OK) return rc;
	const double window[2] = {lo_pos, hi_pos};
	return positions_device(sampletimes, speeds, k, num_input_samples, seg_n, total, pos, cap, m, (cudaStream_t)stream,
	                        nullptr, window, pos_origin, pos_count);
}

// Per-segment totals of the cumsum of 1/speed for segments [seg_begin, seg_end) of a curve: the part of
// util/resampling.py:120-126 that a rank of a time-sharded job contributes (dist.TimeShard.positions
// all-gathers the slices and hands the whole array to par_speed_to_pos_range_sums_f64).
extern "C" PAR_API int par_segment_sums_f64(const double *sampletimes, const double *speeds, int64_t k,
                                            int64_t seg_begin, int64_t seg_end, double *sums, int64_t *seg_n_out,
                                            unsigned flags, int device, void *stream) {
	if (!(flags & PAR_DEVICE_PTRS)) { set_error("segment_sums: device pointers only"); return PAR_EUNSUPPORTED; }
	if (!sampletimes || !speeds || k < 2 || seg_begin < 0 || seg_end < seg_begin || seg_end > k - 1 || (!sums && seg_end > seg_begin)) {
		set_error("segment_sums: bad argument");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	const int64_t cnt = seg_end - seg_begin;
	if (cnt == 0) return PAR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	// the error-diffused segment lengths of the WHOLE curve (a serial host recurrence); handed back so that the
	// caller's next step does not repeat it
	std::vector<int64_t> seg_local;
	int64_t *seg_n = seg_n_out;
	if (!seg_n) { seg_local.resize(k - 1); seg_n = seg_local.data(); }
	if ((rc = par_speed_segments(sampletimes, speeds, k, seg_n, nullptr)) != PAR_OK) return rc;
	char *pin = g_pinned.get(((size_t)cnt * 2 + 1) * 8);
	if (!pin) { set_error("segment_sums: pinned staging allocation failed"); return PAR_ECUDA; }
	double *h_sp = (double *)pin;
	int64_t *h_n = (int64_t *)(h_sp + cnt + 1);
	memcpy(h_sp, speeds + seg_begin, (size_t)(cnt + 1) * 8);
	memcpy(h_n, seg_n + seg_begin, (size_t)cnt * 8);
	DevBuf d_tab(st);
	if ((rc = d_tab.alloc(((size_t)cnt * 2 + 1) * 8)) != PAR_OK) return rc;
	PAR_CUDA(cudaMemcpyAsync(d_tab.p, pin, ((size_t)cnt * 2 + 1) * 8, cudaMemcpyHostToDevice, st));
	if ((rc = launch_segment_sums(d_tab.as<double>(), (const int64_t *)(d_tab.as<double>() + cnt + 1), cnt, sums, st)) != PAR_OK)
		return rc;
	PAR_CUDA(cudaStreamSynchronize(st));
	return PAR_OK;
}

extern "C" PAR_API int par_speed_to_pos_range_sums_f64(const double *sampletimes, const double *speeds, int64_t k,
                                                       double num_input_samples, double lo_pos, double hi_pos,
                                                       const double *seg_sums, const int64_t *seg_n_in, double *pos, int64_t cap,
                                                       int64_t *pos_origin, int64_t *pos_count, int64_t *m,
                                                       unsigned flags, int device, void *stream) {
	if (!(flags & PAR_DEVICE_PTRS)) { set_error("speed_to_pos_range: device pointers only"); return PAR_EUNSUPPORTED; }
	if (!sampletimes || !speeds || k < 2 || !m || !pos_origin || !pos_count || !pos || !seg_sums || !(lo_pos <= hi_pos)) {
		set_error("speed_to_pos_range: bad argument");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	std::vector<int64_t> seg_n(k - 1);
	int64_t total = 0;
	if (seg_n_in) {
		for (int64_t i = 0; i + 1 < k; i++) { seg_n[i] = seg_n_in[i]; if (seg_n_in[i] > 0) total += seg_n_in[i]; }
	} else if ((rc = par_speed_segments(sampletimes, speeds, k, seg_n.data(), &total)) != PAR_OK) {
		return rc;
	}
	const double window[2] = {lo_pos, hi_pos};
	return positions_device(sampletimes, speeds, k, num_input_samples, seg_n, total, pos, cap, m, (cudaStream_t)stream,
	                        nullptr, window, pos_origin, pos_count, seg_sums);
}

// Resample with DEVICE positions; signal / out are host or device according to `flags`.
// Host pointers: when the positions come from a speed curve with positive speeds (`chain`), they are
// monotone and chunk boundaries are put on segment starts, where the host knows the read position:
// chunks of the output then flow through upload -> interpolate -> download on three streams.
static int resample_with_dev_pos(bool sinc, const double *dpos, int64_t m, const float *signal, int64_t n_in,
                                 int64_t sig_stride, int n_ch, int64_t sig_ch_stride, int nt,
                                 float *out, int64_t out_stride, int64_t out_ch_stride,
                                 unsigned flags, int device, cudaStream_t st, const SegChain *chain = nullptr,
                                 bool monotone = false, AudioUploader *pre_up = nullptr, SideStreams *pre_ss = nullptr,
                                 double period_dev = -1.0) {
	int rc;
	SincArgs a;
	a.pos = dpos; a.m = m; a.n_in = n_in; a.n_ch = n_ch; a.nt = nt;
	a.aligned_edges = (flags & PAR_SINC_ALIGNED_EDGES) ? 1 : 0;
	a.kernel = (flags & PAR_SINC_KERNEL_WS) ? 2 : ((flags & PAR_SINC_KERNEL_TILED) ? 1 : 0);
	a.out_begin = 0; a.out_end = m;
	a.pos_origin = a.sig_origin = a.out_origin = 0;
	a.period_dev = period_dev;
	if (flags & PAR_DEVICE_PTRS) {
		a.signal = signal; a.sig_stride = sig_stride; a.sig_ch_stride = sig_ch_stride;
		a.out = out; a.out_stride = out_stride; a.out_ch_stride = out_ch_stride;
		return sinc ? launch_sinc(a, device, st) : launch_linear(a, device, st);
	}
	CallTrace tr("resample");
	DevBuf dout(st);
	AudioUploader up_local(st);
	DevOut dd;
	SideStreams ss_local;
	// par_varispeed_f32 hands in an uploader that is already sending the head of the signal
	AudioUploader &up = pre_up ? *pre_up : up_local;
	SideStreams &ss = pre_ss ? *pre_ss : ss_local;
	if (!pre_up && (rc = up.init(signal, n_in, sig_stride, n_ch, sig_ch_stride)) != PAR_OK) return rc;
	if ((rc = alloc_out(dout, m, out_stride, n_ch, out_ch_stride, st, &dd)) != PAR_OK) return rc;
	if (!pre_ss) {
		if ((rc = ss.init()) != PAR_OK) return rc;
		if ((rc = ss.after(ss.up, st)) != PAR_OK) return rc;
	}
	const DevAudio da = up.view();
	a.signal = da.p; a.sig_stride = da.stride; a.sig_ch_stride = da.ch_stride;
	a.out = dd.p; a.out_stride = dd.stride; a.out_ch_stride = dd.ch_stride;
	EventTimer tm(st);
	tm.start();
	tr.mark("allocated, streams ready");
	int64_t per_chunk = chunk_bytes() / (int64_t)(sizeof(float) * n_ch);
	if (per_chunk < 1024) per_chunk = 1024;
	const bool pipelined = chain && monotone && m > 2 * per_chunk;
	if (!pipelined) {
		if ((rc = up.upload_to(n_in, ss.up)) != PAR_OK) return rc;
		if ((rc = ss.after(st, ss.up)) != PAR_OK) return rc;
		if ((rc = sinc ? launch_sinc(a, device, st) : launch_linear(a, device, st)) != PAR_OK) return rc;
		if ((rc = ss.after(ss.down, st)) != PAR_OK) return rc;
		if ((rc = download_out(dd, out, 0, m, out_stride, n_ch, out_ch_stride, ss.down)) != PAR_OK) return rc;
	} else {
		const int64_t n_seg = (int64_t)chain->start.size();
		int64_t seg = 0, o0 = 0;
		int64_t ramp = per_chunk / 8 > 1024 && ramp_enabled() ? per_chunk / 8 : per_chunk;     // small first chunks: see par_stft_f32
		for (; o0 < m; ramp = ramp * 2 < per_chunk ? ramp * 2 : per_chunk) {
			// advance to the first segment start at least `ramp` outputs ahead
			while (seg < n_seg && chain->start[seg] < o0 + ramp) seg++;
			int64_t o1 = m, need = n_in;
			if (seg < n_seg && chain->start[seg] < m) {
				o1 = chain->start[seg];
				// outputs before segment `seg` read positions <= off[seg]; taps reach NT further
				const double top = chain->off[seg] + (double)nt + 4.0;
				need = top < (double)n_in ? (top > 0.0 ? (int64_t)top : 0) : n_in;
			}
			if ((rc = up.upload_to(need, ss.up)) != PAR_OK) return rc;
			if ((rc = ss.after(st, ss.up)) != PAR_OK) return rc;
			a.out_begin = o0; a.out_end = o1;
			if ((rc = sinc ? launch_sinc(a, device, st) : launch_linear(a, device, st)) != PAR_OK) return rc;
			if ((rc = ss.after(ss.down, st)) != PAR_OK) return rc;
			if ((rc = download_out(dd, out, o0, o1, out_stride, n_ch, out_ch_stride, ss.down)) != PAR_OK) return rc;
			o0 = o1;
		}
	}
	tm.stop();
	tr.mark("all chunks enqueued");
	if ((rc = ss.drain()) != PAR_OK) return rc;
	PAR_CUDA(cudaStreamSynchronize(st));
	tm.finish();
	tr.mark("drained");
	return PAR_OK;
}

static int check_resample_args(bool sinc, int64_t m, const void *pos_or_curve, const float *signal, int64_t n_in,
                               int64_t sig_stride, int n_ch, int nt, const float *out, int64_t out_stride) {
	if (m < 0 || n_in < 0 || n_ch < 1 || sig_stride < 1 || out_stride < 1 || (sinc && (nt < 1 || nt > 512))) {
		set_error("resample: bad size argument");
		return PAR_EINVAL;
	}
	if (m > 0 && (!pos_or_curve || !out)) { set_error("resample: null pointer"); return PAR_EINVAL; }
	if (n_in > 0 && !signal) { set_error("resample: null signal"); return PAR_EINVAL; }
	return PAR_OK;
}

static int resample_common(bool sinc, const double *pos, int64_t m, const float *signal, int64_t n_in,
                           int64_t sig_stride, int n_ch, int64_t sig_ch_stride, int nt,
                           float *out, int64_t out_stride, int64_t out_ch_stride,
                           unsigned flags, int device, void *stream) {
	int rc = check_resample_args(sinc, m, pos, signal, n_in, sig_stride, n_ch, nt, out, out_stride);
	if (rc != PAR_OK) return rc;
	if ((rc = use_device(device)) != PAR_OK) return rc;
	if (m == 0) return PAR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	if (flags & PAR_DEVICE_PTRS)
		return resample_with_dev_pos(sinc, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out, out_stride,
		                             out_ch_stride, flags, device, st);
	DevBuf dpos(st);
	if ((rc = dpos.alloc(m * sizeof(double))) != PAR_OK) return rc;
	PAR_CUDA(cudaMemcpyAsync(dpos.p, pos, m * sizeof(double), cudaMemcpyHostToDevice, st));
	return resample_with_dev_pos(sinc, dpos.as<double>(), m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out,
	                             out_stride, out_ch_stride, flags, device, st);
}

PAR_API int par_sinc_resample_f32(const double *pos, int64_t m, const float *signal, int64_t n_in,
                          int64_t sig_stride, int n_ch, int64_t sig_ch_stride, int nt,
                          float *out, int64_t out_stride, int64_t out_ch_stride,
                          unsigned flags, int device, void *stream) {
	return resample_common(true, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out, out_stride,
	                       out_ch_stride, flags, device, stream);
}

PAR_API int par_linear_resample_f32(const double *pos, int64_t m, const float *signal, int64_t n_in,
                            int64_t sig_stride, int n_ch, int64_t sig_ch_stride,
                            float *out, int64_t out_stride, int64_t out_ch_stride,
                            unsigned flags, int device, void *stream) {
	return resample_common(false, pos, m, signal, n_in, sig_stride, n_ch, sig_ch_stride, 1, out, out_stride,
	                       out_ch_stride, flags, device, stream);
}

PAR_API int par_varispeed_f32(const double *sampletimes, const double *speeds, int64_t k,
                      const float *signal, int64_t n_in, int64_t sig_stride, int n_ch, int64_t sig_ch_stride,
                      int mode, int nt, float *out, int64_t out_cap, int64_t out_stride, int64_t out_ch_stride,
                      int64_t *m, unsigned flags, int device, void *stream) {
	if (!sampletimes || !speeds || k < 2 || !m || (mode != PAR_MODE_LINEAR && mode != PAR_MODE_SINC)) {
		set_error("varispeed: bad argument");
		return PAR_EINVAL;
	}
	const bool sinc = mode == PAR_MODE_SINC;
	int rc = check_resample_args(sinc, out_cap, sampletimes, signal, n_in, sig_stride, n_ch, nt, out, out_stride);
	if (rc != PAR_OK) return rc;
	if ((rc = use_device(device)) != PAR_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	CallTrace tr("varispeed");
	std::vector<int64_t> seg_n(k - 1);
	int64_t total = 0;
	if ((rc = par_speed_segments(sampletimes, speeds, k, seg_n.data(), &total)) != PAR_OK) return rc;
	tr.mark("segments");
	// positive finite speeds and non-empty segments => monotone positions => pipelined host path
	bool monotone = true;
	double period_dev = 0.0;           // largest |read period - 1| on the curve: sizes the resampler's tiles
	for (int64_t i = 0; i < k; i++) {
		monotone = monotone && speeds[i] > 0.0 && speeds[i] < 1e6;
		const double d = fabs(1.0 / speeds[i] - 1.0);
		if (d > period_dev) period_dev = d;
	}
	if (!(period_dev < 1e6)) period_dev = -1.0;
	for (int64_t i = 0; i + 1 < k; i++) monotone = monotone && seg_n[i] >= 2;
	DevBuf dpos(st);
	SegChain chain;
	AudioUploader up(st);          // declared behind dpos: destroyed (and its streams drained) before it is freed
	SideStreams ss;
	const bool host_io = !(flags & PAR_DEVICE_PTRS) && n_in > 0 && signal;
	if ((rc = dpos.alloc((size_t)(total > 0 ? total : 1) * sizeof(double))) != PAR_OK) return rc;
	if (host_io) {
		// the head of the signal crosses PCIe while the positions are being expanded
		if ((rc = up.init(signal, n_in, sig_stride, n_ch, sig_ch_stride)) != PAR_OK) return rc;
		if ((rc = ss.init()) != PAR_OK) return rc;
		if ((rc = ss.after(ss.up, st)) != PAR_OK) return rc;       // allocations are stream-ordered on st
		if (ramp_enabled() && (rc = up.upload_to(4 << 20, ss.up)) != PAR_OK) return rc;
	}
	rc = positions_device(sampletimes, speeds, k, (double)n_in, seg_n, total, dpos.as<double>(), total, m, st, &chain);
	if (rc != PAR_OK) return rc;
	if (*m > out_cap) {
		set_error("varispeed: output capacity too small (need *m samples per channel)");
		return PAR_ECAPACITY;
	}
	if (*m == 0) return PAR_OK;
	tr.mark("positions");
	rc = resample_with_dev_pos(sinc, dpos.as<double>(), *m, signal, n_in, sig_stride, n_ch, sig_ch_stride, nt, out,
	                           out_stride, out_ch_stride, flags, device, st, &chain, monotone, host_io ? &up : nullptr,
	                           host_io ? &ss : nullptr, period_dev);
	tr.mark("done");
	return rc;
}

// ---- trackers ----------------------------------------------------------------------------------
static int trace_on_device(const float *mag_dev, int64_t pitch, int num_bins, int64_t frame0, int64_t count,
                           int fft_size, double sr, double tolerance_st, int mode, double *freqs, cudaStream_t st) {
	int rc;
	DevBuf df(st);
	if ((rc = df.alloc((size_t)count * sizeof(double))) != PAR_OK) return rc;
	PAR_CUDA(cudaMemcpyAsync(df.p, freqs, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, st));
	rc = launch_trace(mag_dev, pitch, num_bins, frame0, count, fft_size, sr, tolerance_st / 12.0, mode, freqs[0],
	                  df.as<double>(), st);
	if (rc != PAR_OK) return rc;
	PAR_CUDA(cudaMemcpyAsync(freqs, df.p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, st));
	PAR_CUDA(cudaStreamSynchronize(st));
	return PAR_OK;
}

static int check_trace_args(int num_bins, int64_t n_frames, int64_t frame0, int64_t count, int fft_size, double sr,
                            double tolerance_st, int mode, const double *freqs) {
	if (num_bins < 8 || frame0 < 0 || count < 0 || frame0 + count > n_frames || fft_size < 2 || !(sr > 0) ||
	    !(tolerance_st >= 0) || mode < PAR_TRACE_PEAK || mode > PAR_TRACE_COG || (count > 0 && !freqs)) {
		set_error("trace: bad argument");
		return PAR_EINVAL;
	}
	for (int64_t i = 0; i < count; i++)
		if (!(freqs[i] > 0.0) || !(freqs[i] < 1e12)) { set_error("trace: trail frequencies must be positive"); return PAR_EINVAL; }
	return PAR_OK;
}

PAR_API int par_trace_f32(const float *mag, int num_bins, int64_t n_frames, int64_t pitch, int64_t frame0,
                  int64_t count, int fft_size, double sr, double tolerance_st, int mode, double *freqs,
                  unsigned flags, int device, void *stream) {
	if (!mag || pitch < num_bins) { set_error("trace: bad spectrogram"); return PAR_EINVAL; }
	int rc = check_trace_args(num_bins, n_frames, frame0, count, fft_size, sr, tolerance_st, mode, freqs);
	if (rc != PAR_OK) return rc;
	if ((rc = use_device(device)) != PAR_OK) return rc;
	if (count == 0) return PAR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	if (flags & PAR_DEVICE_PTRS)
		return trace_on_device(mag, pitch, num_bins, frame0, count, fft_size, sr, tolerance_st, mode, freqs, st);
	DevBuf dm(st);                                       // only the traced frames are uploaded
	if ((rc = dm.alloc((size_t)count * pitch * sizeof(float))) != PAR_OK) return rc;
	PAR_CUDA(cudaMemcpyAsync(dm.p, mag + frame0 * pitch, (size_t)count * pitch * sizeof(float), cudaMemcpyHostToDevice, st));
	return trace_on_device(dm.as<float>(), pitch, num_bins, 0, count, fft_size, sr, tolerance_st, mode, freqs, st);
}

PAR_API int par_stft_trace_f32(const float *x, int64_t n, int64_t x_stride, int n_fft, int hop, int zeropad,
                       const float *window, int64_t frame0, int64_t count, double sr, double tolerance_st,
                       int mode, double *freqs, unsigned flags, int device, void *stream) {
	if (!x || !window || n < 1 || n_fft < 2 || hop < 1 || zeropad < 1 || x_stride < 1) {
		set_error("stft_trace: bad argument");
		return PAR_EINVAL;
	}
	const int64_t T = par_stft_num_frames(n, n_fft, hop);
	const int num_bins = (int)((int64_t)n_fft * zeropad / 2 + 1);
	int rc = check_trace_args(num_bins, T, frame0, count, n_fft * zeropad, sr, tolerance_st, mode, freqs);
	if (rc != PAR_OK) return rc;
	if ((rc = use_device(device)) != PAR_OK) return rc;
	if (count == 0) return PAR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	const float *dwin = device_window(device, window, n_fft, st);
	if (!dwin) return PAR_ECUDA;
	DevBuf dmag(st), dplanar(st);
	AudioUploader up(st);
	if ((rc = dmag.alloc((size_t)count * num_bins * sizeof(float))) != PAR_OK) return rc;
	StftArgs a;
	a.n = n; a.n_ch = 1; a.n_fft = n_fft; a.hop = hop; a.zeropad = zeropad; a.window = dwin; a.magnitude = 1;
	a.frame0 = frame0; a.n_frames = count; a.x_origin = 0; a.x_ch_stride = 0;
	if (flags & PAR_DEVICE_PTRS) {
		a.x = x; a.x_stride = x_stride;
	} else {
		if ((rc = up.init(x, n, x_stride, 1, 0)) != PAR_OK) return rc;
		if ((rc = up.upload_to(n, st)) != PAR_OK) return rc;
		const DevAudio da = up.view();
		if (da.stride != 1) {
			if ((rc = dplanar.alloc((size_t)up.n_al * sizeof(float))) != PAR_OK) return rc;
			if ((rc = launch_deinterleave(da.p, n, da.stride, 1, da.ch_stride, dplanar.as<float>(), up.n_al, device, st)) != PAR_OK)
				return rc;
			a.x = dplanar.as<float>();
		} else {
			a.x = da.p;
		}
		a.x_stride = 1;
	}
	// row 0 of the scratch spectrogram is frame frame0
	a.out = (char *)dmag.p - (size_t)frame0 * num_bins * sizeof(float);
	a.out_pitch = num_bins; a.out_ch_stride = 0;
	if ((rc = launch_stft(a, device, st)) != PAR_OK) return rc;
	return trace_on_device(dmag.as<float>(), num_bins, num_bins, 0, count, n_fft * zeropad, sr, tolerance_st, mode, freqs, st);
}

// ---- shard entry points (device pointers only): one rank's slice of a time-sharded job ----------
PAR_API int par_stft_range_f32(const float *x, int64_t n_local, int64_t x_origin, int64_t n_global, int n_ch,
                       int64_t x_ch_stride, int n_fft, int hop, int zeropad, const float *window,
                       int64_t frame0, int64_t n_frames, void *out, int64_t out_pitch, int64_t out_ch_stride,
                       unsigned flags, int device, void *stream) {
	if (!(flags & PAR_DEVICE_PTRS)) { set_error("stft_range: device pointers only"); return PAR_EUNSUPPORTED; }
	if (!x || !out || !window) { set_error("stft_range: null pointer"); return PAR_EINVAL; }
	const int64_t T = par_stft_num_frames(n_global, n_fft, hop);
	const int64_t F = (int64_t)n_fft * zeropad / 2 + 1;
	if (n_global < 1 || n_local < 1 || n_ch < 1 || n_fft < 2 || hop < 1 || zeropad < 1 || x_origin < 0 ||
	    x_origin + n_local > n_global || frame0 < 0 || n_frames < 0 || frame0 + n_frames > T || out_pitch < F) {
		set_error("stft_range: bad size argument");
		return PAR_EINVAL;
	}
	if (n_frames == 0) return PAR_OK;
	// the slice must hold every sample its frames read: [frame0*hop - n_fft/2, (frame0+n_frames-1)*hop + n_fft/2),
	// reflected into [0, n_global) at the two ends of the signal
	const int64_t half = n_fft / 2;
	int64_t lo = frame0 * hop - half, hi = (frame0 + n_frames - 1) * hop - half + n_fft;
	if (lo < 0) { if (-lo + 1 > hi) hi = -lo + 1; lo = 0; }
	if (hi > n_global) { const int64_t r = 2 * (n_global - 1) - (hi - 1); if (r < lo) lo = r < 0 ? 0 : r; hi = n_global; }
	if (lo < x_origin || hi > x_origin + n_local) {
		set_error("stft_range: the slice does not cover the samples (with halo) its frames read");
		return PAR_EINVAL;
	}
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const float *dwin = device_window(device, window, n_fft, st);
	if (!dwin) return PAR_ECUDA;
	StftArgs a;
	a.x = x; a.n = n_global; a.x_stride = 1; a.x_ch_stride = x_ch_stride; a.x_origin = x_origin;
	a.n_ch = n_ch; a.n_fft = n_fft; a.hop = hop; a.zeropad = zeropad;
	a.n_frames = n_frames; a.frame0 = frame0; a.window = dwin; a.magnitude = (flags & PAR_OUT_MAGNITUDE) ? 1 : 0;
	// row 0 of `out` is frame `frame0`
	const size_t esz = a.magnitude ? sizeof(float) : sizeof(float2);
	a.out = (char *)out - (size_t)frame0 * out_pitch * esz;
	a.out_pitch = out_pitch; a.out_ch_stride = out_ch_stride;
	return launch_stft(a, device, st);
}

PAR_API int par_resample_range_f32(const double *pos, int64_t pos_origin, int64_t pos_count, int64_t m_global,
                           int64_t out_begin, int64_t out_end, const float *signal, int64_t sig_origin,
                           int64_t sig_count, int64_t n_in_global, int n_ch, int64_t sig_ch_stride, int mode, int nt,
                           float *out, int64_t out_stride, int64_t out_ch_stride,
                           unsigned flags, int device, void *stream) {
	if (!(flags & PAR_DEVICE_PTRS)) { set_error("resample_range: device pointers only"); return PAR_EUNSUPPORTED; }
	const bool sinc = mode == PAR_MODE_SINC;
	if ((mode != PAR_MODE_LINEAR && !sinc) || !pos || !signal || !out || n_ch < 1 || out_stride < 1 ||
	    (sinc && (nt < 1 || nt > 512)) || out_begin < 0 || out_end < out_begin || out_end > m_global ||
	    pos_origin > out_begin || pos_origin < 0 || sig_origin < 0 || sig_origin + sig_count > n_in_global) {
		set_error("resample_range: bad argument");
		return PAR_EINVAL;
	}
	// the positions slice must reach one past the last output (period of the last sample) unless that is the end
	const int64_t need_pos_end = out_end < m_global ? out_end + 1 : m_global;
	const int64_t need_pos_begin = (out_end == m_global && m_global >= 2 && out_begin > m_global - 2) ? m_global - 2 : out_begin;
	if (pos_origin > need_pos_begin || pos_origin + pos_count < need_pos_end) {
		set_error("resample_range: the positions slice does not cover [out_begin, out_end] (+1)");
		return PAR_EINVAL;
	}
	if (out_end == out_begin) return PAR_OK;
	int rc = use_device(device);
	if (rc != PAR_OK) return rc;
	SincArgs a;
	a.pos = pos; a.m = m_global; a.signal = signal; a.n_in = n_in_global; a.sig_stride = 1; a.sig_ch_stride = sig_ch_stride;
	a.n_ch = n_ch; a.nt = nt; a.out = out; a.out_stride = out_stride; a.out_ch_stride = out_ch_stride;
	a.aligned_edges = (flags & PAR_SINC_ALIGNED_EDGES) ? 1 : 0;
	a.kernel = (flags & PAR_SINC_KERNEL_WS) ? 2 : ((flags & PAR_SINC_KERNEL_TILED) ? 1 : 0);
	a.out_begin = out_begin; a.out_end = out_end;
	a.pos_origin = pos_origin; a.sig_origin = sig_origin; a.out_origin = out_begin;
	return sinc ? launch_sinc(a, device, (cudaStream_t)stream) : launch_linear(a, device, (cudaStream_t)stream);
}

}  // extern "C"
