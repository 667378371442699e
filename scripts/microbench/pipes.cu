// Throughput of the sm_100a FP32 / packed-FP32 / MUFU / LDS pipes as seen by one SM sub-partition:
// clocks per warp instruction with W warps per sub-partition issuing long runs of independent ops.
// Development aid for csrc/sinc_core.cuh (build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 400

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
	float2 d;
	asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(uint64_t *)&d) : "l"(*(uint64_t *)&a), "l"(*(uint64_t *)&b), "l"(*(uint64_t *)&c));
	return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
	float2 d;
	asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(*(uint64_t *)&d) : "l"(*(uint64_t *)&a), "l"(*(uint64_t *)&b));
	return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
	float2 d;
	asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(*(uint64_t *)&d) : "l"(*(uint64_t *)&a), "l"(*(uint64_t *)&b));
	return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
	float d;
	asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
	return d;
}
__device__ __forceinline__ float rcp(float a) {
	float d;
	asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
	return d;
}
__device__ __forceinline__ float2 lds64(const float2 *p) {
	float2 d;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(d.x), "=f"(d.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
	return d;
}

// MODE: see main().  Every thread keeps 8 independent chains.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float *out, long long *cyc, const float *in, int iters) {
	__shared__ float2 sm[4096];
	for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float2(1.f + i * 1e-6f, 1.f);
	__syncthreads();
	float2 a[8], b = make_float2(in[0], in[1]), c = make_float2(in[2], in[3]);
	float s[8];
#pragma unroll
	for (int i = 0; i < 8; i++) { a[i] = make_float2(in[i], in[i + 1]); s[i] = in[i]; }
	const float2 *xp = sm + (threadIdx.x & 31);
	__syncthreads();
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MODE == 0) s[i] = ffma(s[i], b.x, c.x);
				if (MODE == 1) s[i] = ffma(s[i], a[i].x, a[(i + 1) & 7].y);
				if (MODE == 2) a[i] = ffma2(a[i], b, c);
				if (MODE == 3) a[i] = ffma2(a[i], a[(i + 3) & 7], a[(i + 5) & 7]);
				if (MODE == 4) a[i] = fmul2(a[i], b);
				if (MODE == 5) s[i] = rcp(s[i]);
				if (MODE == 6) { s[i] = rcp(s[i]); a[i] = ffma2(a[i], b, c); }
				if (MODE == 7) { s[i] = rcp(s[i]); a[i] = ffma2(a[i], b, c); a[(i + 1) & 7] = fmul2(a[(i + 1) & 7], c); }
				if (MODE == 8) { s[i] = rcp(s[i]); a[i] = ffma2(a[i], b, c); a[(i + 1) & 7] = fmul2(a[(i + 1) & 7], c); a[(i + 2) & 7] = ffma2(a[(i + 2) & 7], c, b); a[(i + 3) & 7] = fmul2(a[(i + 3) & 7], b); }
				if (MODE == 9) { float2 x = lds64(xp + (r * 8 + i) * 32); a[i] = fadd2(a[i], x); }
				if (MODE == 10) { float2 x = lds64(xp + (r * 8 + i) * 32); a[i] = ffma2(x, make_float2(s[i], s[i]), a[i]); a[(i + 1) & 7] = ffma2(x, make_float2(s[(i + 1) & 7], s[(i + 1) & 7]), a[(i + 1) & 7]); }
				if (MODE == 11) { float2 x = lds64(xp + (r * 8 + i) * 32); a[i] = ffma2(x, make_float2(s[i], s[i]), a[i]); a[(i + 1) & 7] = ffma2(x, make_float2(s[(i + 1) & 7], s[(i + 1) & 7]), a[(i + 1) & 7]);
				                  a[(i + 2) & 7] = ffma2(x, make_float2(s[(i + 2) & 7], s[(i + 2) & 7]), a[(i + 2) & 7]); a[(i + 3) & 7] = ffma2(x, make_float2(s[(i + 3) & 7], s[(i + 3) & 7]), a[(i + 3) & 7]); }
				if (MODE == 12) { float2 x = lds64(xp + (r * 8 + i) * 32); s[i] = ffma(x.x, s[(i + 1) & 7], s[i]); s[(i + 2) & 7] = ffma(x.y, s[(i + 1) & 7], s[(i + 2) & 7]); s[(i + 4) & 7] = ffma(x.x, s[(i + 3) & 7], s[(i + 4) & 7]); s[(i + 6) & 7] = ffma(x.y, s[(i + 3) & 7], s[(i + 6) & 7]); }
			}
		}
	}
	float acc = 0.f;
#pragma unroll
	for (int i = 0; i < 8; i++) acc += a[i].x + a[i].y + s[i];
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double ops_per_slot, int threads, float *out, long long *cyc, const float *in) {
	bench<MODE><<<148, threads>>>(out, cyc, in, ITER);
	cudaDeviceSynchronize();
	bench<MODE><<<148, threads>>>(out, cyc, in, ITER);
	cudaError_t e = cudaDeviceSynchronize();
	long long h[148];
	cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < 148; i++) avg += h[i];
	avg /= 148;
	const double warps_per_smsp = threads / 32 / 4.0;
	const double per = avg / (warps_per_smsp * 32.0 * ITER);          // clocks per "slot" (one pass of the innermost statement) per warp
	printf("%-64s %2.0f w/smsp  %7.2f clk per slot = %6.2f clk per instruction  (%s)\n", name, warps_per_smsp, per, per / ops_per_slot,
	       cudaGetErrorString(e));
}

int main() {
	float *out, *in;
	long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 4);
	cudaMalloc(&cyc, 148 * 8);
	cudaMalloc(&in, 64 * 4);
	float hin[64];
	for (int i = 0; i < 64; i++) hin[i] = 1.0f + i * 1e-3f;
	cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
	for (int threads : {1024, 512}) {
		run<0>("FFMA (acc, 2 shared sources)", 1, threads, out, cyc, in);
		run<1>("FFMA (3 distinct sources)", 1, threads, out, cyc, in);
		run<2>("FFMA2 (acc, 2 shared sources)", 1, threads, out, cyc, in);
		run<3>("FFMA2 (3 distinct sources)", 1, threads, out, cyc, in);
		run<4>("FMUL2", 1, threads, out, cyc, in);
		run<5>("MUFU.RCP", 1, threads, out, cyc, in);
		run<6>("MUFU.RCP + FFMA2", 2, threads, out, cyc, in);
		run<7>("MUFU.RCP + FFMA2 + FMUL2", 3, threads, out, cyc, in);
		run<8>("MUFU.RCP + 2 FFMA2 + 2 FMUL2", 5, threads, out, cyc, in);
		run<9>("LDS.64 + FADD2", 2, threads, out, cyc, in);
		run<10>("LDS.64 + 2 FFMA2 (x pair, w broadcast, acc)", 3, threads, out, cyc, in);
		run<11>("LDS.64 + 4 FFMA2 (x pair, w broadcast, acc)", 5, threads, out, cyc, in);
		run<12>("LDS.64 + 4 FFMA (x, w, acc)", 5, threads, out, cyc, in);
	}
	return 0;
}
