// Throughput of the sm_100a FP32 / packed-FP32 / MUFU / LDS pipes as seen by one SM sub-partition:
// clocks per warp instruction with 8 warps per sub-partition issuing long runs of independent ops.
// Development aid for csrc/sinc_core.cuh (build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
#define ITER 200

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

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float *out, long long *cyc, const float *in) {
	__shared__ float2 sm[2048];
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i * 1e-3f, 1.f);
	__syncthreads();
	float2 a[8], b = make_float2(in[0], in[1]), c = make_float2(in[2], in[3]);
	float s[8];
#pragma unroll
	for (int i = 0; i < 8; i++) { a[i] = make_float2(in[i], in[i + 1]); s[i] = in[i]; }
	const float2 *xp = sm + (threadIdx.x & 31);
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < ITER; it++) {
#pragma unroll
		for (int r = 0; r < REP / 8; r++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MODE == 0) s[i] = ffma(s[i], b.x, c.x);                     // FFMA, 2 of 3 sources shared
				if (MODE == 1) a[i] = ffma2(a[i], b, c);                        // FFMA2
				if (MODE == 2) a[i] = fmul2(a[i], b);                           // FMUL2
				if (MODE == 3) a[i] = fadd2(a[i], b);                           // FADD2
				if (MODE == 4) s[i] = rcp(s[i]);                                // MUFU.RCP
				if (MODE == 5) { a[i] = ffma2(a[i], b, c); s[i] = ffma(s[i], b.x, c.x); }   // FFMA2 + FFMA
				if (MODE == 6) { a[i] = ffma2(a[i], b, c); s[i] = rcp(s[i]); }              // FFMA2 + MUFU
				if (MODE == 7) { a[i] = ffma2(xp[(r * 8 + i) * 32], b, a[i]); }             // LDS.64 + FFMA2
				if (MODE == 8) { s[i] = ffma(s[i], a[i].x, a[(i + 1) & 7].y); }             // FFMA 3 distinct sources
				if (MODE == 9) { a[i] = ffma2(a[i], a[(i + 3) & 7], a[(i + 5) & 7]); }      // FFMA2 3 distinct sources
				if (MODE == 10) { a[i] = ffma2(a[i], make_float2(b.x, b.x), c); }           // FFMA2 with broadcast scalar
				if (MODE == 11) { s[i] = s[i] * b.x; }                                      // FMUL
				if (MODE == 12) { a[i] = ffma2(xp[(r * 8 + i) * 32], b, a[i]); a[(i + 1) & 7] = ffma2(xp[(r * 8 + i) * 32], c, a[(i + 1) & 7]); } // 1 LDS.64 + 2 FFMA2
			}
		}
	}
	const long long t1 = clock64();
	float acc = 0.f;
#pragma unroll
	for (int i = 0; i < 8; i++) acc += a[i].x + a[i].y + s[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops_per_rep, float *out, long long *cyc, const float *in) {
	bench<MODE><<<148, 1024>>>(out, cyc, in);
	cudaDeviceSynchronize();
	bench<MODE><<<148, 1024>>>(out, cyc, in);
	cudaError_t e = cudaDeviceSynchronize();
	long long h[148];
	cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < 148; i++) avg += h[i];
	avg /= 148;
	// 8 warps per sub-partition, each REP*ITER*ops instructions
	const double per = avg / (8.0 * REP * ITER * ops_per_rep);
	printf("%-44s %8.3f clk per warp instruction per sub-partition  (%s)\n", name, per, cudaGetErrorString(e));
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
	run<0>("FFMA (acc, shared b, c)", 1, out, cyc, in);
	run<8>("FFMA (3 distinct sources)", 1, out, cyc, in);
	run<11>("FMUL", 1, out, cyc, in);
	run<1>("FFMA2 (acc, shared b, c)", 1, out, cyc, in);
	run<9>("FFMA2 (3 distinct sources)", 1, out, cyc, in);
	run<10>("FFMA2 (broadcast scalar operand)", 1, out, cyc, in);
	run<2>("FMUL2", 1, out, cyc, in);
	run<3>("FADD2", 1, out, cyc, in);
	run<4>("MUFU.RCP", 1, out, cyc, in);
	run<5>("FFMA2 + FFMA (per instruction)", 2, out, cyc, in);
	run<6>("FFMA2 + MUFU.RCP (per instruction)", 2, out, cyc, in);
	run<7>("LDS.64 + FFMA2 (per pair)", 1, out, cyc, in);
	run<12>("LDS.64 + 2 FFMA2 (per triple)", 1, out, cyc, in);
	return 0;
}
