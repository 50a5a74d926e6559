// Microbenchmark: does a DFMA with three DISTINCT 64-bit register operands issue slower than one with
// shared / immediate operands (register-file bank limit)?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed)
{
	double a[8], x[8], y[8];
#pragma unroll
	for (int q = 0; q < 8; q++) { a[q] = seed + threadIdx.x + q; x[q] = 1.0 - 1e-9 * (q + 1) * seed; y[q] = 1e-9 * (q + 2) * seed; }
	const double m = 0.999999 * seed, b = 1.0e-9 * seed;
	for (int i = 0; i < iters; i += 16) {
#pragma unroll
		for (int u = 0; u < 16; u++) {
#pragma unroll
			for (int q = 0; q < 8; q++) {
				if (MODE == 0) a[q] = fma(a[q], m, b);                 // 1 varying + 2 shared register operands
				if (MODE == 1) a[q] = fma(x[q], y[q], a[q]);           // 3 distinct register operands
				if (MODE == 2) a[q] = fma(x[q], m, a[q]);              // 2 distinct + 1 shared
				if (MODE == 3) a[q] = a[q] * x[q];                     // DMUL 2 distinct
				if (MODE == 4) a[q] = a[q] + x[q];                     // DADD 2 distinct
				if (MODE == 5) a[q] = fma(x[q], y[(q + 1) & 7], a[q]); // 3 distinct, different pairing
			}
		}
	}
	double s = 0;
#pragma unroll
	for (int q = 0; q < 8; q++) s += a[q] + x[q] + y[q];
	if (s == 12345.678) out[0] = s;
}

template <int MODE>
static void run(const char *name)
{
	double *out; cudaMalloc(&out, 8);
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int blocks = p.multiProcessorCount * 8, iters = 1 << 15;
	k<MODE><<<blocks, 256>>>(out, 1024, 1.0);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e30f;
	for (int r = 0; r < 3; r++) {
		cudaEventRecord(a); k<MODE><<<blocks, 256>>>(out, iters, 1.0); cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
	}
	double inst = (double)blocks * 256 * 8.0 * iters;
	printf("%-40s %8.3f ms  %7.2f Ginstr/s  (%.2f TFLOP/s as FMA)\n", name, best, inst / best / 1e6, inst * 2 / best / 1e9);
	cudaFree(out);
}

int main()
{
	run<0>("DFMA a = fma(a, m, b)   1 distinct");
	run<2>("DFMA a = fma(x, m, a)   2 distinct");
	run<1>("DFMA a = fma(x, y, a)   3 distinct");
	run<5>("DFMA a = fma(x, y', a)  3 distinct");
	run<3>("DMUL a = a * x          2 distinct");
	run<4>("DADD a = a + x          2 distinct");
	return 0;
}
