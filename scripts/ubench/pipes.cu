// Micro-benchmark: throughput of the integer instructions the lifting code is made of (dev aid).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int *out, int seed, int iters)
{
	int a[8];
#pragma unroll
	for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 7 + i;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (MODE == 0) a[i] = a[i] + (a[(i + 1) & 7] >> 4);                 // LEA.HI.SX32 (ALU)
			if (MODE == 1) a[i] = __mulhi(a[(i + 1) & 7], 1 << 28) + a[i];      // IMAD.HI (FMA?)
			if (MODE == 2) a[i] = a[i] * 3 + a[(i + 1) & 7];                    // IMAD (FMA)
			if (MODE == 3) a[i] = (int)(short)(a[i] + a[(i + 1) & 7]);          // IADD + PRMT
			if (MODE == 4) { a[i] = a[i] + (a[(i + 1) & 7] >> 4); a[(i + 2) & 7] = a[(i + 2) & 7] * 3 + a[i]; }  // mixed ALU + FMA
			if (MODE == 5) { a[i] = a[i] + (a[(i + 1) & 7] >> 4); a[(i + 2) & 7] = __mulhi(a[(i + 3) & 7], 1 << 28) + a[(i + 2) & 7]; }  // LEA.HI + IMAD.HI
		}
	}
	int s = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int per_iter)
{
	int *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 20000;
	k<MODE><<<148 * 8, 256>>>(d, 1, 100);
	cudaEventRecord(e0);
	k<MODE><<<148 * 8, 256>>>(d, 1, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double ops = (double)148 * 8 * 256 * iters * 8 * per_iter;
	printf("%-28s %8.3f ms  %7.2f Tinstr/s (lane-ops)  = %.1f lane-ops/clk/SM at 1.965 GHz\n", name, ms, ops / ms / 1e9, ops / ms / 1e3 / 148 / 1.965e6);
	cudaFree(d);
}
int main()
{
	run<0>("x + (y>>4)  LEA.HI.SX32", 1);
	run<1>("mulhi(y,2^28)+x IMAD.HI", 1);
	run<2>("x*3+y IMAD", 1);
	run<3>("(short)(x+y) IADD+PRMT", 2);
	run<4>("LEA.HI + IMAD mixed", 2);
	run<5>("LEA.HI + IMAD.HI mixed", 2);
	return 0;
}
