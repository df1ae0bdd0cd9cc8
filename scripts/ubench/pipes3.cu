// Micro-benchmark 3 (round 2): IDP.4A (dp4a) as a byte extractor / small linear combiner -- which pipe does it use
// on sm_100a, and does it overlap with LOP3 (ALU pipe) and IMAD (FMA pipe)?  Dev aid, not product code.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)
{
	int d;
	asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned seed, int iters)
{
	unsigned a[8];
#pragma unroll
	for (int i = 0; i < 8; i++) a[i] = seed * 2654435761u + threadIdx.x * 7 + i * 0x10003;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const unsigned y = a[(i + 1) & 7], z = a[(i + 3) & 7];
			if (MODE == 0) a[i] = (unsigned)dp4a_us(y, 0x00000800, (int)a[i]);                       // IDP.4A, immediate selector
			if (MODE == 1) a[i] = (unsigned)dp4a_us(y, (int)z, (int)a[i]);                            // IDP.4A, register selector
			if (MODE == 2) { a[i] = (unsigned)dp4a_us(y, 0x00000800, (int)a[i]); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; }   // + LOP3
			if (MODE == 3) { a[i] = (unsigned)dp4a_us(y, 0x00000800, (int)a[i]); a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }    // + IMAD
			if (MODE == 4) { a[i] = (unsigned)dp4a_us(y, 0x00000800, (int)a[i]); a[(i + 2) & 7] = __byte_perm(a[(i + 2) & 7], z, 0x5410 + (it & 1)); }  // + PRMT
			if (MODE == 5) a[i] = (a[i] & y) ^ z;                                                     // LOP3 (reference)
			if (MODE == 6) { a[i] = (unsigned)dp4a_us(y, 0x00000800, (int)a[i]); a[(i + 2) & 7] = __vadd2(a[(i + 2) & 7], z); }  // + VIADD.16x2
		}
	}
	unsigned s = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name)
{
	unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 10000;
	k<MODE><<<148 * 8, 256>>>(d, 1, 100);
	cudaEventRecord(e0);
	k<MODE><<<148 * 8, 256>>>(d, 1, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	printf("%-34s %8.3f ms  (one full-rate instruction per slot = 1.33 ms)\n", name, ms);
	cudaFree(d);
}
int main()
{
	run<5>("LOP3"); run<0>("IDP.4A imm"); run<1>("IDP.4A reg"); run<2>("IDP.4A + LOP3"); run<3>("IDP.4A + IMAD");
	run<4>("IDP.4A + PRMT"); run<6>("IDP.4A + VIADD.16x2");
	return 0;
}
