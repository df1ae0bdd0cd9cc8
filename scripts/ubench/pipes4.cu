// Micro-benchmark 4 (round 2): cvt.pack.sat.u8.s32 (I2IP: clip two ints to bytes and pack them) and IDP.2A
// (sign-extend / add a 16-bit half) -- rate and pipe on sm_100a.  Dev aid, not product code.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned pack_sat(int a, int b, unsigned c)
{
	unsigned d;
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ int dp2a_lo(unsigned a, unsigned b, int c)
{
	int d;
	asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
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
			if (MODE == 0) a[i] = pack_sat((int)a[i], (int)y, z);                                         // I2IP
			if (MODE == 1) { a[i] = pack_sat((int)a[i], (int)y, z); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; }   // + LOP3
			if (MODE == 2) { a[i] = pack_sat((int)a[i], (int)y, z); a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }    // + IMAD
			if (MODE == 3) a[i] = (unsigned)dp2a_lo(y, 0x0100u, (int)a[i]);                                 // IDP.2A
			if (MODE == 4) { a[i] = (unsigned)dp2a_lo(y, 0x0100u, (int)a[i]); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; }  // + LOP3
			if (MODE == 5) a[i] = (a[i] & y) ^ z;                                                           // LOP3 (reference)
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
	run<5>("LOP3"); run<0>("I2IP.U8.S32.SAT"); run<1>("I2IP + LOP3"); run<2>("I2IP + IMAD"); run<3>("IDP.2A"); run<4>("IDP.2A + LOP3");
	return 0;
}
