// Micro-benchmark 2 (round 2): which pipe do the packed 16x2 / "video" integer instructions of sm_100a use,
// and what do they sustain alone and mixed with LOP3 (ALU pipe) / IMAD (FMA pipe)?  Dev aid, not product code.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned vadd2(unsigned a, unsigned b) { return __vadd2(a, b); }
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
			if (MODE == 0) a[i] = vadd2(a[i], y);                                   // VIADD.16x2
			if (MODE == 1) a[i] = __viaddmax_s16x2(a[i], y, z);                     // VIADDMNMX.S16x2
			if (MODE == 2) a[i] = __vimax3_s16x2(a[i], y, z);                       // VIMNMX3.S16x2
			if (MODE == 3) a[i] = __vmaxs2(a[i], y);                                // VIMNMX.S16x2
			if (MODE == 4) a[i] = (a[i] & y) ^ z;                                   // LOP3
			if (MODE == 5) a[i] = a[i] * 3u + y;                                    // IMAD
			if (MODE == 6) a[i] = __byte_perm(a[i], y, 0x5410 + (it & 1));          // PRMT
			if (MODE == 7) a[i] = (unsigned)((int)(a[i] ^ y) >> 3);                 // LOP3 + SHF
			if (MODE == 8) { a[i] = vadd2(a[i], y); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; }          // VIADD.16x2 + LOP3
			if (MODE == 9) { a[i] = vadd2(a[i], y); a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }           // VIADD.16x2 + IMAD
			if (MODE == 10) { a[i] = (a[i] & y) ^ z; a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }          // LOP3 + IMAD
			if (MODE == 11) { a[i] = __vimax3_s16x2(a[i], y, z); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; }  // VIMNMX3 + LOP3
			if (MODE == 12) { a[i] = __vimax3_s16x2(a[i], y, z); a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }   // VIMNMX3 + IMAD
			if (MODE == 13) { a[i] = vadd2(a[i], y); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; a[(i + 4) & 7] = a[(i + 4) & 7] * 3u + z; }  // all three
			if (MODE == 14) a[i] = a[i] + y + z;                                    // IADD3
			if (MODE == 15) { a[i] = a[i] + y + z; a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }            // IADD3 + IMAD
			if (MODE == 16) a[i] = a[i] + 0x12345u;                                 // add imm (VIADD / IADD3 / IMAD.IADD: compiler's pick)
			if (MODE == 17) { a[i] = __viaddmax_s16x2(a[i], y, z); a[(i + 2) & 7] = a[(i + 2) & 7] * 3u + z; }  // VIADDMNMX + IMAD
			if (MODE == 18) { a[i] = __viaddmax_s16x2(a[i], y, z); a[(i + 2) & 7] = (a[(i + 2) & 7] & z) ^ y; }  // VIADDMNMX + LOP3
		}
	}
	unsigned s = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int per_iter)
{
	unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 10000;
	k<MODE><<<148 * 8, 256>>>(d, 1, 100);
	cudaEventRecord(e0);
	k<MODE><<<148 * 8, 256>>>(d, 1, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double ops = (double)148 * 8 * 256 * iters * 8 * per_iter;
	printf("%-34s %8.3f ms  %6.1f lane-ops/clk/SM at 1.965 GHz (%d instr/iter-slot)\n", name, ms, ops / ms / 1e3 / 148 / 1.965e6, per_iter);
	cudaFree(d);
}
int main()
{
	run<0>("VIADD.16x2", 1); run<1>("VIADDMNMX.S16x2", 1); run<2>("VIMNMX3.S16x2", 1); run<3>("VIMNMX.S16x2", 1);
	run<4>("LOP3", 1); run<5>("IMAD", 1); run<6>("PRMT", 1); run<7>("LOP3+SHF (both ALU)", 2);
	run<8>("VIADD.16x2 + LOP3", 2); run<9>("VIADD.16x2 + IMAD", 2); run<10>("LOP3 + IMAD", 2);
	run<11>("VIMNMX3 + LOP3", 2); run<12>("VIMNMX3 + IMAD", 2); run<13>("VIADD.16x2 + LOP3 + IMAD", 3);
	run<14>("IADD3", 1); run<15>("IADD3 + IMAD", 2); run<16>("add imm", 1);
	run<17>("VIADDMNMX + IMAD", 2); run<18>("VIADDMNMX + LOP3", 2);
	return 0;
}
