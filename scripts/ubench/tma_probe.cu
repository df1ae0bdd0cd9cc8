// Stand-alone check of the bulk tensor (TMA) load used by ric_fwd0.cuh's experiment: u8 image, box 256 x 2, negative x.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, unsigned char *out, int x, int y)
{
	__shared__ alignas(128) unsigned char tile[512];
	__shared__ unsigned long long bar;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		             ::"r"(smem_u32(tile)), "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
	}
	asm volatile(
	    "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
	for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = tile[i];
}
int main()
{
	const int W = 512, H = 64, pitch = 512;
	std::vector<unsigned char> h(pitch * H);
	for (int i = 0; i < pitch * H; i++) h[i] = (unsigned char)(i * 7 + (i >> 9));
	unsigned char *d, *o;
	cudaMalloc(&d, h.size()); cudaMalloc(&o, 512);
	cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
	typedef CUresult (*fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
	                         const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void *p = nullptr; cudaDriverEntryPointQueryResult q;
	cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
	CUtensorMap map;
	const cuuint64_t dim[2] = {W, H}, stride[1] = {pitch};
	const cuuint32_t box[2] = {256, 2}, es[2] = {1, 1};
	CUresult r = ((fn_t)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode: %d\n", (int)r);
	for (int t = 0; t < 3; t++) {
		const int x = t == 0 ? 0 : t == 1 ? -8 : 264, y = t == 2 ? 63 : 4;
		k<<<1, 128>>>(map, o, x, y);
		e = cudaDeviceSynchronize();
		printf("x=%d y=%d kernel: %s\n", x, y, cudaGetErrorString(e));
		if (e != cudaSuccess) return 1;
		std::vector<unsigned char> g(512);
		cudaMemcpy(g.data(), o, 512, cudaMemcpyDeviceToHost);
		int bad = 0;
		for (int r2 = 0; r2 < 2; r2++)
			for (int c = 0; c < 256; c++) {
				const int xx = x + c, yy = y + r2;
				const unsigned char want = (xx < 0 || xx >= W || yy >= H) ? 0 : h[yy * pitch + xx];
				bad += g[r2 * 256 + c] != want;
			}
		printf("  mismatches %d\n", bad);
	}
	return 0;
}
