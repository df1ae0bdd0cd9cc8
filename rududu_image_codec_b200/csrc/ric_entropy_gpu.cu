// ric_entropy_gpu.cu -- the entropy stage of the .ric format on the device, for large batches.
//
// The bit stream of one image is strictly serial (ric_entropy.h), so the device runs ONE IMAGE PER WARP with
// lane 0 executing the same coder source as the host (ric_entropy_core.h); the parallelism is across images.
// An SM sub-partition issues one warp instruction per clock whatever the number of active lanes, so the stage's
// throughput is images-in-flight x issue slots, and it pays off only when thousands of images are resident
// (BASELINE configs[3]: 4096 x 1080p).  Its point is what it removes: the band arenas (2 bytes per sample) no
// longer cross PCIe, only the finished payload (about 0.06 bytes per sample at q = 9) does.
#include "ric_entropy_gpu.h"

#include "ric_entropy_core.h"

namespace ric {

__global__ void __launch_bounds__(64) entropy_encode_kernel(const HostGeom *gp, const ent::Tables *T, char *arenas, size_t img_ar,
                                                            uint8_t *out, size_t stride, long long *sizes, int n)
{
	const int img = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if (img >= n || (threadIdx.x & 31)) return;
	const HostGeom &g = *gp;
	uint8_t *o = out + (size_t)img * stride;
	ent::MuxWriter w(o, stride);
	ent::WPort io(w, T);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;  // Y, Cg, Co (ric.cpp:163-168)
		ent::walk_plane(io, g, arenas + (size_t)img * img_ar + (size_t)plane * g.arena_bytes);
	}
	uint8_t *end = w.finish();
	sizes[img] = w.overflow() ? -1 : (long long)(end - o);
}

// The arenas must have been cleared by the caller (CBand::Clear of every band, bandcodec.cpp:503).
__global__ void __launch_bounds__(64) entropy_decode_kernel(const HostGeom *gp, const ent::Tables *T, const uint8_t *payloads, size_t stride,
                                                            const long long *sizes, char *arenas, size_t img_ar, int *bad, int n)
{
	const int img = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if (img >= n || (threadIdx.x & 31)) return;
	const HostGeom &g = *gp;
	ent::MuxReader r(payloads + (size_t)img * stride, (size_t)sizes[img]);
	ent::RPort io(r, T);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;
		ent::walk_plane(io, g, arenas + (size_t)img * img_ar + (size_t)plane * g.arena_bytes);
	}
	if (r.overrun()) atomicExch(bad, 1);
}

cudaError_t launch_entropy_encode(const HostGeom *g, const void *tables, char *arenas, size_t img_ar, uint8_t *out, size_t stride,
                                  long long *sizes, int n, cudaStream_t st)
{
	entropy_encode_kernel<<<(n + 1) / 2, 64, 0, st>>>(g, (const ent::Tables *)tables, arenas, img_ar, out, stride, sizes, n);
	return cudaGetLastError();
}

cudaError_t launch_entropy_decode(const HostGeom *g, const void *tables, const uint8_t *payloads, size_t stride, const long long *sizes,
                                  char *arenas, size_t img_ar, int *bad, int n, cudaStream_t st)
{
	entropy_decode_kernel<<<(n + 1) / 2, 64, 0, st>>>(g, (const ent::Tables *)tables, payloads, stride, sizes, arenas, img_ar, bad, n);
	return cudaGetLastError();
}

}  // namespace ric
