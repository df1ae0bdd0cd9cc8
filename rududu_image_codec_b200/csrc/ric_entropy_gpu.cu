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
                                                            const long long *sizes, char *arenas, size_t img_ar, int *bad, int *status, int n)
{
	const int img = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if (img >= n || (threadIdx.x & 31)) return;
	const HostGeom &g = *gp;
	if (status) status[img] = 0;
	if (sizes[img] < 0 || (unsigned long long)sizes[img] > stride) {  // not a length this slot can hold: leave the (cleared) arenas alone
		atomicExch(bad, 1);
		if (status) status[img] = 1;
		return;
	}
	ent::MuxReader r(payloads + (size_t)img * stride, (size_t)sizes[img]);
	ent::RPort io(r, T);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;
		ent::walk_plane(io, g, arenas + (size_t)img * img_ar + (size_t)plane * g.arena_bytes);
	}
	if (r.overrun()) {
		atomicExch(bad, 1);
		if (status) status[img] = 1;
	}
}


// ---- encode with the parallel pre-pass (block hints, ric_entropy_core.h) ---------------------------------
// hint_kernel: one thread per 4x4 block of every D/H/V band of every plane of every image.
__global__ void __launch_bounds__(256) hint_kernel(const HostGeom *gp, const ent::Tables *T, const char *arenas, size_t img_ar,
                                                   ent::BlockHint *hints, int nplanes_total)
{
	const HostGeom &g = *gp;
	const long long slots = (long long)g.flag_bytes;
	const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= slots * nplanes_total) return;
	const int pl = (int)(gid / slots);
	const int slot = (int)(gid % slots);
	// which band / block is this slot?  (bands are laid out back to back, each padded to 16 slots)
	int id = 0;
	while (id + 1 < 3 * g.nlev && slot >= g.flag_off[id + 1]) id++;
	const int rel = slot - g.flag_off[id], bw = g.flag_bw[id];
	const int by = rel / bw, bx = rel % bw;
	if (by >= (g.band[id].dimy + 3) / 4) return;  // padding slot
	const int img = pl / g.channels, plane = pl % g.channels;
	hints[gid] = ent::make_hint(g, *T, arenas + (size_t)img * img_ar + (size_t)plane * g.arena_bytes, id, bx, by);
}

__global__ void __launch_bounds__(64) entropy_encode_hinted_kernel(const HostGeom *gp, const ent::Tables *T, char *arenas, size_t img_ar,
                                                                   const ent::BlockHint *hints, uint8_t *out, size_t stride,
                                                                   long long *sizes, int n)
{
	const int img = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if (img >= n || (threadIdx.x & 31)) return;
	const HostGeom &g = *gp;
	uint8_t *o = out + (size_t)img * stride;
	ent::MuxWriter w(o, stride);
	ent::WPort io(w, T);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;
		ent::walk_plane_hinted(io, g, arenas + (size_t)img * img_ar + (size_t)plane * g.arena_bytes,
		                       hints + ((size_t)img * g.channels + plane) * g.flag_bytes);
	}
	uint8_t *end = w.finish();
	sizes[img] = w.overflow() ? -1 : (long long)(end - o);
}

cudaError_t launch_entropy_encode_hinted(const HostGeom *g, const HostGeom &hg, const void *tables, char *arenas, size_t img_ar, void *hints,
                                         uint8_t *out, size_t stride, long long *sizes, int n, cudaStream_t st)
{
	const long long total = (long long)hg.flag_bytes * hg.channels * n;
	hint_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, (const ent::Tables *)tables, arenas, img_ar, (ent::BlockHint *)hints, hg.channels * n);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	entropy_encode_hinted_kernel<<<(n + 1) / 2, 64, 0, st>>>(g, (const ent::Tables *)tables, arenas, img_ar, (const ent::BlockHint *)hints, out, stride,
	                                                         sizes, n);
	return cudaGetLastError();
}

cudaError_t launch_entropy_encode(const HostGeom *g, const void *tables, char *arenas, size_t img_ar, uint8_t *out, size_t stride,
                                  long long *sizes, int n, cudaStream_t st)
{
	entropy_encode_kernel<<<(n + 1) / 2, 64, 0, st>>>(g, (const ent::Tables *)tables, arenas, img_ar, out, stride, sizes, n);
	return cudaGetLastError();
}

cudaError_t launch_entropy_decode(const HostGeom *g, const void *tables, const uint8_t *payloads, size_t stride, const long long *sizes,
                                  char *arenas, size_t img_ar, int *bad, int *status, int n, cudaStream_t st)
{
	entropy_decode_kernel<<<(n + 1) / 2, 64, 0, st>>>(g, (const ent::Tables *)tables, payloads, stride, sizes, arenas, img_ar, bad, status, n);
	return cudaGetLastError();
}

}  // namespace ric
