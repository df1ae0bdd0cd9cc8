// ric_entropy_gpu.h -- launchers of the device entropy stage (ric_entropy_gpu.cu: one image per warp).
// g and tables are DEVICE pointers (tables: the POD block of ric::entropy_tables()).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ric_host.h"

namespace ric {

cudaError_t launch_entropy_encode(const HostGeom *g, const void *tables, char *arenas, size_t img_ar, uint8_t *out, size_t stride,
                                  long long *sizes, int n, cudaStream_t st);
// Pre-pass + hinted coder (ric_entropy_core.h); hints: device buffer of hint_bytes(hg, n) bytes; hg = host copy of *g.
inline size_t hint_bytes(const HostGeom &hg, int n) { return (size_t)hg.flag_bytes * hg.channels * n * 8; }
cudaError_t launch_entropy_encode_hinted(const HostGeom *g, const HostGeom &hg, const void *tables, char *arenas, size_t img_ar, void *hints,
                                         uint8_t *out, size_t stride, long long *sizes, int n, cudaStream_t st);
cudaError_t launch_entropy_decode(const HostGeom *g, const void *tables, const uint8_t *payloads, size_t stride, const long long *sizes,
                                  char *arenas, size_t img_ar, int *bad, int *status, int n, cudaStream_t st);

}  // namespace ric
