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
cudaError_t launch_entropy_decode(const HostGeom *g, const void *tables, const uint8_t *payloads, size_t stride, const long long *sizes,
                                  char *arenas, size_t img_ar, int *bad, int n, cudaStream_t st);

}  // namespace ric
