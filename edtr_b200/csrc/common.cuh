// Host-side helpers shared by the C-ABI translation units: error reporting,
// the cuTensorMapEncodeTiled driver entry point, launch checks.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/edtr_b200.h"

namespace edtr {

void set_error(const char* fmt, ...);

// Encodes a bf16 tiled tensor map with 128-byte swizzle. dims/box innermost first;
// strides_bytes has rank-1 entries (dims 1..rank-1). Returns 0 on success.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);

int check_launch(const char* what);

// Programmatic dependent launch: every kernel is launched with the stream-serialization attribute so its
// prologue (barrier init, TMEM allocation, descriptor prefetch, index math) overlaps the tail of the previous
// kernel; on the device every kernel calls pdl_launch_dependents() first and pdl_wait() before it touches
// global memory.  On by default, EDTR_PDL=0 disables it (without the attribute the device calls are no-ops).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define EDTR_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)edtr::launch_kernel(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)

#define EDTR_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      edtr::set_error(__VA_ARGS__);  \
      return EDTR_ERR_INVALID;       \
    }                                \
  } while (0)

}  // namespace edtr
