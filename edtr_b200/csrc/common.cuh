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

#define EDTR_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      edtr::set_error(__VA_ARGS__);  \
      return EDTR_ERR_INVALID;       \
    }                                \
  } while (0)

}  // namespace edtr
