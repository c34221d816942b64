// Library plumbing of the C-ABI: error string, device check, TMA descriptor
// encoding through the driver entry point (no link-time libcuda dependency).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace edtr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  // Programmatic dependent launch, with LATE triggers (see ptx.cuh / gemm2.cu).  Measured on B200 inside the
  // CUDA graphs: triggers at kernel entry cost ~3 % (waiting dependents take registers / warp slots a multi-wave
  // predecessor still needs), late triggers gain ~1.3 % on the 4-step sample; all GPU tests pass either way.
  // EDTR_PDL=0 turns it off.
  static const bool on = [] {
    const char* e = getenv("EDTR_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return EDTR_ERR_CUDA;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) {
      set_error("tensor map box[%d] = %u out of range", i, box[i]);
      return EDTR_ERR_INVALID;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) {
      set_error("tensor map stride[%d] = %llu not a multiple of 16 bytes", i,
                static_cast<unsigned long long>(strides_bytes[i]));
      return EDTR_ERR_INVALID;
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)",
              static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
              static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

int prime_gemm_attributes();
int prime_attention_attributes();
int prime_gemm2_attributes();
int prime_norm_attributes();
size_t gemm2_workspace_size(int M, int N, int K);

}  // namespace edtr

extern "C" const char* edtr_last_error(void) { return edtr::g_err; }

extern "C" int edtr_version(void) { return 100; }

extern "C" int edtr_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    edtr::set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

extern "C" size_t edtr_gemm_workspace_size(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return edtr::gemm2_workspace_size(M, N, K);
}

extern "C" int edtr_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    edtr::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    edtr::set_error("edtr_b200 needs an sm_100-class GPU, found sm_%d%d", major, minor);
    return EDTR_ERR_DEVICE;
  }
  int rc = edtr::prime_gemm_attributes();
  if (rc) return rc;
  rc = edtr::prime_attention_attributes();
  if (rc) return rc;
  rc = edtr::prime_gemm2_attributes();
  if (rc) return rc;
  rc = edtr::prime_norm_attributes();
  if (rc) return rc;
  if (edtr::get_encode_fn() == nullptr) return EDTR_ERR_CUDA;
  return EDTR_OK;
}
