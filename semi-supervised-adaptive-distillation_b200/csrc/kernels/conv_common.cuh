// Host-side helpers shared by the tensor-core convolution translation units: tensor-map encoding
// through the driver entry point (no link-time dependency on libcuda) and device queries.
#ifndef SAD_CONV_COMMON_CUH_
#define SAD_CONV_COMMON_CUH_

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// fp32 tensor map with 128-byte swizzle; out-of-bounds elements read as zero
inline int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const char* what, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
                      CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(SAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(SAD_ERR_CUDA, std::string("cuTensorMapEncodeTiled(") + what + ") failed: CUresult " + std::to_string((int)r));
  return SAD_OK;
}

// channels-last activations (N, H, W, C): dims {C, W, H, N}, box {32 channels, box_x, box_y, 1}
inline int encode_nhwc_map(CUtensorMap* m, const float* xt, int N, int C, int H, int W, int box_x, int box_y, const char* what,
                           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  return encode_map(m, xt, 4, dims, str, box, what, swizzle);
}
// the same for fp16 activations: box {64 channels, box_x, box_y, 1} = the same 128-byte rows
inline int encode_nhwc_map_f16(CUtensorMap* m, const void* xt, int N, int C, int H, int W, int box_x, int box_y, const char* what) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  return encode_map(m, xt, 4, dims, str, box, what, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}

inline int sm_count(int* sms) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc != SAD_OK) return rc;
  return check_cuda(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
}

}  // namespace sad
#endif
