// Bring-up probe: one TMA tensor load of a given rank/variant into shared memory, copied back and checked.
// usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kRank>
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int nfloats, int c0, int c1, int c2, int c3, int c4) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(nfloats * 4));
    if (kRank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(smem)), "l"((uint64_t)&map), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    if (kRank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(s32(smem)), "l"((uint64_t)&map), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    if (kRank == 5)
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(s32(smem)), "l"((uint64_t)&map), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = 2, C = 64, H = 16, W = 32;
  std::vector<float> h((size_t)N * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 65536 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap m;
  cuuint64_t dims[5], str[4];
  cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
  int rank = 4, c[5] = {0, 0, 0, 0, 0};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  switch (variant) {
    case 0:  // 4D natural {W,H,C,N}, box {32,1,32,1}
      dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = N; str[0] = W * 4; str[1] = H * W * 4; str[2] = (cuuint64_t)C * H * W * 4;
      box[0] = 32; box[1] = 1; box[2] = 32; box[3] = 1; c[1] = 3; c[2] = 32; c[3] = 1; break;
    case 1:  // 4D permuted {W,C,H,N}, box {32,32,8,1}
      dims[0] = W; dims[1] = C; dims[2] = H; dims[3] = N; str[0] = H * W * 4; str[1] = W * 4; str[2] = (cuuint64_t)C * H * W * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; box[3] = 1; c[1] = 32; c[2] = 3; c[3] = 1; break;
    case 2:  // 3D merged {W,H,C*N} natural, box {32,1,32}
      rank = 3; dims[0] = W; dims[1] = H; dims[2] = C * N; str[0] = W * 4; str[1] = H * W * 4;
      box[0] = 32; box[1] = 1; box[2] = 32; c[1] = 3; c[2] = 64 + 32; break;
    case 3:  // 3D permuted merged {W, C*N, H}, box {32,32,8}
      rank = 3; dims[0] = W; dims[1] = C * N; dims[2] = H; str[0] = H * W * 4; str[1] = W * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; c[1] = 64 + 32; c[2] = 3; break;
    case 4:  // 4D natural, no swizzle
      sw = CU_TENSOR_MAP_SWIZZLE_NONE;
      dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = N; str[0] = W * 4; str[1] = H * W * 4; str[2] = (cuuint64_t)C * H * W * 4;
      box[0] = 32; box[1] = 1; box[2] = 32; box[3] = 1; c[1] = 3; c[2] = 32; c[3] = 1; break;
    case 5:  // 4D natural, small box {32,2,4,1}
      dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = N; str[0] = W * 4; str[1] = H * W * 4; str[2] = (cuuint64_t)C * H * W * 4;
      box[0] = 32; box[1] = 2; box[2] = 4; box[3] = 1; c[1] = 3; c[2] = 32; c[3] = 1; break;
    case 6:  // 3D permuted merged with negative x and y coordinates
      rank = 3; dims[0] = W; dims[1] = C * N; dims[2] = H; str[0] = H * W * 4; str[1] = W * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; c[0] = -1; c[1] = 64 + 32; c[2] = -1; break;
    case 8:  // 3D permuted merged, negative y only
      rank = 3; dims[0] = W; dims[1] = C * N; dims[2] = H; str[0] = H * W * 4; str[1] = W * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; c[0] = 0; c[1] = 64 + 32; c[2] = -1; break;
    case 9:  // inner coordinate +4 elements (16 B aligned)
      rank = 3; dims[0] = W; dims[1] = C * N; dims[2] = H; str[0] = H * W * 4; str[1] = W * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; c[0] = 4; c[1] = 64 + 32; c[2] = 3; break;
    case 10:  // inner coordinate +1 element (unaligned)
      rank = 3; dims[0] = W; dims[1] = C * N; dims[2] = H; str[0] = H * W * 4; str[1] = W * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; c[0] = 1; c[1] = 64 + 32; c[2] = 3; break;
    case 11:  // NHWC view of the same buffer {C,W,H,N}, negative x and y, box {32,32,8,1}
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; str[0] = C * 4; str[1] = (cuuint64_t)W * C * 4; str[2] = (cuuint64_t)H * W * C * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; box[3] = 1; c[0] = 32; c[1] = -1; c[2] = -1; c[3] = 1; break;
    case 12:  // NHWC view, box beyond the high edge
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; str[0] = C * 4; str[1] = (cuuint64_t)W * C * 4; str[2] = (cuuint64_t)H * W * C * 4;
      box[0] = 32; box[1] = 32; box[2] = 8; box[3] = 1; c[0] = 32; c[1] = 1; c[2] = 9; c[3] = 1; break;
    case 7:  // 5D {W,H,C,N,1}
      rank = 5; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = N; dims[4] = 1; str[0] = W * 4; str[1] = H * W * 4; str[2] = (cuuint64_t)C * H * W * 4; str[3] = (cuuint64_t)N * C * H * W * 4;
      box[0] = 32; box[1] = 1; box[2] = 32; box[3] = 1; box[4] = 1; c[1] = 3; c[2] = 32; c[3] = 1; break;
  }
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  int nfl = 1;
  for (int i = 0; i < rank; ++i) nfl *= box[i];
  printf("variant %d rank %d encode=%d box floats %d\n", variant, rank, (int)r, nfl);
  if (r != CUDA_SUCCESS) return 0;
  size_t smem = (size_t)nfl * 4 + 1024;
  if (rank == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<3><<<1, 128, smem>>>(m, out, nfl, c[0], c[1], c[2], c[3], c[4]); }
  if (rank == 4) { cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<4><<<1, 128, smem>>>(m, out, nfl, c[0], c[1], c[2], c[3], c[4]); }
  if (rank == 5) { cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<5><<<1, 128, smem>>>(m, out, nfl, c[0], c[1], c[2], c[3], c[4]); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("  sync: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 0;
  std::vector<float> o(nfl);
  cudaMemcpy(o.data(), out, (size_t)nfl * 4, cudaMemcpyDeviceToHost);
  printf("  first row (unswizzle not applied): ");
  for (int i = 0; i < 8; ++i) printf("%.0f ", o[i]);
  printf(" | row1: ");
  for (int i = 32; i < 40; ++i) printf("%.0f ", o[i]);
  printf("\n");
  return 0;
}
