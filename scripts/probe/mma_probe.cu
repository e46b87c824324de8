// Microbenchmark: tensor-pipe cycles per tcgen05.mma kind::tf32 with operands resident in shared memory (no loads):
//   one CTA  : M = 128, N = 256 / 128, K = 8
//   CTA pair : M = 256 (cta_group::2), N = 256 / 128, K = 8
// Prints cycles per MMA and the implied TFLOP/s per SM pair / GPU at the measured clock.  Build + run on the GPU box:
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I../../semi-supervised-adaptive-distillation_b200/csrc/kernels \
//        -I../../include mma_probe.cu -o mma_probe && ./mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tc_utils.cuh"

using namespace sad;

// stride_mode 0: the same four operand slices over and over; 1: walk a 192 KB ring of distinct stages (A 16 KB + B 16/32 KB each)
template <bool kPair>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, int N, long long* cycles, int stride_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t done_bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      mbar_init(&done_bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    if (kPair) tmem_alloc_2sm<512>(&slot);
    else tmem_alloc<512>(&slot);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_tf32(kPair ? 256 : 128, N, 0, 0);
    const uint32_t a = smem_u32(smem), b = a + 16384;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      // stage bytes: one CTA 48 KB (4 stages), pair 32 KB (6 stages) as in the convolution kernel
      const uint32_t stage_bytes = kPair ? 32768u : 49152u, n_stages = kPair ? 6u : 4u;
      const uint32_t off = stride_mode ? ((uint32_t)(i >> 2) % n_stages) * stage_bytes : 0u;
      const uint64_t ad = umma_smem_desc_sw128(a + off + (i & 3) * 32, 16, 1024), bd = umma_smem_desc_sw128(b + off + (i & 3) * 32, 16, 1024);
      if (kPair) umma_tf32_2sm(tmem, ad, bd, idesc, i != 0);
      else umma_tf32(tmem, ad, bd, idesc, i != 0);
    }
    if (kPair) umma_commit_2sm(&done_bar, (uint16_t)0x1);
    else umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after_sync();
    if (kPair) tmem_dealloc_2sm<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  int dev_clock_khz = 0, sms = 0;
  cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n = 20000;
  const size_t smem = 193 * 1024;
  cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; ++mode)
  for (int pair = 0; pair < 2; ++pair)
    for (int N : {256, 128}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (!pair) {
          probe<false><<<sms, 128, smem>>>(n, N, d, mode);
        } else {
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = dim3(sms / 2 * 2);
          cfg.blockDim = dim3(128);
          cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 2;
          at[0].val.clusterDim.y = 1;
          at[0].val.clusterDim.z = 1;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, probe<true>, n, N, d, mode);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("pair=%d N=%d: %s\n", pair, N, cudaGetErrorString(e));
          return 1;
        }
      }
      long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / n;
      const double M = pair ? 256 : 128;
      const double flop_per_cycle_per_sm = 2.0 * M * N * 8 / per / (pair ? 2 : 1);
      printf("%s %s M=%3d N=%3d K=8: %.1f cycles/MMA  -> %.0f flop/cycle/SM -> %.0f TFLOP/s on %d SMs at %.0f MHz (all SMs issuing)\n",
             mode ? "[distinct stages]" : "[same operands]  ", pair ? "CTA pair" : "one CTA ", (int)M, N, per, flop_per_cycle_per_sm, flop_per_cycle_per_sm * sms * dev_clock_khz * 1e3 / 1e12, sms,
             dev_clock_khz / 1e3);
    }
  return 0;
}
