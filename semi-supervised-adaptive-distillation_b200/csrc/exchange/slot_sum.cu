// Local half of the copy-engine gradient exchange (sad_exchange.cc, include/sad_exchange.h): after the all-gather has filled the
// world slots of a bucket's region, out[i] = slot[0][i] + slot[1][i] + ... + slot[world-1][i], added in RANK ORDER in fp32 — the
// order the reference's single-process NCCLAllreduce test checks against (caffe2/caffe2/contrib/nccl/nccl_ops_test.py:56-79 sums the
// per-GPU inputs in GPU order), and the same bits on every rank because every rank adds the same values in the same order.
//
// Roofline: HBM, (world + 1) * 4 bytes per gradient element (world slot reads, one write).  It runs on the communication stream
// beside the backward pass, so the grid is a fraction of the GPU (kGrid CTAs): enough loads in flight for most of the HBM rate,
// few enough that the compute kernels' waves are not reshaped for long.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sad_exchange.h"

namespace {

constexpr int kThreads = 256;
constexpr int kGrid = 148 * 2;

// kAlignedOut: out is 16-byte aligned (one float4 store); otherwise the four sums leave as scalar stores — the slot loads, which are
// world / (world + 1) of the traffic, stay 16-byte vectors either way (the window's regions and strides are 512-byte multiples).
template <int kWorld, bool kAlignedOut>
__global__ void __launch_bounds__(kThreads) slot_sum_kernel(const float* __restrict__ slots, size_t stride, int world,
                                                             float* __restrict__ out, size_t count) {
  const size_t n4 = count / 4;
  const size_t step = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += step) {
    float4 acc = __ldcs(reinterpret_cast<const float4*>(slots) + i);
    if (kWorld > 0) {
#pragma unroll
      for (int r = 1; r < kWorld; ++r) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(slots + (size_t)r * stride) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    } else {
      for (int r = 1; r < world; ++r) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(slots + (size_t)r * stride) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (kAlignedOut) {
      reinterpret_cast<float4*>(out)[i] = acc;
    } else {
      out[4 * i] = acc.x; out[4 * i + 1] = acc.y; out[4 * i + 2] = acc.z; out[4 * i + 3] = acc.w;
    }
  }
  // the last count % 4 elements
  const size_t tail = n4 * 4 + (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (tail < count) {
    float acc = slots[tail];
    for (int r = 1; r < world; ++r) acc += slots[(size_t)r * stride + tail];
    out[tail] = acc;
  }
}

__global__ void __launch_bounds__(kThreads) slot_sum_scalar_kernel(const float* __restrict__ slots, size_t stride, int world,
                                                                    float* __restrict__ out, size_t count) {
  const size_t step = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += step) {
    float acc = slots[i];
    for (int r = 1; r < world; ++r) acc += slots[(size_t)r * stride + i];
    out[i] = acc;
  }
}

}  // namespace

// returns a cudaError_t (0 = launched)
extern "C" __attribute__((visibility("default"))) int sad_exchange_slot_sum_f32(const float* slots, size_t stride, int world, float* out, size_t count, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (count == 0) return 0;
  if (!slots || !out || world < 1) return (int)cudaErrorInvalidValue;
  const size_t n4 = count / 4;
  int grid = (int)((n4 + kThreads - 1) / kThreads);
  if (grid > kGrid) grid = kGrid;
  if (grid < 1) grid = 1;
  const bool vec = (reinterpret_cast<uintptr_t>(slots) & 15) == 0 && stride % 4 == 0;
  const bool aligned_out = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#define SAD_SLOT_SUM(W)                                                                                          \
  do {                                                                                                           \
    if (aligned_out) slot_sum_kernel<W, true><<<grid, kThreads, 0, stream>>>(slots, stride, world, out, count);  \
    else slot_sum_kernel<W, false><<<grid, kThreads, 0, stream>>>(slots, stride, world, out, count);             \
  } while (0)
  if (!vec) {
    int g = (int)((count + kThreads - 1) / kThreads);
    slot_sum_scalar_kernel<<<g > kGrid ? kGrid : g, kThreads, 0, stream>>>(slots, stride, world, out, count);
  } else if (world == 2) {
    SAD_SLOT_SUM(2);
  } else if (world == 4) {
    SAD_SLOT_SUM(4);
  } else if (world == 8) {
    SAD_SLOT_SUM(8);
  } else {
    SAD_SLOT_SUM(0);
  }
#undef SAD_SLOT_SUM
  return (int)cudaGetLastError();
}
