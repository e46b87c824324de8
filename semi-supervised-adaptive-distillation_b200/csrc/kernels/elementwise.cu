// Relu / ReluGradient as stand-alone operators (caffe2/caffe2/operators/relu_op.cu:22-62): the tower runs them in
// place on the convolution output (retinanet_heads.py:124,209).  HBM-bound, 8 B/element (Relu) and 12 B/element
// (ReluGradient); 128-bit accesses, grid-stride over a multiple of the SM count.  The fused head path (head.cu)
// never launches these: there ReLU lives in the convolution epilogue and ReluGradient in the data-gradient epilogue.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (size_t i = tid; i < n4; i += stride) {
      float4 v = x4[i];
      v.x = v.x > 0.f ? v.x : 0.f;   // X > 0 ? X : 0 (relu_op.cu:24-26): NaN -> 0 like the reference
      v.y = v.y > 0.f ? v.y : 0.f;
      v.z = v.z > 0.f ? v.z : 0.f;
      v.w = v.w > 0.f ? v.w : 0.f;
      y4[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) y[i] = x[i] > 0.f ? x[i] : 0.f;
  } else {
    for (size_t i = tid; i < n; i += stride) y[i] = x[i] > 0.f ? x[i] : 0.f;
  }
}

__global__ void __launch_bounds__(256) relu_grad_kernel(const float* __restrict__ yv, const float* __restrict__ dy, float* __restrict__ dx,
                                                        size_t n, int vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4* y4 = reinterpret_cast<const float4*>(yv);
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    float4* o4 = reinterpret_cast<float4*>(dx);
    for (size_t i = tid; i < n4; i += stride) {
      const float4 a = y4[i], g = d4[i];
      o4[i] = make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);  // relu_op.cu:32-34
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) dx[i] = yv[i] > 0.f ? dy[i] : 0.f;
  } else {
    for (size_t i = tid; i < n; i += stride) dx[i] = yv[i] > 0.f ? dy[i] : 0.f;
  }
}

// Sigmoid as a stand-alone operator (caffe2/caffe2/operators/sigmoid_op.cu:24-29: Y = 1 / (1 + exp(-X))): what the teacher's
// graph appends to retnet_cls_pred_fpnL when model.train is False (retinanet_heads.py:153-163).  8 B/element, HBM-bound.
// The fused head object never launches it: there the Sigmoid is the prediction convolution's epilogue.
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }
__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (size_t i = tid; i < n4; i += stride) {
      const float4 v = x4[i];
      y4[i] = make_float4(sigmoid_ref(v.x), sigmoid_ref(v.y), sigmoid_ref(v.z), sigmoid_ref(v.w));
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) y[i] = sigmoid_ref(x[i]);
  } else {
    for (size_t i = tid; i < n; i += stride) y[i] = sigmoid_ref(x[i]);
  }
}

// Scale (caffe2/caffe2/operators/scale_op.h:31-50 -> math::Scale, caffe2/caffe2/utils/math_gpu.cu:1264-1302: y = x * alpha): what
// _CorrectMomentum runs over every `<param>_momentum` blob when the learning rate changes (detectron/lib/modeling/detector.py:
// 628-648); over the flat momentum buffer it is one launch instead of one per parameter.  8 B/element, HBM-bound.
__global__ void __launch_bounds__(256) scale_kernel(const float* x, float* y, float alpha, size_t n, int vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (size_t i = tid; i < n4; i += stride) {
      const float4 v = x4[i];
      y4[i] = make_float4(v.x * alpha, v.y * alpha, v.z * alpha, v.w * alpha);
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) y[i] = x[i] * alpha;
  } else {
    for (size_t i = tid; i < n; i += stride) y[i] = x[i] * alpha;
  }
}

static unsigned ew_grid(size_t n) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t want = (n / 4 + 255) / 256;
  const size_t cap = (size_t)sms * 8;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace sad

using namespace sad;

extern "C" {

SAD_EXPORT int sad_relu_f32(const float* x, float* y, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!x || !y))) return set_error(SAD_ERR_INVALID, "relu: bad argument");
  if (n == 0) return SAD_OK;
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  relu_kernel<<<ew_grid((size_t)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, (size_t)n, vec);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "relu launch");
}

SAD_EXPORT int sad_sigmoid_f32(const float* x, float* y, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!x || !y))) return set_error(SAD_ERR_INVALID, "sigmoid: bad argument");
  if (n == 0) return SAD_OK;
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  sigmoid_kernel<<<ew_grid((size_t)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, (size_t)n, vec);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "sigmoid launch");
}

SAD_EXPORT int sad_scale_f32(const float* x, float* y, int64_t n, float alpha, void* stream) {
  if (n < 0 || (n > 0 && (!x || !y))) return set_error(SAD_ERR_INVALID, "scale: bad argument");
  if (n == 0) return SAD_OK;
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  scale_kernel<<<ew_grid((size_t)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, alpha, (size_t)n, vec);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "scale launch");
}

SAD_EXPORT int sad_relu_grad_f32(const float* y, const float* dy, float* dx, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!y || !dy || !dx))) return set_error(SAD_ERR_INVALID, "relu gradient: bad argument");
  if (n == 0) return SAD_OK;
  const int vec = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  relu_grad_kernel<<<ew_grid((size_t)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, dy, dx, (size_t)n, vec);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "relu gradient launch");
}

}  // extern "C"
