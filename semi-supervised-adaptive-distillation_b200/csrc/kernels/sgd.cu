// MomentumSGDUpdate with Detectron's per-parameter preamble, one launch for a whole flat parameter buffer
// (SURVEY.md §8f rank 4).  Replaces, per parameter blob of detectron/lib/modeling/optimizer.py:95-130,
//   Scale(param_grad, 2.0)                                  for biases            (optimizer.py:115-121)
//   WeightedSum([param_grad, one, param, wd], param_grad)   for weights           (optimizer.py:122-124)
//   MomentumSGDUpdate([grad, momentum, lr, param])          caffe2/caffe2/sgd/momentum_sgd_op_gpu.cu:23-54
// i.e. ~3 launches per blob and 60 for the head alone.  The flat buffer is described by up to SAD_MAX_SGD_SEGMENTS
// ranges, each with its gradient multiplier (2 for biases) and weight decay (0 for biases).  HBM-bound: reads g, m, p
// and writes g, m, p = 24 B/parameter (the reference's three ops move 8 + 12..16 + 24 B).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

struct SgdArgs {
  float* param;
  float* grad;
  float* mom;
  const float* lr;
  int64_t seg_end[SAD_MAX_SGD_SEGMENTS];
  float grad_mult[SAD_MAX_SGD_SEGMENTS];
  float weight_decay[SAD_MAX_SGD_SEGMENTS];
  int32_t n_segs, nesterov;
  int64_t total;
  float momentum;
  const uint32_t* skip_if_nonzero;   // NULL, or a device word: the launch leaves every buffer untouched when it is not 0
};

__device__ __forceinline__ void sgd_elem(float& p, float& g, float& m, float LR, float momentum, float mult, float wd, bool nesterov) {
  const float gi = mult * g + wd * p;                      // Scale(2.0) for biases / WeightedSum(grad, 1, param, wd) for weights
  if (!nesterov) {
    const float adjusted = LR * gi + momentum * m;         // momentum_sgd_op_gpu.cu:35-41
    m = adjusted;
    g = adjusted;
    p -= adjusted;
  } else {
    const float mi_new = momentum * m + LR * gi;           // :44-51
    g = (1.f + momentum) * mi_new - momentum * m;
    m = mi_new;
    p -= g;
  }
}

__global__ void __launch_bounds__(256) momentum_sgd_kernel(const SgdArgs a) {
  if (a.skip_if_nonzero && __ldg(a.skip_if_nonzero) != 0u) return;   // a gradient overflowed (mixed fp16): skip the step
  const float LR = __ldg(a.lr);
  const int64_t n4 = a.total >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const bool nest = a.nesterov != 0;
  for (int64_t q = tid; q < n4; q += stride) {
    const int64_t i = q << 2;
    float4 p = reinterpret_cast<float4*>(a.param)[q], g = reinterpret_cast<float4*>(a.grad)[q], m = reinterpret_cast<float4*>(a.mom)[q];
    int s = 0;
    while (s + 1 < a.n_segs && i >= a.seg_end[s]) ++s;
    if (i + 3 < a.seg_end[s]) {   // the whole quad lies in one segment (the common case)
      const float mult = a.grad_mult[s], wd = a.weight_decay[s];
      sgd_elem(p.x, g.x, m.x, LR, a.momentum, mult, wd, nest);
      sgd_elem(p.y, g.y, m.y, LR, a.momentum, mult, wd, nest);
      sgd_elem(p.z, g.z, m.z, LR, a.momentum, mult, wd, nest);
      sgd_elem(p.w, g.w, m.w, LR, a.momentum, mult, wd, nest);
    } else {
      float* pp = &p.x;
      float* gp = &g.x;
      float* mp = &m.x;
      for (int k = 0; k < 4; ++k) {
        int sk = s;
        while (sk + 1 < a.n_segs && i + k >= a.seg_end[sk]) ++sk;
        sgd_elem(pp[k], gp[k], mp[k], LR, a.momentum, a.grad_mult[sk], a.weight_decay[sk], nest);
      }
    }
    reinterpret_cast<float4*>(a.param)[q] = p;
    reinterpret_cast<float4*>(a.grad)[q] = g;
    reinterpret_cast<float4*>(a.mom)[q] = m;
  }
  for (int64_t i = (n4 << 2) + tid; i < a.total; i += stride) {
    int s = 0;
    while (s + 1 < a.n_segs && i >= a.seg_end[s]) ++s;
    sgd_elem(a.param[i], a.grad[i], a.mom[i], LR, a.momentum, a.grad_mult[s], a.weight_decay[s], nest);
  }
}

// ---------------------------------------------------------------------------------------------
// The two optimiser operators as the reference graph names them, one blob per call (operator classes in csrc/ops/sgd_ops.cc).
// The element expressions are written exactly as in the reference sources so that nvcc contracts them into the same FMAs
// (momentum_sgd_op_gpu.cu:35-51; math_gpu.cu ScaleKernelDeviceAlpha / AxpyKernel behind WeightedSumOp, utility_ops.h:333-378):
// results are bit-identical to the reference operators (tests/test_sgd_gpu.py runs both).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) momentum_sgd_update_kernel(const int64_t N, const float* __restrict__ g, const float* __restrict__ m,
                                                                  float* ng, float* nm, const float* __restrict__ lr, const float momentum,
                                                                  const bool nesterov, const float* param_in, float* param) {
  const float LR = lr[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    if (!nesterov) {
      const float adjusted_gradient = LR * g[i] + momentum * m[i];
      nm[i] = adjusted_gradient;
      ng[i] = adjusted_gradient;
      if (param) param[i] = param_in[i] - adjusted_gradient;
    } else {
      const float mi = m[i];
      const float mi_new = momentum * mi + LR * g[i];
      nm[i] = mi_new;
      ng[i] = (1 + momentum) * mi_new - momentum * mi;
      if (param) param[i] = param_in[i] - ng[i];
    }
  }
}

struct WsumArgs {
  const float* x[SAD_MAX_INPUTS];
  const float* w[SAD_MAX_INPUTS];
  int32_t n_inputs;
};
__global__ void __launch_bounds__(256) weighted_sum_kernel(const WsumArgs a, float* out, const int64_t N) {
  float w[SAD_MAX_INPUTS];
#pragma unroll
  for (int k = 0; k < SAD_MAX_INPUTS; ++k) w[k] = k < a.n_inputs ? __ldg(a.w[k]) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float y = a.x[0][i] * w[0];                                   // math::Scale (ScaleKernelDeviceAlpha): y = x * (*alpha)
#pragma unroll
    for (int k = 1; k < SAD_MAX_INPUTS; ++k)
      if (k < a.n_inputs) y = fmaf(a.x[k][i], w[k], y);           // math::Axpy (AxpyKernel): y += x * (*a), one FMA as nvcc contracts it
    out[i] = y;
  }
}

// ---- *flag |= (any element of x is inf or NaN): the overflow test of mixed-precision training --------------------------------
__global__ void __launch_bounds__(256) nonfinite_flag_kernel(const float* __restrict__ x, int64_t n, uint32_t* flag) {
  const int64_t n4 = n >> 2, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t q = tid; q < n4; q += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + q);
    // exponent all ones <=> inf or NaN
    bad |= ((__float_as_uint(v.x) & 0x7f800000u) == 0x7f800000u) | ((__float_as_uint(v.y) & 0x7f800000u) == 0x7f800000u) |
           ((__float_as_uint(v.z) & 0x7f800000u) == 0x7f800000u) | ((__float_as_uint(v.w) & 0x7f800000u) == 0x7f800000u);
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) bad |= (__float_as_uint(x[i]) & 0x7f800000u) == 0x7f800000u;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

}  // namespace sad

using namespace sad;

extern "C" {

SAD_EXPORT int sad_nonfinite_flag_f32(const float* x, int64_t n, uint32_t* flag, void* stream) {
  if (n < 0 || !flag) return set_error(SAD_ERR_INVALID, "nonfinite flag: bad argument");
  if (n == 0) return SAD_OK;
  if (!x || (reinterpret_cast<uintptr_t>(x) & 15)) return set_error(SAD_ERR_INVALID, "nonfinite flag: x must be a 16-byte aligned device buffer");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n / 4 + 255) / 256 + 1, cap = (int64_t)sms * 8;
  nonfinite_flag_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, flag);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "nonfinite flag launch");
}

SAD_EXPORT int sad_momentum_sgd_update_f32(const float* grad, const float* mom, const float* lr, const float* param, float* grad_out,
                                           float* mom_out, float* param_out, int64_t n, float momentum, int nesterov, void* stream) {
  if (n < 0) return set_error(SAD_ERR_INVALID, "MomentumSGDUpdate: negative size");
  if (n == 0) return SAD_OK;
  if (!grad || !mom || !lr || !grad_out || !mom_out || (!param != !param_out))
    return set_error(SAD_ERR_INVALID, "MomentumSGDUpdate: null tensor (param and param_out come together, or both NULL for MomentumSGD)");
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + 255) / 256, cap = (int64_t)sms * 8;
  momentum_sgd_update_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, grad, mom, grad_out, mom_out, lr, momentum, nesterov != 0, param, param_out);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "MomentumSGDUpdate launch");
}

SAD_EXPORT int sad_weighted_sum_f32(const float* const* xs, const float* const* ws, int n_inputs, float* out, int64_t n, void* stream) {
  if (!xs || !ws || n_inputs < 1 || n_inputs > SAD_MAX_INPUTS) return set_error(SAD_ERR_INVALID, "WeightedSum: 1 .. SAD_MAX_INPUTS (tensor, weight) pairs");
  if (n < 0) return set_error(SAD_ERR_INVALID, "WeightedSum: negative size");
  if (n == 0) return SAD_OK;
  if (!out) return set_error(SAD_ERR_INVALID, "WeightedSum: null output");
  WsumArgs a{};
  for (int k = 0; k < n_inputs; ++k) {
    if (!xs[k] || !ws[k]) return set_error(SAD_ERR_INVALID, "WeightedSum: null input");
    if (k > 0 && xs[k] == out) return set_error(SAD_ERR_INVALID, "WeightedSum: in-place only with input 0 (utility_ops.h:357-364)");
    a.x[k] = xs[k];
    a.w[k] = ws[k];
  }
  a.n_inputs = n_inputs;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + 255) / 256, cap = (int64_t)sms * 8;
  weighted_sum_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, out, n);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "WeightedSum launch");
}

static int momentum_sgd_impl(float* param, float* grad, float* momentum_buf, const sad_sgd_segment* segments, int n_segments,
                             const float* lr, float momentum, int nesterov, const uint32_t* skip_if_nonzero, void* stream) {
  if (!segments || n_segments < 1 || n_segments > SAD_MAX_SGD_SEGMENTS)
    return set_error(SAD_ERR_INVALID, "momentum sgd: n_segments must be in [1, SAD_MAX_SGD_SEGMENTS]");
  if (!lr) return set_error(SAD_ERR_INVALID, "momentum sgd: null learning rate");
  SgdArgs a{};
  int64_t end = 0;
  for (int s = 0; s < n_segments; ++s) {
    if (segments[s].count < 0) return set_error(SAD_ERR_INVALID, "momentum sgd: negative segment length");
    end += segments[s].count;
    a.seg_end[s] = end;
    a.grad_mult[s] = segments[s].grad_multiplier;
    a.weight_decay[s] = segments[s].weight_decay;
  }
  if (end == 0) return SAD_OK;
  if (!param || !grad || !momentum_buf) return set_error(SAD_ERR_INVALID, "momentum sgd: null buffer");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(momentum_buf)) & 15)
    return set_error(SAD_ERR_INVALID, "momentum sgd: buffers must be 16-byte aligned");
  a.param = param;
  a.grad = grad;
  a.mom = momentum_buf;
  a.lr = lr;
  a.n_segs = n_segments;
  a.nesterov = nesterov;
  a.total = end;
  a.momentum = momentum;
  a.skip_if_nonzero = skip_if_nonzero;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (end / 4 + 255) / 256;
  const int64_t cap = (int64_t)sms * 8;
  const unsigned blocks = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
  momentum_sgd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "momentum sgd launch");
}

SAD_EXPORT int sad_momentum_sgd_f32(float* param, float* grad, float* momentum_buf, const sad_sgd_segment* segments, int n_segments,
                                    const float* lr, float momentum, int nesterov, void* stream) {
  return momentum_sgd_impl(param, grad, momentum_buf, segments, n_segments, lr, momentum, nesterov, nullptr, stream);
}
SAD_EXPORT int sad_momentum_sgd_guarded_f32(float* param, float* grad, float* momentum_buf, const sad_sgd_segment* segments, int n_segments,
                                            const float* lr, float momentum, int nesterov, const uint32_t* skip_if_nonzero, void* stream) {
  return momentum_sgd_impl(param, grad, momentum_buf, segments, n_segments, lr, momentum, nesterov, skip_if_nonzero, stream);
}

}  // extern "C"
