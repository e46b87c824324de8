// AffineChannel (+Gradient) and UpsampleNearest (+Gradient): the two Detectron operators of the ResNet / FPN body that sit
// between the convolutions (caffe2/modules/detectron/affine_channel_op.cu:22-98,
// caffe2/modules/detectron/upsample_nearest_op.cu:62-217; callers detectron/lib/modeling/ResNet.py:219-278 and
// FPN.py:230-249).  All four are HBM-bound streams:
//   AffineChannel           y = x * scale[c] + bias[c]        8 B/element  (one FMA, as nvcc contracts the reference's expression)
//   AffineChannelGradient   dx = dy * scale[c]                8 B/element
//   UpsampleNearest         y[.., Y, X] = x[.., Y/s, X/s]     4 + 4 s^2 B per input element (20 B at s = 2)
//   UpsampleNearestGradient dx = sum of the s x s block of dy, added in the reference's order (x offset outer, y offset
//                           inner, starting from the 0 the reference's math::Set leaves), so the result is bit-identical
// 128-bit accesses where the row length allows (W % 4 == 0, 16-byte aligned pointers), one 32-bit division per float4,
// grid-stride over a multiple of the SM count.  Index arithmetic is 32-bit like the reference's (int index / int ii);
// the entry points refuse tensors of 2^31 elements or more instead of wrapping.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

// ---- AffineChannel ------------------------------------------------------------------------------------------------
// x viewed as [N*C rows][HW]; channel of a row = row % C.  kBias: forward (fma with bias) or gradient (plain product).
template <bool kBias>
__global__ void __launch_bounds__(256) affine_channel_vec4_kernel(const float4* x, const float* __restrict__ scale,
                                                                  const float* __restrict__ bias, float4* y, uint32_t n4,
                                                                  uint32_t hw4, uint32_t C) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint32_t c = (i / hw4) % C;
    const float s = __ldg(scale + c);
    float4 v = x[i];
    if (kBias) {
      const float b = __ldg(bias + c);
      v.x = fmaf(v.x, s, b); v.y = fmaf(v.y, s, b); v.z = fmaf(v.z, s, b); v.w = fmaf(v.w, s, b);
    } else {
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    }
    y[i] = v;
  }
}

template <bool kBias>
__global__ void __launch_bounds__(256) affine_channel_scalar_kernel(const float* x, const float* __restrict__ scale,
                                                                    const float* __restrict__ bias, float* y, uint32_t n,
                                                                    uint32_t hw, uint32_t C) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t c = (i / hw) % C;   // affine_channel_op.cu:32 / :46
    y[i] = kBias ? fmaf(x[i], __ldg(scale + c), __ldg(bias + c)) : x[i] * __ldg(scale + c);
  }
}

// ---- UpsampleNearest, scale 2, rows of W % 4 == 0 -------------------------------------------------------------------
// One thread per INPUT float4 (row r = outer*H + y, x4): writes the two output rows 2r and 2r+1, two float4 each.
__global__ void __launch_bounds__(256) upsample2_vec4_kernel(const float4* __restrict__ x, float4* __restrict__ y, uint32_t n4,
                                                             uint32_t w4) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint32_t r = i / w4, c4 = i - r * w4;
    const float4 v = x[i];
    const float4 lo = make_float4(v.x, v.x, v.y, v.y), hi = make_float4(v.z, v.z, v.w, v.w);
    const size_t o = (size_t)(2 * r) * (2 * w4) + 2 * c4;
    __stcs(y + o, lo);
    __stcs(y + o + 1, hi);
    __stcs(y + o + 2 * w4, lo);
    __stcs(y + o + 2 * w4 + 1, hi);
  }
}

// One thread per dX float4: reads the 2 x 8 block of dY.  Order of the four additions per element as in downscale()
// (upsample_nearest_op.cu:102-113): x offset i outer, y offset j inner, accumulator starting at 0.
__global__ void __launch_bounds__(256) upsample2_grad_vec4_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, uint32_t n4,
                                                                  uint32_t w4) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint32_t r = i / w4, c4 = i - r * w4;
    const size_t o = (size_t)(2 * r) * (2 * w4) + 2 * c4;
    const float4 a0 = __ldcs(dy + o), a1 = __ldcs(dy + o + 1);                      // y offset 0: x 0..7
    const float4 b0 = __ldcs(dy + o + 2 * w4), b1 = __ldcs(dy + o + 2 * w4 + 1);    // y offset 1
    float4 g;
    g.x = (((0.f + a0.x) + b0.x) + a0.y) + b0.y;
    g.y = (((0.f + a0.z) + b0.z) + a0.w) + b0.w;
    g.z = (((0.f + a1.x) + b1.x) + a1.y) + b1.y;
    g.w = (((0.f + a1.z) + b1.z) + a1.w) + b1.w;
    dx[i] = g;
  }
}

// ---- UpsampleNearest, any integer scale / row length ---------------------------------------------------------------
// per OUTPUT element, the reference's translate_idx (upsample_nearest_op.cu:66-80) with (d1 folded into outer, d2, d3)
__global__ void __launch_bounds__(256) upsample_generic_kernel(const float* __restrict__ x, float* __restrict__ y, uint32_t n_out,
                                                               uint32_t Ho, uint32_t Wo, uint32_t s) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x; ii < n_out; ii += stride) {
    const uint32_t w = ii % Wo, t = ii / Wo, z = t % Ho, o = t / Ho;
    y[ii] = x[((size_t)o * (Ho / s) + z / s) * (Wo / s) + w / s];
  }
}

// per INPUT-GRADIENT element, translate_idx_inv (:82-100) and the i / j loop of downscale (:102-113)
__global__ void __launch_bounds__(256) upsample_grad_generic_kernel(const float* __restrict__ dy, float* __restrict__ dx, uint32_t n_in,
                                                                    uint32_t H, uint32_t W, uint32_t s) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t ii = blockIdx.x * blockDim.x + threadIdx.x; ii < n_in; ii += stride) {
    const uint32_t w = ii % W, t = ii / W, z = t % H, o = t / H;
    float acc = 0.f;
    for (uint32_t i = 0; i < s; ++i)
      for (uint32_t j = 0; j < s; ++j) acc += dy[((size_t)o * (H * s) + z * s + j) * (W * s) + w * s + i];
    dx[ii] = acc;
  }
}

static unsigned stream_grid(size_t items) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t want = (items + 255) / 256, cap = (size_t)sms * 8;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

static bool aligned16(const void* a, const void* b) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

// ---- FPN top-down merge: UpsampleNearest(top, 2) + Sum with the lateral (FPN.py:230-249) in one pass ------------------------
// out[o][Y][X][c] = lateral[o][Y][X][c] + top[o][Y/2][X/2][c] over a tensor viewed as (outer, H_out, W_out, inner):
// inner = 1 is NCHW (outer = N*C planes), inner = C is channels-last (outer = N).  9 B per output element (4 lateral + 1 top +
// 4 out) against 17 for the two operators (UpsampleNearest writes 4, Sum reads 4 + 4 and writes 4).  One thread per output
// float4 along the innermost axis that allows it; exact (a single fp32 add, like math::Add behind SumOp).
template <bool kVecInner>
__global__ void __launch_bounds__(256) upsample2_add_kernel(const float* __restrict__ top, const float* __restrict__ lateral,
                                                            float* __restrict__ out, uint32_t n_vec, uint32_t Wo, uint32_t Ho,
                                                            uint32_t inner) {
  const uint32_t stride = gridDim.x * blockDim.x;
  if (kVecInner) {   // inner % 4 == 0: a float4 never crosses a pixel
    const uint32_t inner4 = inner >> 2, Wi = Wo >> 1, Hi = Ho >> 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint32_t c4 = i % inner4;
      uint32_t r = i / inner4;
      const uint32_t X = r % Wo;
      r /= Wo;
      const uint32_t Y = r % Ho, o = r / Ho;
      const float4 a = reinterpret_cast<const float4*>(lateral)[i];
      const float4 b = __ldg(reinterpret_cast<const float4*>(top) + ((size_t)(o * Hi + (Y >> 1)) * Wi + (X >> 1)) * inner4 + c4);
      reinterpret_cast<float4*>(out)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  } else {           // inner == 1, W_out % 4 == 0: a float4 of the output row reads two elements of the top row
    const uint32_t Wo4 = Wo >> 2, Wi = Wo >> 1, Hi = Ho >> 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint32_t x4 = i % Wo4;
      uint32_t r = i / Wo4;
      const uint32_t Y = r % Ho, o = r / Ho;
      const float4 a = reinterpret_cast<const float4*>(lateral)[i];
      const float2 b = __ldg(reinterpret_cast<const float2*>(top + ((size_t)o * Hi + (Y >> 1)) * Wi) + x4);
      reinterpret_cast<float4*>(out)[i] = make_float4(a.x + b.x, a.y + b.x, a.z + b.y, a.w + b.y);
    }
  }
}
__global__ void __launch_bounds__(256) upsample2_add_scalar_kernel(const float* __restrict__ top, const float* __restrict__ lateral,
                                                                   float* __restrict__ out, uint32_t n, uint32_t Wo, uint32_t Ho,
                                                                   uint32_t inner) {
  const uint32_t stride = gridDim.x * blockDim.x, Wi = Wo >> 1, Hi = Ho >> 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t c = i % inner;
    uint32_t r = i / inner;
    const uint32_t X = r % Wo;
    r /= Wo;
    const uint32_t Y = r % Ho, o = r / Ho;
    out[i] = lateral[i] + __ldg(top + ((size_t)(o * Hi + (Y >> 1)) * Wi + (X >> 1)) * inner + c);
  }
}

}  // namespace sad


using namespace sad;

extern "C" {

SAD_EXPORT int sad_affine_channel_f32(const float* x, const float* scale, const float* bias, float* y, int N, int C, int64_t HW,
                                      void* stream) {
  if (N < 0 || C <= 0 || HW < 0) return set_error(SAD_ERR_INVALID, "affine channel: bad shape");
  const int64_t n = (int64_t)N * C * HW;
  if (n == 0) return SAD_OK;
  if (!x || !scale || !y) return set_error(SAD_ERR_INVALID, "affine channel: null pointer");
  if (n >= ((int64_t)1 << 31)) return set_error(SAD_ERR_UNSUPPORTED, "affine channel: 2^31 elements or more (the reference indexes with int)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (HW % 4 == 0 && aligned16(x, y)) {
    const uint32_t n4 = (uint32_t)(n / 4), hw4 = (uint32_t)(HW / 4);
    if (bias)
      affine_channel_vec4_kernel<true><<<stream_grid(n4), 256, 0, st>>>(reinterpret_cast<const float4*>(x), scale, bias,
                                                                       reinterpret_cast<float4*>(y), n4, hw4, (uint32_t)C);
    else
      affine_channel_vec4_kernel<false><<<stream_grid(n4), 256, 0, st>>>(reinterpret_cast<const float4*>(x), scale, nullptr,
                                                                        reinterpret_cast<float4*>(y), n4, hw4, (uint32_t)C);
  } else {
    if (bias)
      affine_channel_scalar_kernel<true><<<stream_grid((size_t)n), 256, 0, st>>>(x, scale, bias, y, (uint32_t)n, (uint32_t)HW, (uint32_t)C);
    else
      affine_channel_scalar_kernel<false><<<stream_grid((size_t)n), 256, 0, st>>>(x, scale, nullptr, y, (uint32_t)n, (uint32_t)HW, (uint32_t)C);
  }
  count_launch(1);
  return check_cuda(cudaGetLastError(), "affine channel launch");
}

SAD_EXPORT int sad_upsample_nearest_f32(const float* x, float* y, int64_t outer, int H, int W, int scale, void* stream) {
  if (outer < 0 || H < 0 || W < 0 || scale < 1) return set_error(SAD_ERR_INVALID, "upsample nearest: bad shape or scale");
  const int64_t n_in = outer * H * W, n_out = n_in * scale * scale;
  if (n_in == 0) return SAD_OK;
  if (!x || !y) return set_error(SAD_ERR_INVALID, "upsample nearest: null pointer");
  if (n_out >= ((int64_t)1 << 31)) return set_error(SAD_ERR_UNSUPPORTED, "upsample nearest: 2^31 output elements or more");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (scale == 2 && W % 4 == 0 && aligned16(x, y))
    upsample2_vec4_kernel<<<stream_grid((size_t)n_in / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y),
                                                                         (uint32_t)(n_in / 4), (uint32_t)(W / 4));
  else
    upsample_generic_kernel<<<stream_grid((size_t)n_out), 256, 0, st>>>(x, y, (uint32_t)n_out, (uint32_t)(H * scale), (uint32_t)(W * scale),
                                                                        (uint32_t)scale);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "upsample nearest launch");
}

SAD_EXPORT int sad_upsample_nearest_grad_f32(const float* dy, float* dx, int64_t outer, int H, int W, int scale, void* stream) {
  if (outer < 0 || H < 0 || W < 0 || scale < 1) return set_error(SAD_ERR_INVALID, "upsample nearest gradient: bad shape or scale");
  const int64_t n_in = outer * H * W, n_out = n_in * scale * scale;
  if (n_in == 0) return SAD_OK;
  if (!dy || !dx) return set_error(SAD_ERR_INVALID, "upsample nearest gradient: null pointer");
  if (n_out >= ((int64_t)1 << 31)) return set_error(SAD_ERR_UNSUPPORTED, "upsample nearest gradient: 2^31 elements or more");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (scale == 2 && W % 4 == 0 && aligned16(dy, dx))
    upsample2_grad_vec4_kernel<<<stream_grid((size_t)n_in / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx),
                                                                              (uint32_t)(n_in / 4), (uint32_t)(W / 4));
  else
    upsample_grad_generic_kernel<<<stream_grid((size_t)n_in), 256, 0, st>>>(dy, dx, (uint32_t)n_in, (uint32_t)H, (uint32_t)W, (uint32_t)scale);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "upsample nearest gradient launch");
}

SAD_EXPORT int sad_upsample_nearest_add_f32(const float* top, const float* lateral, float* out, int64_t outer, int H_out, int W_out,
                                            int64_t inner, void* stream) {
  if (outer < 0 || H_out < 0 || W_out < 0 || inner < 1) return set_error(SAD_ERR_INVALID, "UpsampleNearest+Sum: bad dimension");
  if ((H_out & 1) || (W_out & 1)) return set_error(SAD_ERR_INVALID, "UpsampleNearest+Sum: the output plane must be twice the top plane (even H, W)");
  const uint64_t n = (uint64_t)outer * H_out * W_out * inner;
  if (n == 0) return SAD_OK;
  if (n >= 0x7fffffffull) return set_error(SAD_ERR_INVALID, "UpsampleNearest+Sum: tensor too large for 32-bit indexing");
  if (!top || !lateral || !out) return set_error(SAD_ERR_INVALID, "UpsampleNearest+Sum: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool aligned = ((reinterpret_cast<uintptr_t>(top) | reinterpret_cast<uintptr_t>(lateral) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  auto blocks = [&](uint64_t work) {
    const uint64_t want = (work + 255) / 256, cap = (uint64_t)sms * 16;
    return (unsigned)(want < cap ? want : cap);
  };
  if (aligned && (inner & 3) == 0) {
    upsample2_add_kernel<true><<<blocks(n / 4), 256, 0, st>>>(top, lateral, out, (uint32_t)(n / 4), (uint32_t)W_out, (uint32_t)H_out, (uint32_t)inner);
  } else if (aligned && inner == 1 && (W_out & 3) == 0) {
    upsample2_add_kernel<false><<<blocks(n / 4), 256, 0, st>>>(top, lateral, out, (uint32_t)(n / 4), (uint32_t)W_out, (uint32_t)H_out, 1u);
  } else {
    upsample2_add_scalar_kernel<<<blocks(n), 256, 0, st>>>(top, lateral, out, (uint32_t)n, (uint32_t)W_out, (uint32_t)H_out, (uint32_t)inner);
  }
  count_launch(1);
  return check_cuda(cudaGetLastError(), "UpsampleNearest+Sum launch");
}

}  // extern "C"
