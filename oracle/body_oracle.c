/* TEST INFRASTRUCTURE — CPU restatement of the reference's AffineChannel(+Gradient) and UpsampleNearest(+Gradient)
 * CUDA kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may use it; never the product.
 *
 * Follows, expression by expression:
 *   caffe2/modules/detectron/affine_channel_op.cu:22-35   ScaleBiasForward  out = in * scale[c] + bias[c], c = (i / hxw) % C
 *   caffe2/modules/detectron/affine_channel_op.cu:37-48   ScaleForward      out = in * scale[c]
 *   caffe2/modules/detectron/upsample_nearest_op.cu:66-80   translate_idx      (output index -> input index)
 *   caffe2/modules/detectron/upsample_nearest_op.cu:82-100  translate_idx_inv
 *   caffe2/modules/detectron/upsample_nearest_op.cu:102-113 upscale / downscale (math::Set 0, then += over x offset i, y offset j)
 * nvcc contracts `in * scale + bias` into one FMA (default -fmad=true), so the restatement calls fmaf; this file is built
 * with -ffp-contract=off so nothing else is contracted.  Pinning: tests/test_body_ops_gpu.py runs the reference's own .cu
 * (oracle/_ref) beside it on the GPU; tests/test_body_oracle.py checks it against numpy on CPU.  Parity is bit-exact. */
#include <math.h>
#include <stdint.h>

void oracle_affine_channel(int64_t n, int C, int64_t hxw, const float* in, const float* scale, const float* bias, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t index = 0; index < n; ++index) {
    const int scale_index = (int)((index / hxw) % C);
    out[index] = bias ? fmaf(in[index], scale[scale_index], bias[scale_index]) : in[index] * scale[scale_index];
  }
}

static int translate_idx(int ii, int d1, int d2, int d3, int scale_factor) {
  int x, y, z, w;
  w = ii % d3; ii = ii / d3;
  z = ii % d2; ii = ii / d2;
  y = ii % d1; ii = ii / d1;
  x = ii;
  w = w / scale_factor;
  z = z / scale_factor;
  d2 /= scale_factor;
  d3 /= scale_factor;
  return (((x * d1 + y) * d2) + z) * d3 + w;
}

static int translate_idx_inv(int ii, int d1, int d2, int d3, int scale_factor, int off_x, int off_y) {
  int x, y, z, w;
  w = ii % d3; ii = ii / d3;
  z = ii % d2; ii = ii / d2;
  y = ii % d1; ii = ii / d1;
  x = ii;
  w = w * scale_factor + off_x;
  z = z * scale_factor + off_y;
  d2 *= scale_factor;
  d3 *= scale_factor;
  return (((x * d1 + y) * d2) + z) * d3 + w;
}

/* output (.., d1, d2, d3) = the UPSAMPLED dims, as the op passes them (upsample_nearest_op.cu:129-138) */
void oracle_upsample_nearest(const float* input, float* output, int64_t no_elements, int scale_factor, int d1, int d2, int d3) {
#pragma omp parallel for schedule(static)
  for (int64_t ii = 0; ii < no_elements; ++ii) output[ii] = input[translate_idx((int)ii, d1, d2, d3, scale_factor)];
}

/* gradInput (.., d1, d2, d3) = the INPUT dims (upsample_nearest_op.cu:176-185) */
void oracle_upsample_nearest_grad(float* gradInput, const float* gradOutput, int64_t no_elements, int scale_factor, int d1, int d2,
                                  int d3) {
#pragma omp parallel for schedule(static)
  for (int64_t ii = 0; ii < no_elements; ++ii) {
    gradInput[ii] = 0.f; /* math::Set, :208 */
    for (int i = 0; i < scale_factor; i++)
      for (int j = 0; j < scale_factor; j++) gradInput[ii] += gradOutput[translate_idx_inv((int)ii, d1, d2, d3, scale_factor, i, j)];
  }
}

/* FPN top-down merge: UpsampleNearest(top, 2) (upsample_nearest_op.cu:102-108) followed by Sum([lateral, td]) (FPN.py:230-249;
 * math::Add: one fp32 add per element).  Tensors viewed as (outer, H_out, W_out, inner): inner = 1 NCHW, inner = C channels-last. */
void oracle_upsample2_add(const float* top, const float* lateral, float* out, int64_t outer, int Ho, int Wo, int64_t inner) {
  const int Hi = Ho / 2, Wi = Wo / 2;
  for (int64_t o = 0; o < outer; ++o)
    for (int Y = 0; Y < Ho; ++Y)
      for (int X = 0; X < Wo; ++X)
        for (int64_t c = 0; c < inner; ++c) {
          const int64_t i = ((o * Ho + Y) * Wo + X) * inner + c;
          out[i] = lateral[i] + top[((o * Hi + Y / 2) * Wi + X / 2) * inner + c];
        }
}
