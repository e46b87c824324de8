// TEST INFRASTRUCTURE (oracle/_ref build only; nothing here is linked into the product).
//
// CPU definitions of the caffe2::math primitives the reference's own CPU convolution calls
// (caffe2/caffe2/operators/conv_op_impl.h:31-180 forward, :357-560 gradient), so that conv_op.cc,
// conv_gradient_op.cc and conv_op_shared.cc compile UNMODIFIED from /root/reference into oracle/_ref/libref_ops.so
// and pin oracle/conv_oracle.c and the tcgen05 kernels to the reference's operator.
//
// What is restated and from where:
//   Set      caffe2/caffe2/utils/math_cpu.cc:759-775       fill
//   Im2col   caffe2/caffe2/utils/math_cpu.cc:1058-1160     NCHW; the general ("Baseline") loop, which produces the
//                                                          same column buffer as the two fast paths above it
//   Col2im   caffe2/caffe2/utils/math_cpu.cc:1222-1338     NCHW; the general ("Fallback") loop: for one image pixel the
//                                                          contributions arrive in (channel, kernel row, kernel col)
//                                                          order exactly as in the equal-padding fast path
//   Gemm     caffe2/caffe2/utils/math_cpu.cc:84-141,297-335  row-major C = alpha op(A) op(B) + beta C.  The reference
//   Gemv     caffe2/caffe2/utils/math_cpu.cc:205-245,337-352 forwards to Eigen or cblas_sgemm / cblas_sgemv (third
//                                                          party, neither vendored source is built here); both are
//                                                          fp32 dot products whose summation ORDER is the library's
//                                                          own.  Restated as fp32 accumulation in ascending k —
//                                                          equal to the library result to fp32 round-off
//                                                          (~sqrt(K) ulp), far below the 1e-4 gate it anchors.
//   NHWC Im2col / Col2im and the N-d forms: the head path is NCHW 2-D (retinanet_heads.py:105-152); these exist
//   only so the templates link, and throw if reached.
#include "caffe2/core/context.h"
#include "caffe2/utils/math.h"

namespace caffe2 {
namespace math {

template <>
void Set<float, CPUContext>(const size_t N, const float alpha, float* Y, CPUContext*) {
  for (size_t i = 0; i < N; ++i) Y[i] = alpha;
}
template <>
void Set<int, CPUContext>(const size_t N, const int alpha, int* Y, CPUContext*) {
  for (size_t i = 0; i < N; ++i) Y[i] = alpha;
}

template <>
void Gemm<float, CPUContext, DefaultEngine>(const CBLAS_TRANSPOSE TransA, const CBLAS_TRANSPOSE TransB, const int M,
                                            const int N, const int K, const float alpha, const float* A,
                                            const float* B, const float beta, float* C, CPUContext*, int) {
  // row-major: op(A) is M x K, op(B) is K x N, C is M x N (math_cpu.cc:297-335: lda = TransA ? M : K, ldb = TransB ? K : N)
  const bool ta = TransA != CblasNoTrans, tb = TransB != CblasNoTrans;
#pragma omp parallel for schedule(static)
  for (int m = 0; m < M; ++m) {
    float* c = C + (size_t)m * N;
    if (beta == 0.f) {
      for (int n = 0; n < N; ++n) c[n] = 0.f;   // BLAS semantics: C is not read when beta == 0
    } else if (beta != 1.f) {
      for (int n = 0; n < N; ++n) c[n] *= beta;
    }
    if (!tb) {
      // ascending k outer, contiguous n inner: each c[n] still accumulates its products in ascending k
      for (int k = 0; k < K; ++k) {
        const float a = alpha * (ta ? A[(size_t)k * M + m] : A[(size_t)m * K + k]);
        const float* b = B + (size_t)k * N;
        for (int n = 0; n < N; ++n) c[n] += a * b[n];
      }
    } else {
      for (int n = 0; n < N; ++n) {
        const float* b = B + (size_t)n * K;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += (ta ? A[(size_t)k * M + m] : A[(size_t)m * K + k]) * b[k];
        c[n] += alpha * acc;
      }
    }
  }
}

template <>
void Gemv<float, CPUContext, DefaultEngine>(const CBLAS_TRANSPOSE TransA, const int M, const int N, const float alpha,
                                            const float* A, const float* x, const float beta, float* y, CPUContext*,
                                            int) {
  // A is M x N row-major.  NoTrans: y[M] = alpha A x[N] + beta y;  Trans: y[N] = alpha A^T x[M] + beta y
  if (TransA == CblasNoTrans) {
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m) {
      float acc = 0.f;
      for (int n = 0; n < N; ++n) acc += A[(size_t)m * N + n] * x[n];
      y[m] = alpha * acc + (beta == 0.f ? 0.f : beta * y[m]);
    }
  } else {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int m = 0; m < M; ++m) acc += A[(size_t)m * N + n] * x[m];
      y[n] = alpha * acc + (beta == 0.f ? 0.f : beta * y[n]);
    }
  }
}

template <>
void Im2col<float, CPUContext, StorageOrder::NCHW>(const float* data_im, const int channels, const int height,
                                                   const int width, const int kernel_h, const int kernel_w,
                                                   const int dilation_h, const int dilation_w, const int pad_t,
                                                   const int pad_l, const int pad_b, const int pad_r,
                                                   const int stride_h, const int stride_w, float* data_col,
                                                   CPUContext*) {
  const int dkernel_h = dilation_h * (kernel_h - 1) + 1;
  const int dkernel_w = dilation_w * (kernel_w - 1) + 1;
  const int height_col = (height + pad_t + pad_b - dkernel_h) / stride_h + 1;
  const int width_col = (width + pad_l + pad_r - dkernel_w) / stride_w + 1;
  const int channels_col = channels * kernel_h * kernel_w;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < channels_col; ++c) {
    const int w_offset = c % kernel_w;
    const int h_offset = (c / kernel_w) % kernel_h;
    const int c_im = c / kernel_h / kernel_w;
    for (int h = 0; h < height_col; ++h) {
      for (int w = 0; w < width_col; ++w) {
        const int h_pad = h * stride_h - pad_t + h_offset * dilation_h;
        const int w_pad = w * stride_w - pad_l + w_offset * dilation_w;
        const bool inside = h_pad >= 0 && h_pad < height && w_pad >= 0 && w_pad < width;
        data_col[((size_t)c * height_col + h) * width_col + w] =
            inside ? data_im[((size_t)c_im * height + h_pad) * width + w_pad] : 0.f;
      }
    }
  }
}

template <>
void Col2im<float, CPUContext, StorageOrder::NCHW>(const float* data_col, const int channels, const int height,
                                                   const int width, const int kernel_h, const int kernel_w,
                                                   const int dilation_h, const int dilation_w, const int pad_t,
                                                   const int pad_l, const int pad_b, const int pad_r,
                                                   const int stride_h, const int stride_w, float* data_im,
                                                   CPUContext* context) {
  Set<float, CPUContext>((size_t)height * width * channels, 0, data_im, context);
  const int dkernel_h = dilation_h * (kernel_h - 1) + 1;
  const int dkernel_w = dilation_w * (kernel_w - 1) + 1;
  const int height_col = (height + pad_t + pad_b - dkernel_h) / stride_h + 1;
  const int width_col = (width + pad_l + pad_r - dkernel_w) / stride_w + 1;
  // one image channel per thread: within a channel the adds keep the reference's (kernel row, kernel col, h, w) order
#pragma omp parallel for schedule(static)
  for (int c_im = 0; c_im < channels; ++c_im) {
    for (int kk = 0; kk < kernel_h * kernel_w; ++kk) {
      const int c = c_im * kernel_h * kernel_w + kk;
      const int w_offset = c % kernel_w;
      const int h_offset = (c / kernel_w) % kernel_h;
      for (int h = 0; h < height_col; ++h) {
        for (int w = 0; w < width_col; ++w) {
          const int h_pad = h * stride_h - pad_t + h_offset * dilation_h;
          const int w_pad = w * stride_w - pad_l + w_offset * dilation_w;
          if (h_pad >= 0 && h_pad < height && w_pad >= 0 && w_pad < width)
            data_im[((size_t)c_im * height + h_pad) * width + w_pad] +=
                data_col[((size_t)c * height_col + h) * width_col + w];
        }
      }
    }
  }
}

template <>
void Im2col<float, CPUContext, StorageOrder::NHWC>(const float*, const int, const int, const int, const int, const int,
                                                   const int, const int, const int, const int, const int, const int,
                                                   const int, const int, float*, CPUContext*) {
  CAFFE_THROW("oracle/_ref: NHWC Im2col is outside the head path (retinanet_heads.py uses NCHW)");
}
template <>
void Col2im<float, CPUContext, StorageOrder::NHWC>(const float*, const int, const int, const int, const int, const int,
                                                   const int, const int, const int, const int, const int, const int,
                                                   const int, const int, float*, CPUContext*) {
  CAFFE_THROW("oracle/_ref: NHWC Col2im is outside the head path (retinanet_heads.py uses NCHW)");
}
template <>
void Im2colNd<float, CPUContext, StorageOrder::NCHW>(const float*, const int*, const int*, const int, const int,
                                                     const int*, const int*, const int*, const int*, const int, float*,
                                                     CPUContext*, bool) {
  CAFFE_THROW("oracle/_ref: N-d Im2col is outside the head path (2-D kernels only)");
}
template <>
void Col2imNd<float, CPUContext, StorageOrder::NCHW>(const float*, const int*, const int*, const int, const int,
                                                     const int*, const int*, const int*, const int*, const int, float*,
                                                     CPUContext*) {
  CAFFE_THROW("oracle/_ref: N-d Col2im is outside the head path (2-D kernels only)");
}

}  // namespace math
}  // namespace caffe2
