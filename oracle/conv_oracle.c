/* TEST INFRASTRUCTURE — CPU oracle for the RetinaNet head's Conv / ConvGradient / Relu.
 *
 * The reference runs these through cuDNN (closed source, not under /root/reference:
 * caffe2/caffe2/operators/conv_op_cudnn.cc:567-617,1011-1059); the semantic definition the oracle
 * restates is the reference's own device-independent implementation:
 *
 *   oracle_conv2d_fwd   <- caffe2/caffe2/operators/conv_op_impl.h:31-180  (per image: Im2col, then
 *                          Y = filter[M x K] * col[K x HoWo], bias added as b[M] x ones[HoWo])
 *   oracle_conv2d_bwd   <- caffe2/caffe2/operators/conv_op_impl.h (RunOnDeviceWithOrderNCHW of the
 *                          gradient op): db = sum dY, dW += dY_n * col_n^T, dX_n = Col2im(W^T * dY_n)
 *   oracle_relu / _grad <- caffe2/caffe2/operators/relu_op.cu:22-35  (Y = X>0 ? X : 0,
 *                          dX = Y>0 ? dY : 0)
 *
 * NCHW, fp32 storage, group 1, dilation 1, cross-correlation (no kernel flip,
 * conv_op_cudnn.cc:487).  Accumulation is float like the reference's sgemm; the association order
 * inside a BLAS is unspecified, so parity against this oracle is tolerance-based (stated in the
 * tests), never bit-exact.
 *
 * Parity pinning: the reference's conv tests (caffe2/caffe2/python/operator_test/conv_test.py:55-119)
 * are device-vs-device and numeric-gradient checks, not stored vectors, and cannot run here (no
 * caffe2 module).  This oracle is pinned by the same two properties in tests/test_conv_oracle.py:
 * a central-difference gradient check of oracle_conv2d_bwd against oracle_conv2d_fwd, and
 * agreement with torch.nn.functional.conv2d (an independent implementation) — "parity pinned by
 * property, not by reference vectors".
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static void im2col(const float* x, int C, int H, int W, int kh, int kw, int pad, int stride, int Ho, int Wo,
                   float* col) {
  /* col is (C*kh*kw) x (Ho*Wo), row index = (c*kh + ky)*kw + kx */
  for (int c = 0; c < C; ++c)
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx) {
        float* row = col + ((size_t)(c * kh + ky) * kw + kx) * Ho * Wo;
        for (int oy = 0; oy < Ho; ++oy) {
          int iy = oy * stride - pad + ky;
          for (int ox = 0; ox < Wo; ++ox) {
            int ix = ox * stride - pad + kx;
            row[oy * Wo + ox] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[((size_t)c * H + iy) * W + ix] : 0.f;
          }
        }
      }
}

static void col2im_add(const float* col, int C, int H, int W, int kh, int kw, int pad, int stride, int Ho, int Wo,
                       float* x) {
  for (int c = 0; c < C; ++c)
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx) {
        const float* row = col + ((size_t)(c * kh + ky) * kw + kx) * Ho * Wo;
        for (int oy = 0; oy < Ho; ++oy) {
          int iy = oy * stride - pad + ky;
          if (iy < 0 || iy >= H) continue;
          for (int ox = 0; ox < Wo; ++ox) {
            int ix = ox * stride - pad + kx;
            if (ix >= 0 && ix < W) x[((size_t)c * H + iy) * W + ix] += row[oy * Wo + ox];
          }
        }
      }
}

ORACLE_API int oracle_conv_out_dim(int in, int k, int pad, int stride) { return (in + 2 * pad - k) / stride + 1; }

/* Y[N,M,Ho,Wo] = conv(X[N,C,H,W], Wt[M,C,kh,kw]) + b[M] (b may be NULL) */
ORACLE_API void oracle_conv2d_fwd(const float* X, const float* Wt, const float* b, float* Y, int N, int C, int H,
                                  int W, int M, int kh, int kw, int pad, int stride) {
  const int Ho = oracle_conv_out_dim(H, kh, pad, stride), Wo = oracle_conv_out_dim(W, kw, pad, stride);
  const int K = C * kh * kw, P = Ho * Wo;
#pragma omp parallel
  {
    float* col = (float*)malloc((size_t)K * P * sizeof(float));
#pragma omp for schedule(dynamic)
    for (int n = 0; n < N; ++n) {
      im2col(X + (size_t)n * C * H * W, C, H, W, kh, kw, pad, stride, Ho, Wo, col);
      float* y = Y + (size_t)n * M * P;
      for (int m = 0; m < M; ++m) {
        float* yr = y + (size_t)m * P;
        for (int p = 0; p < P; ++p) yr[p] = 0.f;
        for (int k = 0; k < K; ++k) {
          const float w = Wt[(size_t)m * K + k];
          const float* cr = col + (size_t)k * P;
          for (int p = 0; p < P; ++p) yr[p] += w * cr[p];
        }
        if (b)
          for (int p = 0; p < P; ++p) yr[p] += b[m] * 1.0f;
      }
    }
    free(col);
  }
}

/* Outputs (any may be NULL): dW[M,C,kh,kw], db[M], dX[N,C,H,W]. */
ORACLE_API void oracle_conv2d_bwd(const float* X, const float* Wt, const float* dY, float* dW, float* db, float* dX,
                                  int N, int C, int H, int W, int M, int kh, int kw, int pad, int stride) {
  const int Ho = oracle_conv_out_dim(H, kh, pad, stride), Wo = oracle_conv_out_dim(W, kw, pad, stride);
  const int K = C * kh * kw, P = Ho * Wo;
  float* col = (float*)malloc((size_t)K * P * sizeof(float));
  if (dW) memset(dW, 0, (size_t)M * K * sizeof(float));
  if (db) memset(db, 0, (size_t)M * sizeof(float));
  for (int n = 0; n < N; ++n) {
    const float* dy = dY + (size_t)n * M * P;
    if (db)
      for (int m = 0; m < M; ++m) {
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += dy[(size_t)m * P + p];
        db[m] += s;
      }
    if (dW) {
      im2col(X + (size_t)n * C * H * W, C, H, W, kh, kw, pad, stride, Ho, Wo, col);
#pragma omp parallel for schedule(static)
      for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
          const float* dr = dy + (size_t)m * P;
          const float* cr = col + (size_t)k * P;
          float s = 0.f;
          for (int p = 0; p < P; ++p) s += dr[p] * cr[p];
          dW[(size_t)m * K + k] += s;
        }
    }
    if (dX) {
      /* col = W^T (K x M) * dY_n (M x P) */
#pragma omp parallel for schedule(static)
      for (int k = 0; k < K; ++k) {
        float* cr = col + (size_t)k * P;
        for (int p = 0; p < P; ++p) cr[p] = 0.f;
        for (int m = 0; m < M; ++m) {
          const float w = Wt[(size_t)m * K + k];
          const float* dr = dy + (size_t)m * P;
          for (int p = 0; p < P; ++p) cr[p] += w * dr[p];
        }
      }
      float* dx = dX + (size_t)n * C * H * W;
      memset(dx, 0, (size_t)C * H * W * sizeof(float));
      col2im_add(col, C, H, W, kh, kw, pad, stride, Ho, Wo, dx);
    }
  }
  free(col);
}

ORACLE_API void oracle_relu(const float* X, float* Y, int64_t n) {
  for (int64_t i = 0; i < n; ++i) Y[i] = X[i] > 0 ? X[i] : 0;
}
ORACLE_API void oracle_relu_grad(const float* Y, const float* dY, float* dX, int64_t n) {
  for (int64_t i = 0; i < n; ++i) dX[i] = Y[i] > 0 ? dY[i] : 0;
}
