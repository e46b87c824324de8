/* TEST INFRASTRUCTURE — CPU oracle for SigmoidFocalLoss / SigmoidFocalLossGradient (the classification loss that
 * shares logits and labels with the distillation loss).  Same rules as distill_oracle.c:
 * only tests/, smoke() and bench.py's CPU-baseline legs may load it.
 *
 * One C function per reference function, every implicit float<->double promotion of the device expressions written
 * out as an explicit cast:
 *   oracle_focal_loss_elem  <- caffe2/modules/detectron/sigmoid_focal_loss_op.cu:26-66   (SigmoidFocalLossKernel)
 *   oracle_focal_grad_elem  <- caffe2/modules/detectron/sigmoid_focal_loss_op.cu:68-109  (SigmoidFocalLossGradientKernel)
 *   oracle_focal_loss       <- ...:112-144 (kernel, math::Sum without scratch, math::Scale)
 *   oracle_focal_grad       <- ...:147-173 (kernel, math::Scale over the tensor)
 * Pinned against the reference itself: the unmodified sigmoid_focal_loss_op.{cc,cu} are part of oracle/_ref/libref_ops.so
 * and are run side by side on the GPU (tests/test_focal_gpu.py), and their outputs are committed as golden vectors.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define ORACLE_API __attribute__((visibility("default")))

extern float oracle_ref_order_sum(const float* x, int64_t n); /* distill_oracle.c: math_gpu.cu:1021-1058 */

/* label of element i of an (N, D, H, W) tensor and its class index d (sigmoid_focal_loss_op.cu:32-40) */
static inline int focal_target(const int32_t* targets, int64_t i, int D, int H, int W, int num_classes, int* d_out) {
  int x = (int)(i % W);
  int y = (int)((i / W) % H);
  int c = (int)((i / ((int64_t)W * H)) % D);
  int n = (int)(i / ((int64_t)W * H * D));
  int A = D / num_classes;
  int a = c / num_classes;
  *d_out = c % num_classes;
  return targets[(int64_t)n * (H * W * A) + (int64_t)a * (H * W) + y * W + x];
}

ORACLE_API float oracle_focal_loss_elem(float logit, int t, int d, float weight_pos, float gamma, float alpha) {
  float c1 = (float)(t == (d + 1));
  float c2 = (float)((t != -1) & (t != (d + 1)));
  float Np = (float)fmax((double)weight_pos, 1.0);
  float zn = (float)((1.0 - (double)alpha) / (double)Np);
  float zp = alpha / Np;
  float p = (float)(1. / (1. + (double)expf(-logit)));
  float term1 = powf((float)(1. - (double)p), gamma) * logf(fmaxf(p, FLT_MIN));
  int ge = logit >= 0;
  float term2 = (float)((double)powf(p, gamma) *
                        (-1. * (double)logit * (double)ge - (double)logf((float)(1. + (double)expf((float)((double)logit - 2. * (double)logit * (double)ge))))));
  float loss = (float)0.0;
  loss += -c1 * term1 * zp;
  loss += -c2 * term2 * zn;
  return loss;
}

ORACLE_API float oracle_focal_grad_elem(float logit, int t, int d, float weight_pos, float gamma, float alpha, float a_loss) {
  float Np = (float)fmax((double)weight_pos, 1.0);
  float zn = (float)((1.0 - (double)alpha) / (double)Np);
  float zp = alpha / Np;
  float c1 = (float)(t == (d + 1));
  float c2 = (float)((t != -1) & (t != (d + 1)));
  float p = (float)(1. / (1. + (double)expf(-logit)));
  float term1 = (float)((double)powf((float)(1. - (double)p), gamma) * (1. - (double)p - (double)(p * gamma * logf(fmaxf(p, FLT_MIN)))));
  int ge = logit >= 0;
  float term2 = (float)((double)powf(p, gamma) *
                        ((-1. * (double)logit * (double)ge - (double)logf((float)(1. + (double)expf((float)((double)logit - 2. * (double)logit * (double)ge))))) *
                             (1. - (double)p) * (double)gamma -
                         (double)p));
  float dx = (float)0.0;
  dx += -c1 * zp * term1;
  dx += -c2 * zn * term2;
  dx = dx * a_loss;
  return dx;
}

ORACLE_API float oracle_focal_loss(int N, int D, int H, int W, const float* logits, const int32_t* targets, const float* weight_pos,
                                   float gamma, float alpha, int num_classes, float scale, float* losses) {
  const int64_t n = (int64_t)N * D * H * W;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int d;
    int t = focal_target(targets, i, D, H, W, num_classes, &d);
    losses[i] = oracle_focal_loss_elem(logits[i], t, d, weight_pos[0], gamma, alpha);
  }
  float s = oracle_ref_order_sum(losses, n);
  return s * scale; /* math::Scale(1, scale_, avg_loss, avg_loss) */
}

ORACLE_API void oracle_focal_grad(int N, int D, int H, int W, const float* logits, const int32_t* targets, const float* weight_pos,
                                  float gamma, float alpha, int num_classes, float scale, const float* d_loss, float* dX) {
  const int64_t n = (int64_t)N * D * H * W;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int d;
    int t = focal_target(targets, i, D, H, W, num_classes, &d);
    float g = oracle_focal_grad_elem(logits[i], t, d, weight_pos[0], gamma, alpha, d_loss[0]);
    dX[i] = g * scale; /* math::Scale(n, scale_, dX, dX): a separate rounding */
  }
}

/* SelectSmoothL1Loss / Gradient: caffe2/modules/detectron/select_smooth_l1_loss_op.cu:23-54 (kernel), :90-143 (zero-filled buffer as
 * large as Y_hat, kernel, math::Sum over the WHOLE buffer, math::Scale), :57-86 and :145-181 (gradient: zero fill + scatter). */
ORACLE_API float oracle_select_smooth_l1_loss(int N, int D, int H, int W, int M, const float* Y_hat, const float* Y, const float* L,
                                              const float* S, float beta, float scale, float* buff) {
  const int64_t total = (int64_t)N * D * H * W;
  if (M == 0) return 0.f;
  for (int64_t i = 0; i < total; ++i) buff[i] = 0.f;
  for (int i = 0; i < M; ++i) {
    int n = (int)L[i * 4], c = (int)L[i * 4 + 1], y = (int)L[i * 4 + 2], x = (int)L[i * 4 + 3];
    for (int j = 0; j < 4; ++j) {
      int ind = n * (D * H * W) + (c + j) * (H * W) + y * W + x;
      float val = Y_hat[ind] - Y[i * 4 + j];
      float abs_val = fabsf(val);
      if (abs_val < beta) buff[ind] = (float)((0.5 * (double)val * (double)val / (double)beta) / fmax((double)S[0], 1.0));
      else buff[ind] = (float)(((double)abs_val - 0.5 * (double)beta) / fmax((double)S[0], 1.0));
    }
  }
  return oracle_ref_order_sum(buff, total) * scale;
}

ORACLE_API void oracle_select_smooth_l1_grad(int N, int D, int H, int W, int M, const float* Y_hat, const float* Y, const float* L,
                                             const float* S, float beta, float scale, const float* d_loss, float* out) {
  const int64_t total = (int64_t)N * D * H * W;
  for (int64_t i = 0; i < total; ++i) out[i] = 0.f;
  for (int i = 0; i < M; ++i) {
    int n = (int)L[i * 4], c = (int)L[i * 4 + 1], y = (int)L[i * 4 + 2], x = (int)L[i * 4 + 3];
    for (int j = 0; j < 4; ++j) {
      int ind = n * (D * H * W) + (c + j) * (H * W) + y * W + x;
      float val = Y_hat[ind] - Y[i * 4 + j];
      float abs_val = fabsf(val);
      if (abs_val < beta) out[ind] = (float)((double)(scale * d_loss[0] * val / beta) / fmax((double)S[0], 1.0));
      else out[ind] = (float)((double)(scale * d_loss[0] * (float)((0.f < val) - (val < 0.f))) / fmax((double)S[0], 1.0));
    }
  }
}

/* MomentumSGDUpdate with Detectron's preamble per parameter (detectron/lib/modeling/optimizer.py:115-130):
 * bias: Scale(grad, 2.0); weight: WeightedSum(grad, 1, param, wd); then MomentumSGDKernel (momentum_sgd_op_gpu.cu:23-54). */
ORACLE_API void oracle_momentum_sgd(int64_t n, float* param, float* grad, float* mom, float lr, float momentum, int nesterov,
                                    float grad_mult, float wd) {
  for (int64_t i = 0; i < n; ++i) {
    float g = grad[i];
    if (grad_mult != 1.f) g = g * grad_mult;              /* Scale */
    if (wd != 0.f) g = g * 1.f + param[i] * wd;           /* WeightedSum: sum of w_k * x_k */
    if (!nesterov) {
      const float adjusted_gradient = lr * g + momentum * mom[i];
      mom[i] = adjusted_gradient;
      grad[i] = adjusted_gradient;
      param[i] -= adjusted_gradient;
    } else {
      const float mi = mom[i];
      const float mi_new = momentum * mi + lr * g;
      mom[i] = mi_new;
      grad[i] = (1 + momentum) * mi_new - momentum * mi;
      param[i] -= grad[i];
    }
  }
}

/* WeightedSum (caffe2/caffe2/operators/utility_ops.h:333-378): out = X_0 * w_0 (math::Scale), then out += X_k * w_k (math::Axpy,
 * which nvcc contracts into one FMA per element on the GPU: caffe2/caffe2/utils/math_gpu.cu AxpyKernel `y[i] += x[i] * (*a)`). */
ORACLE_API void oracle_weighted_sum(int64_t n, int n_inputs, const float* const* xs, const float* ws, float* out) {
  for (int64_t i = 0; i < n; ++i) {
    float y = xs[0][i] * ws[0];
    for (int k = 1; k < n_inputs; ++k) y = fmaf(xs[k][i], ws[k], y);
    out[i] = y;
  }
}
