/* TEST INFRASTRUCTURE — CPU oracle for the adaptive-distillation operators.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library; the product path never does.
 *
 * It is a CPU restatement of the reference's CUDA operators (the reference has NO CPU
 * implementation: pow_sum_op.h:33-36, sigmoid_adaptive_distillation_loss_op.h:42-45,73-76), one
 * C function per reference function, with every implicit float<->double promotion of the device
 * expressions written out as an explicit cast:
 *
 *   oracle_pow_sum            <- caffe2/modules/detectron/pow_sum_op.cu:25-43
 *   oracle_distill_loss_elem  <- caffe2/modules/detectron/sigmoid_adaptive_distillation_loss_op.cu:33-66
 *   oracle_distill_grad_elem  <- caffe2/modules/detectron/sigmoid_adaptive_distillation_loss_op.cu:74-104
 *   oracle_distill_loss       <- ...loss_op.cu:108-141 (kernel, math::Sum, math::Scale)
 *   oracle_distill_grad       <- ...loss_op.cu:144-171 (kernel, math::Scale)
 *   ref_order_sum             <- caffe2/caffe2/utils/math_gpu.cu:1021-1058 (SumKernel<<<1,128>>>)
 *
 * Parity pinning: the reference ships no test, golden vector or fixture for these operators
 * (SURVEY.md §4), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: the
 * unmodified reference .cu files built as oracle/_ref/libref_ops.so and run on the B200
 * (tests/test_ref_gpu_oracle.py), plus hand-computed known-answer cases in tests/golden/.
 *
 * Differences that remain between this host code and the device code: libm vs libdevice ulps
 * and FMA contraction (nvcc fuses a*b+c, this file is built with -ffp-contract=off).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* math_gpu.cu:1021-1058 — the scratch-less Sum the ops use: ONE block of 128 threads.
 * lane j accumulates x[j], x[j+128], ... in float; lanes 0..31 then add lanes +32,+64,+96 as
 * (p[j+32] + p[j+64]) + p[j+96] added to p[j]; lane 0 finally adds p[0..31] in order. */
static float ref_order_sum(const float* x, int64_t n) {
  float part[128];
  for (int j = 0; j < 128; ++j) part[j] = 0.f;
  int64_t full = n / 128 * 128;
  for (int64_t base = 0; base < full; base += 128)
    for (int j = 0; j < 128; ++j) part[j] += x[base + j];
  for (int64_t i = full; i < n; ++i) part[i - full] += x[i];
  for (int j = 0; j < 32; ++j) part[j] += part[j + 32] + part[j + 64] + part[j + 96];
  float total = 0.f;
  for (int j = 0; j < 32; ++j) total += part[j];
  return total;
}
ORACLE_API float oracle_ref_order_sum(const float* x, int64_t n) { return ref_order_sum(x, n); }

/* pow_sum_op.cu:25-43.  res = 0; for each input: buff = powf(in, power) (math_gpu.cu:1257-1262);
 * s = Sum(buff); res = res + s (float adds, input order).  `scratch` needs max(sizes) floats. */
ORACLE_API float oracle_pow_sum(const float* const* inputs, const int64_t* sizes, int n_inputs, float power,
                                float* scratch) {
  float res = 0.f;
  for (int k = 0; k < n_inputs; ++k) {
    const float* in = inputs[k];
    const int64_t n = sizes[k];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) scratch[i] = powf(in[i], power);
    float s = ref_order_sum(scratch, n);
    res = res + s;
  }
  return res;
}

/* bool -> int -> float, as `(logits[i] >= 0)` converts inside a float expression */
static inline float ge0f(float x) { return (float)(x >= 0); }
static inline double ge0d(float x) { return (double)(x >= 0); }

/* CUDA's max(float, double) overload evaluates in double (fmax); NaN loses to the number. */
static inline float norm_clamp(float wp) { return (float)fmax((double)wp, 1.0); }

/* ...loss_op.cu:35-42: which label the flat element index i of an (N, D, H, W) tensor reads.
 * All of it is int arithmetic on the int-narrowed loop variable. */
ORACLE_API int oracle_label_index(int i, int D, int H, int W, int num_classes) {
  int x = i % W;
  int y = (i / W) % H;
  int c = (i / (W * H)) % D;
  int n = i / (W * H * D);
  int A = D / num_classes;
  int a = c / num_classes;
  return n * (H * W * A) + a * (H * W) + y * W + x;
}

/* ...loss_op.cu:49-64, one element.  `keep` = (t != ignored_label). */
ORACLE_API float oracle_distill_loss_elem(float x, float pt, int keep, float wp, float gamma, float alpha,
                                          float beta) {
  float Np = norm_clamp(wp);                                   /* :49  double max -> float      */
  float zn = (float)((1.0 - (double)alpha) / (double)Np);      /* :50  double                   */
  float zp = alpha / Np;                                       /* :51  float                    */
  float p = (float)(1. / (1. + (double)expf(-x)));             /* :55  exp(float) is the float  */
                                                               /*      overload; add/div double */
  float inner = x - 2 * x * ge0f(x);                           /* :58  float (int literal 2)    */
  double dl = -1. * (double)x * (double)(pt - ge0f(x))         /* :58  double (literal -1.)     */
              + (double)logf(fmaxf(FLT_MIN, 1 + expf(inner)));
  dl = dl + (double)(beta * (pt * logf(pt) + (1 - pt) * logf(1 - pt))); /* :59 float term       */
  float D_loss = (float)dl;
  float adaptive_target = 1 - expf(-D_loss);                   /* :61  float                    */

  /* :63-64.  The parenthesis holding `-1.*x*(x>=0)` and `2.*x` is double; expf/logf narrow
   * their argument to float and return float. */
  double arg = (double)x - 2. * (double)x * ge0d(x);
  double neg = -1. * (double)x * ge0d(x) - (double)logf((float)(1. + (double)expf((float)arg)));
  double term = (double)(pt * logf(fmaxf(FLT_MIN, p)) * zp) + (double)(1 - pt) * neg * (double)zn;
  double out = (double)(-powf(adaptive_target, gamma)) * term * (double)keep;
  return (float)out;
}

/* ...loss_op.cu:87-102 and the Scale at :167-168, one element.  Three roundings are kept:
 * store, divide by Np, multiply by scale. */
ORACLE_API float oracle_distill_grad_elem(float x, float pt, int keep, float wp, float gamma, float alpha,
                                          float beta, float d_loss, float scale) {
  float Np = norm_clamp(wp);                                   /* :87 */
  float p = (float)(1. / (1. + (double)expf(-x)));             /* :89 */
  float inner = x - 2 * x * ge0f(x);
  double dl = -1. * (double)x * (double)(pt - ge0f(x))         /* :92  no FLT_MIN clamp here    */
              + (double)logf(1 + expf(inner));
  dl = dl + (double)(beta * (pt * logf(pt) + (1 - pt) * logf(1 - pt))); /* :93 */
  float DL = (float)dl;
  float expDL = expf(-DL);                                     /* :94 */
  float adaptive_target = 1 - expDL;                           /* :95 */

  double arg = (double)x - 2. * (double)x * ge0d(x);           /* :97 */
  double neg = -1. * (double)x * ge0d(x) - (double)logf((float)(1. + (double)expf((float)arg)));
  float DLoss = (float)((double)(alpha * pt * logf(fmaxf(FLT_MIN, p))) +
                        (double)((1 - alpha) * (1 - pt)) * neg);
  float dx = -(-(pt - p) * gamma * powf(adaptive_target, gamma - 1) * expDL * DLoss +   /* :98-99 */
               powf(adaptive_target, gamma) * (alpha * (pt - p) - (1 - 2 * alpha) * (1 - pt) * p)) *
             d_loss * (float)keep;
  dx = dx / Np;                                                /* :102 */
  dx = dx * scale;                                             /* :167-168 math::Scale */
  return dx;
}

/* ...loss_op.cu:108-141.  losses (N*D*H*W floats) is the op's member scratch `losses_`; the
 * result is Sum(losses) in the reference's summation order, then one float multiply by scale. */
ORACLE_API float oracle_distill_loss(int N, int D, int H, int W, int ignored_label, const float* logits,
                                     const float* targets, const int32_t* gt, const float* weight_pos,
                                     float gamma, float alpha, float beta, int num_classes, float scale,
                                     float* losses) {
  const int64_t total = (int64_t)N * D * H * W;
  const float wp = weight_pos[0];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    int t = gt[oracle_label_index((int)i, D, H, W, num_classes)];
    losses[i] = oracle_distill_loss_elem(logits[i], targets[i], t != ignored_label, wp, gamma, alpha, beta);
  }
  float avg_loss = ref_order_sum(losses, total);
  avg_loss = avg_loss * scale;
  return avg_loss;
}

/* ...loss_op.cu:144-171. */
ORACLE_API void oracle_distill_grad(int N, int D, int H, int W, int ignored_label, const float* logits,
                                    const float* targets, const int32_t* gt, const float* weight_pos,
                                    float gamma, float alpha, float beta, int num_classes, float scale,
                                    const float* d_avg_loss, float* dX) {
  const int64_t total = (int64_t)N * D * H * W;
  const float wp = weight_pos[0];
  const float a_loss = d_avg_loss[0];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    int t = gt[oracle_label_index((int)i, D, H, W, num_classes)];
    dX[i] = oracle_distill_grad_elem(logits[i], targets[i], t != ignored_label, wp, gamma, alpha, beta, a_loss,
                                     scale);
  }
}

/* Not in the reference: a float64 evaluation of the same formulas (no float roundings), used by
 * tests to separate "differs from the reference" from "the reference itself is ill-conditioned
 * here" (SURVEY.md Appendix D item 2). */
ORACLE_API void oracle_distill_elem_f64(double x, double pt, int keep, double wp, double gamma, double alpha,
                                        double beta, double d_loss, double scale, double* loss_out,
                                        double* grad_out) {
  double Np = fmax(wp, 1.0);
  double s = x >= 0 ? 1.0 : 0.0;
  double e = exp(-fabs(x));
  double L = log1p(e);
  double p = x >= 0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
  double ent = beta * (pt * log(pt) + (1 - pt) * log(1 - pt));
  double DL = -x * (pt - s) + L + ent;
  double E = exp(-DL);
  double AT = 1 - E;
  double logp = fmax(log(DBL_MIN), (x < 0 ? x : 0.0) - L);
  double lq = -x * s - L;
  double DLoss = alpha * pt * logp + (1 - alpha) * (1 - pt) * lq;
  *loss_out = -pow(AT, gamma) * DLoss / Np * keep * scale;
  *grad_out = -(-(pt - p) * gamma * pow(AT, gamma - 1) * E * DLoss +
                pow(AT, gamma) * (alpha * (pt - p) - (1 - 2 * alpha) * (1 - pt) * p)) *
              d_loss * keep / Np * scale;
}
