// SigmoidFocalLoss (+Gradient) — the classification loss that reads the same logits and labels as the distillation loss
// (retinanet_heads.py:277-293).  Replaces
//   SigmoidFocalLossOp<float, CUDAContext>::RunOnDevice          caffe2/modules/detectron/sigmoid_focal_loss_op.cu:112-144
//   SigmoidFocalLossGradientOp<float, CUDAContext>::RunOnDevice  caffe2/modules/detectron/sigmoid_focal_loss_op.cu:147-173
// and their kernels (:26-66, :68-109).  The reference writes a full-size per-element loss tensor, sums it with one
// 128-thread block (math::Sum without scratch), and rescales the gradient tensor in a second pass (math::Scale); here
// loss and gradient come from ONE pass (8.05 B/element: X 4 + dX 4 + labels 4/80), the loss is reduced in two
// deterministic stages, and `accumulate` adds the gradient into an existing d_logits — the autograd Sum of the two
// consumers of retnet_cls_pred_fpnL (caffe2/caffe2/python/core.py:695,792-842) without a separate pass.
// HBM-bound; fp32 arithmetic with one exp, one log and one reciprocal per element (gamma == 2: the RetinaNet
// default RETINANET.LOSS_GAMMA; any other gamma takes two more exp).
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "distill_math.cuh"
#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

constexpr int kFoThreads = 256;

struct FocalArgs {
  const float* X;
  const int32_t* G;
  const float* fg_num;
  const float* d_loss;
  float* dX;
  float* loss;
  float* partials;        // [gridDim.x][1]
  unsigned int* counter;
  uint32_t rows, HW, D, C;  // rows = N * D
  float gamma, alpha, scale;
  int32_t accumulate;
};

// loss summand WITHOUT the 1 / Np factor, and d(loss)/dx without 1 / Np and the upstream gradient
template <bool kGamma2, bool kLoss, bool kGrad>
__device__ __forceinline__ void focal_elem(float x, int t, int d, float gamma, float alpha, float& loss, float& grad) {
  const float e = __expf(-fabsf(x));
  const float u = 1.f + e;
  const float l = log1pf(e);   // accurate for small e (x << 0), where log(1 - p) = -l carries the whole negative-class term
  const float mx = fmaxf(x, 0.f);
  const float logp_raw = (x - mx) - l;                // log p
  const float logp = fmaxf(logp_raw, -87.3365447f);   // log(max(p, FLT_MIN)), sigmoid_focal_loss_op.cu:50,92
  const float log1mp = -(mx + l);                     // -x*[x>=0] - log(1 + exp(x - 2x[x>=0])), :54-55
  const float r = __frcp_rn(u);
  const float p = x >= 0.f ? r : e * r;
  const float q = 1.f - p;
  float qg, pg;  // (1-p)^gamma, p^gamma
  if (kGamma2) {
    qg = q * q;
    pg = p * p;
  } else if (gamma == 0.f) {
    qg = pg = 1.f;
  } else {
    qg = __expf(gamma * log1mp);
    pg = __expf(gamma * logp_raw);
  }
  const bool c1 = t == d + 1;
  const bool c2 = (t != -1) & (t != d + 1);
  if (kLoss) {
    const float term1 = qg * logp, term2 = pg * log1mp;
    loss = -((c1 ? alpha * term1 : 0.f) + (c2 ? (1.f - alpha) * term2 : 0.f));
  }
  if (kGrad) {
    const float term1 = qg * (q - gamma * p * logp);           // (1-p)^g (1 - p - g p log p), :90-92
    const float term2 = pg * (gamma * q * log1mp - p);         // p^g (g (1-p) log(1-p) - p), :94-98
    grad = -((c1 ? alpha * term1 : 0.f) + (c2 ? (1.f - alpha) * term2 : 0.f));
  }
}

template <bool kGamma2, bool kLoss, bool kGrad, bool kVec>
__global__ void __launch_bounds__(kFoThreads) focal_kernel(const FocalArgs a) {
  __shared__ float red_f[kFoThreads / 32];
  __shared__ float blk_sum[1];
  __shared__ bool is_last;
  const float Np = fmaxf(__ldg(a.fg_num), 1.0f);   // max(weight_pos[0], 1.0), :44,84
  const float kg = kGrad ? (a.d_loss ? __ldg(a.d_loss) : 1.f) * a.scale / Np : 0.f;
  float acc = 0.f;
  const uint32_t A = a.D / a.C;
  if (kVec) {
    const uint32_t qpr = a.HW >> 2;  // quads per row
    const uint64_t total = (uint64_t)a.rows * qpr;
    for (uint64_t i = (uint64_t)blockIdx.x * kFoThreads + threadIdx.x; i < total; i += (uint64_t)gridDim.x * kFoThreads) {
      const uint32_t row = (uint32_t)(i / qpr), hq = (uint32_t)(i - (uint64_t)row * qpr);
      const uint32_t n = row / a.D, c = row - n * a.D;
      const uint32_t an = c / a.C, d = c - an * a.C;
      // label index = n*H*W*A + a*H*W + y*W + x (:37-40)
      const int4 t = __ldg(reinterpret_cast<const int4*>(a.G + ((size_t)n * A + an) * a.HW) + hq);
      const float4 x = __ldg(reinterpret_cast<const float4*>(a.X + (size_t)row * a.HW) + hq);
      const float xs[4] = {x.x, x.y, x.z, x.w};
      const int ts[4] = {t.x, t.y, t.z, t.w};
      float g[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float li = 0.f, gi = 0.f;
        focal_elem<kGamma2, kLoss, kGrad>(xs[v], ts[v], (int)d, a.gamma, a.alpha, li, gi);
        if (kLoss) acc += li;
        g[v] = gi * kg;
      }
      if (kGrad) {
        float4* out = reinterpret_cast<float4*>(a.dX + (size_t)row * a.HW) + hq;
        float4 o = make_float4(g[0], g[1], g[2], g[3]);
        if (a.accumulate) {
          const float4 prev = *out;
          o.x += prev.x;
          o.y += prev.y;
          o.z += prev.z;
          o.w += prev.w;
        }
        *out = o;
      }
    }
  } else {
    const uint64_t total = (uint64_t)a.rows * a.HW;
    for (uint64_t i = (uint64_t)blockIdx.x * kFoThreads + threadIdx.x; i < total; i += (uint64_t)gridDim.x * kFoThreads) {
      const uint32_t row = (uint32_t)(i / a.HW), hw = (uint32_t)(i - (uint64_t)row * a.HW);
      const uint32_t n = row / a.D, c = row - n * a.D;
      const uint32_t an = c / a.C, d = c - an * a.C;
      const int t = __ldg(a.G + ((size_t)n * A + an) * a.HW + hw);
      float li = 0.f, gi = 0.f;
      focal_elem<kGamma2, kLoss, kGrad>(__ldg(a.X + i), t, (int)d, a.gamma, a.alpha, li, gi);
      if (kLoss) acc += li;
      if (kGrad) a.dX[i] = a.accumulate ? a.dX[i] + gi * kg : gi * kg;
    }
  }
  if (kLoss) {
    const float s = group_sum<kFoThreads>(acc, red_f, threadIdx.x, 1);
    if (threadIdx.x == 0) blk_sum[0] = s;
    if (publish_and_ticket<1>(blk_sum, a.partials, a.counter, &is_last)) {
      __threadfence();
      if (threadIdx.x < 32) {
        const double t = warp_sum_partials<1>(a.partials, 0, threadIdx.x);
        if (threadIdx.x == 0) {
          a.loss[0] = (float)(t / (double)Np) * a.scale;
          *a.counter = 0u;
        }
      }
    }
  }
}

}  // namespace sad

using namespace sad;

extern "C" {

SAD_EXPORT void sad_focal_default_params(sad_focal_params* p) {
  if (!p) return;
  p->gamma = 1.f;   // sigmoid_focal_loss_op.h:33-36
  p->alpha = 0.25f;
  p->scale = 1.f;
  p->num_classes = 80;
}

SAD_EXPORT size_t sad_focal_workspace_bytes(void) { return 256 + (size_t)kMaxRingCtas * sizeof(float); }

SAD_EXPORT int sad_sigmoid_focal_loss_f32(const float* logits, const int32_t* labels, const float* fg_num, int N, int D, int H, int W,
                                          const sad_focal_params* params, float* loss, const float* d_loss, float* d_logits,
                                          int accumulate_grad, void* workspace, size_t workspace_bytes, void* stream) {
  if (!params) return set_error(SAD_ERR_INVALID, "focal loss: null params");
  if (N < 0 || D < 0 || H < 0 || W < 0) return set_error(SAD_ERR_INVALID, "focal loss: negative dimension");
  if (params->num_classes < 1 || D % params->num_classes) return set_error(SAD_ERR_INVALID, "focal loss: D must be a multiple of num_classes");
  if (!(params->scale >= 0.f)) return set_error(SAD_ERR_INVALID, "focal loss: scale must be >= 0");   // CAFFE_ENFORCE(scale_ >= 0)
  if (!loss && !d_logits) return set_error(SAD_ERR_INVALID, "focal loss: neither loss nor d_logits requested");
  if (!fg_num) return set_error(SAD_ERR_INVALID, "focal loss: null normaliser");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t total = (uint64_t)N * D * H * W;
  if (total == 0) {
    if (loss) return check_cuda(cudaMemsetAsync(loss, 0, sizeof(float), st), "focal loss memset");
    return SAD_OK;
  }
  if (!logits || !labels) return set_error(SAD_ERR_INVALID, "focal loss: null tensor");
  if ((uint64_t)N * D > 0xffffffffull || (uint64_t)H * W > 0xffffffffull) return set_error(SAD_ERR_INVALID, "focal loss: tensor too large");
  FocalArgs a{};
  a.X = logits;
  a.G = labels;
  a.fg_num = fg_num;
  a.d_loss = d_loss;
  a.dX = d_logits;
  a.loss = loss;
  a.rows = (uint32_t)((uint64_t)N * D);
  a.HW = (uint32_t)((uint64_t)H * W);
  a.D = (uint32_t)D;
  a.C = (uint32_t)params->num_classes;
  a.gamma = params->gamma;
  a.alpha = params->alpha;
  a.scale = params->scale;
  a.accumulate = accumulate_grad;
  if (loss) {
    if (!workspace || workspace_bytes < sad_focal_workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 255))
      return set_error(SAD_ERR_WORKSPACE, "focal loss: workspace must be 256-byte aligned and >= sad_focal_workspace_bytes()");
    a.counter = static_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  }
  int dev = 0, sms = 0, rc;
  if ((rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) != SAD_OK) return rc;
  if ((rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) != SAD_OK) return rc;
  const bool vec = (a.HW % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(d_logits)) & 15) == 0);
  const uint64_t work = vec ? total / 4 : total;
  uint64_t blocks = (work + kFoThreads - 1) / kFoThreads;
  const uint64_t cap = (uint64_t)sms * 8 < (uint64_t)kMaxRingCtas ? (uint64_t)sms * 8 : (uint64_t)kMaxRingCtas;
  if (blocks > cap) blocks = cap;
  const bool g2 = params->gamma == 2.0f;
  const bool want_loss = loss != nullptr, want_grad = d_logits != nullptr;
#define SAD_FOCAL_LAUNCH(G2, LS, GR, V) focal_kernel<G2, LS, GR, V><<<(unsigned)blocks, kFoThreads, 0, st>>>(a)
#define SAD_FOCAL_OUT(G2, V)                                        \
  do {                                                              \
    if (want_loss && want_grad) SAD_FOCAL_LAUNCH(G2, true, true, V); \
    else if (want_loss) SAD_FOCAL_LAUNCH(G2, true, false, V);        \
    else SAD_FOCAL_LAUNCH(G2, false, true, V);                       \
  } while (0)
  if (g2 && vec) SAD_FOCAL_OUT(true, true);
  else if (g2) SAD_FOCAL_OUT(true, false);
  else if (vec) SAD_FOCAL_OUT(false, true);
  else SAD_FOCAL_OUT(false, false);
#undef SAD_FOCAL_OUT
#undef SAD_FOCAL_LAUNCH
  count_launch(1);
  return check_cuda(cudaGetLastError(), "focal loss launch");
}

}  // extern "C"
