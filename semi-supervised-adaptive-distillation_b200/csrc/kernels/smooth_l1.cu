// SelectSmoothL1Loss (+Gradient) — RetinaNet's box-regression loss over the M foreground anchors of one FPN level
// (retinanet_heads.py:258-272).  Replaces
//   SelectSmoothL1LossOp<float, CUDAContext>::RunOnDevice          caffe2/modules/detectron/select_smooth_l1_loss_op.cu:90-143
//   SelectSmoothL1LossGradientOp<float, CUDAContext>::RunOnDevice  caffe2/modules/detectron/select_smooth_l1_loss_op.cu:145-181
// and their kernels (:23-54, :57-86).  Inputs as the reference: Y_hat (N, D = A*4, H, W) predicted deltas, Y (M, 4)
// targets, L (M, 4) FLOAT locations {image n, first channel c, y, x}, S the foreground count.
// The reference fills a tensor as large as Y_hat with zeros, scatters 4*M terms into it and sums ALL of it with one
// 128-thread block; here the 4*M terms are summed directly (two deterministic stages).  The gradient is a zero fill of
// d_Y_hat plus a 4*M-element scatter, as in the reference.  Tiny, latency-bound work: no roofline claim.
#include <cuda_runtime.h>
#include <stdint.h>

#include "distill_math.cuh"
#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

constexpr int kSlThreads = 256;

struct SmoothL1Args {
  const float* y_hat;
  const float* y;
  const float* locs;
  const float* fg_num;
  const float* d_loss;
  float* d_y_hat;
  float* loss;
  float* partials;
  unsigned int* counter;
  int32_t D, H, W, M;
  float beta, scale;
};

template <bool kLoss, bool kGrad>
__global__ void __launch_bounds__(kSlThreads) select_smooth_l1_kernel(const SmoothL1Args a) {
  __shared__ float red_f[kSlThreads / 32];
  __shared__ float blk_sum[1];
  __shared__ bool is_last;
  const float S = fmaxf(__ldg(a.fg_num), 1.0f);                                   // max(S[0], 1.0), :43,81
  const float gk = kGrad ? a.scale * (a.d_loss ? __ldg(a.d_loss) : 1.f) : 0.f;     // norm * d_loss, :78
  float acc = 0.f;
  const int total = a.M * 4;
  for (int e = blockIdx.x * kSlThreads + threadIdx.x; e < total; e += gridDim.x * kSlThreads) {
    const int i = e >> 2, j = e & 3;
    const int n = (int)__ldg(a.locs + i * 4), c = (int)__ldg(a.locs + i * 4 + 1);   // float -> int truncation, :31-34
    const int y = (int)__ldg(a.locs + i * 4 + 2), x = (int)__ldg(a.locs + i * 4 + 3);
    const size_t ind = (size_t)n * a.D * a.H * a.W + (size_t)(c + j) * a.H * a.W + (size_t)y * a.W + x;   // :38
    const float val = __ldg(a.y_hat + ind) - __ldg(a.y + e);
    const float av = fabsf(val);
    if (kLoss) acc += (av < a.beta ? 0.5f * val * val / a.beta : av - 0.5f * a.beta) / S;   // :43-47
    if (kGrad) a.d_y_hat[ind] = av < a.beta ? gk * val / a.beta / S : gk * (float)((0.f < val) - (val < 0.f)) / S;   // :80-84
  }
  if (kLoss) {
    const float s = group_sum<kSlThreads>(acc, red_f, threadIdx.x, 1);
    if (threadIdx.x == 0) blk_sum[0] = s;
    if (publish_and_ticket<1>(blk_sum, a.partials, a.counter, &is_last)) {
      __threadfence();
      if (threadIdx.x < 32) {
        const double t = warp_sum_partials<1>(a.partials, 0, threadIdx.x);
        if (threadIdx.x == 0) {
          a.loss[0] = (float)t * a.scale;   // math::Scale(1, scale_, ...), :140-141
          *a.counter = 0u;
        }
      }
    }
  }
}

}  // namespace sad

using namespace sad;

extern "C" {

SAD_EXPORT size_t sad_smooth_l1_workspace_bytes(void) { return 256 + (size_t)kMaxRingCtas * sizeof(float); }

SAD_EXPORT int sad_select_smooth_l1_loss_f32(const float* y_hat, const float* y, const float* locs, const float* fg_num, int N, int D, int H,
                                             int W, int M, float beta, float scale, float* loss, const float* d_loss, float* d_y_hat,
                                             void* workspace, size_t workspace_bytes, void* stream) {
  if (N < 0 || D < 0 || H < 0 || W < 0 || M < 0) return set_error(SAD_ERR_INVALID, "smooth l1: negative dimension");
  if (!(beta > 0.f)) return set_error(SAD_ERR_INVALID, "smooth l1: beta must be > 0");     // CAFFE_ENFORCE(beta_ > 0)
  if (!(scale >= 0.f)) return set_error(SAD_ERR_INVALID, "smooth l1: scale must be >= 0");  // CAFFE_ENFORCE(scale_ >= 0)
  if (!loss && !d_y_hat) return set_error(SAD_ERR_INVALID, "smooth l1: neither loss nor gradient requested");
  if (!fg_num) return set_error(SAD_ERR_INVALID, "smooth l1: null normaliser");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  const size_t total = (size_t)N * D * H * W;
  // the gradient of every position that is not a foreground anchor is zero (select_smooth_l1_loss_op.cu:155-157)
  if (d_y_hat && total && (rc = check_cuda(cudaMemsetAsync(d_y_hat, 0, total * sizeof(float), st), "smooth l1: zero gradient")) != SAD_OK) return rc;
  if (M == 0) {   // no foreground box on this level: loss 0 (:100-104), gradient all zero
    if (loss) return check_cuda(cudaMemsetAsync(loss, 0, sizeof(float), st), "smooth l1: zero loss");
    return SAD_OK;
  }
  if (!y_hat || !y || !locs) return set_error(SAD_ERR_INVALID, "smooth l1: null tensor");
  if ((size_t)M * 4 > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "smooth l1: too many boxes");
  SmoothL1Args a{};
  a.y_hat = y_hat;
  a.y = y;
  a.locs = locs;
  a.fg_num = fg_num;
  a.d_loss = d_loss;
  a.d_y_hat = d_y_hat;
  a.loss = loss;
  a.D = D;
  a.H = H;
  a.W = W;
  a.M = M;
  a.beta = beta;
  a.scale = scale;
  if (loss) {
    if (!workspace || workspace_bytes < sad_smooth_l1_workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 255))
      return set_error(SAD_ERR_WORKSPACE, "smooth l1: workspace must be 256-byte aligned and >= sad_smooth_l1_workspace_bytes()");
    a.counter = static_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  }
  unsigned blocks = (unsigned)(((size_t)M * 4 + kSlThreads - 1) / kSlThreads);
  if (blocks > (unsigned)kMaxRingCtas) blocks = kMaxRingCtas;
  if (loss && d_y_hat) select_smooth_l1_kernel<true, true><<<blocks, kSlThreads, 0, st>>>(a);
  else if (loss) select_smooth_l1_kernel<true, false><<<blocks, kSlThreads, 0, st>>>(a);
  else select_smooth_l1_kernel<false, true><<<blocks, kSlThreads, 0, st>>>(a);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "smooth l1 launch");
}

}  // extern "C"
