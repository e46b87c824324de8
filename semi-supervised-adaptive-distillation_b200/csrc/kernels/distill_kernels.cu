// sm_100a kernels for the adaptive-distillation loss path + their C-ABI launchers.
//
//   distill_kernel   fused SigmoidAdaptiveDistillLoss (+Gradient) over up to 8 FPN levels in ONE
//                    launch: reads logits X, teacher probs T (and the int32 labels once per
//                    (n, anchor) row, reused for all classes), writes dX and/or the per-level loss.
//                    Replaces reference kernels sigmoid_adaptive_distillation_loss_op.cu:28-67 and
//                    :69-105 plus math::Sum (single 128-thread block, math_gpu.cu:1021-1058) and
//                    the extra full-tensor math::Scale pass (:167-168).
//   pow_sum_kernel   sum_k sum_j in_k[j]^power in ONE launch, no temporary; replaces
//                    pow_sum_op.cu:25-43 (1 + 3*n_inputs launches, a full temp write + re-read).
//
// Both are HBM-bound streaming kernels: 128-bit coalesced loads/stores, several independent
// 16-byte loads in flight per thread, fp32-only arithmetic with 4 MUFU ops per element on the
// gamma=2/beta=0 path (the reference mixes in FP64 divisions), warp-shuffle
// + shared-memory block reduction, and a deterministic "last CTA finishes" second stage in fp64.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "distill_math.cuh"
#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

// -------------------------------------------------------------------------------------------
// fused multi-level loss + gradient
// -------------------------------------------------------------------------------------------
// Work decomposition.  For one level, X viewed as [NA = N*A][C][HW]: the label of element
// (na, c, hw) is G[na*HW + hw] for every class c (…loss_op.cu:35-42 with a = c / num_classes).
// A "row" is one (na, hw-vector) pair (kVec consecutive hw positions); a tile is kThreads rows x
// kClassesPerTile classes.  Each thread owns one row: it loads its label vector ONCE, turns it
// into a keep-mask held in registers, then walks its classes with stride HW, kUnroll planes at a
// time (2*kUnroll independent 16-byte loads in flight per thread).
constexpr int kThreads = 128;
constexpr int kClassesPerTile = 8;
constexpr int kUnroll = 4;

struct LevelDesc {
  const float* X;
  const float* T;
  const int32_t* G;
  float* dX;
  float* loss;
  const float* d_loss;
  uint32_t HW;          // H*W
  uint32_t hw_vecs;     // HW / kVec
  uint32_t rows;        // NA * hw_vecs
  uint32_t tile_begin;  // first tile id of this level
};
struct DistillArgs {
  LevelDesc lv[SAD_MAX_LEVELS];
  uint32_t tile_end[SAD_MAX_LEVELS];
  int32_t n_levels;
  int32_t num_classes;
  int32_t class_groups;  // ceil(num_classes / kClassesPerTile)
  int32_t ignored_label;
  float gamma, alpha, beta, scale;
  const float* normalizer;
  float* partials;         // one float per tile
  unsigned int* counter;   // zero before launch; reset by the last CTA
};

template <int kVec, bool kFast, bool kLoss, bool kGrad>
__global__ void __launch_bounds__(kThreads, 8)
distill_kernel(const __grid_constant__ DistillArgs args) {
  __shared__ float red_f[kThreads / 32];
  __shared__ double red_d[kThreads / 32];
  __shared__ bool is_last;

  const uint32_t tile = blockIdx.x;
  int l = 0;
#pragma unroll 1
  while (l + 1 < args.n_levels && tile >= args.tile_end[l]) ++l;
  const LevelDesc& lv = args.lv[l];

  const uint32_t local = tile - lv.tile_begin;
  const uint32_t row_tile = local / (uint32_t)args.class_groups;
  const uint32_t cgrp = local - row_tile * (uint32_t)args.class_groups;
  const uint32_t row = row_tile * kThreads + threadIdx.x;
  const bool valid = row < lv.rows;

  const float Np = fmaxf(__ldg(args.normalizer), 1.0f);
  const float alpha = args.alpha, gamma = args.gamma, beta = args.beta;
  const float one_m_alpha = 1.f - alpha, one_m_2alpha = 1.f - 2.f * alpha;

  float keep[kVec];
  size_t base = 0;
  if (valid) {
    const uint32_t na = row / lv.hw_vecs;
    const uint32_t hv = row - na * lv.hw_vecs;
    const size_t goff = (size_t)na * lv.HW + (size_t)hv * kVec;
    if constexpr (kVec == 4) {
      const int4 t = __ldg(reinterpret_cast<const int4*>(lv.G + goff));
      keep[0] = t.x != args.ignored_label ? 1.f : 0.f;
      keep[1] = t.y != args.ignored_label ? 1.f : 0.f;
      keep[2] = t.z != args.ignored_label ? 1.f : 0.f;
      keep[3] = t.w != args.ignored_label ? 1.f : 0.f;
    } else {
      keep[0] = __ldg(lv.G + goff) != args.ignored_label ? 1.f : 0.f;
    }
    base = ((size_t)na * args.num_classes + (size_t)cgrp * kClassesPerTile) * lv.HW + (size_t)hv * kVec;
  } else {
#pragma unroll
    for (int v = 0; v < kVec; ++v) keep[v] = 0.f;
  }
  int n_cls = args.num_classes - (int)cgrp * kClassesPerTile;
  n_cls = n_cls < kClassesPerTile ? n_cls : kClassesPerTile;
  if (!valid) n_cls = 0;

  float kg = 0.f;
  if (kGrad) {
    const float dl = lv.d_loss ? __ldg(lv.d_loss) : 1.f;
    kg = dl * args.scale / Np;
  }
  float acc = 0.f;

  const float* __restrict__ Xp = lv.X + base;
  const float* __restrict__ Tp = lv.T + base;
  float* __restrict__ dXp = kGrad ? lv.dX + base : nullptr;
  const size_t plane = lv.HW;

#pragma unroll 1
  for (int c0 = 0; c0 < n_cls; c0 += kUnroll) {
    float xv[kUnroll][kVec], tv[kUnroll][kVec];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (c0 + u < n_cls) {
        if constexpr (kVec == 4) {
          const float4 a = ld_stream4(Xp + (size_t)(c0 + u) * plane);
          const float4 b = ld_stream4(Tp + (size_t)(c0 + u) * plane);
          xv[u][0] = a.x; xv[u][1] = a.y; xv[u][2] = a.z; xv[u][3] = a.w;
          tv[u][0] = b.x; tv[u][1] = b.y; tv[u][2] = b.z; tv[u][3] = b.w;
        } else {
          xv[u][0] = ld_stream1(Xp + (size_t)(c0 + u) * plane);
          tv[u][0] = ld_stream1(Tp + (size_t)(c0 + u) * plane);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (c0 + u < n_cls) {
        float gv[kVec];
#pragma unroll
        for (int v = 0; v < kVec; ++v) {
          float li = 0.f, g = 0.f;
          distill_elem<kFast, kLoss, kGrad>(xv[u][v], tv[u][v], gamma, alpha, beta, one_m_alpha, one_m_2alpha, li, g);
          if (kLoss) acc = fmaf(li, keep[v], acc);   // NaN * 0 stays NaN, like the reference's `* (t != ignored)`
          if (kGrad) gv[v] = g * (keep[v] * kg);
        }
        if (kGrad) {
          if constexpr (kVec == 4) {
            *reinterpret_cast<float4*>(dXp + (size_t)(c0 + u) * plane) = make_float4(gv[0], gv[1], gv[2], gv[3]);
          } else {
            dXp[(size_t)(c0 + u) * plane] = gv[0];
          }
        }
      }
    }
  }

  if (kLoss) {
    const float tile_sum = block_sum<kThreads>(acc, red_f);
    if (threadIdx.x == 0) {
      args.partials[tile] = tile_sum;
      __threadfence();
      const unsigned int ticket = atomicAdd(args.counter, 1u);
      is_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
      // second stage: fixed order (tile index), fp64, independent of CTA scheduling
      __threadfence();
      for (int k = 0; k < args.n_levels; ++k) {
        double s = 0.0;
        for (uint32_t t = args.lv[k].tile_begin + threadIdx.x; t < args.tile_end[k]; t += kThreads)
          s += (double)__ldcg(args.partials + t);
        s = block_sum<kThreads>(s, red_d);
        if (threadIdx.x == 0) args.lv[k].loss[0] = (float)(s / (double)Np) * args.scale;
      }
      if (threadIdx.x == 0) *args.counter = 0u;
    }
  }
}

// -------------------------------------------------------------------------------------------
// PowSum
// -------------------------------------------------------------------------------------------
constexpr int kPsThreads = 256;
constexpr int kPsUnroll = 8;                                  // float4 loads in flight per thread
constexpr int kPsChunk = kPsThreads * 4 * kPsUnroll;          // elements per CTA (8192 = 32 KB)

enum PowMode { kPowGeneric = 0, kPowOne = 1, kPowTwo = 2, kPowAccurate = 3 };

struct PowSumArgs {
  const float* in[SAD_MAX_INPUTS];
  int64_t n[SAD_MAX_INPUTS];
  uint32_t cta_begin[SAD_MAX_INPUTS];
  uint32_t cta_end[SAD_MAX_INPUTS];
  int32_t n_inputs;
  float power;
  float* partials;        // one float per CTA
  unsigned int* counter;
  float* out;
};

template <int kMode>
__device__ __forceinline__ float pow_elem(float x, float power) {
  if (kMode == kPowOne) return x;
  if (kMode == kPowTwo) return x * x;
  if (kMode == kPowAccurate) return powf(x, power);
  return ex2_approx(power * lg2_approx(x));  // x^p for x >= 0; NaN for x < 0 like powf with non-integer p
}

template <int kMode>
__global__ void __launch_bounds__(kPsThreads, 4) pow_sum_kernel(const __grid_constant__ PowSumArgs args) {
  __shared__ float red_f[kPsThreads / 32];
  __shared__ double red_d[kPsThreads / 32];
  __shared__ bool is_last;

  const uint32_t cta = blockIdx.x;
  int k = 0;
#pragma unroll 1
  while (k + 1 < args.n_inputs && cta >= args.cta_end[k]) ++k;
  const float* __restrict__ in = args.in[k];
  const int64_t n = args.n[k];
  const int64_t start = (int64_t)(cta - args.cta_begin[k]) * kPsChunk;
  const float power = args.power;

  float acc = 0.f;
  const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (aligned && start + kPsChunk <= n) {
    float4 v[kPsUnroll];
#pragma unroll
    for (int u = 0; u < kPsUnroll; ++u) v[u] = ld_stream4(in + start + ((int64_t)u * kPsThreads + threadIdx.x) * 4);
#pragma unroll
    for (int u = 0; u < kPsUnroll; ++u) {
      acc += pow_elem<kMode>(v[u].x, power) + pow_elem<kMode>(v[u].y, power);
      acc += pow_elem<kMode>(v[u].z, power) + pow_elem<kMode>(v[u].w, power);
    }
  } else {
    const int64_t stop = start + kPsChunk < n ? start + kPsChunk : n;
    for (int64_t i = start + threadIdx.x; i < stop; i += kPsThreads) acc += pow_elem<kMode>(ld_stream1(in + i), power);
  }

  const float cta_sum = block_sum<kPsThreads>(acc, red_f);
  if (threadIdx.x == 0) {
    args.partials[cta] = cta_sum;
    __threadfence();
    const unsigned int ticket = atomicAdd(args.counter, 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // per input: fp64 sum of its CTA partials in index order, rounded to float; then the
    // reference's running float add over inputs (pow_sum_op.cu:39)
    float res = 0.f;
    for (int j = 0; j < args.n_inputs; ++j) {
      double s = 0.0;
      for (uint32_t t = args.cta_begin[j] + threadIdx.x; t < args.cta_end[j]; t += kPsThreads) s += (double)__ldcg(args.partials + t);
      s = block_sum<kPsThreads>(s, red_d);
      res = res + (float)s;
    }
    if (threadIdx.x == 0) {
      args.out[0] = res;
      *args.counter = 0u;
    }
  }
}

}  // namespace sad

// ===========================================================================================
// C ABI
// ===========================================================================================
using namespace sad;

namespace {
thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace sad {
int set_error(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SAD_OK;
  return set_error(SAD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sad

extern "C" {

SAD_EXPORT const char* sad_last_error(void) { return g_error.c_str(); }
SAD_EXPORT const char* sad_version(void) { return "sad_b200 0.1.0 sm_100a"; }
SAD_EXPORT uint64_t sad_launch_count(void) { return g_launches.load(); }

SAD_EXPORT void sad_distill_default_params(sad_distill_params* p) {
  if (!p) return;
  p->gamma = 1.0f;
  p->alpha = 0.25f;
  p->beta = 0.0f;
  p->scale = 1.0f;
  p->num_classes = 80;
  p->ignored_label = -1;
}

// The first 256 bytes of a workspace hold the "CTAs finished" ticket counter.  It must be zero when a
// kernel starts; the last CTA of every launch puts it back to zero, so a workspace is initialised
// ONCE after allocation and then reused launch after launch (also inside CUDA graphs) by one op
// instance on one stream at a time, like the reference's per-op member scratch tensors.
SAD_EXPORT int sad_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || workspace_bytes < 256 || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(SAD_ERR_WORKSPACE, "workspace must be 256-byte aligned and at least 256 bytes");
  // the whole workspace: control words sit at its start and, for the one-launch step, in a block behind the two-launch scratch
  return check_cuda(cudaMemsetAsync(workspace, 0, workspace_bytes, static_cast<cudaStream_t>(stream)), "workspace init");
}

// ---- PowSum -------------------------------------------------------------------------------
static int pow_sum_plan(const int64_t* sizes, int n_inputs, uint32_t* begin, uint32_t* end, uint32_t* total) {
  if (!sizes || n_inputs < 1 || n_inputs > SAD_MAX_INPUTS)
    return set_error(SAD_ERR_INVALID, "PowSum: n_inputs must be in [1, " + std::to_string(SAD_MAX_INPUTS) + "]");
  uint64_t t = 0;
  for (int k = 0; k < n_inputs; ++k) {
    if (sizes[k] < 0) return set_error(SAD_ERR_INVALID, "PowSum: negative input size");
    begin[k] = (uint32_t)t;
    t += (uint64_t)((sizes[k] + kPsChunk - 1) / kPsChunk);
    end[k] = (uint32_t)t;
  }
  if (t > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "PowSum: inputs too large for one launch");
  *total = (uint32_t)t;
  return SAD_OK;
}

SAD_EXPORT size_t sad_pow_sum_workspace_bytes(const int64_t* sizes, int n_inputs) {
  uint32_t b[SAD_MAX_INPUTS], e[SAD_MAX_INPUTS], total = 0;
  if (pow_sum_plan(sizes, n_inputs, b, e, &total) != SAD_OK) return 0;
  const size_t simt = (size_t)(total ? total : 1) * sizeof(float);
  const size_t ring = (size_t)kMaxRingCtas * SAD_MAX_INPUTS * sizeof(float);
  return 256 + (simt > ring ? simt : ring);
}

SAD_EXPORT int sad_pow_sum_f32(const float* const* inputs, const int64_t* sizes, int n_inputs, float power,
                               float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!inputs || !out) return set_error(SAD_ERR_INVALID, "PowSum: null inputs/out");
  PowSumArgs a{};
  uint32_t total = 0;
  int rc = pow_sum_plan(sizes, n_inputs, a.cta_begin, a.cta_end, &total);
  if (rc != SAD_OK) return rc;
  for (int k = 0; k < n_inputs; ++k) {
    if (sizes[k] > 0 && !inputs[k]) return set_error(SAD_ERR_INVALID, "PowSum: null input pointer");
    a.in[k] = inputs[k];
    a.n[k] = sizes[k];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (total == 0) {  // every input empty: the reference's running sum stays 0
    return check_cuda(cudaMemsetAsync(out, 0, sizeof(float), st), "PowSum memset");
  }
  if (pow_sum_ring_supported(inputs, sizes, n_inputs))  // 16-byte aligned inputs: persistent TMA ring
    return launch_pow_sum_ring(inputs, sizes, n_inputs, power, out, workspace, workspace_bytes, st);
  const size_t need = 256 + (size_t)total * sizeof(float);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(SAD_ERR_WORKSPACE, "PowSum: workspace must be 256-byte aligned and >= sad_pow_sum_workspace_bytes()");
  a.n_inputs = n_inputs;
  a.power = power;
  a.counter = static_cast<unsigned int*>(workspace);
  a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  a.out = out;

  int mode = kPowGeneric;
  if (power == 1.0f) mode = kPowOne;
  else if (power == 2.0f) mode = kPowTwo;
  else if (power == floorf(power) || !(power > 0.f)) mode = kPowAccurate;  // integer / non-positive exponents: full powf semantics
  switch (mode) {
    case kPowOne: pow_sum_kernel<kPowOne><<<total, kPsThreads, 0, st>>>(a); break;
    case kPowTwo: pow_sum_kernel<kPowTwo><<<total, kPsThreads, 0, st>>>(a); break;
    case kPowAccurate: pow_sum_kernel<kPowAccurate><<<total, kPsThreads, 0, st>>>(a); break;
    default: pow_sum_kernel<kPowGeneric><<<total, kPsThreads, 0, st>>>(a); break;
  }
  count_launch(1);
  return check_cuda(cudaGetLastError(), "PowSum launch");
}

// ---- distillation loss ---------------------------------------------------------------------
static int distill_plan(const sad_distill_level* levels, int n_levels, int num_classes, int vec, DistillArgs* a,
                        uint32_t* total_tiles) {
  const uint32_t cgroups = (uint32_t)((num_classes + kClassesPerTile - 1) / kClassesPerTile);
  uint64_t t = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    const uint64_t HW = (uint64_t)L.H * L.W;
    const uint64_t NA = (uint64_t)L.N * (L.D / num_classes);
    const uint64_t hw_vecs = HW / vec;
    const uint64_t rows = NA * hw_vecs;
    if (HW > 0xffffffffull || rows > 0xffffffffull) return set_error(SAD_ERR_INVALID, "distill: level too large");
    if (a) {
      a->lv[l].HW = (uint32_t)HW;
      a->lv[l].hw_vecs = (uint32_t)(hw_vecs ? hw_vecs : 1);
      a->lv[l].rows = (uint32_t)rows;
      a->lv[l].tile_begin = (uint32_t)t;
    }
    t += (rows + kThreads - 1) / kThreads * cgroups;
    if (a) a->tile_end[l] = (uint32_t)t;
    if (t > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "distill: too many tiles for one launch");
  }
  *total_tiles = (uint32_t)t;
  return SAD_OK;
}

static int distill_validate(const sad_distill_level* levels, int n_levels, int num_classes) {
  if (!levels || n_levels < 1 || n_levels > SAD_MAX_LEVELS)
    return set_error(SAD_ERR_INVALID, "distill: n_levels must be in [1, " + std::to_string(SAD_MAX_LEVELS) + "]");
  if (num_classes < 1) return set_error(SAD_ERR_INVALID, "distill: num_classes must be >= 1");
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    if (L.N < 0 || L.D < 0 || L.H < 0 || L.W < 0) return set_error(SAD_ERR_INVALID, "distill: negative dimension");
    if (L.D % num_classes != 0)
      return set_error(SAD_ERR_INVALID, "distill: channel dim D=" + std::to_string(L.D) + " is not a multiple of num_classes=" + std::to_string(num_classes));
  }
  return SAD_OK;
}

SAD_EXPORT size_t sad_distill_workspace_bytes(const sad_distill_level* levels, int n_levels) {
  // class-group count depends on num_classes; size for the worst case (1 class per group)
  if (!levels || n_levels < 1 || n_levels > SAD_MAX_LEVELS) return 0;
  uint64_t tiles = 0;
  for (int l = 0; l < n_levels; ++l) {
    const uint64_t elems = (uint64_t)levels[l].N * levels[l].D * levels[l].H * levels[l].W;
    // a tile covers >= 1 class x up to kThreads rows; rows*classes <= elems, plus ragged last tiles
    tiles += elems / kThreads + (uint64_t)levels[l].D + 1;
  }
  const size_t simt = (size_t)tiles * sizeof(float);
  const size_t ring = (size_t)kMaxRingCtas * SAD_MAX_LEVELS * sizeof(float);
  return 256 + (simt > ring ? simt : ring);
}

SAD_EXPORT int sad_distill_f32(const sad_distill_level* levels, int n_levels, const float* normalizer,
                               const sad_distill_params* params, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (!params) return set_error(SAD_ERR_INVALID, "distill: null params");
  if (!(params->scale >= 0.f)) return set_error(SAD_ERR_INVALID, "distill: scale must be >= 0 (reference: CAFFE_ENFORCE(scale_ >= 0))");
  int rc = distill_validate(levels, n_levels, params->num_classes);
  if (rc != SAD_OK) return rc;
  if (!normalizer) return set_error(SAD_ERR_INVALID, "distill: null normalizer");

  const bool want_loss = levels[0].loss != nullptr, want_grad = levels[0].d_logits != nullptr;
  if (!want_loss && !want_grad) return set_error(SAD_ERR_INVALID, "distill: neither loss nor d_logits requested");
  int vec = 4;
  uint64_t total_elems = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    const uint64_t elems = (uint64_t)L.N * L.D * L.H * L.W;
    total_elems += elems;
    if ((L.loss != nullptr) != want_loss || (L.d_logits != nullptr) != want_grad)
      return set_error(SAD_ERR_INVALID, "distill: all levels of one call must request the same outputs");
    if (elems && (!L.logits || !L.teacher_prob || !L.labels)) return set_error(SAD_ERR_INVALID, "distill: null input pointer");
    const uintptr_t bits = reinterpret_cast<uintptr_t>(L.logits) | reinterpret_cast<uintptr_t>(L.teacher_prob) |
                           reinterpret_cast<uintptr_t>(L.labels) | reinterpret_cast<uintptr_t>(L.d_logits);
    if (((uint64_t)L.H * L.W) % 4 != 0 || (bits & 15)) vec = 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (total_elems == 0) {  // empty tensors: loss 0, nothing to write
    if (want_loss)
      for (int l = 0; l < n_levels; ++l) {
        rc = check_cuda(cudaMemsetAsync(levels[l].loss, 0, sizeof(float), st), "distill memset");
        if (rc != SAD_OK) return rc;
      }
    return SAD_OK;
  }

  if (vec == 4 && distill_ring_supported(levels, n_levels, params->num_classes))  // persistent TMA ring
    return launch_distill_ring(levels, n_levels, normalizer, params, workspace, workspace_bytes, st);
  if (vec == 1 && n_levels > 1) {
    // Dispatch per level: one level with H*W % 4 != 0 (P7 = 5 x 7 of a 640 x 896 input) must not push the other levels — 99.9 %
    // of the elements — off the ring.  Ring-capable levels go to the ring kernel, the rest to the scalar SIMT kernel below:
    // two launches on the same stream sharing the workspace one after the other.
    sad_distill_level ring_lv[SAD_MAX_LEVELS], rest_lv[SAD_MAX_LEVELS];
    int n_ring = 0, n_rest = 0;
    for (int l = 0; l < n_levels; ++l) {
      if (distill_ring_supported(levels + l, 1, params->num_classes)) ring_lv[n_ring++] = levels[l];
      else rest_lv[n_rest++] = levels[l];
    }
    if (n_ring > 0 && n_rest > 0) {
      rc = launch_distill_ring(ring_lv, n_ring, normalizer, params, workspace, workspace_bytes, st);
      if (rc != SAD_OK) return rc;
      return sad_distill_f32(rest_lv, n_rest, normalizer, params, workspace, workspace_bytes, stream);
    }
  }

  DistillArgs a{};
  uint32_t tiles = 0;
  rc = distill_plan(levels, n_levels, params->num_classes, vec, &a, &tiles);
  if (rc != SAD_OK) return rc;
  for (int l = 0; l < n_levels; ++l) {
    a.lv[l].X = levels[l].logits;
    a.lv[l].T = levels[l].teacher_prob;
    a.lv[l].G = levels[l].labels;
    a.lv[l].dX = levels[l].d_logits;
    a.lv[l].loss = levels[l].loss;
    a.lv[l].d_loss = levels[l].d_loss;
  }
  a.n_levels = n_levels;
  a.num_classes = params->num_classes;
  a.class_groups = (params->num_classes + kClassesPerTile - 1) / kClassesPerTile;
  a.ignored_label = params->ignored_label;
  a.gamma = params->gamma;
  a.alpha = params->alpha;
  a.beta = params->beta;
  a.scale = params->scale;
  a.normalizer = normalizer;
  if (want_loss) {
    const size_t need = 256 + (size_t)tiles * sizeof(float);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
      return set_error(SAD_ERR_WORKSPACE, "distill: workspace must be 256-byte aligned and >= sad_distill_workspace_bytes()");
    a.counter = static_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  }
  if (tiles == 0) return SAD_OK;

  const bool fast = params->gamma == 2.0f && params->beta == 0.0f;
#define SAD_LAUNCH(V, F, LS, GR) distill_kernel<V, F, LS, GR><<<tiles, kThreads, 0, st>>>(a)
#define SAD_DISPATCH_OUT(V, F)                                       \
  do {                                                               \
    if (want_loss && want_grad) SAD_LAUNCH(V, F, true, true);        \
    else if (want_loss) SAD_LAUNCH(V, F, true, false);               \
    else SAD_LAUNCH(V, F, false, true);                              \
  } while (0)
  if (vec == 4) {
    if (fast) SAD_DISPATCH_OUT(4, true); else SAD_DISPATCH_OUT(4, false);
  } else {
    if (fast) SAD_DISPATCH_OUT(1, true); else SAD_DISPATCH_OUT(1, false);
  }
#undef SAD_DISPATCH_OUT
#undef SAD_LAUNCH
  count_launch(1);
  return check_cuda(cudaGetLastError(), "distill launch");
}

}  // extern "C"
