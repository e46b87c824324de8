// Persistent bulk-copy-ring kernels (sm_100a): the production path of the two HBM-bound operators.
//
//   distill_ring_kernel   fused SigmoidAdaptiveDistillLoss (+Gradient), all FPN levels in one launch
//   pow_sum_ring_kernel   PowSum over all inputs in one launch
//
// Why a ring: the first (SIMT, one-tile-per-CTA) kernels in distill_kernels.cu left every SM idle for a
// third of the launch (ncu profiles/r01a_*: 4.06 / 2.70 waves, sm__cycles_active 66 % / 61 % of
// elapsed) because the last partial wave costs a full CTA lifetime.  Here the grid is exactly
// 2 CTAs x #SMs, every CTA owns an equal contiguous range of fixed-size work units, and one
// producer thread per CTA streams the units through a kStages-deep shared-memory ring with TMA 1-D
// bulk copies (cp.async.bulk + mbarrier complete_tx), so tens of KB stay in flight per SM without
// occupying registers and all CTAs finish together.  8 consumer warps per CTA do the arithmetic from
// shared memory (conflict-free LDS.128) and write the gradient with coalesced 128-bit stores.
//
// Unit of work (distill): one (image*anchor, hw-tile of <= 512 positions, group of <= 8 classes):
// <= 8 rows of X, <= 8 rows of T (<= 2 KB each, contiguous in NCHW) and the <= 2 KB label row they
// share (...loss_op.cu:35-42: label index = n*H*W*A + a*H*W + y*W + x for every class of anchor a).
// Requires HW % 4 == 0 and 16-byte aligned tensors (bulk copies move multiples of 16 bytes); other
// shapes take the SIMT kernels.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "distill_math.cuh"
#include "ring.cuh"
#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

// ---------------------------------------------------------------------------------------------
// fused multi-level loss + gradient
// ---------------------------------------------------------------------------------------------
constexpr int kRHW = 512;        // hw positions per unit: 128 lanes x float4
constexpr int kRCT = 8;          // classes per unit
constexpr int kRStages = 3;
constexpr int kRConsumers = 256;  // 8 warps: thread -> (hw quad = tid & 127, class half = tid >> 7)
constexpr int kRThreads = kRConsumers + 32;
constexpr int kRPer = kRCT / 2;  // classes per consumer thread per unit

struct RingLevel {
  const float* X;
  const float* T;
  const int32_t* G;
  float* dX;
  float* loss;
  const float* d_loss;
  uint32_t HW, hw_tiles, unit_begin, unit_end;
};
struct RingArgs {
  RingLevel lv[SAD_MAX_LEVELS];
  int32_t n_levels, num_classes, class_groups, ignored_label;
  uint32_t total_units;
  float gamma, alpha, beta, scale;
  const float* normalizer;
  float* partials;        // [gridDim.x][SAD_MAX_LEVELS]
  unsigned int* counter;  // zero before launch; reset by the last CTA
};

struct __align__(16) UnitDesc {
  float* dX;       // gradient address of (class 0 of this unit, hw 0 of this tile)
  uint32_t n_hw;   // valid hw positions (multiple of 4)
  uint32_t n_cls;  // valid classes
  uint32_t plane;  // H*W
  int32_t level;
  float kg;        // d_loss * scale / Np of this level
  uint32_t pad;
};
struct __align__(128) RingStage {
  float X[kRCT][kRHW];
  float T[kRCT][kRHW];
  int32_t G[kRHW];
};
constexpr size_t kRingSmemBytes = sizeof(RingStage) * kRStages;

template <bool kFast, bool kAlphaHalf, bool kLoss, bool kGrad>
__global__ void __launch_bounds__(kRThreads, 2) distill_ring_kernel(const __grid_constant__ RingArgs args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RingStage* stages = reinterpret_cast<RingStage*>(smem_raw);
  __shared__ UnitDesc desc[kRStages];
  __shared__ __align__(8) uint64_t full_bar[kRStages];
  __shared__ __align__(8) uint64_t empty_bar[kRStages];
  __shared__ float lvl_sum[SAD_MAX_LEVELS];
  __shared__ float red_f[kRConsumers / 32];
  __shared__ bool is_last;

  const int tid = threadIdx.x;
  // units are dealt round-robin (CTA c takes c, c + grid, ...): the small units of the coarse FPN levels, which sit at
  // the end of the unit list, are spread over all CTAs instead of giving the last CTAs almost nothing to do, and at
  // any instant the grid works on consecutive units
  const uint32_t u0 = blockIdx.x, u1 = args.total_units, ustep = gridDim.x;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kRConsumers / 32);
    }
    mbar_fence_init();
  }
  if (tid < SAD_MAX_LEVELS) lvl_sum[tid] = 0.f;
  __syncthreads();

  const float Np = fmaxf(__ldg(args.normalizer), 1.0f);

  if (tid >= kRConsumers) {
    // ===== producer: one thread streams this CTA's units through the ring =====
    if (tid == kRConsumers) {
      const uint64_t pol = policy_evict_first();
      const uint32_t C = (uint32_t)args.num_classes, cg = (uint32_t)args.class_groups;
      RingState rs;
      int l = 0;
      float kg = 0.f;
      int kg_level = -1;
#pragma unroll 1
      for (uint32_t u = u0; u < u1; u += ustep) {
        while (u >= args.lv[l].unit_end) ++l;
        const RingLevel& L = args.lv[l];
        if (kGrad && kg_level != l) {
          kg = (L.d_loss ? __ldg(L.d_loss) : 1.f) * args.scale / Np;
          kg_level = l;
        }
        const uint32_t local = u - L.unit_begin;
        const uint32_t item = local / cg, chunk = local - item * cg;
        const uint32_t na = item / L.hw_tiles, ht = item - na * L.hw_tiles;
        const uint32_t hw0 = ht * kRHW;
        const uint32_t n_hw = min((uint32_t)kRHW, L.HW - hw0);
        const uint32_t c0 = chunk * kRCT;
        const uint32_t n_cls = min((uint32_t)kRCT, C - c0);
        const size_t off = ((size_t)na * C + c0) * L.HW + hw0;
        RingStage& st = stages[rs.stage];
        mbar_wait(&empty_bar[rs.stage], rs.phase ^ 1u);
        UnitDesc d;
        d.dX = kGrad ? L.dX + off : nullptr;
        d.n_hw = n_hw;
        d.n_cls = n_cls;
        d.plane = L.HW;
        d.level = l;
        d.kg = kg;
        d.pad = 0;
        desc[rs.stage] = d;
        const uint32_t row_bytes = n_hw * 4u;
        mbar_arrive_expect_tx(&full_bar[rs.stage], (2u * n_cls + 1u) * row_bytes);
        const float* xs = L.X + off;
        const float* ts = L.T + off;
        for (uint32_t c = 0; c < n_cls; ++c) {
          bulk_g2s(st.X[c], xs + (size_t)c * L.HW, row_bytes, &full_bar[rs.stage], pol);
          bulk_g2s(st.T[c], ts + (size_t)c * L.HW, row_bytes, &full_bar[rs.stage], pol);
        }
        bulk_g2s(st.G, L.G + (size_t)na * L.HW + hw0, row_bytes, &full_bar[rs.stage], pol);
        rs.advance<kRStages>();
      }
    }
  } else {
    // ===== consumers =====
    const int lane = tid & 31;
    const uint32_t h = (uint32_t)(tid & 127) * 4u;
    const uint32_t cbase = (uint32_t)(tid >> 7) * kRPer;
    const float alpha = args.alpha, gamma = args.gamma, beta = args.beta;
    FastConsts fc;
    fc.ca2 = 2.f * alpha * kLn2;
    fc.cb2 = 2.f * (1.f - alpha) * kLn2;
    fc.alpha = alpha;
    fc.om2a = 1.f - 2.f * alpha;
    const float one_m_alpha = 1.f - alpha, one_m_2alpha = 1.f - 2.f * alpha;
    const int32_t ignored = args.ignored_label;
    float acc = 0.f;
    int cur_level = -1;
    RingState rs;
#pragma unroll 1
    for (uint32_t u = u0; u < u1; u += ustep) {
      mbar_wait(&full_bar[rs.stage], rs.phase);
      const UnitDesc d = desc[rs.stage];
      const RingStage& st = stages[rs.stage];
      if (kLoss && d.level != cur_level) {
        if (cur_level >= 0) {
          const float s = group_sum<kRConsumers>(acc, red_f, tid, 1);
          if (tid == 0) lvl_sum[cur_level] = s;
          acc = 0.f;
        }
        cur_level = d.level;
      }
      if (h < d.n_hw) {
        const int4 g = *reinterpret_cast<const int4*>(&st.G[h]);
        float keep[4], kk[4];
        keep[0] = g.x != ignored ? 1.f : 0.f;
        keep[1] = g.y != ignored ? 1.f : 0.f;
        keep[2] = g.z != ignored ? 1.f : 0.f;
        keep[3] = g.w != ignored ? 1.f : 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) kk[v] = keep[v] * d.kg;
        float4 xv[kRPer], tv[kRPer];
#pragma unroll
        for (int j = 0; j < kRPer; ++j) {
          xv[j] = *reinterpret_cast<const float4*>(&st.X[cbase + j][h]);
          tv[j] = *reinterpret_cast<const float4*>(&st.T[cbase + j][h]);
        }
        float* out = kGrad ? d.dX + (size_t)cbase * d.plane + h : nullptr;
#pragma unroll
        for (int j = 0; j < kRPer; ++j) {
          if (cbase + j < d.n_cls) {
            const float xs[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
            const float ts[4] = {tv[j].x, tv[j].y, tv[j].z, tv[j].w};
            float gv[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (kFast) {
                distill_elem_fast<kAlphaHalf, kLoss, kGrad>(xs[v], ts[v], keep[v], kk[v], fc, acc, gv[v]);
              } else {
                float li = 0.f, gg = 0.f;
                distill_elem<false, kLoss, kGrad>(xs[v], ts[v], gamma, alpha, beta, one_m_alpha, one_m_2alpha, li, gg);
                if (kLoss) acc = fmaf(li, keep[v], acc);
                if (kGrad) gv[v] = gg * kk[v];
              }
            }
            if (kGrad) *reinterpret_cast<float4*>(out + (size_t)j * d.plane) = make_float4(gv[0], gv[1], gv[2], gv[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[rs.stage]);
      rs.advance<kRStages>();
    }
    if (kLoss && cur_level >= 0) {
      const float s = group_sum<kRConsumers>(acc, red_f, tid, 1);
      if (tid == 0) lvl_sum[cur_level] = s;
    }
  }

  if (kLoss) {
    if (publish_and_ticket<SAD_MAX_LEVELS>(lvl_sum, args.partials, args.counter, &is_last)) {
      __threadfence();
      const double half = kFast ? 0.5 : 1.0;  // the fast path accumulates twice the summand
      const int lane = tid & 31;
      for (int k = tid >> 5; k < args.n_levels; k += kRThreads / 32) {
        const double s = warp_sum_partials<SAD_MAX_LEVELS>(args.partials, k, lane);
        if (lane == 0) args.lv[k].loss[0] = (float)(half * s / (double)Np) * args.scale;
      }
      if (tid == 0) *args.counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PowSum
// ---------------------------------------------------------------------------------------------
constexpr int kPRChunk = 4096;  // floats per unit (16 KB)
constexpr int kPRStages = 4;
constexpr int kPRConsumers = 256;
constexpr int kPRThreads = kPRConsumers + 32;
constexpr size_t kPowRingSmemBytes = (size_t)kPRChunk * 4 * kPRStages;

enum PowMode { kPowGeneric = 0, kPowOne = 1, kPowTwo = 2, kPowAccurate = 3 };

struct PowRingArgs {
  const float* in[SAD_MAX_INPUTS];
  int64_t n[SAD_MAX_INPUTS];
  uint32_t unit_begin[SAD_MAX_INPUTS];
  uint32_t unit_end[SAD_MAX_INPUTS];
  int32_t n_inputs;
  uint32_t total_units;
  float power;
  float* partials;  // [gridDim.x][SAD_MAX_INPUTS]
  unsigned int* counter;
  float* out;
};
struct __align__(16) PowDesc {
  const float* tail;  // up to 3 trailing elements of the input that a 16-byte bulk copy cannot carry
  uint32_t count;     // floats in the stage (multiple of 4)
  uint32_t tail_n;
  int32_t input;
  uint32_t pad[3];
};

template <int kMode>
__device__ __forceinline__ float pow_elem(float x, float power) {
  if (kMode == kPowOne) return x;
  if (kMode == kPowTwo) return x * x;
  if (kMode == kPowAccurate) return powf(x, power);
  return ex2_approx(power * lg2_approx(x));  // x^p for x >= 0; NaN for x < 0 like powf with non-integer p
}

template <int kMode>
__global__ void __launch_bounds__(kPRThreads, 2) pow_sum_ring_kernel(const __grid_constant__ PowRingArgs args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float(*stages)[kPRChunk] = reinterpret_cast<float(*)[kPRChunk]>(smem_raw);
  __shared__ PowDesc desc[kPRStages];
  __shared__ __align__(8) uint64_t full_bar[kPRStages];
  __shared__ __align__(8) uint64_t empty_bar[kPRStages];
  __shared__ float in_sum[SAD_MAX_INPUTS];
  __shared__ float red_f[kPRConsumers / 32];
  __shared__ bool is_last;

  const int tid = threadIdx.x;
  const uint32_t u0 = (uint32_t)((uint64_t)blockIdx.x * args.total_units / gridDim.x);
  const uint32_t u1 = (uint32_t)((uint64_t)(blockIdx.x + 1) * args.total_units / gridDim.x);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kPRStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kPRConsumers / 32);
    }
    mbar_fence_init();
  }
  if (tid < SAD_MAX_INPUTS) in_sum[tid] = 0.f;
  __syncthreads();

  if (tid >= kPRConsumers) {
    if (tid == kPRConsumers) {
      const uint64_t pol = policy_evict_first();
      RingState rs;
      int k = 0;
#pragma unroll 1
      for (uint32_t u = u0; u < u1; ++u) {
        while (u >= args.unit_end[k]) ++k;
        const int64_t n = args.n[k];
        const int64_t n4 = n & ~(int64_t)3;
        const int64_t start = (int64_t)(u - args.unit_begin[k]) * kPRChunk;
        int64_t cnt = n4 - start;
        cnt = cnt < 0 ? 0 : (cnt > kPRChunk ? kPRChunk : cnt);
        const bool last_of_input = u + 1 == args.unit_end[k];
        mbar_wait(&empty_bar[rs.stage], rs.phase ^ 1u);
        PowDesc d;
        d.tail = args.in[k] + n4;
        d.count = (uint32_t)cnt;
        d.tail_n = last_of_input ? (uint32_t)(n - n4) : 0u;
        d.input = k;
        d.pad[0] = d.pad[1] = d.pad[2] = 0;
        desc[rs.stage] = d;
        mbar_arrive_expect_tx(&full_bar[rs.stage], (uint32_t)cnt * 4u);
        if (cnt > 0) bulk_g2s(stages[rs.stage], args.in[k] + start, (uint32_t)cnt * 4u, &full_bar[rs.stage], pol);
        rs.advance<kPRStages>();
      }
    }
  } else {
    const int lane = tid & 31;
    const float power = args.power;
    float acc = 0.f;
    int cur = -1;
    RingState rs;
#pragma unroll 1
    for (uint32_t u = u0; u < u1; ++u) {
      mbar_wait(&full_bar[rs.stage], rs.phase);
      const PowDesc d = desc[rs.stage];
      if (d.input != cur) {
        if (cur >= 0) {
          const float s = group_sum<kPRConsumers>(acc, red_f, tid, 1);
          if (tid == 0) in_sum[cur] = s;
          acc = 0.f;
        }
        cur = d.input;
      }
      const float4* src = reinterpret_cast<const float4*>(stages[rs.stage]);
      const uint32_t n4 = d.count >> 2;
      float4 v[kPRChunk / 4 / kPRConsumers];
#pragma unroll
      for (int j = 0; j < kPRChunk / 4 / kPRConsumers; ++j) v[j] = src[tid + j * kPRConsumers];
#pragma unroll
      for (int j = 0; j < kPRChunk / 4 / kPRConsumers; ++j) {
        if ((uint32_t)(tid + j * kPRConsumers) < n4) {
          acc += pow_elem<kMode>(v[j].x, power) + pow_elem<kMode>(v[j].y, power);
          acc += pow_elem<kMode>(v[j].z, power) + pow_elem<kMode>(v[j].w, power);
        }
      }
      if ((uint32_t)tid < d.tail_n) acc += pow_elem<kMode>(__ldg(d.tail + tid), power);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[rs.stage]);
      rs.advance<kPRStages>();
    }
    if (cur >= 0) {
      const float s = group_sum<kPRConsumers>(acc, red_f, tid, 1);
      if (tid == 0) in_sum[cur] = s;
    }
  }

  if (publish_and_ticket<SAD_MAX_INPUTS>(in_sum, args.partials, args.counter, &is_last)) {
    __threadfence();
    // per input: fp64 sum of the CTA partials in a fixed order, rounded to float (one warp per input);
    // then the reference's running float add over inputs (pow_sum_op.cu:39)
    const int lane = tid & 31;
    for (int j = tid >> 5; j < args.n_inputs; j += kPRThreads / 32) {
      const double s = warp_sum_partials<SAD_MAX_INPUTS>(args.partials, j, lane);
      if (lane == 0) in_sum[j] = (float)s;
    }
    __syncthreads();
    if (tid == 0) {
      float res = 0.f;
      for (int j = 0; j < args.n_inputs; ++j) res = res + in_sum[j];
      args.out[0] = res;
      *args.counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int ring_grid(uint32_t units, int ctas_per_sm, uint32_t* grid) {
  int dev = 0, sms = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc != SAD_OK) return rc;
  static int cached_sms[64] = {0};
  if (dev >= 0 && dev < 64 && cached_sms[dev] > 0) {
    sms = cached_sms[dev];
  } else {
    rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
    if (rc != SAD_OK) return rc;
    if (dev >= 0 && dev < 64) cached_sms[dev] = sms;
  }
  uint32_t g = (uint32_t)(sms * ctas_per_sm);
  if (g > (uint32_t)kMaxRingCtas) g = kMaxRingCtas;
  if (g > units) g = units;
  *grid = g ? g : 1;
  return SAD_OK;
}

template <typename K>
static int ring_set_smem(K kernel, size_t bytes) {
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "cudaFuncSetAttribute");
}

bool distill_ring_supported(const sad_distill_level* levels, int n_levels, int num_classes) {
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    const uint64_t HW = (uint64_t)L.H * L.W;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(L.logits) | reinterpret_cast<uintptr_t>(L.teacher_prob) |
                           reinterpret_cast<uintptr_t>(L.labels) | reinterpret_cast<uintptr_t>(L.d_logits);
    if (HW % 4 != 0 || (bits & 15) || HW > 0xffffffffull) return false;
    const uint64_t units = (uint64_t)L.N * (L.D / num_classes) * ((HW + kRHW - 1) / kRHW) * ((num_classes + kRCT - 1) / kRCT);
    if (units > 0x3fffffffull) return false;
  }
  return true;
}

int launch_distill_ring(const sad_distill_level* levels, int n_levels, const float* normalizer, const sad_distill_params* p,
                        void* workspace, size_t workspace_bytes, cudaStream_t st) {
  RingArgs a{};
  const uint32_t C = (uint32_t)p->num_classes;
  const uint32_t cg = (C + kRCT - 1) / kRCT;
  uint64_t t = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    const uint64_t HW = (uint64_t)L.H * L.W;
    const uint64_t NA = (uint64_t)L.N * ((uint32_t)L.D / C);
    const uint64_t tiles = (HW + kRHW - 1) / kRHW;
    a.lv[l].X = L.logits;
    a.lv[l].T = L.teacher_prob;
    a.lv[l].G = L.labels;
    a.lv[l].dX = L.d_logits;
    a.lv[l].loss = L.loss;
    a.lv[l].d_loss = L.d_loss;
    a.lv[l].HW = (uint32_t)HW;
    a.lv[l].hw_tiles = (uint32_t)(tiles ? tiles : 1);
    a.lv[l].unit_begin = (uint32_t)t;
    t += NA * tiles * cg;
    a.lv[l].unit_end = (uint32_t)t;
    if (t > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "distill: too many work units for one launch");
  }
  if (t == 0) return SAD_OK;
  const bool want_loss = levels[0].loss != nullptr, want_grad = levels[0].d_logits != nullptr;
  a.n_levels = n_levels;
  a.num_classes = p->num_classes;
  a.class_groups = (int32_t)cg;
  a.ignored_label = p->ignored_label;
  a.total_units = (uint32_t)t;
  a.gamma = p->gamma;
  a.alpha = p->alpha;
  a.beta = p->beta;
  a.scale = p->scale;
  a.normalizer = normalizer;
  if (want_loss) {
    const size_t need = 256 + (size_t)kMaxRingCtas * SAD_MAX_LEVELS * sizeof(float);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
      return set_error(SAD_ERR_WORKSPACE, "distill: workspace must be 256-byte aligned and >= sad_distill_workspace_bytes()");
    a.counter = static_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  }
  uint32_t grid = 1;
  int rc = ring_grid(a.total_units, 2, &grid);
  if (rc != SAD_OK) return rc;
  const bool fast = p->gamma == 2.0f && p->beta == 0.0f;
  const bool half = fast && p->alpha == 0.5f;
#define SAD_RING_LAUNCH(F, H, LS, GR)                                                   \
  do {                                                                                  \
    auto kern = distill_ring_kernel<F, H, LS, GR>;                                      \
    if ((rc = ring_set_smem(kern, kRingSmemBytes)) != SAD_OK) return rc;                \
    kern<<<grid, kRThreads, kRingSmemBytes, st>>>(a);                                   \
  } while (0)
#define SAD_RING_OUT(F, H)                                             \
  do {                                                                 \
    if (want_loss && want_grad) SAD_RING_LAUNCH(F, H, true, true);     \
    else if (want_loss) SAD_RING_LAUNCH(F, H, true, false);            \
    else SAD_RING_LAUNCH(F, H, false, true);                           \
  } while (0)
  if (half) SAD_RING_OUT(true, true);
  else if (fast) SAD_RING_OUT(true, false);
  else SAD_RING_OUT(false, false);
#undef SAD_RING_OUT
#undef SAD_RING_LAUNCH
  count_launch(1);
  return check_cuda(cudaGetLastError(), "distill ring launch");
}

bool pow_sum_ring_supported(const float* const* inputs, const int64_t* sizes, int n_inputs) {
  uint64_t units = 0;
  for (int k = 0; k < n_inputs; ++k) {
    if (sizes[k] > 0 && (reinterpret_cast<uintptr_t>(inputs[k]) & 15)) return false;
    units += (uint64_t)((sizes[k] + kPRChunk - 1) / kPRChunk);
  }
  return units <= 0x3fffffffull;
}

int launch_pow_sum_ring(const float* const* inputs, const int64_t* sizes, int n_inputs, float power, float* out,
                        void* workspace, size_t workspace_bytes, cudaStream_t st) {
  PowRingArgs a{};
  uint64_t t = 0;
  for (int k = 0; k < n_inputs; ++k) {
    a.in[k] = inputs[k];
    a.n[k] = sizes[k];
    a.unit_begin[k] = (uint32_t)t;
    const int64_t n4 = sizes[k] & ~(int64_t)3;
    uint64_t units = (uint64_t)((n4 + kPRChunk - 1) / kPRChunk);
    if (units == 0 && sizes[k] > 0) units = 1;  // fewer than 4 elements: a unit that carries only the tail
    t += units;
    a.unit_end[k] = (uint32_t)t;
  }
  if (t == 0) return check_cuda(cudaMemsetAsync(out, 0, sizeof(float), st), "PowSum memset");
  a.n_inputs = n_inputs;
  a.total_units = (uint32_t)t;
  a.power = power;
  a.out = out;
  const size_t need = 256 + (size_t)kMaxRingCtas * SAD_MAX_INPUTS * sizeof(float);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(SAD_ERR_WORKSPACE, "PowSum: workspace must be 256-byte aligned and >= sad_pow_sum_workspace_bytes()");
  a.counter = static_cast<unsigned int*>(workspace);
  a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  uint32_t grid = 1;
  int rc = ring_grid(a.total_units, 2, &grid);
  if (rc != SAD_OK) return rc;

  int mode = kPowGeneric;
  if (power == 1.0f) mode = kPowOne;
  else if (power == 2.0f) mode = kPowTwo;
  else if (power == floorf(power) || !(power > 0.f)) mode = kPowAccurate;  // integer / non-positive exponents: full powf semantics
#define SAD_POW_LAUNCH(M)                                                       \
  do {                                                                          \
    auto kern = pow_sum_ring_kernel<M>;                                         \
    if ((rc = ring_set_smem(kern, kPowRingSmemBytes)) != SAD_OK) return rc;     \
    kern<<<grid, kPRThreads, kPowRingSmemBytes, st>>>(a);                       \
  } while (0)
  switch (mode) {
    case kPowOne: SAD_POW_LAUNCH(kPowOne); break;
    case kPowTwo: SAD_POW_LAUNCH(kPowTwo); break;
    case kPowAccurate: SAD_POW_LAUNCH(kPowAccurate); break;
    default: SAD_POW_LAUNCH(kPowGeneric); break;
  }
#undef SAD_POW_LAUNCH
  count_launch(1);
  return check_cuda(cudaGetLastError(), "PowSum ring launch");
}

}  // namespace sad
