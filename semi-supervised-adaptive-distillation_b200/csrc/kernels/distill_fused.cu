// PowSum + SigmoidAdaptiveDistillLoss + SigmoidAdaptiveDistillLossGradient for every FPN level in ONE cooperative
// launch — the whole loss step of add_distill_loss (detectron/lib/modeling/retinanet_heads.py:313-352):
//
//   phase 1  normaliser = sum over levels of teacher_prob ^ power          (pow_sum_op.cu:25-43)
//   ------   grid-wide barrier: every CTA adds the per-CTA partial sums in the same fixed order
//   phase 2  per element loss and d(logits), loss reduced per level        (...loss_op.cu:28-67, 69-105, 108-171)
//
// Why one launch (measured on B200): the two-kernel step is 22 us (PowSum) + 51 us (loss + gradient); every SM is empty
// for ~4 us of each launch (ramp + drain) and the teacher probabilities are read from DRAM twice.  Here there is one
// ramp/drain, and phase 1 walks T backwards with an L2 evict-last policy while phase 2 walks forwards with evict-first
// loads and streaming stores, so phase 2 finds T still in the 126 MB L2 (ncu: 172 MB read from DRAM per launch instead of
// 236 MB; with evict-first in phase 1 it is 208 MB).
// Round 2 (profiles/r02_*): (1) the per-element arithmetic is packed (fma/add/mul.rn.f32x2: two elements per FMA-pipe
// instruction, distill_math.cuh) and the teacher-probability NaN rule is checked once per thread and unit instead of per element:
// 34.2 M -> 25.3 M warp instructions per launch, phase 2 is now DRAM-bound (157 MB in 26.5 us); (2) phase-2 loads do not depend
// on the normaliser, so the producer requests the first three units while the grid barrier is being crossed; (3) the barrier's
// words carry the PowSum partials themselves (bar_deliver), which takes three global round trips off its critical path; (4) levels
// whose H*W is not a multiple of 4 (P7 = 5 x 7 of a 640 x 896 input) are done by a scalar tail pass of the same launch instead of
// pushing every level to the two-launch SIMT path.
// Work units are dealt round-robin (static) in both phases.  Dynamic hand-outs through atomic counters were measured in round 1
// (phase 2) and again in round 2 (both phases, with schedule-independent exact sums so that results stay bit-identical): they
// even out the CTAs' finish times but neither phase gets shorter (phase 1 18.5 vs 16.6 us, phase 2 31.1 vs 30.7 us to the last CTA).
// 16 instead of 8 consumer warps: no change; 3 CTAs per SM with 2-stage rings (72 registers, no spills): 69.5 us (round 1).
//
// Determinism: static unit assignment; the normaliser and the per-level losses are integer (fixed-point) sums of per-CTA fp32
// partials, i.e. independent of arrival order: bit-identical run to run.
//
// Restricted to the arithmetic fast path (gamma == 2, beta == 0: the reference's headline configs) with both outputs
// requested; everything else runs as the two ring kernels of distill_ring.cu behind the same entry point.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <string>

#include "distill_math.cuh"
#include "ring.cuh"
#include "sad_b200.h"
#include "sad_internal.h"

namespace sad {

constexpr int kFHW = 512;           // hw positions per phase-2 unit
constexpr int kFCT = 8;             // classes per phase-2 unit
constexpr int kFStages = 3;
constexpr int kFConsumers = 256;
constexpr int kFWarps = kFConsumers / 32;
constexpr int kFThreads = kFConsumers + 32;
constexpr int kFPer = kFCT / 2;
constexpr int kF1Chunk = 4096;      // floats per phase-1 unit (16 KB)
constexpr int kF1Stages = 6;         // the phase-1 ring (96 KB) reuses the whole phase-2 ring's shared memory

struct FusedLevel {
  const float* X;
  const float* T;
  const int32_t* G;
  float* dX;
  float* loss;
  const float* d_loss;
  uint32_t HW, hw_tiles, unit_begin, unit_end;  // phase-2 units
  uint32_t p1_begin, p1_end;                    // phase-1 units (chunks of T)
  uint32_t elems;                               // N * D * HW
  uint32_t D;                                   // channels (A * num_classes)
  uint32_t tail;                                // 1: H*W % 4 != 0 — no ring units, done by the scalar tail pass
};
struct FusedArgs {
  FusedLevel lv[SAD_MAX_LEVELS];
  int32_t n_levels, num_classes, class_groups, ignored_label;
  uint32_t total_units, p1_total;
  float alpha, scale, power;
  uint32_t dbg;  // experiment switches (SAD_FUSED_DEBUG): 1 = evict-first in phase 1, 2 = plain stores, 8 = globaltimer stamps
  float* norm_out;
  unsigned long long* stamps;  // debug (dbg & 8): [gridDim.x][5] globaltimer values
  // control block (zero between launches; the last CTA leaves it zeroed):
  unsigned int* ctrl;          // [1] final ticket
  unsigned int* flags;         // [8..16) phase-2 (loss) per level: kFxNaN / kFxPosInf / kFxNegInf
  unsigned long long* bar;     // [SAD_MAX_LEVELS] grid-barrier words, one per input: arrival count and PowSum partials in ONE atomic (below)
  unsigned long long* p2_acc;  // [SAD_MAX_LEVELS][4]: exact fixed-point sum of the loss terms per level (distill_math.cuh, Fx128) in four 32-bit pieces
};

// Grid barrier that carries the data it is there for.  The only thing the CTAs exchange between the phases is the normaliser, so
// each CTA delivers its PowSum partial of input k and its arrival with a single fire-and-forget 64-bit atomic on word k:
//     bits 0..53   the partial in fixed point, 22 fractional bits (a sum of probabilities^power: 0 <= sum < 2^32)
//     bits 54..62  +1 (arrival count; the grid has at most 2 x 148 CTAs)
//     bit  63      sticky "not representable" (NaN, negative, or beyond the per-CTA cap that keeps the sum out of the count field)
// and then polls the words: the load that shows the last arrival also returns the complete sum.  Against "write partials, fence,
// count, poll, fence, read 296 x 5 partials" this takes three global round trips off the critical path (measured: 3.9 -> x us from
// the slowest CTA's end of phase 1 to the start of phase 2); integer addition makes the sum independent of the arrival order.
constexpr int kBarFrac = 22;
constexpr int kBarCountShift = 54;
constexpr unsigned long long kBarBad = 1ull << 63;
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bar_deliver(unsigned long long* word, float partial, uint32_t grid) {
  const double cap = (double)((1ull << kBarCountShift) / grid - 1ull);   // sum over all CTAs stays below 2^54
  const double q = (double)partial * (double)(1u << kBarFrac);
  if (!(q >= 0.0 && q <= cap)) {   // NaN, negative or too large
    atomicOr(word, kBarBad);       // same word, same thread: ordered before the arrival below
    atomicAdd(word, 1ull << kBarCountShift);
  } else {
    atomicAdd(word, (1ull << kBarCountShift) + (unsigned long long)__double2ll_rn(q));
  }
}

struct __align__(16) FUnitDesc {
  float* dX;
  uint32_t n_hw, n_cls, plane, unit;
  int32_t level;
  float kg;
};
struct __align__(128) FStage {
  float X[kFCT][kFHW];
  float T[kFCT][kFHW];
  int32_t G[kFHW];
};
constexpr size_t kFusedSmemBytes = sizeof(FStage) * kFStages;
static_assert(kFusedSmemBytes >= (size_t)kF1Chunk * 4 * kF1Stages, "phase-1 ring must fit in the phase-2 ring");

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// deliver one value per warp (already reduced over the warp, valid in lane 0) into the CTA's exact accumulator `slot`
__device__ __forceinline__ void fx_deliver(float v, unsigned long long* acc_smem /* [lo, hi] */, unsigned int* flag_smem) {
  Fx128 x;
  const unsigned int f = fx_from_float(v, x);
  if (f) atomicOr(flag_smem, f);
  else fx_atomic_add(acc_smem, x);
}

template <bool kAlphaHalf, bool kPowAccurate>
__global__ void __launch_bounds__(kFThreads, 2) distill_fused_kernel(const __grid_constant__ FusedArgs args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FStage* stages = reinterpret_cast<FStage*>(smem_raw);
  float(*p1_stages)[kF1Chunk] = reinterpret_cast<float(*)[kF1Chunk]>(smem_raw);   // the phase-1 ring reuses the phase-2 ring's memory
  __shared__ FUnitDesc desc[kFStages];
  __shared__ int32_t p1_desc[kF1Stages][2];  // {input, count}; count < 0: no more chunks
  __shared__ __align__(8) uint64_t full_bar[kFStages], empty_bar[kFStages], p1_full[kF1Stages], p1_empty[kF1Stages];
  __shared__ __align__(16) unsigned long long acc_s[2][SAD_MAX_LEVELS][2];   // [phase][input / level][lo, hi]: this CTA's exact sums
  __shared__ unsigned int flag_s[2][SAD_MAX_LEVELS];
  __shared__ float in_sum[SAD_MAX_LEVELS];
  __shared__ float red_f[kFWarps];
  __shared__ float np_smem;
  __shared__ bool is_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kFWarps);
    }
#pragma unroll
    for (int s = 0; s < kF1Stages; ++s) {
      mbar_init(&p1_full[s], 1);
      mbar_init(&p1_empty[s], kFWarps);
    }
    // The phase-2 ring's memory holds the phase-1 ring first.  A stage joins phase 2 when the consumer warps have arrived on
    // its empty barrier after phase 1 — that barrier's FIRST completion — so the phase-2 ring state starts with phase bit 1
    // (first wait: parity 0), and this dummy first completion of every full barrier keeps the consumer side in step.
#pragma unroll
    for (int s = 0; s < kFStages; ++s) mbar_arrive(&full_bar[s]);
    mbar_fence_init();
  }
  if (tid < SAD_MAX_LEVELS) in_sum[tid] = 0.f;
  if (tid < 2 * SAD_MAX_LEVELS) {
    acc_s[tid / SAD_MAX_LEVELS][tid % SAD_MAX_LEVELS][0] = 0ull;
    acc_s[tid / SAD_MAX_LEVELS][tid % SAD_MAX_LEVELS][1] = 0ull;
    flag_s[tid / SAD_MAX_LEVELS][tid % SAD_MAX_LEVELS] = 0u;
  }
  __syncthreads();
  const bool stamp = (args.dbg & 8u) && tid == 0;
  if (stamp) args.stamps[blockIdx.x * 5 + 0] = gtimer();

  if (tid >= kFConsumers) {
    // ================= producer thread: phase-1 chunks, then (without waiting for the normaliser) phase-2 units ==========
    if (tid == kFConsumers) {
      {
        // unit j of this CTA = global chunk (p1_total - 1 - (blockIdx.x + j * gridDim.x)): last chunk first (phase 2 walks forwards
        // and then finds the head of T most recently used in L2).  Static deal: handing the chunks out through a counter was
        // measured in round 2 (per-chunk exact sums so that the result stays deterministic) and made the phase slower, 18.5 vs 16.6 us.
        const uint64_t pol = (args.dbg & 1u) ? policy_evict_first() : policy_evict_last();
        RingState rs;
        int k = args.n_levels - 1;
#pragma unroll 1
        for (uint32_t r = blockIdx.x; r < args.p1_total; r += gridDim.x) {
          const uint32_t u = args.p1_total - 1u - r;
          while (u < args.lv[k].p1_begin) --k;
          const uint32_t start = (u - args.lv[k].p1_begin) * (uint32_t)kF1Chunk;
          const uint32_t left = args.lv[k].elems - start;
          const uint32_t cnt = left < (uint32_t)kF1Chunk ? left : (uint32_t)kF1Chunk;
          mbar_wait(&p1_empty[rs.stage], rs.phase ^ 1u);
          p1_desc[rs.stage][0] = k;
          p1_desc[rs.stage][1] = (int32_t)cnt;
          mbar_arrive_expect_tx(&p1_full[rs.stage], cnt * 4u);
          bulk_g2s(p1_stages[rs.stage], args.lv[k].T + start, cnt * 4u, &p1_full[rs.stage], pol);
          rs.advance<kF1Stages>();
        }
      }
      // phase 2: X and T rows and the label row of each unit, units dealt round-robin (static: a dynamic hand-out was measured
      // again in round 2 — it evens out the finish times but the phase as a whole is not shorter, 31.1 vs 30.7 us).  Nothing
      // here depends on the normaliser, so the first three units are requested the moment the consumers are through with
      // phase 1 and land while the grid barrier is crossed.
      const uint64_t pol = policy_evict_first();
      const uint32_t C = (uint32_t)args.num_classes, cg = (uint32_t)args.class_groups;
      RingState rs;
      rs.phase = 1u;
      int l = 0;
#pragma unroll 1
      for (uint32_t u = blockIdx.x; u < args.total_units; u += gridDim.x) {
        mbar_wait(&empty_bar[rs.stage], rs.phase ^ 1u);
        while (u >= args.lv[l].unit_end) ++l;
        const FusedLevel& L = args.lv[l];
        const uint32_t local = u - L.unit_begin;
        const uint32_t item = local / cg, chunk = local - item * cg;
        const uint32_t na = item / L.hw_tiles, ht = item - na * L.hw_tiles;
        const uint32_t hw0 = ht * kFHW;
        const uint32_t n_hw = min((uint32_t)kFHW, L.HW - hw0);
        const uint32_t c0 = chunk * kFCT;
        const uint32_t n_cls = min((uint32_t)kFCT, C - c0);
        const size_t off = ((size_t)na * C + c0) * L.HW + hw0;
        FStage& st = stages[rs.stage];
        FUnitDesc d;
        d.dX = L.dX + off;
        d.n_hw = n_hw;
        d.n_cls = n_cls;
        d.plane = L.HW;
        d.unit = u;
        d.level = l;
        d.kg = 0.f;
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
        rs.advance<kFStages>();
      }
    }
  } else {
    // ================= consumers, phase 1: PowSum over the teacher probabilities =================
    {
      const float power = args.power;
      const f32x2 power2 = pk2(power, power);
      float acc = 0.f;
      int cur = -1;
      RingState rs;
#pragma unroll 1
      for (uint32_t r = blockIdx.x; r < args.p1_total; r += gridDim.x) {
        mbar_wait(&p1_full[rs.stage], rs.phase);
        const int input = p1_desc[rs.stage][0];
        const uint32_t n4 = (uint32_t)p1_desc[rs.stage][1] >> 2;
        if (input != cur) {
          if (cur >= 0) {
            const float s = group_sum<kFConsumers>(acc, red_f, tid, 1);
            if (tid == 0) in_sum[cur] = s;
            acc = 0.f;
          }
          cur = input;
        }
        const float4* src = reinterpret_cast<const float4*>(p1_stages[rs.stage]);
#pragma unroll
        for (int j = 0; j < kF1Chunk / 4 / kFConsumers; ++j) {
          const uint32_t i = (uint32_t)tid + j * kFConsumers;
          if (i < n4) {
            const float4 v = src[i];
            if (kPowAccurate) {
              acc += powf(v.x, power) + powf(v.y, power);
              acc += powf(v.z, power) + powf(v.w, power);
            } else {  // x^p = 2^(p log2 x) for x >= 0 (NaN for x < 0, like powf with a non-integer exponent)
              float a, b, c, d;
              upk2(mul2(power2, pk2(lg2_approx(v.x), lg2_approx(v.y))), a, b);
              upk2(mul2(power2, pk2(lg2_approx(v.z), lg2_approx(v.w))), c, d);
              acc += ex2_approx(a) + ex2_approx(b);
              acc += ex2_approx(c) + ex2_approx(d);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&p1_empty[rs.stage]);
        rs.advance<kF1Stages>();
      }
      if (cur >= 0) {
        const float s = group_sum<kFConsumers>(acc, red_f, tid, 1);
        if (tid == 0) in_sum[cur] = s;
      }
    }
    // phase 1 consumed: the ring's memory is free for phase 2
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kFStages; ++s) mbar_arrive(&empty_bar[s]);
    }
    named_bar_sync(2, kFConsumers);

    // ================= grid barrier that carries the normaliser (bar_deliver above) =================
    if (stamp) args.stamps[blockIdx.x * 5 + 1] = gtimer();
    if (warp == 0) {
      const bool mine = lane < args.n_levels;
      if (mine) bar_deliver(args.bar + lane, in_sum[lane], gridDim.x);
      unsigned long long v = 0ull;
      for (;;) {
        if (mine) v = ld_acquire_gpu_u64(args.bar + lane);
        const bool done = !mine || ((v >> kBarCountShift) & 0x1ffull) >= (unsigned long long)gridDim.x;
        if (__all_sync(0xffffffffu, done)) break;
        __nanosleep(64);
      }
      // per input: the exact sum rounded to float once; then the reference's running float add over the inputs (pow_sum_op.cu:39)
      if (mine)
        in_sum[lane] = (v & kBarBad) ? __int_as_float(0x7fffffff)
                                     : (float)((double)(v & ((1ull << kBarCountShift) - 1ull)) * (1.0 / (double)(1u << kBarFrac)));
      __syncwarp();
      if (lane == 0) {
        float res = 0.f;
        for (int j = 0; j < args.n_levels; ++j) res = res + in_sum[j];
        np_smem = res;
        if (blockIdx.x == 0) args.norm_out[0] = res;
      }
    }
    named_bar_sync(2, kFConsumers);
    const float Np = fmaxf(np_smem, 1.0f);   // max(weight_pos[0], 1.0): ...loss_op.cu:49,87
    if (stamp) args.stamps[blockIdx.x * 5 + 2] = gtimer();

    // ================= consumers, phase 2: loss + gradient, units dealt round-robin =================
    const uint32_t h = (uint32_t)(tid & 127) * 4u;
    const uint32_t cbase = (uint32_t)(tid >> 7) * kFPer;
    const float alpha = args.alpha;
    FastConsts fc;
    fc.ca2 = 2.f * alpha * kLn2;
    fc.cb2 = 2.f * (1.f - alpha) * kLn2;
    fc.alpha = alpha;
    fc.om2a = 1.f - 2.f * alpha;
    const PairConsts pc = pair_consts();
    const int32_t ignored = args.ignored_label;
    const float kscale = args.scale / Np;
    RingState rs;
    rs.phase = 1u;
    float acc = 0.f, kg = 0.f;
    int cur_level = -1;
#pragma unroll 1
    for (uint32_t u = blockIdx.x; u < args.total_units; u += gridDim.x) {
      mbar_wait(&full_bar[rs.stage], rs.phase);
      const FUnitDesc d = desc[rs.stage];
      const FStage& st = stages[rs.stage];
      if (d.level != cur_level) {
        if (cur_level >= 0) {   // the level's share of this CTA: thread sums over its units (static deal), shuffle tree, exact delivery
          const float ws = warp_sum(acc);
          if (lane == 0) fx_deliver(ws, acc_s[1][cur_level], &flag_s[1][cur_level]);
          acc = 0.f;
        }
        cur_level = d.level;
        const float* dl = args.lv[cur_level].d_loss;
        kg = (dl ? __ldg(dl) : 1.f) * kscale;   // d_loss * scale / Np of this level
      }
      if (h < d.n_hw) {
        const int4 g = *reinterpret_cast<const int4*>(&st.G[h]);
        float keep[4], kk[4];
        keep[0] = g.x != ignored ? 1.f : 0.f;
        keep[1] = g.y != ignored ? 1.f : 0.f;
        keep[2] = g.z != ignored ? 1.f : 0.f;
        keep[3] = g.w != ignored ? 1.f : 0.f;
        float* out = d.dX + (size_t)cbase * d.plane + h;
        if (kAlphaHalf && cbase + kFPer <= d.n_cls) {
          // packed arithmetic (distill_math.cuh, distill_pair_half): 2 elements per FMA-pipe instruction
#pragma unroll
          for (int v = 0; v < 4; ++v) kk[v] = keep[v] * (kg * kLn2);
          const f32x2 kk01 = pk2(kk[0], kk[1]), kk23 = pk2(kk[2], kk[3]);
          f32x2 a01 = pk2(0.f, 0.f), a23 = a01;   // sum over this thread's classes of AT^2 * D per hw position
          float vmin = 1.f;
#pragma unroll
          for (int j = 0; j < kFPer; ++j) {
            const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&st.X[cbase + j][h]);
            const ulonglong2 tv = *reinterpret_cast<const ulonglong2*>(&st.T[cbase + j][h]);
            const f32x2 g01 = distill_pair_half(xv.x, tv.x, kk01, pc, a01, vmin);
            const f32x2 g23 = distill_pair_half(xv.y, tv.y, kk23, pc, a23, vmin);
            float4 o;
            upk2(g01, o.x, o.y);
            upk2(g23, o.z, o.w);
            __stcs(reinterpret_cast<float4*>(out + (size_t)j * d.plane), o);
          }
          if (vmin > 0.f) {
            float s0, s1, s2, s3;
            upk2(a01, s0, s1);
            upk2(a23, s2, s3);
            // acc gathers twice the (positive) loss summand: -(AT^2) * D2 * keep with D2 = ln2 * D
            acc = fmaf(-kLn2 * keep[0], s0, acc);
            acc = fmaf(-kLn2 * keep[1], s1, acc);
            acc = fmaf(-kLn2 * keep[2], s2, acc);
            acc = fmaf(-kLn2 * keep[3], s3, acc);
          } else {
            // some teacher probability of this thread's 16 elements is <= 0, >= 1 or NaN: the reference's result there is NaN
            // (...loss_op.cu:59,93).  Redo them with the scalar function, which carries that rule per element.
#pragma unroll
            for (int v = 0; v < 4; ++v) kk[v] = keep[v] * kg;
#pragma unroll 1
            for (int j = 0; j < kFPer; ++j) {
              const float4 xq = *reinterpret_cast<const float4*>(&st.X[cbase + j][h]);
              const float4 tq = *reinterpret_cast<const float4*>(&st.T[cbase + j][h]);
              const float xs[4] = {xq.x, xq.y, xq.z, xq.w};
              const float ts[4] = {tq.x, tq.y, tq.z, tq.w};
              float gv[4];
#pragma unroll
              for (int v = 0; v < 4; ++v) distill_elem_fast<true, true, true>(xs[v], ts[v], keep[v], kk[v], fc, acc, gv[v]);
              __stcs(reinterpret_cast<float4*>(out + (size_t)j * d.plane), make_float4(gv[0], gv[1], gv[2], gv[3]));
            }
          }
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) kk[v] = keep[v] * kg;
          float4 xv[kFPer], tv[kFPer];
#pragma unroll
          for (int j = 0; j < kFPer; ++j) {
            xv[j] = *reinterpret_cast<const float4*>(&st.X[cbase + j][h]);
            tv[j] = *reinterpret_cast<const float4*>(&st.T[cbase + j][h]);
          }
#pragma unroll
          for (int j = 0; j < kFPer; ++j) {
            if (cbase + j < d.n_cls) {
              const float xs[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
              const float ts[4] = {tv[j].x, tv[j].y, tv[j].z, tv[j].w};
              float gv[4];
#pragma unroll
              for (int v = 0; v < 4; ++v) distill_elem_fast<kAlphaHalf, true, true>(xs[v], ts[v], keep[v], kk[v], fc, acc, gv[v]);
              __stcs(reinterpret_cast<float4*>(out + (size_t)j * d.plane), make_float4(gv[0], gv[1], gv[2], gv[3]));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[rs.stage]);
      rs.advance<kFStages>();
    }
    if (cur_level >= 0) {
      const float ws = warp_sum(acc);
      if (lane == 0) fx_deliver(ws, acc_s[1][cur_level], &flag_s[1][cur_level]);
    }

    // ================= tail levels: H*W % 4 != 0 (e.g. P7 = 5 x 7 of a 640 x 896 input) =================
    // Their rows are not 16-byte aligned, so they cannot ride the bulk-copy ring; they are a fraction of a per cent of the
    // elements (coarsest levels only) and are done here with scalar accesses, strided over the whole grid (static), in the
    // same launch.  Index arithmetic of ...loss_op.cu:35-42.
    for (int l = 0; l < args.n_levels; ++l) {
      const FusedLevel& L = args.lv[l];
      if (!L.tail) continue;
      const float* dl = L.d_loss;
      const float kgl = (dl ? __ldg(dl) : 1.f) * kscale;
      const uint32_t C = (uint32_t)args.num_classes, A = L.D / C;
      float tacc = 0.f;
      for (uint32_t i = blockIdx.x * kFConsumers + tid; i < L.elems; i += gridDim.x * kFConsumers) {
        const uint32_t hw = i % L.HW, c = (i / L.HW) % L.D, n = i / (L.HW * L.D);
        const int32_t t = __ldg(L.G + ((size_t)n * A + c / C) * L.HW + hw);
        const float keep = t != ignored ? 1.f : 0.f;
        float gv;
        distill_elem_fast<kAlphaHalf, true, true>(ld_stream1(L.X + i), ld_stream1(L.T + i), keep, keep * kgl, fc, tacc, gv);
        L.dX[i] = gv;
      }
      const float ws = warp_sum(tacc);
      if (lane == 0) fx_deliver(ws, acc_s[1][l], &flag_s[1][l]);
    }
  }

  // ================= every CTA adds its exact level sums to the global ones; the last CTA rounds them once =================
  // The 128-bit sums travel as four 32-bit pieces, each added into its own 64-bit word (32 bits of headroom: no carries to
  // propagate, so the adds are fire-and-forget reductions instead of two dependent atomic round trips); the reader puts
  // the pieces back together modulo 2^128.
  if (stamp) args.stamps[blockIdx.x * 5 + 3] = gtimer();
  __syncthreads();
  if (tid < 4 * args.n_levels) {
    const int k = tid >> 2, piece = tid & 3;
    const unsigned long long limb = acc_s[1][k][piece >> 1];
    const unsigned long long v = (piece & 1) ? (limb >> 32) : (limb & 0xffffffffull);
    if (v) atomicAdd(args.p2_acc + 4 * k + piece, v);
    if (piece == 0 && flag_s[1][k]) atomicOr(&args.flags[SAD_MAX_LEVELS + k], flag_s[1][k]);
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&args.ctrl[1], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const float Np = fmaxf(np_smem, 1.0f);
    if (tid < args.n_levels) {
      unsigned long long w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = __ldcg(args.p2_acc + 4 * tid + i);
      // value = w0 + w1 * 2^32 + w2 * 2^64 + w3 * 2^96  (mod 2^128)
      const unsigned long long lo = w[0] + (w[1] << 32);
      const unsigned long long c0 = lo < w[0] ? 1ull : 0ull;
      const unsigned long long hi = (w[1] >> 32) + w[2] + (w[3] << 32) + c0;
      const double s = fx_to_double(lo, hi, __ldcg(args.flags + SAD_MAX_LEVELS + tid));
      // the arithmetic accumulates twice the summand
      args.lv[tid].loss[0] = (float)(0.5 * s / (double)Np) * args.scale;
    }
    __syncthreads();
    if (tid < SAD_MAX_LEVELS) {   // leave the control block zeroed for the next launch
      args.bar[tid] = 0ull;
#pragma unroll
      for (int i = 0; i < 4; ++i) args.p2_acc[4 * tid + i] = 0ull;
      args.flags[tid] = args.flags[SAD_MAX_LEVELS + tid] = 0u;
    }
    if (tid == 0) args.ctrl[1] = 0u;
    if (stamp) args.stamps[blockIdx.x * 5 + 4] = gtimer();
  }
}

// a level whose rows the bulk-copy ring cannot address (H*W % 4 != 0): scalar tail pass
static bool fused_tail_level(const sad_distill_level& L) { return (((uint64_t)L.H * L.W) & 3) != 0; }

static size_t fused_units(const sad_distill_level* levels, int n_levels, int num_classes, uint64_t* p1_units) {
  uint64_t t = 0, p1 = 0;
  const uint32_t C = (uint32_t)num_classes, cg = (C + kFCT - 1) / kFCT;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    const uint64_t HW = (uint64_t)L.H * L.W, NA = (uint64_t)L.N * ((uint32_t)L.D / C);
    if (!fused_tail_level(L)) t += NA * ((HW + kFHW - 1) / kFHW) * cg;
    p1 += ((uint64_t)L.N * L.D * HW + kF1Chunk - 1) / kF1Chunk;
  }
  if (p1_units) *p1_units = p1;
  return (size_t)t;
}

// layout of the fused workspace: control block [0, 1024): ctrl (16 B) | flags (64 B at 64) | barrier words (64 B at 256) |
// phase-2 sums (256 B at 512); debug stamps behind it.  sad_workspace_init zeroes it once; every launch leaves it zeroed.
constexpr size_t kFusedCtrlBytes = 1024;
static size_t fused_ws_bytes(size_t) { return kFusedCtrlBytes + (size_t)kMaxRingCtas * 5 * sizeof(unsigned long long); }

bool distill_fused_supported(const sad_distill_level* levels, int n_levels, const sad_distill_params* p, float power) {
  if (!(p->gamma == 2.0f && p->beta == 0.0f)) return false;
  uint64_t ring_elems = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    if (!L.loss || !L.d_logits) return false;
    const uint64_t HW = (uint64_t)L.H * L.W;
    const uint64_t n = (uint64_t)L.N * L.D * HW;
    if (n == 0 || n > 0xfffffff0ull) return false;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(L.logits) | reinterpret_cast<uintptr_t>(L.teacher_prob) |
                           reinterpret_cast<uintptr_t>(L.labels) | reinterpret_cast<uintptr_t>(L.d_logits);
    if (bits & 15) return false;
    // phase 1 streams T in 16-byte multiples whatever the row length: the level's element count must be a multiple of 4
    // (true for every D % 4 == 0, e.g. 9 anchors x 80 classes)
    if (n & 3) return false;
    if (!fused_tail_level(L)) ring_elems += n;
  }
  if (ring_elems == 0) return false;   // nothing for the ring: the SIMT kernels are the better fit
  uint64_t p1 = 0;
  const size_t units = fused_units(levels, n_levels, p->num_classes, &p1);
  return units > 0 && units < 0x3fffffffull && p1 < 0x3fffffffull && power == power;
}

int launch_distill_fused(const sad_distill_level* levels, int n_levels, float power, float* norm_out, const sad_distill_params* p,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
  FusedArgs a{};
  const uint32_t C = (uint32_t)p->num_classes, cg = (C + kFCT - 1) / kFCT;
  uint64_t t = 0, p1 = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    const uint64_t HW = (uint64_t)L.H * L.W, NA = (uint64_t)L.N * ((uint32_t)L.D / C);
    const uint64_t tiles = (HW + kFHW - 1) / kFHW;
    FusedLevel& D = a.lv[l];
    D.X = L.logits;
    D.T = L.teacher_prob;
    D.G = L.labels;
    D.dX = L.d_logits;
    D.loss = L.loss;
    D.d_loss = L.d_loss;
    D.HW = (uint32_t)HW;
    D.hw_tiles = (uint32_t)tiles;
    D.tail = fused_tail_level(L) ? 1u : 0u;
    D.D = (uint32_t)L.D;
    D.unit_begin = (uint32_t)t;
    if (!D.tail) t += NA * tiles * cg;
    D.unit_end = (uint32_t)t;
    D.elems = (uint32_t)((uint64_t)L.N * L.D * HW);
    D.p1_begin = (uint32_t)p1;
    p1 += ((uint64_t)D.elems + kF1Chunk - 1) / kF1Chunk;
    D.p1_end = (uint32_t)p1;
  }
  a.n_levels = n_levels;
  a.num_classes = p->num_classes;
  a.class_groups = (int32_t)cg;
  a.ignored_label = p->ignored_label;
  a.total_units = (uint32_t)t;
  a.p1_total = (uint32_t)p1;
  a.alpha = p->alpha;
  a.scale = p->scale;
  a.power = power;
  a.norm_out = norm_out;
  if (const char* e = getenv("SAD_FUSED_DEBUG")) a.dbg = (uint32_t)atoi(e);
  if (!workspace || workspace_bytes < fused_ws_bytes((size_t)t) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(SAD_ERR_WORKSPACE, "distill fused: workspace must be 256-byte aligned and >= sad_distill_fused_workspace_bytes()");
  char* wsb = static_cast<char*>(workspace);
  a.ctrl = reinterpret_cast<unsigned int*>(wsb);
  a.flags = reinterpret_cast<unsigned int*>(wsb + 64);
  a.bar = reinterpret_cast<unsigned long long*>(wsb + 256);
  a.p2_acc = reinterpret_cast<unsigned long long*>(wsb + 512);
  a.stamps = reinterpret_cast<unsigned long long*>(wsb + kFusedCtrlBytes);

  int dev = 0, sms = 0, rc;
  if ((rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) != SAD_OK) return rc;
  if ((rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) != SAD_OK) return rc;
  const bool half = p->alpha == 0.5f;
  const bool accurate = power == floorf(power) || !(power > 0.f);  // integer / non-positive exponents: full powf semantics
  const void* kern = half ? (accurate ? (const void*)distill_fused_kernel<true, true> : (const void*)distill_fused_kernel<true, false>)
                          : (accurate ? (const void*)distill_fused_kernel<false, true> : (const void*)distill_fused_kernel<false, false>);
  if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes), "cudaFuncSetAttribute")) != SAD_OK)
    return rc;
  int per_sm = 0;
  if ((rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFThreads, kFusedSmemBytes), "occupancy query")) != SAD_OK) return rc;
  if (per_sm < 1) return set_error(SAD_ERR_CUDA, "distill fused: the kernel does not fit on an SM");
  if (per_sm > 2) per_sm = 2;
  uint32_t grid = (uint32_t)(sms * per_sm);   // every CTA must be resident: the kernel contains a grid-wide barrier
  if (grid > (uint32_t)kMaxRingCtas) grid = kMaxRingCtas;
  if (grid > 500u) grid = 500u;   // the barrier words count arrivals in 9 bits
  void* params[] = {&a};
  rc = check_cuda(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kFThreads), params, kFusedSmemBytes, st), "distill fused launch");
  if (rc != SAD_OK) return rc;
  count_launch(1);
  return SAD_OK;
}

size_t distill_fused_workspace_bytes(const sad_distill_level* levels, int n_levels, int num_classes) {
  return fused_ws_bytes(fused_units(levels, n_levels, num_classes, nullptr));
}

}  // namespace sad

using namespace sad;

extern "C" {

// bytes the two-launch form (PowSum, then loss + gradient) needs at the start of the workspace; the one-launch kernel's control
// block lives behind it, so a workspace can serve either form call after call
static size_t fused_two_launch_bytes(const sad_distill_level* levels, int n_levels) {
  int64_t sizes[SAD_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) sizes[l] = (int64_t)levels[l].N * levels[l].D * levels[l].H * levels[l].W;
  const size_t a = sad_pow_sum_workspace_bytes(sizes, n_levels), b = sad_distill_workspace_bytes(levels, n_levels);
  return (((a > b ? a : b) + 255) / 256) * 256;
}

SAD_EXPORT size_t sad_distill_fused_workspace_bytes(const sad_distill_level* levels, int n_levels, int num_classes) {
  if (!levels || n_levels < 1 || n_levels > SAD_MAX_LEVELS || num_classes < 1) return 0;
  for (int l = 0; l < n_levels; ++l)
    if (levels[l].N < 0 || levels[l].D < 0 || levels[l].H < 0 || levels[l].W < 0 || levels[l].D % num_classes) return 0;
  return fused_two_launch_bytes(levels, n_levels) + distill_fused_workspace_bytes(levels, n_levels, num_classes);
}

SAD_EXPORT int sad_distill_fused_f32(const sad_distill_level* levels, int n_levels, float power, float* normalizer_out,
                                     const sad_distill_params* params, void* workspace, size_t workspace_bytes, void* stream) {
  if (!levels || !params || !normalizer_out || n_levels < 1 || n_levels > SAD_MAX_LEVELS)
    return set_error(SAD_ERR_INVALID, "distill fused: bad argument");
  if (params->num_classes < 1) return set_error(SAD_ERR_INVALID, "distill fused: num_classes must be >= 1");
  if (!(params->scale >= 0.f)) return set_error(SAD_ERR_INVALID, "distill fused: scale must be >= 0");
  bool dims_ok = true;
  for (int l = 0; l < n_levels; ++l) {
    const sad_distill_level& L = levels[l];
    if (L.N < 0 || L.D < 0 || L.H < 0 || L.W < 0 || L.D % params->num_classes) dims_ok = false;
    else if ((uint64_t)L.N * L.D * L.H * L.W && (!L.logits || !L.teacher_prob || !L.labels)) dims_ok = false;
  }
  if (!dims_ok) return set_error(SAD_ERR_INVALID, "distill fused: bad level (negative size, D % num_classes != 0 or null tensor)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (distill_fused_supported(levels, n_levels, params, power)) {
    const size_t off = fused_two_launch_bytes(levels, n_levels);
    if (!workspace || workspace_bytes < off) return set_error(SAD_ERR_WORKSPACE, "distill fused: workspace smaller than sad_distill_fused_workspace_bytes()");
    return launch_distill_fused(levels, n_levels, power, normalizer_out, params, static_cast<char*>(workspace) + off, workspace_bytes - off, st);
  }
  // general arguments / shapes: the same contract as two launches sharing the workspace one after the other
  const float* ins[SAD_MAX_LEVELS];
  int64_t sizes[SAD_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) {
    ins[l] = levels[l].teacher_prob;
    sizes[l] = (int64_t)levels[l].N * levels[l].D * levels[l].H * levels[l].W;
  }
  int rc = sad_pow_sum_f32(ins, sizes, n_levels, power, normalizer_out, workspace, workspace_bytes, stream);
  if (rc != SAD_OK) return rc;
  return sad_distill_f32(levels, n_levels, normalizer_out, params, workspace, workspace_bytes, stream);
}

}  // extern "C"
