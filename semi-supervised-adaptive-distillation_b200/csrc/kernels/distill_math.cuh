// Per-element arithmetic and reduction helpers shared by the SIMT kernels (distill_kernels.cu) and the
// persistent bulk-copy-ring kernels (distill_ring.cu).
#ifndef SAD_DISTILL_MATH_CUH_
#define SAD_DISTILL_MATH_CUH_

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

namespace sad {

// -------------------------------------------------------------------------------------------
// small device helpers
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// streaming 128-bit load: read-only path, do not allocate in L1 (every X/T byte is used once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of one value per thread over the CTA; result valid in thread 0.  kThreads % 32 == 0.
template <int kThreads, typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* kThreads/32 entries */) {
  constexpr int kWarps = kThreads / 32;
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = 0;
  if (warp == 0) {
    r = lane < kWarps ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// -------------------------------------------------------------------------------------------
// per-element math
// -------------------------------------------------------------------------------------------
// Given logit x and teacher probability pt, produce (both WITHOUT the ignore mask, 1/Np, scale):
//   li = -AT^gamma * DLoss          (>= 0; the loss summand is li * keep / Np)
//   g  = d*gamma*AT^(gamma-1)*E*DLoss - AT^gamma*g2   (the gradient is g * keep * d_loss / Np)
// with, as in ...loss_op.cu:54-64 / :88-99,
//   e = exp(-|x|), L = log(1+e), p = sigmoid(x), DL = -x*(pt-[x>=0]) + L + beta*(-H(pt)),
//   E = exp(-DL), AT = 1-E, DLoss = alpha*pt*log(max(FLT_MIN,p)) + (1-alpha)(1-pt)*log(1-p),
//   d = pt-p, g2 = alpha*d - (1-2alpha)(1-pt)p.
// kFast = (gamma == 2 && beta == 0): 4 MUFU (ex2, lg2, rcp, ex2), no powf.  The beta term is
// dropped from the arithmetic but its NaN is kept: the reference evaluates
// beta*(pt*logf(pt)+(1-pt)*logf(1-pt)) even for beta == 0, which is NaN unless 0 < pt < 1.
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLogFltMin = -87.33654475055310898657f;  // logf(FLT_MIN)

template <bool kFast, bool kLoss, bool kGrad>
__device__ __forceinline__ void distill_elem(float x, float pt, float gamma, float alpha, float beta,
                                             float one_m_alpha, float one_m_2alpha, float& li, float& g) {
  const float e = ex2_approx(-fabsf(x) * kLog2e);
  const float u = 1.f + e;
  const float L = lg2_approx(u) * kLn2;
  const float mx = fmaxf(x, 0.f);
  const float logp = fmaxf((x - mx) - L, kLogFltMin);
  const float lq = -(mx + L);
  float DL = fmaf(-x, pt, mx) + L;
  const float q = 1.f - pt;
  if (kFast) {
    DL = (pt > 0.f && pt < 1.f) ? DL : __int_as_float(0x7fffffff);
  } else {
    DL += beta * (pt * logf(pt) + q * logf(q));
  }
  const float E = ex2_approx(-DL * kLog2e);
  const float AT = 1.f - E;
  const float DLoss = fmaf(alpha, pt * logp, one_m_alpha * (q * lq));
  if (kFast) {
    if (kLoss) li = -(AT * AT) * DLoss;
    if (kGrad) {
      const float r = rcp_approx(u);
      const float p = x >= 0.f ? r : e * r;
      const float d = pt - p;
      const float g2 = fmaf(alpha, d, -one_m_2alpha * (q * p));
      g = AT * fmaf(2.f * d * E, DLoss, -AT * g2);
    }
  } else {
    const float atg = powf(AT, gamma);
    if (kLoss) li = -atg * DLoss;
    if (kGrad) {
      const float r = rcp_approx(u);
      const float p = x >= 0.f ? r : e * r;
      const float d = pt - p;
      const float g2 = fmaf(alpha, d, -one_m_2alpha * (q * p));
      g = d * gamma * powf(AT, gamma - 1.f) * E * DLoss - atg * g2;
    }
  }
}

// -------------------------------------------------------------------------------------------
// lean fast path (gamma == 2, beta == 0), base-2 logs: 4 MUFU + ~29 FP32 instructions per element
// for loss + gradient with alpha == 0.5 (the headline configs), ~33 for general alpha.
//   x2 = x*log2(e), e = 2^-|x2|, l = log2(1+e), a0 = log2 p, b = log2(1-p)
//   S  = pt*a0 + (1-pt)*b = -DL*log2(e)  ->  E = 2^S = exp(-DL), AT = 1-E
//   D2 = 2*DLoss = 2*ln2*(alpha*pt*a + (1-alpha)*(1-pt)*b),  a = max(a0, log2 FLT_MIN)  (...loss_op.cu:63)
//   acc2 += -(AT^2)*D2*keep            (twice the loss summand; the caller halves the total)
//   g    = AT*(d*E*D2 - AT*g2),  d = pt-p,  g2 = alpha*d - (1-2alpha)(1-pt)p   (...loss_op.cu:97-99)
// a0 and b are formed from max(x2,0) separately (not b = a0 - x2) so neither cancels.
struct FastConsts {
  float ca2;   // 2*alpha*ln2
  float cb2;   // 2*(1-alpha)*ln2
  float alpha;
  float om2a;  // 1 - 2*alpha
};

template <bool kAlphaHalf, bool kLoss, bool kGrad>
__device__ __forceinline__ void distill_elem_fast(float x, float pt, float keep, float kk, const FastConsts& k,
                                                  float& acc2, float& gout) {
  const float x2 = x * kLog2e;
  const float e = ex2_approx(-fabsf(x2));
  const float u = 1.f + e;
  const float l = lg2_approx(u);
  const float mx = fmaxf(x2, 0.f);
  const float a0 = (x2 - mx) - l;
  const float b = -(mx + l);
  const float q = 1.f - pt;
  const float B = q * b;
  float S = fmaf(pt, a0, B);
  // the reference evaluates beta*(pt*logf(pt)+(1-pt)*logf(1-pt)) even for beta == 0: NaN unless 0 < pt < 1
  S = (pt > 0.f && pt < 1.f) ? S : __int_as_float(0x7fffffff);
  const float E = ex2_approx(S);
  const float AT = 1.f - E;
  const float a = fmaxf(a0, -126.f);
  float D2;
  if (kAlphaHalf) D2 = kLn2 * fmaf(pt, a, B);
  else D2 = fmaf(k.ca2 * pt, a, k.cb2 * B);
  if (kLoss) acc2 = fmaf(-(AT * AT) * D2, keep, acc2);  // NaN * 0 stays NaN, like the reference's `* (t != ignored)`
  if (kGrad) {
    const float r = rcp_approx(u);
    const float p = x >= 0.f ? r : e * r;
    const float d = pt - p;
    float g;
    if (kAlphaHalf) {
      g = (AT * d) * fmaf(E, D2, -0.5f * AT);
    } else {
      const float g2 = fmaf(k.alpha, d, -k.om2a * (q * p));
      g = AT * fmaf(d * E, D2, -AT * g2);
    }
    gout = g * kk;
  }
}

// -------------------------------------------------------------------------------------------
// packed fast path (gamma == 2, beta == 0, alpha == 0.5): TWO elements per instruction on the FMA pipe
// (fma/add/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2, sm_100), the 4 MUFU per element stay scalar.
// The r01 kernel was issue-bound (ncu profiles/r01h_distill_fused: issue active 54 %, 31 FP + 4 MUFU
// instructions per element in this function alone); per element this form issues
//   4.5 FMUL2 + 3.5 FADD2 + 2 FFMA2 + 4 MUFU + 4.5 scalar (2 FMNMX, FSETP + FSEL, half a FMNMX3) = 18.5.
// Same formulas as distill_elem_fast (comments there), with
//   * the "teacher probability outside (0, 1) -> NaN" rule (the reference evaluates pt*logf(pt) + (1-pt)*logf(1-pt)
//     even for beta == 0, ...loss_op.cu:59,93) taken out of the per-element code: vmin collects min(pt * (1 - pt))
//     over everything a thread touches (FMNMX3.NAN), and the caller re-runs the scalar function for the unit when
//     !(vmin > 0), i.e. when some pt was <= 0, >= 1 or NaN;
//   * ln2 and the ignore mask folded into the per-position factors: acc2 gathers AT^2 * D with D = D2 / ln2 per hw
//     position (the caller applies keep, -ln2/2 and 1/Np), the gradient is multiplied by kk2 = keep * kg * ln2.
// -------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float min3_nan(float a, float b, float c) {
  float d;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

struct PairConsts {
  f32x2 log2e, one, neg_c;   // (log2 e, log2 e), (1, 1), (-0.5 / ln2, -0.5 / ln2)
};
__device__ __forceinline__ PairConsts pair_consts() {
  PairConsts k;
  k.log2e = pk2(kLog2e, kLog2e);
  k.one = pk2(1.f, 1.f);
  k.neg_c = pk2(-0.5f / kLn2, -0.5f / kLn2);
  return k;
}

// two elements (x, pt packed): returns the two gradients times kk2; acc2 += AT^2 * D; vmin = min(vmin, pt * (1 - pt))
__device__ __forceinline__ f32x2 distill_pair_half(f32x2 x, f32x2 pt, f32x2 kk2, const PairConsts& k, f32x2& acc2, float& vmin) {
  const f32x2 x2 = mul2(x, k.log2e);
  float x2a, x2b;
  upk2(x2, x2a, x2b);
  const float ea = ex2_approx(-fabsf(x2a)), eb = ex2_approx(-fabsf(x2b));
  const f32x2 e = pk2(ea, eb);
  const f32x2 u = add2(e, k.one);
  float ua, ub;
  upk2(u, ua, ub);
  const f32x2 l = pk2(lg2_approx(ua), lg2_approx(ub));
  const f32x2 nmx = pk2(fminf(-x2a, 0.f), fminf(-x2b, 0.f));   // -max(x2, 0)
  const f32x2 a0 = sub2(add2(x2, nmx), l);                     // min(x2, 0) - l = log2 p
  const f32x2 b = sub2(nmx, l);                                // -(max(x2, 0) + l) = log2(1 - p)
  const f32x2 q = sub2(k.one, pt);
  const f32x2 B = mul2(q, b);
  const f32x2 S = fma2(pt, a0, B);                             // -DL * log2 e
  float va, vb;
  upk2(mul2(pt, q), va, vb);
  vmin = min3_nan(vmin, va, vb);
  float Sa, Sb;
  upk2(S, Sa, Sb);
  const f32x2 E = pk2(ex2_approx(Sa), ex2_approx(Sb));
  const f32x2 AT = sub2(k.one, E);
  float a0a, a0b;
  upk2(a0, a0a, a0b);
  const f32x2 D = fma2(pt, pk2(fmaxf(a0a, -126.f), fmaxf(a0b, -126.f)), B);   // FLT_MIN clamp on log p: loss term only (...loss_op.cu:63)
  acc2 = fma2(mul2(AT, AT), D, acc2);
  const float ra = rcp_approx(ua), rb = rcp_approx(ub);
  float era, erb;
  upk2(mul2(e, pk2(ra, rb)), era, erb);
  const f32x2 p = pk2(x2a >= 0.f ? ra : era, x2b >= 0.f ? rb : erb);
  const f32x2 d = sub2(pt, p);
  const f32x2 t = fma2(E, D, mul2(AT, k.neg_c));               // (E * D2 - 0.5 * AT) / ln2
  return mul2(mul2(mul2(AT, d), t), kk2);
}

// -------------------------------------------------------------------------------------------
// Exact accumulation of fp32 partial sums: a signed 128-bit fixed-point number, LSB = 2^-64, in two 64-bit limbs.
// Integer addition is associative, so the total is the same bits whatever the ORDER in which warps / CTAs deliver their
// partials — which is what lets work be handed out dynamically and still be bit-identical run to run — and a CTA that
// needs the total reads two words instead of re-adding hundreds of per-CTA partials.  Every fp32 value with |v| in
// [2^-40, 2^62) is represented exactly (24-bit mantissa shifted into place); smaller ones lose bits below 2^-64; larger
// ones, infinities and NaN raise a flag that the reader turns back into inf / NaN.
// -------------------------------------------------------------------------------------------
struct Fx128 {
  unsigned long long lo;
  unsigned long long hi;   // two's complement
};
constexpr unsigned int kFxNaN = 1u, kFxPosInf = 2u, kFxNegInf = 4u;

// returns 0, or the flag to raise when v is not representable
__device__ __forceinline__ unsigned int fx_from_float(float v, Fx128& out) {
  const uint32_t b = __float_as_uint(v);
  const uint32_t e = (b >> 23) & 0xffu;
  if (e >= 127u + 62u) {   // |v| >= 2^62, inf, NaN
    out.lo = out.hi = 0ull;
    if (e == 0xffu && (b & 0x7fffffu)) return kFxNaN;
    return (b >> 31) ? kFxNegInf : kFxPosInf;
  }
  const unsigned long long m = (unsigned long long)((b & 0x7fffffu) | (e ? 0x800000u : 0u));
  const int shift = (int)(e ? e : 1u) - 86;   // v = m * 2^(e - 150) = (m << (e - 86)) * 2^-64
  unsigned long long l, h;
  if (shift <= -24) {
    l = 0ull;
    h = 0ull;
  } else if (shift < 0) {
    l = m >> (-shift);
    h = 0ull;
  } else if (shift == 0) {
    l = m;
    h = 0ull;
  } else if (shift < 64) {
    l = m << shift;
    h = m >> (64 - shift);
  } else {
    l = 0ull;
    h = m << (shift - 64);
  }
  if (b >> 31) {   // negate (two's complement over 128 bits)
    l = ~l + 1ull;
    h = ~h + (l == 0ull ? 1ull : 0ull);
  }
  out.lo = l;
  out.hi = h;
  return 0u;
}
// acc += x (acc in shared or global memory); exact whatever the interleaving with other adders: the low limbs add up modulo
// 2^64 and every wrap is seen by exactly one adder, which carries it into the high limb
__device__ __forceinline__ void fx_atomic_add(unsigned long long* acc /* [lo, hi] */, const Fx128& x) {
  unsigned long long carry = 0ull;
  if (x.lo) {
    const unsigned long long old = atomicAdd(acc, x.lo);
    carry = (old + x.lo < old) ? 1ull : 0ull;
  }
  if (x.hi + carry) atomicAdd(acc + 1, x.hi + carry);
}
__device__ __forceinline__ double fx_to_double(unsigned long long lo, unsigned long long hi, unsigned int flags) {
  if (flags & kFxNaN) return __longlong_as_double(0x7ff8000000000000ll);
  if ((flags & kFxPosInf) && (flags & kFxNegInf)) return __longlong_as_double(0x7ff8000000000000ll);
  if (flags & kFxPosInf) return __longlong_as_double(0x7ff0000000000000ll);
  if (flags & kFxNegInf) return __longlong_as_double(0xfff0000000000000ll);
  return (double)(long long)hi + (double)lo * 5.421010862427522e-20;   // 2^-64
}

// Sum over a group of kThreads threads that share named barrier `bar_id`; result valid in the group's thread 0.
template <int kThreads, typename T>
__device__ __forceinline__ T group_sum(T v, T* smem, int tid, uint32_t bar_id) {
  constexpr int kWarps = kThreads / 32;
  v = warp_sum(v);
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) smem[warp] = v;
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(kThreads) : "memory");
  T r = 0;
  if (warp == 0) {
    r = lane < kWarps ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(kThreads) : "memory");
  return r;
}

// ---------------------------------------------------------------------------------------------
// Two-stage deterministic reduction tail shared by the ring kernels.
// Stage 1: every CTA publishes kSlots sums (from shared memory) and takes a ticket.
// Stage 2: the CTA holding the last ticket adds the published values per slot in a fixed order
// (lane-strided over CTA index, then a butterfly), one WARP PER SLOT in parallel, in fp64.
// The first version of this tail walked the slots one after the other with two block barriers and
// block-wide fp64 shuffles each: ~12 k cycles (6 us) during which 147 SMs idled (ncu r01c:
// sm__cycles_active max 89.6 k vs avg 77.6 k of 97.6 k elapsed).
// ---------------------------------------------------------------------------------------------
template <int kSlots>
__device__ __forceinline__ bool publish_and_ticket(const float* vals_smem, float* partials, unsigned int* counter,
                                                   bool* is_last_smem) {
  __syncthreads();  // vals_smem complete
  const int tid = threadIdx.x;
  if (tid < 32) {
    if (tid < kSlots) {
      partials[(size_t)blockIdx.x * kSlots + tid] = vals_smem[tid];
      __threadfence();
    }
    __syncwarp();
    if (tid == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(counter, 1u);
      *is_last_smem = ticket == gridDim.x - 1;
    }
  }
  __syncthreads();
  return *is_last_smem;
}

// fp64 sum over CTAs of slot k, executed by one full warp; result valid in every lane.  The loads of up to 384 CTAs'
// partials are issued together (one L2 round trip instead of one per 32 CTAs); the additions keep a fixed order.
template <int kSlots>
__device__ __forceinline__ double warp_sum_partials(const float* partials, int k, int lane) {
  constexpr int kBatch = 12;
  float v[kBatch];
#pragma unroll
  for (int i = 0; i < kBatch; ++i) {
    const uint32_t b = (uint32_t)lane + 32u * i;
    v[i] = b < gridDim.x ? __ldcg(partials + (size_t)b * kSlots + k) : 0.f;
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kBatch; ++i) s += (double)v[i];
  for (uint32_t b = (uint32_t)lane + 32u * kBatch; b < gridDim.x; b += 32) s += (double)__ldcg(partials + (size_t)b * kSlots + k);
  return warp_sum(s);
}

}  // namespace sad
#endif
