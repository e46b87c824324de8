// Weight and bias gradient of the RetinaNet head's 3x3 / stride 1 / pad 1 convolution on the tcgen05
// tensor cores (kind::tf32, fp32 accumulators in TMEM), every FPN level in ONE launch.
//
// Replaces, for the head's shapes, the filter/bias half of
//   CudnnConvGradientOp::DoRunWithType   caffe2/caffe2/operators/conv_op_cudnn.cc:645-1100
//     (cudnnConvolutionBackwardBias :1011-1020, cudnnConvolutionBackwardFilter :1022-1040)
// and the autograd `Sum` that adds the five per-level dW / db of a weight shared across levels
// (caffe2/caffe2/python/core.py:695,706-842; retinanet_heads.py:90-152: levels 4..7 use ConvShared),
// with the semantics of ConvGradientOp NCHW (conv_op_impl.h:182-420):
//     dW[co][ci][ky][kx] = sum_{l,n,y,x} dY_l[n][co][y][x] * X_l[n][ci][y + ky - 1][x + kx - 1]   (zero padding)
//     db[co]             = sum_{l,n,y,x} dY_l[n][co][y][x]
//
// Formulation: per filter tap one GEMM whose reduction axis is the pixel axis,
//     D_tap[co, ci] = sum_pixels dYt[pixel][co] * Xt[pixel + shift(tap)][ci]
//   * both operands are read from the channels-last tensors the forward / data-gradient kernels already
//     produce ((N, H, W, C), tf32-rounded), so both are "MN-major" UMMA operands: a TMA box
//     {32 channels, 32 x, 1 y, 1 n} lands in shared memory as 32 pixel rows of 128 B with the
//     "128-byte swizzle, 32-byte atom" pattern (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) = the only MN-major
//     layout tf32 operands have (UMMA layout type SWIZZLE_128B_BASE32B: 4 K-rows x 128 B groups 512 B
//     apart, channel chunks LBO = 4096 B apart; with the plain 128-byte swizzle the MMA returns zeros, and
//     (LBO, SBO, K step) = (4096, 512, 1024) is the only combination of the probed ones that is exact —
//     measured on B200, scripts/wgrad_probe.py history in DESIGN.md).  The tap shift is applied to the X box coordinate (x + dx, y + dy); the TMA
//     unit zero-fills out-of-bounds pixels, which implements the padding and ragged image edges.
//   * tile: M = 128 output channels (4 boxes) x N = 256 input channels (8 boxes), K = 32 pixels per
//     pipeline stage (48 KB), 4 stages; one elected thread issues 4 tcgen05.mma 128x256x8 per stage.
//   * work item = (tap, M tile, N tile, K split): the pixel blocks of all levels form one K axis which
//     is cut into `splits` equal ranges so that items ~ 1 or 2 waves of the 148 SMs.  Each item's
//     accumulator is written to a partial buffer [split][tap][co][ci]; conv3x3_wgrad_finish_kernel adds
//     the splits in a fixed order (deterministic, no atomics) and writes dW in the operator's
//     (Cout, Cin, 3, 3) layout.  db partials come from bias_grad_partial_kernel and are finished by
//     the same kernel.
//
// fp16 operands (template parameter kF16, entry point sad_conv3x3_wgrad_f16; BASELINE.json configs[4]: mixed fp16 compute): the same
// formulation on tcgen05.mma kind::f16.  16-bit MN-major operands use the PLAIN 128-byte swizzle (TMA CU_TENSOR_MAP_SWIZZLE_128B,
// UMMA layout type 2): chunk rows hold 64 channels, 8-pixel swizzle atoms are 1024 B apart (SBO), chunks kWgChunkBytes apart
// (LBO), one MMA covers 16 pixels (+2048 B).  Accumulators, partial buffers and the finished gradients stay fp32.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_common.cuh"
#include "sad_b200.h"
#include "sad_internal.h"
#include "tc_utils.cuh"

namespace sad {

constexpr int kWgM = 128;                 // output channels per tile
constexpr int kWgN = 256;                 // input channels per tile
constexpr int kWgKP = 32;                 // pixels per stage (one row segment)
constexpr int kWgStages = 4;
constexpr int kWgChunkBytes = kWgKP * 128;               // one {32 ch x 32 px} box: 4 KB
constexpr int kWgABytes = (kWgM / 32) * kWgChunkBytes;   // 16 KB
constexpr int kWgBBytes = (kWgN / 32) * kWgChunkBytes;   // 32 KB
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;     // 48 KB
// fp16 operands (kF16): a 128-byte chunk row holds 64 channels and one MMA covers 16 pixels.  A stage therefore takes TWO image
// rows of a 32-pixel segment (TMA box {64 ch, 32 x, 2 y}): 64 pixels x (2 + 4 chunks) = the same 48 KB and 4 MMAs per stage as the
// tf32 form.  (With 32-pixel stages — 24 KB, 2 MMAs — the kernel was bound by the producer's per-stage work, not by the tensor
// pipe: 38 % active, 48 us against the tf32 form's 51 us; profiles/r01m_f16_kernels_digest.txt.)
template <bool kF16>
struct WgTraits {
  static constexpr int kCh = kF16 ? 64 : 32;                        // channels per 128-byte chunk row
  static constexpr int kRows = kF16 ? 2 : 1;                        // image rows per stage
  static constexpr int kKP = kWgKP * kRows;                         // pixels per stage
  static constexpr int kChunkBytes = kKP * 128;                     // one {kCh channels x kKP pixels} box
  static constexpr int kAChunks = kWgM / kCh, kBChunks = kWgN / kCh;
  static constexpr int kABytes = kAChunks * kChunkBytes, kBBytes = kBChunks * kChunkBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kWgStages;
  static constexpr int kKStep = kF16 ? 16 : 8;                      // pixels per MMA
};
static_assert(WgTraits<true>::kStageBytes == kWgStageBytes && WgTraits<false>::kStageBytes == kWgStageBytes, "same stage bytes");
constexpr int kWgEpiWarps = 8;                           // two warps per TMEM lane quarter, each draining 4 of the 8 column blocks
constexpr int kWgThreads = 64 + 32 * kWgEpiWarps;        // warp 0: TMA, warp 1: MMA + TMEM, warps 2-9: epilogue
constexpr int kWgTmemCols = 512;                         // 2 accumulator buffers x 256 columns
constexpr size_t kWgSmemBytes = (size_t)kWgStages * kWgStageBytes + 1024 + 256;
constexpr int kWgMaxSplits = 64;
constexpr int kBgThreads = 256;
constexpr int kBgMaxBlocks = 1024;

struct WgLevel {
  int32_t N, H, W;
  uint32_t xsegs;                     // ceil(W / 32)
  uint32_t yblocks;                   // ceil(H / rows per stage)
  uint32_t block_begin, block_end;    // pixel blocks (n, y block, xseg) of this level on the global K axis
};
struct alignas(64) WgArgs {
  CUtensorMap tmap_dy[SAD_MAX_LEVELS];   // 4-D {C, W, H, N}, box {32, 32, 1, 1}: one 32-channel chunk (tail tiles); fp16: 64 channels
  CUtensorMap tmap_x[SAD_MAX_LEVELS];
  CUtensorMap tmap_dy5[SAD_MAX_LEVELS];  // 5-D {32, W, H, N, C/32}, box {32, 32, 1, 1, 4}: a whole M tile in one copy
  CUtensorMap tmap_x5[SAD_MAX_LEVELS];   // 5-D, box {32, 32, 1, 1, 8}: a whole N tile in one copy
  WgLevel lv[SAD_MAX_LEVELS];
  float* partial;                     // [splits][9][cout][cin]
  int32_t n_levels, cin, cout;
  uint32_t m_tiles, n_tiles, splits, total_blocks, total_items;
  // 3xTF32 (sad_conv3x3_wgrad_f32x3): dY and X are split tensors [hi | lo] (conv3x3.cu, kX3).  Every pixel block is staged npass = 3
  // times — (dY_hi, X_hi), (dY_hi, X_lo), (dY_lo, X_hi) — into the same accumulator; only the producer's channel coordinates move.
  uint32_t npass;            // 1, or 3
  int32_t dy_lo, x_lo;       // channel offset of the lo half: round_up(cout, 32) / round_up(cin, 32)
};

struct WgItem {
  int tap, m0, n0;
  uint32_t split, kb_begin, kb_end;
};
// item index: tile fastest (so CTAs running concurrently read the same pixel range), split slowest
__device__ __forceinline__ WgItem wg_decode_item(const WgArgs& a, uint32_t item) {
  const uint32_t tiles = 9u * a.m_tiles * a.n_tiles;
  uint32_t r = item % tiles;
  WgItem it;
  it.split = item / tiles;
  it.n0 = (int)(r % a.n_tiles) * kWgN;
  r /= a.n_tiles;
  it.m0 = (int)(r % a.m_tiles) * kWgM;
  it.tap = (int)(r / a.m_tiles);
  it.kb_begin = (uint32_t)(((uint64_t)a.total_blocks * it.split) / a.splits);
  it.kb_end = (uint32_t)(((uint64_t)a.total_blocks * (it.split + 1)) / a.splits);
  return it;
}

template <bool kF16>
__global__ void __launch_bounds__(kWgThreads, 1) conv3x3_wgrad_tf32_kernel(const __grid_constant__ WgArgs args) {
  using TR = WgTraits<kF16>;
  constexpr int kCh = TR::kCh, kStages = TR::kStages, kStageBytes = TR::kStageBytes, kABytes = TR::kABytes;
  constexpr int kChunkBytes = TR::kChunkBytes, kRows = TR::kRows;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < args.n_levels; ++l) {
      tma_prefetch_desc(&args.tmap_dy[l]);
      tma_prefetch_desc(&args.tmap_x[l]);
      tma_prefetch_desc(&args.tmap_dy5[l]);
      tma_prefetch_desc(&args.tmap_x5[l]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full[b], 1);
        mbar_init(&tmem_empty[b], 32 * kWgEpiWarps);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<kWgTmemCols>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      RingState rs;
      for (uint32_t item = blockIdx.x; item < args.total_items; item += gridDim.x) {
        const WgItem it = wg_decode_item(args, item);
        const int dy = it.tap / 3 - 1, dx = it.tap % 3 - 1;
        // locate the first pixel block, then walk (xseg, y, n, level) incrementally
        if (it.kb_begin >= it.kb_end) continue;
        int l = 0;
        while (it.kb_begin >= args.lv[l].block_end) ++l;
        uint32_t r = it.kb_begin - args.lv[l].block_begin;
        int xs = (int)(r % args.lv[l].xsegs);
        r /= args.lv[l].xsegs;
        int y = (int)(r % args.lv[l].yblocks);   // row block: image rows y * kRows .. + kRows - 1 (rows past H are zero-filled)
        int n = (int)(r / args.lv[l].yblocks);
        const int a_left = (args.cout - it.m0 + kCh - 1) / kCh, b_left = (args.cin - it.n0 + kCh - 1) / kCh;
        const int a_chunks = a_left < TR::kAChunks ? a_left : TR::kAChunks, b_chunks = b_left < TR::kBChunks ? b_left : TR::kBChunks;
        const bool a_full = it.m0 + kWgM <= args.cout, b_full = it.n0 + kWgN <= args.cin;  // whole tile inside the tensor
        for (uint32_t kb = it.kb_begin; kb < it.kb_end; ++kb) {
         for (uint32_t pass = 0; pass < args.npass; ++pass) {
          const int ma = it.m0 + (pass == 2 ? args.dy_lo : 0);   // first dY channel of this stage (lo half on pass 2)
          const int nb = it.n0 + (pass == 1 ? args.x_lo : 0);    // first X channel of this stage (lo half on pass 1)
          mbar_wait(&empty_bar[rs.stage], rs.phase ^ 1u);
          uint8_t* sa = smem + (size_t)rs.stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          // 32-channel chunks that lie entirely past Cout / Cin are not loaded: whatever the stage holds there
          // only reaches accumulator rows / columns the epilogue never writes out
          mbar_arrive_expect_tx(&full_bar[rs.stage], (uint32_t)(a_chunks + b_chunks) * kChunkBytes);
          if (a_full) {  // the tile's 4 chunks in one 5-D copy: [chunk][pixel][32 channels], chunks 4 KB apart
            tma_load_5d(sa, &args.tmap_dy5[l], &full_bar[rs.stage], 0, xs * kWgKP, y * kRows, n, ma / kCh);
          } else {
#pragma unroll
            for (int j = 0; j < TR::kAChunks; ++j)
              if (j < a_chunks)
                tma_load_4d(sa + j * kChunkBytes, &args.tmap_dy[l], &full_bar[rs.stage], ma + kCh * j, xs * kWgKP, y * kRows, n);
          }
          if (b_full) {
            tma_load_5d(sb, &args.tmap_x5[l], &full_bar[rs.stage], 0, xs * kWgKP + dx, y * kRows + dy, n, nb / kCh);
          } else {
#pragma unroll
            for (int j = 0; j < TR::kBChunks; ++j)
              if (j < b_chunks)
                tma_load_4d(sb + j * kChunkBytes, &args.tmap_x[l], &full_bar[rs.stage], nb + kCh * j, xs * kWgKP + dx, y * kRows + dy, n);
          }
          rs.advance<kStages>();
         }
          if (++xs == (int)args.lv[l].xsegs) {
            xs = 0;
            if (++y == (int)args.lv[l].yblocks) {
              y = 0;
              if (++n == args.lv[l].N) {
                n = 0;
                ++l;
                while (l < args.n_levels && args.lv[l].block_end == args.lv[l].block_begin) ++l;  // empty levels
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = kF16 ? umma_idesc_f16(kWgM, kWgN, /*A MN-major*/ 1, /*B MN-major*/ 1)
                                      : umma_idesc_tf32(kWgM, kWgN, /*A MN-major*/ 1, /*B MN-major*/ 1);
      RingState rs;
      uint32_t itn = 0;
      for (uint32_t item = blockIdx.x; item < args.total_items; item += gridDim.x, ++itn) {
        const WgItem it = wg_decode_item(args, item);
        const uint32_t buf = itn & 1u, aphase = (itn >> 1) & 1u;
        mbar_wait(&tmem_empty[buf], aphase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * kWgN;
        uint32_t first = 1;
        const uint32_t n_stages = (it.kb_end - it.kb_begin) * args.npass;
        for (uint32_t kb = 0; kb < n_stages; ++kb) {
          mbar_wait(&full_bar[rs.stage], rs.phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + (size_t)rs.stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < TR::kKP / TR::kKStep; ++k) {
            if (kF16) {
              // MN-major 16-bit operands take the plain 128-byte swizzle (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
              // 16-byte units): one K step = 16 pixel rows of 128 B = two 8-row swizzle atoms 1024 B apart (SBO); the
              // 64-channel chunks of the M / N axis are kChunkBytes (64 pixel rows) apart (LBO)
              const uint64_t adesc = umma_smem_desc_sw128(a_addr + k * 2048, kChunkBytes, 1024);
              const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 2048, kChunkBytes, 1024);
              umma_f16(d_tmem, adesc, bdesc, idesc, first ? 0u : 1u);
            } else {
              // MN-major tf32 operands: one K step = 8 pixel rows of 128 B = two 4-row swizzle groups 512 B apart (SBO);
              // the 32-channel chunks of the M / N axis are kWgChunkBytes apart (LBO)
              const uint64_t adesc = umma_smem_desc_sw128_base32(a_addr + k * 1024, kWgChunkBytes, 512);
              const uint64_t bdesc = umma_smem_desc_sw128_base32(b_addr + k * 1024, kWgChunkBytes, 512);
              umma_tf32(d_tmem, adesc, bdesc, idesc, first ? 0u : 1u);
            }
            first = 0;
          }
          umma_commit(&empty_bar[rs.stage]);
          rs.advance<kStages>();
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> partial[split][tap][co][ci] =====================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int jhalf = (warp - 2) >> 2;   // which half of the accumulator's column blocks this warp drains
    const int row = q * 32 + lane;
    uint32_t itn = 0;
    for (uint32_t item = blockIdx.x; item < args.total_items; item += gridDim.x, ++itn) {
      const WgItem it = wg_decode_item(args, item);
      const uint32_t buf = itn & 1u, aphase = (itn >> 1) & 1u;
      const int co = it.m0 + row;
      const bool co_ok = co < args.cout;
      mbar_wait(&tmem_full[buf], aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kWgN;
      float* dst = args.partial + (((size_t)it.split * 9 + it.tap) * args.cout + (co_ok ? co : 0)) * args.cin + it.n0;
      const bool empty = it.kb_end == it.kb_begin;  // a split with no pixel blocks: the accumulator was never written
#pragma unroll 1
      for (int j = jhalf * (kWgN / 64); j < (jhalf + 1) * (kWgN / 64); ++j) {
        float v[32];
        tmem_ld_32x32(taddr + j * 32, v);
        if (co_ok) {
          if (((args.cin & 3) == 0) && it.n0 + j * 32 + 32 <= args.cin) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + j * 32 + i) =
                  empty ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (it.n0 + j * 32 + i < args.cin) dst[j * 32 + i] = empty ? 0.f : v[i];
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&tmem_empty[buf]);
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<kWgTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT path for channel counts the TMA maps cannot address (C % 4 != 0) or unaligned tensors: one thread
// per (tap, co, ci), pixels summed serially.  Same partial layout with splits = 1.
// ---------------------------------------------------------------------------------------------
struct WgSimtArgs {
  const float* x[SAD_MAX_LEVELS];
  const float* dy[SAD_MAX_LEVELS];
  int32_t N[SAD_MAX_LEVELS], H[SAD_MAX_LEVELS], W[SAD_MAX_LEVELS];
  float* partial;
  int32_t n_levels, cin, cout;
};
__global__ void conv3x3_wgrad_simt_kernel(const WgSimtArgs a) {
  const size_t total = (size_t)9 * a.cout * a.cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % a.cin);
    const int co = (int)((i / a.cin) % a.cout);
    const int tap = (int)(i / ((size_t)a.cin * a.cout));
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    float acc = 0.f;
    for (int l = 0; l < a.n_levels; ++l) {
      const int H = a.H[l], W = a.W[l];
      for (int n = 0; n < a.N[l]; ++n)
        for (int y = 0; y < H; ++y) {
          const int sy = y + dy;
          if (sy < 0 || sy >= H) continue;
          for (int x = 0; x < W; ++x) {
            const int sx = x + dx;
            if (sx < 0 || sx >= W) continue;
            acc = fmaf(a.dy[l][(((size_t)n * H + y) * W + x) * a.cout + co], a.x[l][(((size_t)n * H + sy) * W + sx) * a.cin + ci], acc);
          }
        }
    }
    a.partial[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// db partials: block b sums the pixels [b * chunk, (b + 1) * chunk) of the concatenated levels for every
// channel.  256 threads = 4 pixel lanes x 64 channel quads: 128-bit loads along the channels-last rows,
// the 4 pixel lanes are combined through shared memory in a fixed order.
// ---------------------------------------------------------------------------------------------
struct BgArgs {
  const float* dy[SAD_MAX_LEVELS];         // fp16 instantiation: __half storage
  uint32_t pix_begin[SAD_MAX_LEVELS + 1];  // prefix of N*H*W over levels
  float* partial;                          // [blocks][cout]
  int32_t n_levels, cout;
  uint32_t chunk;
  int32_t row, lo_off;   // floats per pixel row of dY (= cout unless the tensor is a split one: 2 * round_up(cout, 32)); lo_off != 0: offset
                         // of the lo half, which is added in (dY = hi + lo)
};
__device__ __forceinline__ float4 bg_load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 bg_load4(const __half* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float bg_load1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float bg_load1(const __half* p) { return __half2float(*p); }

template <typename InT>
__global__ void __launch_bounds__(kBgThreads) bias_grad_partial_kernel(const BgArgs a) {
  __shared__ float4 red[kBgThreads];
  const uint32_t total = a.pix_begin[a.n_levels];
  const uint32_t p0 = blockIdx.x * a.chunk;
  const uint32_t p1 = p0 + a.chunk < total ? p0 + a.chunk : total;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  if ((a.cout & 3) == 0) {
    const int quads = a.cout >> 2;
    for (int q0 = 0; q0 < quads; q0 += 64) {
      const int q = q0 + tx;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < quads) {
        for (int l = 0; l < a.n_levels; ++l) {
          const uint32_t lo = p0 > a.pix_begin[l] ? p0 : a.pix_begin[l];
          const uint32_t hi = p1 < a.pix_begin[l + 1] ? p1 : a.pix_begin[l + 1];
          const InT* base = reinterpret_cast<const InT*>(a.dy[l]) - (size_t)a.pix_begin[l] * a.row + (size_t)q * 4;
#pragma unroll 4
          for (uint32_t p = lo + ty; p < hi; p += 4) {
            float4 v = bg_load4(base + (size_t)p * a.row);
            if (a.lo_off) {   // split tensor: the element is hi + lo (exact in fp32)
              const float4 w = bg_load4(base + (size_t)p * a.row + a.lo_off);
              v.x += w.x;
              v.y += w.y;
              v.z += w.z;
              v.w += w.w;
            }
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
          }
        }
      }
      red[threadIdx.x] = acc;
      __syncthreads();
      if (ty == 0 && q < quads) {
        float4 r = red[tx];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          const float4 o = red[tx + 64 * k];
          r.x += o.x;
          r.y += o.y;
          r.z += o.z;
          r.w += o.w;
        }
        *reinterpret_cast<float4*>(a.partial + (size_t)blockIdx.x * a.cout + (size_t)q * 4) = r;
      }
      __syncthreads();
    }
  } else {  // channel counts that are not a multiple of 4: scalar loads, one thread per channel
    for (int c = threadIdx.x; c < a.cout; c += kBgThreads) {
      float acc = 0.f;
      int l = 0;
      for (uint32_t p = p0; p < p1; ++p) {
        while (p >= a.pix_begin[l + 1]) ++l;
        const InT* e = reinterpret_cast<const InT*>(a.dy[l]) + (size_t)(p - a.pix_begin[l]) * a.row + c;
        acc += a.lo_off ? bg_load1(e) + bg_load1(e + a.lo_off) : bg_load1(e);
      }
      a.partial[(size_t)blockIdx.x * a.cout + c] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// finish: dW[co][ci][tap] (+)= sum_s partial[s][tap][co][ci];  db[co] (+)= sum_b bias_partial[b][co]
// A block of 288 threads = 9 taps (one warp each) x 32 consecutive (co, ci): coalesced 128-byte partial
// reads, fixed summation order, and the 288 results leave through shared memory as one contiguous
// run of the (Cout, Cin, 3, 3) tensor.
// ---------------------------------------------------------------------------------------------
constexpr int kFinThreads = 288;
// cout_src >= cout: channel count of the partial buffers (the fp16 path pads the gradient tensors' channels to a multiple of 8);
// out_scale: 1 / loss scale of the fp16 path (a power of two), 1 otherwise.
__global__ void __launch_bounds__(kFinThreads) conv3x3_wgrad_finish_kernel(const float* __restrict__ partial, int splits, int cin, int cout,
                                                                          int cout_src, float out_scale, float* __restrict__ d_weight,
                                                                          const float* __restrict__ bias_partial, int bias_blocks,
                                                                          float* __restrict__ d_bias, int accumulate) {
  __shared__ float out[kFinThreads];
  const size_t plane = (size_t)cout * cin, plane_src = (size_t)cout_src * cin;
  const int tap = threadIdx.x >> 5, ii = threadIdx.x & 31;
  const size_t groups = (plane + 31) / 32;
  for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
    const size_t i = g * 32 + ii;  // i = co * cin + ci
    float acc = 0.f;
    if (i < plane) {
      const float* src = partial + (size_t)tap * plane_src + i;   // i = co * cin + ci is the same in both planes (co < cout)
#pragma unroll 4
      for (int s = 0; s < splits; ++s) acc += __ldg(src + (size_t)s * 9 * plane_src);
    }
    out[ii * 9 + tap] = acc * out_scale;
    __syncthreads();
    const size_t o = g * 288 + threadIdx.x;
    if (o < 9 * plane) d_weight[o] = accumulate ? d_weight[o] + out[threadIdx.x] : out[threadIdx.x];
    __syncthreads();
  }
  if (d_bias) {
    // one warp per channel: lanes stride over the partial blocks, then a fixed-order shuffle tree
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t c = warp; c < (size_t)cout; c += warps) {
      float acc = 0.f;
      for (int b = lane; b < bias_blocks; b += 32) acc += __ldg(bias_partial + (size_t)b * cout_src + c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      acc *= out_scale;
      if (lane == 0) d_bias[c] = accumulate ? d_bias[c] + acc : acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct WgPlan {
  uint32_t m_tiles, n_tiles, tiles, splits, total_blocks, bias_blocks, bias_chunk;
  uint64_t pixels;
  size_t partial_bytes, bias_partial_bytes;
};

static int wg_plan(const sad_wgrad_level* levels, int n_levels, int cin, int cout, int sms, WgPlan* p, int rows = 1) {
  if (!levels || n_levels < 1 || n_levels > SAD_MAX_LEVELS) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: n_levels must be in [1, 8]");
  if (cin < 1 || cout < 1) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: bad channel counts");
  uint64_t blocks = 0, pixels = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_wgrad_level& L = levels[l];
    if (L.N < 0 || L.H < 0 || L.W < 0) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: negative dimension");
    blocks += (uint64_t)L.N * ((L.H + rows - 1) / rows) * ((L.W + kWgKP - 1) / kWgKP);
    pixels += (uint64_t)L.N * L.H * L.W;
  }
  if (blocks > 0x7fffffffull || pixels > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: too many pixels");
  p->m_tiles = (uint32_t)((cout + kWgM - 1) / kWgM);
  p->n_tiles = (uint32_t)((cin + kWgN - 1) / kWgN);
  p->tiles = 9u * p->m_tiles * p->n_tiles;
  p->total_blocks = (uint32_t)blocks;
  p->pixels = pixels;
  // K splits: items = tiles * splits should fill one or two waves of the SMs; prefer the fewer splits
  // (less partial traffic) unless two waves use the machine clearly better
  auto util = [&](uint32_t s) {
    const uint64_t items = (uint64_t)p->tiles * s;
    const uint64_t waves = (items + sms - 1) / sms;
    return (double)items / (double)(waves * sms);
  };
  uint32_t s1 = (uint32_t)sms / p->tiles, s2 = (uint32_t)(2 * sms) / p->tiles;
  if (s1 < 1) s1 = 1;
  if (s2 < 1) s2 = 1;
  uint32_t s = util(s2) > util(s1) + 0.05 ? s2 : s1;
  const uint32_t by_work = p->total_blocks / 8 ? p->total_blocks / 8 : 1;  // at least 8 pixel blocks per item
  if (s > by_work) s = by_work;
  if (s > (uint32_t)kWgMaxSplits) s = kWgMaxSplits;
  p->splits = s;
  p->partial_bytes = (size_t)s * 9 * cout * cin * sizeof(float);
  uint32_t bb = (uint32_t)((pixels + 31) / 32);
  if (bb > (uint32_t)kBgMaxBlocks) bb = kBgMaxBlocks;
  if (bb < 1) bb = 1;
  p->bias_chunk = (uint32_t)((pixels + bb - 1) / bb);
  if (p->bias_chunk < 1) p->bias_chunk = 1;
  p->bias_blocks = (uint32_t)((pixels + p->bias_chunk - 1) / p->bias_chunk);
  if (p->bias_blocks < 1) p->bias_blocks = 1;
  p->bias_partial_bytes = (size_t)kBgMaxBlocks * cout * sizeof(float);
  return SAD_OK;
}

// channels-last (N, H, W, C) viewed as {32 c_lo, W, H, N, C/32 c_hi}: only whole 32-channel chunks are addressable,
// a box {32, box_x, 1, 1, chunks} lands as [chunk][pixel][32 channels]
static int encode_nhwc5_map(CUtensorMap* m, const float* xt, int N, int C, int H, int W, int box_x, int chunks, const char* what) {
  const cuuint64_t dims[5] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(C / 32)};
  const cuuint64_t str[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, 128};
  const cuuint32_t box[5] = {32, (cuuint32_t)box_x, 1, 1, (cuuint32_t)chunks};
  return encode_map(m, xt, 5, dims, str, box, what, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

// the same for fp16: {64 c_lo, W, H, N, C/64 c_hi}, plain 128-byte swizzle
static int encode_nhwc5_map_f16(CUtensorMap* m, const void* xt, int N, int C, int H, int W, int box_x, int box_y, int chunks, const char* what) {
  const cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(C / 64)};
  const cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, 128};
  const cuuint32_t box[5] = {64, (cuuint32_t)box_x, (cuuint32_t)box_y, 1, (cuuint32_t)chunks};
  return encode_map(m, xt, 5, dims, str, box, what, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}

static int wg_sms(int* sms) {
  // sizing must not depend on a device being present (workspace_bytes is callable on a CPU-only host):
  // fall back to the B200's 148 SMs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || *sms < 1) {
    cudaGetLastError();
    *sms = 148;
  }
  return SAD_OK;
}

}  // namespace sad

using namespace sad;

extern "C" {

SAD_EXPORT size_t sad_conv3x3_wgrad_workspace_bytes(const sad_wgrad_level* levels, int n_levels, int cin, int cout) {
  int sms = 0;
  wg_sms(&sms);
  WgPlan p;
  if (wg_plan(levels, n_levels, cin, cout, sms, &p) != SAD_OK) return 0;
  return ((p.partial_bytes + 255) / 256) * 256 + p.bias_partial_bytes;
}

// cout_out: channels of d_weight / d_bias; cout (>= cout_out): channel count of the dY tensors (the fp16 path pads it to a multiple
// of 8; the extra channels never leave the partial buffers); out_scale multiplies the finished gradients (1 / loss scale).
static int wgrad_impl(const sad_wgrad_level* levels, int n_levels, int cin, int cout, int cout_out, float out_scale, float* d_weight, float* d_bias,
                      int accumulate, void* workspace, size_t workspace_bytes, void* stream, bool f16, bool x3 = false) {
  int sms = 0, rc;
  if ((rc = sm_count(&sms)) != SAD_OK) return rc;
  WgPlan p;
  const int rows = f16 ? WgTraits<true>::kRows : 1;
  if ((rc = wg_plan(levels, n_levels, cin, cout, sms, &p, rows)) != SAD_OK) return rc;
  if (!d_weight) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: d_weight is null");
  if (cout_out < 1 || cout_out > cout) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: output channels must be in [1, dY channels]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool tma_ok = f16 ? (cin % 8 == 0) && (cout % 8 == 0) : (cin % 4 == 0) && (cout % 4 == 0);
  for (int l = 0; l < n_levels; ++l) {
    const sad_wgrad_level& L = levels[l];
    if ((uint64_t)L.N * L.H * L.W && (!L.x_nhwc || !L.dy_nhwc)) return set_error(SAD_ERR_INVALID, "conv3x3 wgrad: null tensor");
    if ((reinterpret_cast<uintptr_t>(L.x_nhwc) | reinterpret_cast<uintptr_t>(L.dy_nhwc)) & 15) tma_ok = false;
  }
  if (!tma_ok && f16)
    return set_error(SAD_ERR_UNSUPPORTED, "conv3x3 wgrad fp16: needs channel counts that are multiples of 8 and 16-byte aligned tensors");
  // 3xTF32: the tensors are split rows [hi | lo] of 2 * round_up(C, 32) floats; cin / cout stay the logical channel counts
  const int cin_s = (cin + 31) & ~31, cout_s = (cout + 31) & ~31;
  const int cin_map = x3 ? 2 * cin_s : cin, cout_map = x3 ? 2 * cout_s : cout;
  if (x3) {
    bool ok = true;
    for (int l = 0; l < n_levels; ++l)
      if ((reinterpret_cast<uintptr_t>(levels[l].x_nhwc) | reinterpret_cast<uintptr_t>(levels[l].dy_nhwc)) & 15) ok = false;
    if (!ok) return set_error(SAD_ERR_UNSUPPORTED, "conv3x3 wgrad f32x3: needs 16-byte aligned tensors (there is no SIMT split path)");
    tma_ok = true;
  }
  if (!tma_ok) p.splits = 1, p.partial_bytes = (size_t)9 * cout * cin * sizeof(float);
  const size_t partial_padded = ((p.partial_bytes + 255) / 256) * 256;
  if (!workspace || workspace_bytes < partial_padded + p.bias_partial_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(SAD_ERR_WORKSPACE, "conv3x3 wgrad: workspace too small or not 256-byte aligned (see sad_conv3x3_wgrad_workspace_bytes)");
  float* partial = static_cast<float*>(workspace);
  float* bias_partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + partial_padded);

  if (p.pixels == 0) {  // nothing to reduce: the gradients are zero
    if (!accumulate) {
      if ((rc = check_cuda(cudaMemsetAsync(d_weight, 0, (size_t)9 * cin * cout_out * sizeof(float), st), "memset dW")) != SAD_OK) return rc;
      if (d_bias && (rc = check_cuda(cudaMemsetAsync(d_bias, 0, (size_t)cout_out * sizeof(float), st), "memset db")) != SAD_OK) return rc;
    }
    return SAD_OK;
  }

  if (tma_ok) {
    WgArgs a{};
    uint64_t blocks = 0;
    int first_valid = -1;
    for (int l = 0; l < n_levels; ++l) {
      const sad_wgrad_level& L = levels[l];
      WgLevel& D = a.lv[l];
      D.N = L.N;
      D.H = L.H;
      D.W = L.W;
      D.xsegs = (uint32_t)((L.W + kWgKP - 1) / kWgKP);
      D.yblocks = (uint32_t)((L.H + rows - 1) / rows);
      D.block_begin = (uint32_t)blocks;
      blocks += (uint64_t)L.N * D.yblocks * D.xsegs;
      D.block_end = (uint32_t)blocks;
      if (D.block_end == D.block_begin) continue;
      if (f16) {
        if ((rc = encode_nhwc_map_f16(&a.tmap_dy[l], L.dy_nhwc, L.N, cout, L.H, L.W, kWgKP, rows, "wgrad fp16 dY {C,W,H,N}")) != SAD_OK) return rc;
        if ((rc = encode_nhwc_map_f16(&a.tmap_x[l], L.x_nhwc, L.N, cin, L.H, L.W, kWgKP, rows, "wgrad fp16 X {C,W,H,N}")) != SAD_OK) return rc;
        if (cout >= kWgM && (rc = encode_nhwc5_map_f16(&a.tmap_dy5[l], L.dy_nhwc, L.N, cout, L.H, L.W, kWgKP, rows, kWgM / 64, "wgrad fp16 dY 5-D")) != SAD_OK) return rc;
        if (cin >= kWgN && (rc = encode_nhwc5_map_f16(&a.tmap_x5[l], L.x_nhwc, L.N, cin, L.H, L.W, kWgKP, rows, kWgN / 64, "wgrad fp16 X 5-D")) != SAD_OK) return rc;
        if (cout < kWgM) a.tmap_dy5[l] = a.tmap_dy[l];
        if (cin < kWgN) a.tmap_x5[l] = a.tmap_x[l];
        if (first_valid < 0) first_valid = l;
        continue;
      }
      if ((rc = encode_nhwc_map(&a.tmap_dy[l], L.dy_nhwc, L.N, cout_map, L.H, L.W, kWgKP, 1, "wgrad dY {C,W,H,N}", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != SAD_OK) return rc;
      if ((rc = encode_nhwc_map(&a.tmap_x[l], L.x_nhwc, L.N, cin_map, L.H, L.W, kWgKP, 1, "wgrad X {C,W,H,N}", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != SAD_OK) return rc;
      // 5-D views {32 c_lo, W, H, N, C/32 c_hi} (c_hi stride 128 B) so one copy brings a whole tile of full chunks
      if (cout >= kWgM && (rc = encode_nhwc5_map(&a.tmap_dy5[l], L.dy_nhwc, L.N, cout_map, L.H, L.W, kWgKP, kWgM / 32, "wgrad dY 5-D")) != SAD_OK) return rc;
      if (cin >= kWgN && (rc = encode_nhwc5_map(&a.tmap_x5[l], L.x_nhwc, L.N, cin_map, L.H, L.W, kWgKP, kWgN / 32, "wgrad X 5-D")) != SAD_OK) return rc;
      if (cout < kWgM) a.tmap_dy5[l] = a.tmap_dy[l];  // never used (no full tile)
      if (cin < kWgN) a.tmap_x5[l] = a.tmap_x[l];
      if (first_valid < 0) first_valid = l;
    }
    for (int l = 0; l < n_levels; ++l)
      if (a.lv[l].block_end == a.lv[l].block_begin) {  // empty level: valid dummy maps, never dereferenced
        a.tmap_dy[l] = a.tmap_dy[first_valid];
        a.tmap_x[l] = a.tmap_x[first_valid];
        a.tmap_dy5[l] = a.tmap_dy5[first_valid];
        a.tmap_x5[l] = a.tmap_x5[first_valid];
      }
    a.partial = partial;
    a.n_levels = n_levels;
    a.cin = cin;
    a.cout = cout;
    a.m_tiles = p.m_tiles;
    a.n_tiles = p.n_tiles;
    a.splits = p.splits;
    a.total_blocks = p.total_blocks;
    a.total_items = p.tiles * p.splits;
    a.npass = x3 ? 3 : 1;
    a.dy_lo = cout_s;
    a.x_lo = cin_s;
    auto kern = f16 ? conv3x3_wgrad_tf32_kernel<true> : conv3x3_wgrad_tf32_kernel<false>;
    if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemBytes),
                         "cudaFuncSetAttribute(wgrad)")) != SAD_OK)
      return rc;
    const uint32_t grid = a.total_items < (uint32_t)sms ? a.total_items : (uint32_t)sms;
    kern<<<grid, kWgThreads, kWgSmemBytes, st>>>(a);
    count_launch(1);
    if ((rc = check_cuda(cudaGetLastError(), "conv3x3 wgrad launch")) != SAD_OK) return rc;
  } else {
    WgSimtArgs a{};
    for (int l = 0; l < n_levels; ++l) {
      a.x[l] = levels[l].x_nhwc;
      a.dy[l] = levels[l].dy_nhwc;
      a.N[l] = levels[l].N;
      a.H[l] = levels[l].H;
      a.W[l] = levels[l].W;
    }
    a.partial = partial;
    a.n_levels = n_levels;
    a.cin = cin;
    a.cout = cout;
    const size_t total = (size_t)9 * cin * cout;
    conv3x3_wgrad_simt_kernel<<<(unsigned)((total + 127) / 128 < 65535 ? (total + 127) / 128 : 65535), 128, 0, st>>>(a);
    count_launch(1);
    if ((rc = check_cuda(cudaGetLastError(), "conv3x3 wgrad simt launch")) != SAD_OK) return rc;
  }

  if (d_bias) {
    BgArgs b{};
    uint64_t pix = 0;
    for (int l = 0; l < n_levels; ++l) {
      b.dy[l] = levels[l].dy_nhwc;
      b.pix_begin[l] = (uint32_t)pix;
      pix += (uint64_t)levels[l].N * levels[l].H * levels[l].W;
    }
    b.pix_begin[n_levels] = (uint32_t)pix;
    b.partial = bias_partial;
    b.n_levels = n_levels;
    b.cout = cout;
    b.chunk = p.bias_chunk;
    b.row = cout_map;
    b.lo_off = x3 ? cout_s : 0;
    if (f16) bias_grad_partial_kernel<__half><<<p.bias_blocks, kBgThreads, 0, st>>>(b);
    else bias_grad_partial_kernel<float><<<p.bias_blocks, kBgThreads, 0, st>>>(b);
    count_launch(1);
    if ((rc = check_cuda(cudaGetLastError(), "bias grad launch")) != SAD_OK) return rc;
  }
  const size_t groups = ((size_t)cin * cout_out + 31) / 32;
  const unsigned fblocks = (unsigned)(groups < (size_t)sms * 16 ? groups : (size_t)sms * 16);
  conv3x3_wgrad_finish_kernel<<<fblocks, kFinThreads, 0, st>>>(partial, (int)p.splits, cin, cout_out, cout, out_scale, d_weight, bias_partial,
                                                       (int)p.bias_blocks, d_bias, accumulate);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "conv3x3 wgrad finish launch");
}

SAD_EXPORT int sad_conv3x3_wgrad_f32(const sad_wgrad_level* levels, int n_levels, int cin, int cout, float* d_weight, float* d_bias,
                                     int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return wgrad_impl(levels, n_levels, cin, cout, cout, 1.f, d_weight, d_bias, accumulate, workspace, workspace_bytes, stream, false);
}

SAD_EXPORT int sad_conv3x3_wgrad_f32x3(const sad_wgrad_level* levels, int n_levels, int cin, int cout, float* d_weight, float* d_bias,
                                       int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return wgrad_impl(levels, n_levels, cin, cout, cout, 1.f, d_weight, d_bias, accumulate, workspace, workspace_bytes, stream, false, true);
}

SAD_EXPORT int sad_conv3x3_wgrad_f16(const sad_wgrad_level* levels, int n_levels, int cin, int dy_channels, int cout, float out_scale,
                                     float* d_weight, float* d_bias, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return wgrad_impl(levels, n_levels, cin, dy_channels, cout, out_scale, d_weight, d_bias, accumulate, workspace, workspace_bytes, stream, true);
}

}  // extern "C"
