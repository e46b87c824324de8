// 3x3 / stride 1 / pad 1 NCHW fp32 convolution of the RetinaNet head on the 5th-generation tensor
// cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM), operands staged by TMA.
//
// Replaces, for the head's shapes, the reference's cuDNN calls
//   CudnnConvOp::RunOnDevice            caffe2/caffe2/operators/conv_op_cudnn.cc:293-643  (Y = conv(X, W) + b)
//   CudnnConvGradientOp (data gradient) caffe2/caffe2/operators/conv_op_cudnn.cc:645-1100 (dX = dY * W^T)
// with the semantics of ConvOp<T>::RunOnDeviceWithOrderNCHW (conv_op_impl.h:31-180: cross-correlation,
// zero padding) — weights (Cout, Cin, 3, 3), activations (N, C, H, W), all levels of the FPN pyramid in
// ONE launch because the head's weights are shared across levels (retinanet_heads.py:90-152,171-245).
//
// Formulation: implicit GEMM per filter tap, no im2col buffer:
//     D[co, pixel] = sum_tap sum_ci  Wp[tap][co][ci] * Xt[n][y + dy(tap)][x + dx(tap)][ci]
//   * M = 128 output channels per tile, N = 256 pixels per tile (8 image rows x 32 columns),
//     K = 32 input channels per pipeline stage, 9 * ceil(Cin/32) stages per tile.
//   * A (weights): repacked once per step by conv3x3_pack_kernel to [tap][Cout][Cin] (K-major) and
//     rounded to tf32 (round-to-nearest); one TMA box {32 ci, 128 co, 1 tap} per stage, 128B swizzle.
//   * B (activations): a channels-last (N, H, W, C) copy of the input (nchw_to_nhwc_kernel, or the
//     channels-last output a previous convolution wrote from its epilogue), read by ONE 4-D TMA box
//     {32 ci, 32 x, 8 y, 1 n} per stage at the tap-shifted coordinate (c0, x0+dx, y0+dy, n).
//     Out-of-bounds elements (negative or past the edge) are zero-filled by the TMA unit, which
//     implements the padding and the ragged tile edges.  In shared memory that is 256 pixel rows of
//     128 B = the canonical K-major 128B-swizzle UMMA operand.
//     Why not straight from NCHW: measured on B200 (scripts/probe/tma_probe.cu), a TMA tile whose
//     INNERMOST coordinate is not a multiple of 16 bytes raises "illegal instruction"; with NCHW the
//     dx = +-1 taps shift the innermost (x) coordinate by 4 bytes.  Channels-last puts the shifts on
//     outer dimensions, where any (also negative) coordinate is legal.
//   * One elected thread issues 4 tcgen05.mma (128x256x8) per stage; accumulators are double-buffered
//     in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   * Epilogue warps: tcgen05.ld 32 lanes x 32 columns -> + bias (-> ReLU) -> NCHW (the operator's
//     output blob; 128-bit stores) and/or channels-last (input of the next convolution / of the weight
//     gradient; one coalesced 128 B line per pixel and warp).
//   * Persistent grid (one CTA per SM), static round-robin tile order with the Cout-tile index fastest
//     so concurrently running CTAs share activation tiles in L2.
// Numerics: tf32 operands (10-bit mantissa, both rounded to nearest: weights in the pack kernel,
// activations when the channels-last copy is written), fp32 accumulation. Tolerance vs the fp32
// oracle is stated in tests/test_conv_gpu.py.
//
// Channel counts that are not a multiple of 4 (TMA needs 16-byte strides) take conv3x3_simt_kernel.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <mutex>
#include <string>

#include "conv_common.cuh"
#include "sad_b200.h"
#include "sad_internal.h"
#include "tc_utils.cuh"

namespace sad {

#ifndef SAD_CONV_ROWS
#define SAD_CONV_ROWS 8   // image rows per pixel tile (x 32 columns).  4 was measured too: see there
#endif
constexpr int kCvM = 128;
constexpr int kCvRows = SAD_CONV_ROWS;
constexpr int kCvCols = 32;
constexpr int kCvN = kCvRows * kCvCols;  // 256
static_assert(kCvRows == 8 || kCvRows == 4, "pixel tile: 8 or 4 rows of 32 pixels");
constexpr int kCvKC = 32;
constexpr int kCvABytes = kCvM * kCvKC * 4;            // 16 KB
constexpr int kCvBBytes = kCvN * kCvKC * 4;            // 32 KB
constexpr int kCvStageBytes = kCvABytes + kCvBBytes;   // 48 KB
#ifndef SAD_CONV_RING_KB
#define SAD_CONV_RING_KB 192
#endif
constexpr int kCvRingBytes = SAD_CONV_RING_KB * 1024;   // operand ring of either kernel (+ 1.25 KB of slack and barriers <= 227 KB per CTA)
static_assert(kCvRingBytes + 1024 + 256 <= 227 * 1024, "shared memory per CTA");
constexpr int kCvStages = kCvRingBytes / kCvStageBytes;                       // 4 (8 rows) / 6 (4 rows)
constexpr int kCvStagesPair = kCvRingBytes / (kCvABytes + kCvBBytes / 2);     // CTA-pair kernel, half the pixel tile per CTA: 6 / 8
constexpr int kCvEpiWarps = 8;                         // two warps per TMEM lane quarter, each draining 4 of the tile's 8 pixel rows
                                                       // (measured head step bs=2 / bs=16: 4 warps 2.10 / 14.99 ms, 8: 2.02 / 14.61, 16: 2.02 / 14.39)
constexpr int kCvThreads = 64 + 32 * kCvEpiWarps;      // warp 0: TMA, warp 1: MMA + TMEM, warps 2-9: epilogue
constexpr int kCvTmemCols = 512;                       // 2 accumulator buffers x 256 columns
constexpr size_t kCvSmemBytes = (size_t)kCvRingBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
static_assert((3 * kCvStagesPair + 4) * 8 + 4 <= 256 && kCvStagesPair >= kCvStages, "barrier block");

struct ConvLevel {
  float* y_nchw;            // may be null
  float* y_nhwc;            // may be null
  const float* mask_nhwc;   // may be null: ReluGradient fused into the data-gradient pass (out = mask > 0 ? out : 0)
  uint32_t* bits_out;       // may be null: sign bits of the output, [N][H][ceil(W/32)][Cout], bit i <-> x = 32 * xb + i
  const uint32_t* bits_in;  // may be null: the same layout, applied like mask_nhwc (bit set = pass)
  int32_t N, H, W;
  int32_t accumulate;       // y_nchw += result (the autograd Sum of two consumers' gradients, core.py:695,792-842)
  uint32_t tiles_x, tiles_y, tile_begin, tile_end;
};
struct alignas(64) ConvArgs {
  CUtensorMap tmap_w;
  CUtensorMap tmap_x[SAD_MAX_LEVELS];    // box {32 ci, 32 x, 8 y, 1 n}: a whole pixel tile
  CUtensorMap tmap_xh[SAD_MAX_LEVELS];   // box {32 ci, 32 x, 4 y, 1 n}: half a pixel tile (CTA-pair kernel)
  ConvLevel lv[SAD_MAX_LEVELS];
  const float* bias;
  int32_t n_levels, cin, cout, relu;
  uint32_t m_tiles, k_blocks, total_tiles;
  float nchw_scale;   // fp16 instantiations only: multiplies what is stored to y_nchw (1 / loss scale on the last data-gradient pass)
  // 3xTF32 ("f32x3") instantiations only: operands are split tensors [hi | lo] (see kX3 below)
  uint32_t k_part;      // K blocks of ONE part (k_blocks = 3 * k_part)
  int32_t k_split;      // channel offset of the lo part in the packed weights and the input rows: round_up(cin, 32)
  int32_t cout_split;   // channel offset of the lo part in y_nhwc rows: round_up(cout, 32); the rows are 2 * cout_split long
};

struct ConvTile {
  int l, n, y0, x0, m0;
};
__device__ __forceinline__ ConvTile conv_decode_tile(const ConvArgs& a, uint32_t tile, int& l_hint) {
  while (tile >= a.lv[l_hint].tile_end) ++l_hint;
  const ConvLevel& L = a.lv[l_hint];
  uint32_t r = tile - L.tile_begin;
  const uint32_t mb = r % a.m_tiles;
  r /= a.m_tiles;
  const uint32_t xb = r % L.tiles_x;
  r /= L.tiles_x;
  const uint32_t yb = r % L.tiles_y;
  const uint32_t n = r / L.tiles_y;
  ConvTile t;
  t.l = l_hint;
  t.n = (int)n;
  t.y0 = (int)yb * kCvRows;
  t.x0 = (int)xb * kCvCols;
  t.m0 = (int)mb * kCvM;
  return t;
}

// kC2: launched as clusters of 2 CTAs (the two SMs of one TPC) that execute ONE 256-channel x 256-pixel MMA per step
// (tcgen05.mma.cta_group::2, SASS UTCHMMA.2CTA): each CTA stages its own 128-channel weight tile (16 KB) and HALF of the
// pixel tile (4 of the 8 rows, 16 KB) per stage — 32 instead of 48 KB of TMA traffic per CTA and stage, 6 instead of 4
// stages in the same shared memory — the leader CTA's MMA thread issues for both, and each CTA's 128 x 256 fp32
// accumulator sits in its own TMEM.  m_tiles then counts PAIRS of 128-channel tiles; both CTAs walk the same tile sequence.
// Measured on B200 (bs = 16, 256 -> 256 / 720 channels): 737 / 744 TF/s against 739 / 695 TF/s for one-CTA tiles; whole head
// step 14.99 vs 15.60 ms at bs = 16 and 2.131 vs 2.152 ms at bs = 2.  Cycle counters in the MMA thread (12 tiles of 288
// MMAs per CTA): 78 % of its time it is blocked issuing into a busy tensor pipe (128 cycles per MMA, the rate
// scripts/probe/mma_probe.cu measures for both forms), 22 % waiting for operands.
// Dead ends measured on the way: stream-K scheduling (-10..-17 %), multicasting the activation tile to two
// independent one-CTA MMAs (+-0), L2 prefetch of the next tile (+-0), and `mbarrier.arrive.release.cluster` for the
// cross-CTA hand-over, which costs ~1400 cycles per call and halved the throughput until replaced by the default form.
//   barriers   full[s]      one per CTA: its own copies (32 KB)
//              peer_full[s] in the LEADER: the partner's relay thread arrives when the partner's full[s] completed (letting the
//                           partner's copies signal the leader's barrier directly — the cta_group::2 form of the TMA load — was
//                           measured 2x slower: 1650 instead of 780 cycles per stage)
//              empty[s]     one per CTA: the leader's commit multicasts "stage read" to both producers
//              tmem_full[b] one per CTA: the leader's commit multicasts "accumulator complete" to both epilogues
//              tmem_empty[b] lives in the LEADER: 256 arrivals, both CTAs' epilogue threads (the partner's remotely)
// kX3: "3xTF32", the fp32-accurate mode (the reference's head convolution is fp32: conv_op_cudnn.cc:494-498 enables tensor-op math
// for fp16 only).  Every fp32 operand v is carried as two tf32 numbers hi = rna_tf32(v), lo = rna_tf32(v - hi) (v = hi + lo to
// 2^-22 |v|), stored side by side: channels-last rows [hi(0..C) pad | lo(0..C) pad] of 2 * round_up(C, 32) floats, packed weights
// [tap][M][hi(0..K) pad | lo(0..K) pad].  The K loop of a tap runs three passes over the same accumulator,
//     W_hi * X_hi  +  W_hi * X_lo  +  W_lo * X_hi        (W_lo * X_lo ~ 2^-22 is dropped),
// which only moves the producer's TMA coordinates (pass 1: activations' lo half, pass 2: weights' lo half); the MMA issuer
// just sees 3x as many stages.  The channels-last output is written as such a split row.  3x the tensor-core work of tf32.
// kF16: fp16 operands (tcgen05.mma kind::f16, K = 16 per instruction): the channels-last input, the packed weights and the
// channels-last output are fp16; a stage still holds 128-byte rows, i.e. 64 instead of 32 input channels, so a tile takes
// half as many stages and MMAs.  Accumulation, bias, activation and the NCHW output stay fp32 (BASELINE.json configs[4]:
// "mixed fp16 compute / fp32 loss accumulate").  fp16 keeps the 10-bit mantissa of tf32; values beyond 65504 become inf.
template <bool kC2, bool kSigmoid, bool kF16 = false, bool kX3 = false>
__global__ void __launch_bounds__(kCvThreads, 1) conv3x3_tf32_kernel(const __grid_constant__ ConvArgs args) {
  static_assert(!(kF16 && kX3), "the split mode is a tf32 mode");
  constexpr int kKCe = kF16 ? 2 * kCvKC : kCvKC;   // input channels per stage (one 128-byte row)
  constexpr int kStages = kC2 ? kCvStagesPair : kCvStages;
  constexpr int kBBytes = kC2 ? kCvBBytes / 2 : kCvBBytes;
  constexpr int kStageBytes = kCvABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned operand ring (128B-swizzle atoms are 1024 B), barriers behind it
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* full_bar = bars;                  // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;       // [kStages] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;   // [2] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;       // [2] epilogue -> MMA
  uint64_t* peer_full = tmem_empty + 2;       // [kStages] pair only: partner's stage landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = kC2 ? cluster_ctarank() : 0u;            // which 128-channel half of the pair
  const uint32_t wstream = kC2 ? blockIdx.x >> 1 : blockIdx.x;    // work stream (both CTAs of a pair walk the same one)
  const uint32_t nstreams = kC2 ? gridDim.x >> 1 : gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmap_w);
    for (int l = 0; l < args.n_levels; ++l)
      tma_prefetch_desc(kC2 ? &args.tmap_xh[l] : &args.tmap_x[l]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
        mbar_init(&peer_full[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full[b], 1);
        mbar_init(&tmem_empty[b], (kC2 ? 2 : 1) * 32 * kCvEpiWarps);
      }
      mbar_fence_init();
    }
    __syncwarp();
    if (kC2) tmem_alloc_2sm<kCvTmemCols>(tmem_slot);
    else tmem_alloc<kCvTmemCols>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kC2) cluster_sync_all();   // the partner's mbarriers exist before anything of ours can signal them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t k_blocks = args.k_blocks;  // ceil(Cin / kKCe); kX3: 3 passes of k_part blocks

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    if (lane == 0) {
      RingState rs;
      int l_hint = 0;
      for (uint32_t tile = wstream; tile < args.total_tiles; tile += nstreams) {
        ConvTile t = conv_decode_tile(args, tile, l_hint);
        if (kC2) t.m0 = (2 * (t.m0 / kCvM) + (int)crank) * kCvM;
        const CUtensorMap* mx = kC2 ? &args.tmap_xh[t.l] : &args.tmap_x[t.l];
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (uint32_t kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&empty_bar[rs.stage], rs.phase ^ 1u);
            uint8_t* sa = smem + (size_t)rs.stage * kStageBytes;
            uint8_t* sb = sa + kCvABytes;
            int ka = (int)kb * kKCe, kx = ka;   // K coordinate in the packed weights / channel coordinate in the input rows
            if (kX3) {
              const uint32_t pass = kb / args.k_part;   // 0: hi x hi, 1: W_hi x X_lo, 2: W_lo x X_hi
              const int kk = (int)(kb - pass * args.k_part) * kKCe;
              ka = kk + (pass == 2 ? args.k_split : 0);
              kx = kk + (pass == 1 ? args.k_split : 0);
            }
            if (kC2) {
              mbar_arrive_expect_tx(&full_bar[rs.stage], kStageBytes);
              tma_load_3d(sa, &args.tmap_w, &full_bar[rs.stage], ka, t.m0, tap);
              // my half of the pixel tile: rows y0 + 4 * rank .. + 3 (= accumulator columns 128 * rank .. + 127)
              tma_load_4d(sb, mx, &full_bar[rs.stage], kx, t.x0 + dx, t.y0 + (int)crank * (kCvRows / 2) + dy, t.n);
            } else {
              mbar_arrive_expect_tx(&full_bar[rs.stage], kStageBytes);
              tma_load_3d(sa, &args.tmap_w, &full_bar[rs.stage], ka, t.m0, tap);
              tma_load_4d(sb, mx, &full_bar[rs.stage], kx, t.x0 + dx, t.y0 + dy, t.n);
            }
            rs.advance<kStages>();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; pair: of the leader CTA only) =====================
    if (kC2 && lane == 0 && crank != 0) {
      // partner CTA: relay "my stage landed" to the leader's MMA thread
      RingState rs;
      for (uint32_t tile = wstream; tile < args.total_tiles; tile += nstreams) {
        const uint32_t n_kb = 9u * k_blocks;
        for (uint32_t kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&full_bar[rs.stage], rs.phase);
          mbar_arrive_cluster(map_to_cta(&peer_full[rs.stage], 0));
          rs.advance<kStages>();
        }
      }
    }
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = kF16 ? umma_idesc_f16(kC2 ? 2 * kCvM : kCvM, kCvN, /*A K-major*/ 0, /*B K-major*/ 0)
                                      : umma_idesc_tf32(kC2 ? 2 * kCvM : kCvM, kCvN, /*A K-major*/ 0, /*B K-major*/ 0);
      RingState rs;
      uint32_t it = 0;
      for (uint32_t tile = wstream; tile < args.total_tiles; tile += nstreams, ++it) {
        // kX3: the tile's stage sequence is cut into TWO accumulation chains, one per TMEM buffer, which the epilogue adds in
        // fp32 (round to nearest): the tensor core's accumulator add rounds toward zero (measured: the result of a K = 2304
        // chain is 1.6e-5 too small in magnitude, growing linearly with the number of MMAs in the chain), so halving the
        // chains halves that bias.  The price is the epilogue / main-loop overlap, < 10 % of a 3x longer main loop.
        const uint32_t buf = kX3 ? 0u : (it & 1u), aphase = kX3 ? (it & 1u) : ((it >> 1) & 1u);
        mbar_wait(&tmem_empty[buf], aphase ^ 1u);
        tc_fence_after_sync();
        const uint32_t n_kb = 9u * k_blocks;
        const uint32_t chain_b = kX3 ? n_kb / 2u : n_kb;   // first stage of the second chain
        for (uint32_t kb = 0; kb < n_kb; ++kb) {
          const uint32_t d_tmem = tmem_base + ((kX3 ? (kb >= chain_b ? 1u : 0u) : buf) * kCvN);
          const uint32_t kb_chain = (kX3 && kb >= chain_b) ? kb - chain_b : kb;   // 0 at the first stage of a chain: overwrite
          mbar_wait(&full_bar[rs.stage], rs.phase);
          if (kC2) mbar_wait_cluster(&peer_full[rs.stage], rs.phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + (size_t)rs.stage * kStageBytes);
          const uint32_t b_addr = a_addr + kCvABytes;
#pragma unroll
          for (int k = 0; k < kCvKC / 8; ++k) {
            // both operands K-major: rows of 128 B (32 tf32 or 64 fp16), 8-row groups 1024 B apart (SBO);
            // one K step (8 tf32 or 16 fp16) = +32 B inside the swizzled row
            const uint64_t adesc = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            if (kF16) {
              if (kC2) umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb_chain | (uint32_t)k) != 0u);
              else umma_f16(d_tmem, adesc, bdesc, idesc, (kb_chain | (uint32_t)k) != 0u);
            } else {
              if (kC2) umma_tf32_2sm(d_tmem, adesc, bdesc, idesc, (kb_chain | (uint32_t)k) != 0u);
              else umma_tf32(d_tmem, adesc, bdesc, idesc, (kb_chain | (uint32_t)k) != 0u);
            }
          }
          // frees the smem stage when these MMAs have read it (pair: in both CTAs)
          if (kC2) umma_commit_2sm(&empty_bar[rs.stage], (uint16_t)0x3);
          else umma_commit(&empty_bar[rs.stage]);
          rs.advance<kStages>();
        }
        // accumulator complete (pair: both CTAs' halves)
        if (kC2) umma_commit_2sm(&tmem_full[buf], (uint16_t)0x3);
        else umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (4 warps = 128 accumulator rows) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access (hardware rule: lanes 32 * (warp % 4) .. + 31)
    constexpr int kRowsPerWarp = kCvRows / (kCvEpiWarps / 4);
    const int jhalf = (warp - 2) >> 2;   // which of the tile's pixel rows (accumulator column blocks) this warp drains
    const int row = q * 32 + lane;
    uint32_t it = 0;
    int l_hint = 0;
    for (uint32_t tile = wstream; tile < args.total_tiles; tile += nstreams, ++it) {
      const uint32_t buf = kX3 ? 0u : (it & 1u), aphase = kX3 ? (it & 1u) : ((it >> 1) & 1u);
      ConvTile t = conv_decode_tile(args, tile, l_hint);
      if (kC2) t.m0 = (2 * (t.m0 / kCvM) + (int)crank) * kCvM;
      const ConvLevel& L = args.lv[t.l];
      const int co = t.m0 + row;
      const bool co_ok = co < args.cout;
      const float b = (co_ok && args.bias) ? __ldg(args.bias + co) : 0.f;
      mbar_wait(&tmem_full[buf], aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kCvN;
      const size_t HW = (size_t)L.H * L.W;
      float* yrow = L.y_nchw ? L.y_nchw + ((size_t)t.n * args.cout + (co_ok ? co : 0)) * HW : nullptr;
      // channels-last output rows: cout floats; kX3: split rows of 2 * cout_split floats, hi at co, lo at cout_split + co
      const size_t ycl_row = kX3 ? 2 * (size_t)args.cout_split : (size_t)args.cout;
      float* ycl = L.y_nhwc ? L.y_nhwc + (size_t)t.n * HW * ycl_row + (co_ok ? co : 0) : nullptr;
      const float* mcl = L.mask_nhwc ? L.mask_nhwc + (size_t)t.n * HW * args.cout + (co_ok ? co : 0) : nullptr;
      // sign-bit planes: one word per (image row, 32-pixel segment, channel); this tile owns segment t.x0 / 32
      const size_t bits_row0 = (((size_t)t.n * L.H) * L.tiles_x + (t.x0 >> 5)) * args.cout + (co_ok ? co : 0);
      const bool vec_ok = (L.W & 3) == 0;
#pragma unroll 1
      for (int j = jhalf * kRowsPerWarp; j < (jhalf + 1) * kRowsPerWarp; ++j) {
        float v[32];
        tmem_ld_32x32(taddr + j * kCvCols, v);
        if (kX3) {   // second accumulation chain (the other TMEM buffer)
          float w2[32];
          tmem_ld_32x32(taddr + kCvN + j * kCvCols, w2);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += w2[i];
        }
        const int y = t.y0 + j;
        if (co_ok && y < L.H) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] += b;
            // the Sigmoid epilogue is its own instantiation: as a run-time branch next to ReLU it slowed EVERY convolution by 7 %
            // (head forward 0.69 vs 0.62 ms, measured A/B on one box)
            if (kSigmoid) v[i] = __frcp_rn(1.f + __expf(-v[i]));   // Sigmoid (sigmoid_op.cu:24-29): the teacher's class probabilities
            else if (args.relu) v[i] = fmaxf(v[i], 0.f);
          }
          if (L.bits_in) {
            // ReluGradient (relu_op.cu:29-35: dX = Y > 0 ? dY : 0) from the sign bits the forward pass left: 1 word per row
            const uint32_t bits = __ldg(L.bits_in + bits_row0 + (size_t)y * L.tiles_x * args.cout);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (bits >> i) & 1u ? v[i] : 0.f;
          } else if (mcl) {
            // the same from a channels-last float tensor Y.  The 32 loads are issued back to back (volatile asm keeps
            // the compiler from chaining them through two registers, which serialises the memory latency)
            const float* msk = mcl + ((size_t)y * L.W + t.x0) * args.cout;
            float m[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              m[i] = 1.f;
              if (t.x0 + i < L.W) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(m[i]) : "l"(msk + (size_t)i * args.cout));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = m[i] > 0.f ? v[i] : 0.f;
          }
          if (L.bits_out) {
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
            L.bits_out[bits_row0 + (size_t)y * L.tiles_x * args.cout] = bits;
          }
          if (yrow) {
            float* dst = yrow + (size_t)y * L.W + t.x0;
            const float ns = kF16 ? args.nchw_scale : 1.f;
            if (vec_ok) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (t.x0 + i < L.W) {
                  float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                  if (kF16) o = make_float4(o.x * ns, o.y * ns, o.z * ns, o.w * ns);
                  if (L.accumulate) {
                    const float4 p = *reinterpret_cast<const float4*>(dst + i);
                    o.x += p.x;
                    o.y += p.y;
                    o.z += p.z;
                    o.w += p.w;
                  }
                  *reinterpret_cast<float4*>(dst + i) = o;
                }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (t.x0 + i < L.W) {
                  if (kF16) dst[i] = L.accumulate ? dst[i] + v[i] * ns : v[i] * ns;
                  else dst[i] = L.accumulate ? dst[i] + v[i] : v[i];
                }
            }
          }
          if (ycl) {
            if (kF16) {
              // fp16 channels-last: the same element offsets in 2-byte elements (the warp's 32 lanes write one 64 B line)
              __half* dst = reinterpret_cast<__half*>(L.y_nhwc) + (size_t)t.n * HW * args.cout + co + ((size_t)y * L.W + t.x0) * args.cout;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (t.x0 + i < L.W) dst[(size_t)i * args.cout] = __float2half_rn(v[i]);
            } else if (kX3) {
              // split rows: two 128 B lines per pixel and warp (hi, lo)
              float* dst = ycl + ((size_t)y * L.W + t.x0) * ycl_row;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (t.x0 + i < L.W) {
                  const float hi = to_tf32_rna(v[i]);
                  dst[(size_t)i * ycl_row] = hi;
                  dst[(size_t)i * ycl_row + args.cout_split] = to_tf32_rna(v[i] - hi);
                }
            } else {
              // channels-last: for a fixed pixel the warp's 32 lanes (consecutive co) write one 128 B line
              float* dst = ycl + ((size_t)y * L.W + t.x0) * args.cout;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (t.x0 + i < L.W) dst[(size_t)i * args.cout] = to_tf32_rna(v[i]);
            }
          }
        }
      }
      tc_fence_before_sync();
      if (kC2) mbar_arrive_cluster(map_to_cta(&tmem_empty[buf], 0));   // the leader's MMA thread owns the accumulator hand-over
      else mbar_arrive(&tmem_empty[buf]);
    }
  }

  __syncthreads();
  if (kC2) cluster_sync_all();   // nothing of the partner's (copies, barrier arrivals, MMA reads of my tiles) may still be in flight
  if (warp == 1) {
    tc_fence_after_sync();
    if (kC2) tmem_dealloc_2sm<kCvTmemCols>(tmem_base);
    else tmem_dealloc<kCvTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// weight repack: (Cout, Cin, 3, 3) fp32 -> [tap][M][K] tf32-rounded fp32
//   mode 0 (forward):        M = Cout, K = Cin,  Wp[tap][co][ci] = W[co][ci][tap]
//   mode 1 (data gradient):  M = Cin,  K = Cout, Wp[tap][ci][co] = W[co][ci][8 - tap]   (taps flipped)
// ---------------------------------------------------------------------------------------------
__global__ void conv3x3_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int mode) {
  const int M = mode == 0 ? cout : cin, K = mode == 0 ? cin : cout;
  const size_t total = (size_t)9 * M * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int m = (int)((i / K) % M);
    const int tap = (int)(i / ((size_t)K * M));
    const float v = mode == 0 ? w[((size_t)m * cin + k) * 9 + tap] : w[((size_t)k * cin + m) * 9 + (8 - tap)];
    out[i] = to_tf32_rna(v);
  }
}

__device__ __forceinline__ void store_operand(float* p, float v) { *p = to_tf32_rna(v); }
__device__ __forceinline__ void store_operand(__half* p, float v) { *p = __float2half_rn(v); }

struct PackMulti {
  const float* src[SAD_MAX_PACK_ITEMS];
  float* dst[SAD_MAX_PACK_ITEMS];   // fp16 instantiation: __half storage
  int32_t cin[SAD_MAX_PACK_ITEMS], cout[SAD_MAX_PACK_ITEMS], mode[SAD_MAX_PACK_ITEMS];
  int32_t split;   // float output only: 3xTF32 rows [hi(0..K) pad | lo(0..K) pad] of 2 * round_up(K, 32) floats
};
// every weight tensor of the head in one launch: blockIdx.y selects the tensor; one thread per (row m, column k)
// of the packed planes moves the 9 taps (writes coalesced along k in each tap plane)
template <typename OutT>
__global__ void conv3x3_pack_multi_kernel(const PackMulti p) {
  const int k_ = blockIdx.y;
  const float* __restrict__ w = p.src[k_];
  OutT* __restrict__ out = reinterpret_cast<OutT*>(p.dst[k_]);
  const int cout = p.cout[k_], cin = p.cin[k_], mode = p.mode[k_];
  const uint32_t M = mode == 0 ? cout : cin, K = mode == 0 ? cin : cout;
  // 16-bit output: the K axis is padded with zeros to a multiple of 8 (16-byte tensor-map strides), e.g. the 36 box-regression
  // channels of the data-gradient pack become 40; the convolution is then called with cin = the padded K
  const bool split = sizeof(OutT) == 4 && p.split != 0;
  const uint32_t Ks = (K + 31u) & ~31u;   // offset of the lo half of a split row
  const uint32_t Kp = sizeof(OutT) == 2 ? (K + 7u) & ~7u : (split ? 2u * Ks : K);
  const uint32_t plane = M * Kp;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const uint32_t kcol = i % Kp, m = i / Kp;
    const bool lo = split && kcol >= Ks;
    const uint32_t k = lo ? kcol - Ks : kcol;
    const float* src = mode == 0 ? w + ((size_t)m * cin + k) * 9 : w + ((size_t)k * cin + m) * 9;
    float v[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      v[tap] = k < K ? __ldg(src + tap) : 0.f;
      if (lo) v[tap] -= to_tf32_rna(v[tap]);   // exact in fp32; store_operand rounds it to tf32
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) store_operand(out + (size_t)tap * plane + i, mode == 0 ? v[tap] : v[8 - tap]);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT fallback for channel counts the TMA path cannot address (C % 4 != 0): direct convolution from
// the same packed weights and channels-last input.  One thread per output element.
// ---------------------------------------------------------------------------------------------
__global__ void conv3x3_simt_kernel(const float* __restrict__ xt, const float* __restrict__ wp, const float* __restrict__ bias,
                                    float* __restrict__ y_nchw, float* __restrict__ y_nhwc, const float* __restrict__ mask_nhwc,
                                    uint32_t* __restrict__ bits_out, const uint32_t* __restrict__ bits_in, int N, int cin, int cout,
                                    int H, int W, int relu, int accumulate) {
  const size_t total = (size_t)N * cout * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const int xx = (int)((i / cout) % W);
    const int yy = (int)((i / ((size_t)cout * W)) % H);
    const int n = (int)(i / ((size_t)cout * W * H));
    float acc = bias ? bias[co] : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      const float* wrow = wp + ((size_t)tap * cout + co) * cin;
      const float* xp = xt + (((size_t)n * H + sy) * W + sx) * cin;
      for (int ci = 0; ci < cin; ++ci) acc = fmaf(wrow[ci], xp[ci], acc);
    }
    if (relu == 1) acc = fmaxf(acc, 0.f);
    else if (relu == 2) acc = __frcp_rn(1.f + __expf(-acc));
    if (mask_nhwc && !(mask_nhwc[i] > 0.f)) acc = 0.f;
    const size_t word = (((size_t)n * H + yy) * ((W + 31) / 32) + (xx >> 5)) * cout + co;
    if (bits_in && !((bits_in[word] >> (xx & 31)) & 1u)) acc = 0.f;
    if (bits_out && acc > 0.f) atomicOr(bits_out + word, 1u << (xx & 31));  // zeroed by the launcher
    if (y_nchw) {
      float* o = y_nchw + (((size_t)n * cout + co) * H + yy) * W + xx;
      *o = accumulate ? *o + acc : acc;
    }
    if (y_nhwc) y_nhwc[i] = to_tf32_rna(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> channels-last (N, H, W, C) fp32 rounded to tf32 (round-to-nearest), all levels in one
// launch: per image a [C][HW] -> [HW][C] transpose in 32 x 32 tiles through shared memory.
// ---------------------------------------------------------------------------------------------
struct LayoutLevel {
  const float* src;
  float* dst;
  uint32_t HW, tiles_hw, tile_begin, tile_end;  // tiles = N * tiles_c * tiles_hw
};
struct LayoutArgs {
  LayoutLevel lv[SAD_MAX_LEVELS];
  int32_t n_levels, C;
  uint32_t tiles_c, total_tiles;
  int32_t C_dst;   // channels of the destination rows (>= C: zero padding, fp16 gradient tensors); tiles_c covers C_dst
  float scale;     // fp16 instantiation only: values are multiplied before rounding (the loss scale of the gradient tensors)
  int32_t split;   // float output only: 3xTF32 rows of 2 * C_dst floats, hi at c, lo at C_dst + c (C_dst = round_up(C, 32))
};
template <typename OutT>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const __grid_constant__ LayoutArgs a) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  int l = 0;
  for (uint32_t t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
    while (t >= a.lv[l].tile_end) ++l;
    const LayoutLevel& L = a.lv[l];
    uint32_t r = t - L.tile_begin;
    const uint32_t th = r % L.tiles_hw;
    r /= L.tiles_hw;
    const uint32_t tc = r % a.tiles_c;
    const uint32_t n = r / a.tiles_c;
    const uint32_t hw0 = th * 32, c0 = tc * 32;
    const float* src = L.src + (size_t)n * a.C * L.HW;
    const uint32_t row = (sizeof(OutT) == 4 && a.split) ? 2u * (uint32_t)a.C_dst : (uint32_t)a.C_dst;
    OutT* dst = reinterpret_cast<OutT*>(L.dst) + (size_t)n * row * L.HW;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t c = c0 + ty + k * 8, hw = hw0 + tx;
      tile[ty + k * 8][tx] = (c < (uint32_t)a.C && hw < L.HW) ? __ldg(src + (size_t)c * L.HW + hw) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t hw = hw0 + ty + k * 8, c = c0 + tx;
      if (c < (uint32_t)a.C_dst && hw < L.HW) {
        const float v = tile[tx][ty + k * 8];
        store_operand(dst + (size_t)hw * row + c, sizeof(OutT) == 2 ? v * a.scale : v);
        if (sizeof(OutT) == 4 && a.split) store_operand(dst + (size_t)hw * row + a.C_dst + c, v - to_tf32_rna(v));
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
}  // namespace sad

using namespace sad;

extern "C" {

SAD_EXPORT size_t sad_conv3x3_sign_bits_bytes(int N, int channels, int H, int W) {
  if (N < 0 || channels < 0 || H < 0 || W < 0) return 0;
  return (size_t)N * H * ((W + 31) / 32) * channels * sizeof(uint32_t);
}

SAD_EXPORT size_t sad_conv3x3_packed_bytes(int cin, int cout) {
  if (cin < 1 || cout < 1) return 0;
  return (size_t)9 * cin * cout * sizeof(float);
}

SAD_EXPORT int sad_conv3x3_pack_weights_f32(const float* weight, int cin, int cout, int mode, float* packed, void* stream) {
  if (!weight || !packed || cin < 1 || cout < 1 || (mode != 0 && mode != 1))
    return set_error(SAD_ERR_INVALID, "conv3x3 pack: bad argument");
  const size_t total = (size_t)9 * cin * cout;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 4096 ? (total + threads - 1) / threads : 4096);
  conv3x3_pack_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(weight, packed, cout, cin, mode);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "conv3x3 pack launch");
}

static int pack_weights_multi_impl(const sad_pack_item* items, int n_items, void* stream, bool f16, bool split = false) {
  if (!items || n_items < 1 || n_items > SAD_MAX_PACK_ITEMS) return set_error(SAD_ERR_INVALID, "conv3x3 pack multi: n_items must be in [1, 32]");
  PackMulti p{};
  size_t most = 0;
  for (int i = 0; i < n_items; ++i) {
    const sad_pack_item& it = items[i];
    if (!it.weight || !it.packed || it.cin < 1 || it.cout < 1 || (it.mode != 0 && it.mode != 1))
      return set_error(SAD_ERR_INVALID, "conv3x3 pack multi: bad item");
    p.src[i] = it.weight;
    p.dst[i] = it.packed;
    p.cin[i] = it.cin;
    p.cout[i] = it.cout;
    p.mode[i] = it.mode;
    const size_t total = (size_t)it.cin * it.cout * (split ? 4 : 1);   // split rows are 2 * round_up(K, 32) long (the kernel is grid-stride: this only sizes the grid)
    if (total > most) most = total;
  }
  p.split = split ? 1 : 0;
  const unsigned bx = (unsigned)((most + 255) / 256 < 1024 ? (most + 255) / 256 : 1024);
  if (f16) conv3x3_pack_multi_kernel<__half><<<dim3(bx, (unsigned)n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else conv3x3_pack_multi_kernel<float><<<dim3(bx, (unsigned)n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "conv3x3 pack multi launch");
}

static int nchw_to_nhwc_impl(const sad_layout_level* levels, int n_levels, int channels, void* stream, bool f16, int channels_dst = 0,
                             float scale = 1.f, bool split = false) {
  if (!levels || n_levels < 1 || n_levels > SAD_MAX_LEVELS || channels < 1) return set_error(SAD_ERR_INVALID, "nchw_to_nhwc: bad argument");
  if (channels_dst == 0) channels_dst = channels;
  if (channels_dst < channels) return set_error(SAD_ERR_INVALID, "nchw_to_nhwc: destination channels < source channels");
  LayoutArgs a{};
  a.n_levels = n_levels;
  a.C = channels;
  a.C_dst = channels_dst;
  a.scale = scale;
  a.split = split ? 1 : 0;
  a.tiles_c = (uint32_t)((channels_dst + 31) / 32);
  uint64_t tiles = 0;
  for (int l = 0; l < n_levels; ++l) {
    const sad_layout_level& L = levels[l];
    if (L.N < 0 || L.H < 0 || L.W < 0) return set_error(SAD_ERR_INVALID, "nchw_to_nhwc: negative dimension");
    const uint64_t HW = (uint64_t)L.H * L.W;
    if ((uint64_t)L.N * HW && (!L.src_nchw || !L.dst_nhwc)) return set_error(SAD_ERR_INVALID, "nchw_to_nhwc: null tensor");
    a.lv[l].src = L.src_nchw;
    a.lv[l].dst = L.dst_nhwc;
    a.lv[l].HW = (uint32_t)HW;
    a.lv[l].tiles_hw = (uint32_t)((HW + 31) / 32);
    a.lv[l].tile_begin = (uint32_t)tiles;
    tiles += (uint64_t)L.N * a.tiles_c * a.lv[l].tiles_hw;
    a.lv[l].tile_end = (uint32_t)tiles;
    if (tiles > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "nchw_to_nhwc: too many tiles");
  }
  if (tiles == 0) return SAD_OK;
  a.total_tiles = (uint32_t)tiles;
  int sms = 0, rc;
  if ((rc = sm_count(&sms)) != SAD_OK) return rc;
  const uint32_t grid = a.total_tiles < (uint32_t)sms * 8u ? a.total_tiles : (uint32_t)sms * 8u;
  if (f16) nchw_to_nhwc_kernel<__half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else nchw_to_nhwc_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "nchw_to_nhwc launch");
}

// `packed` has layout [tap][cout][cin] (sad_conv3x3_pack_weights_f32 mode 0 for the forward operator,
// mode 1 — with cin/cout swapped by the caller — for the data gradient).
static int conv3x3_fwd_impl(const sad_conv_level* levels, int n_levels, const float* packed, const float* bias, int cin,
                            int cout, int relu, void* stream, bool f16, float nchw_scale = 1.f, bool x3 = false) {
  if (!levels || n_levels < 1 || n_levels > SAD_MAX_LEVELS) return set_error(SAD_ERR_INVALID, "conv3x3: n_levels must be in [1, 8]");
  if (!packed || cin < 1 || cout < 1) return set_error(SAD_ERR_INVALID, "conv3x3: bad weights/channels");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvArgs a{};
  uint64_t tiles = 0;
  // two or more 128-channel tiles: CTA pairs (clusters of 2) share each pixel tile's activations; m_tiles then counts pairs
  const uint32_t m_single = (uint32_t)((cout + kCvM - 1) / kCvM);
  const bool pair = m_single >= 2;
  const uint32_t m_tiles = pair ? (m_single + 1) / 2 : m_single;
  bool tma_ok = (cin % (f16 ? 8 : 4) == 0) && ((reinterpret_cast<uintptr_t>(packed) & 15) == 0);   // 16-byte tensor-map strides
  int rc;
  for (int l = 0; l < n_levels; ++l) {
    const sad_conv_level& L = levels[l];
    if (L.N < 0 || L.H < 0 || L.W < 0) return set_error(SAD_ERR_INVALID, "conv3x3: negative dimension");
    const uint64_t elems = (uint64_t)L.N * L.H * L.W;
    if (elems && (!L.x_nhwc || (!L.y_nchw && !L.y_nhwc))) return set_error(SAD_ERR_INVALID, "conv3x3: null tensor");
    if ((reinterpret_cast<uintptr_t>(L.x_nhwc) | reinterpret_cast<uintptr_t>(L.y_nchw)) & 15) tma_ok = false;
    ConvLevel& D = a.lv[l];
    D.y_nchw = L.y_nchw;
    D.y_nhwc = L.y_nhwc;
    D.mask_nhwc = L.relu_mask_nhwc;
    D.bits_out = L.relu_bits_out;
    D.bits_in = L.relu_bits_in;
    D.accumulate = L.accumulate_nchw;
    D.N = L.N;
    D.H = L.H;
    D.W = L.W;
    D.tiles_x = (uint32_t)((L.W + kCvCols - 1) / kCvCols);
    D.tiles_y = (uint32_t)((L.H + kCvRows - 1) / kCvRows);
    D.tile_begin = (uint32_t)tiles;
    tiles += (uint64_t)L.N * D.tiles_y * D.tiles_x * m_tiles;
    D.tile_end = (uint32_t)tiles;
    if (tiles > 0x7fffffffull) return set_error(SAD_ERR_INVALID, "conv3x3: too many tiles");
  }
  if (tiles == 0) return SAD_OK;
  if (!tma_ok && f16)
    return set_error(SAD_ERR_UNSUPPORTED, "conv3x3 fp16: needs Cin % 8 == 0 and 16-byte aligned tensors (there is no SIMT fp16 path)");
  const int cin_s = (cin + 31) & ~31;   // x3: the lo half of a split row / packed K row starts here; rows are 2 * cin_s long
  if (x3) {
    bool ok = (reinterpret_cast<uintptr_t>(packed) & 15) == 0;
    for (int l = 0; l < n_levels; ++l) {
      if ((reinterpret_cast<uintptr_t>(levels[l].x_nhwc) | reinterpret_cast<uintptr_t>(levels[l].y_nchw)) & 15) ok = false;
      if (levels[l].relu_mask_nhwc)
        return set_error(SAD_ERR_UNSUPPORTED, "conv3x3 f32x3: ReluGradient comes from sign bits (relu_bits_in), not from a float mask tensor");
    }
    if (!ok) return set_error(SAD_ERR_UNSUPPORTED, "conv3x3 f32x3: needs 16-byte aligned tensors (there is no SIMT split path)");
    tma_ok = true;
  }
  const int cin_map = x3 ? 2 * cin_s : cin;   // innermost extent of the packed-weight and activation tensor maps
  if (f16) {
    for (int l = 0; l < n_levels; ++l)
      if (levels[l].relu_mask_nhwc)
        return set_error(SAD_ERR_UNSUPPORTED, "conv3x3 fp16: ReluGradient comes from sign bits (relu_bits_in), not from a float mask tensor");
  }
  if (!tma_ok) {  // channel count / alignment the TMA path cannot address
    for (int l = 0; l < n_levels; ++l) {
      const sad_conv_level& L = levels[l];
      const size_t total = (size_t)L.N * cout * L.H * L.W;
      if (total == 0) continue;
      const size_t blocks = (total + 127) / 128;
      if (L.relu_bits_out &&
          (rc = check_cuda(cudaMemsetAsync(L.relu_bits_out, 0, (size_t)L.N * L.H * ((L.W + 31) / 32) * cout * sizeof(uint32_t), st),
                           "conv3x3 simt: clear sign bits")) != SAD_OK)
        return rc;
      conv3x3_simt_kernel<<<(unsigned)(blocks < 1048576 ? blocks : 1048576), 128, 0, st>>>(L.x_nhwc, packed, bias, L.y_nchw, L.y_nhwc,
                                                                                           L.relu_mask_nhwc, L.relu_bits_out, L.relu_bits_in, L.N, cin, cout, L.H,
                                                                                           L.W, relu,
                                                                                           L.accumulate_nchw);
      count_launch(1);
      if ((rc = check_cuda(cudaGetLastError(), "conv3x3 simt launch")) != SAD_OK) return rc;
    }
    return SAD_OK;
  }
  {
    const cuuint64_t esz = f16 ? 2 : 4;
    const cuuint64_t dims[3] = {(cuuint64_t)cin_map, (cuuint64_t)cout, 9};
    const cuuint64_t str[2] = {(cuuint64_t)cin_map * esz, (cuuint64_t)cin_map * cout * esz};
    const cuuint32_t box[3] = {(cuuint32_t)(f16 ? 2 * kCvKC : kCvKC), kCvM, 1};
    if ((rc = encode_map(&a.tmap_w, packed, 3, dims, str, box, "packed weights", CU_TENSOR_MAP_SWIZZLE_128B,
                         f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) != SAD_OK)
      return rc;
  }
  for (int l = 0; l < n_levels; ++l) {
    const sad_conv_level& L = levels[l];
    if ((uint64_t)L.N * L.H * L.W == 0) {  // empty level: a valid dummy map (never used, no tiles)
      a.tmap_x[l] = a.tmap_w;
      a.tmap_xh[l] = a.tmap_w;
      continue;
    }
    if (f16) {
      if ((rc = encode_nhwc_map_f16(&a.tmap_x[l], L.x_nhwc, L.N, cin, L.H, L.W, kCvCols, kCvRows, "fp16 activations {C,W,H,N}")) != SAD_OK) return rc;
      if ((rc = encode_nhwc_map_f16(&a.tmap_xh[l], L.x_nhwc, L.N, cin, L.H, L.W, kCvCols, kCvRows / 2, "fp16 activations {C,W,H,N}, half tile")) !=
          SAD_OK)
        return rc;
      continue;
    }
    if ((rc = encode_nhwc_map(&a.tmap_x[l], L.x_nhwc, L.N, cin_map, L.H, L.W, kCvCols, kCvRows, "activations {C,W,H,N}")) != SAD_OK) return rc;
    if ((rc = encode_nhwc_map(&a.tmap_xh[l], L.x_nhwc, L.N, cin_map, L.H, L.W, kCvCols, kCvRows / 2, "activations {C,W,H,N}, half tile")) != SAD_OK)
      return rc;
  }
  a.bias = bias;
  a.nchw_scale = nchw_scale;
  a.n_levels = n_levels;
  a.cin = cin;
  a.cout = cout;
  a.relu = relu;
  a.m_tiles = m_tiles;
  const int kc = f16 ? 2 * kCvKC : kCvKC;
  a.k_blocks = (uint32_t)((cin + kc - 1) / kc);
  if (x3) {
    a.k_part = (uint32_t)(cin_s / kCvKC);
    a.k_blocks = 3 * a.k_part;
    a.k_split = cin_s;
    a.cout_split = (cout + 31) & ~31;
  }
  a.total_tiles = (uint32_t)tiles;

  int sms = 0;
  if ((rc = sm_count(&sms)) != SAD_OK) return rc;
  if (pair) {
    auto kern = f16  ? (relu == 2 ? conv3x3_tf32_kernel<true, true, true> : conv3x3_tf32_kernel<true, false, true>)
                : x3 ? (relu == 2 ? conv3x3_tf32_kernel<true, true, false, true> : conv3x3_tf32_kernel<true, false, false, true>)
                     : (relu == 2 ? conv3x3_tf32_kernel<true, true, false> : conv3x3_tf32_kernel<true, false, false>);
    if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCvSmemBytes),
                         "cudaFuncSetAttribute(conv3x3 pair)")) != SAD_OK)
      return rc;
    uint32_t pairs = (uint32_t)sms / 2;
    if (pairs > a.total_tiles) pairs = a.total_tiles;
    if (pairs < 1) pairs = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kCvThreads);
    cfg.dynamicSmemBytes = kCvSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if ((rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, a), "conv3x3 pair launch")) != SAD_OK) return rc;
  } else {
    auto kern = f16  ? (relu == 2 ? conv3x3_tf32_kernel<false, true, true> : conv3x3_tf32_kernel<false, false, true>)
                : x3 ? (relu == 2 ? conv3x3_tf32_kernel<false, true, false, true> : conv3x3_tf32_kernel<false, false, false, true>)
                     : (relu == 2 ? conv3x3_tf32_kernel<false, true, false> : conv3x3_tf32_kernel<false, false, false>);
    if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCvSmemBytes),
                         "cudaFuncSetAttribute(conv3x3)")) != SAD_OK)
      return rc;
    const uint32_t grid = a.total_tiles < (uint32_t)sms ? a.total_tiles : (uint32_t)sms;
    kern<<<grid, kCvThreads, kCvSmemBytes, st>>>(a);
  }
  count_launch(1);
  return check_cuda(cudaGetLastError(), "conv3x3 launch");
}

SAD_EXPORT int sad_conv3x3_pack_weights_multi_f32(const sad_pack_item* items, int n_items, void* stream) {
  return pack_weights_multi_impl(items, n_items, stream, false);
}
SAD_EXPORT int sad_conv3x3_pack_weights_multi_f16(const sad_pack_item* items, int n_items, void* stream) {
  return pack_weights_multi_impl(items, n_items, stream, true);
}
SAD_EXPORT int sad_nchw_to_nhwc_f32(const sad_layout_level* levels, int n_levels, int channels, void* stream) {
  return nchw_to_nhwc_impl(levels, n_levels, channels, stream, false);
}
SAD_EXPORT int sad_nchw_to_nhwc_f16(const sad_layout_level* levels, int n_levels, int channels, int channels_dst, float scale, void* stream) {
  return nchw_to_nhwc_impl(levels, n_levels, channels, stream, true, channels_dst, scale);
}
SAD_EXPORT int sad_conv3x3_fwd_f32(const sad_conv_level* levels, int n_levels, const float* packed, const float* bias, int cin,
                                   int cout, int relu, void* stream) {
  return conv3x3_fwd_impl(levels, n_levels, packed, bias, cin, cout, relu, stream, false);
}
// 3xTF32: split operands (sad_b200.h)
SAD_EXPORT int sad_conv3x3_split_channels(int channels) { return channels < 1 ? 0 : (channels + 31) & ~31; }
SAD_EXPORT size_t sad_conv3x3_packed_bytes_f32x3(int cin, int cout, int mode) {
  if (cin < 1 || cout < 1 || (mode != 0 && mode != 1)) return 0;
  const int M = mode == 0 ? cout : cin, K = mode == 0 ? cin : cout;
  return (size_t)9 * M * 2 * sad_conv3x3_split_channels(K) * sizeof(float);
}
SAD_EXPORT int sad_nchw_to_nhwc_f32x3(const sad_layout_level* levels, int n_levels, int channels, void* stream) {
  return nchw_to_nhwc_impl(levels, n_levels, channels, stream, false, sad_conv3x3_split_channels(channels), 1.f, true);
}
SAD_EXPORT int sad_conv3x3_pack_weights_multi_f32x3(const sad_pack_item* items, int n_items, void* stream) {
  return pack_weights_multi_impl(items, n_items, stream, false, true);
}
SAD_EXPORT int sad_conv3x3_fwd_f32x3(const sad_conv_level* levels, int n_levels, const float* packed, const float* bias, int cin,
                                     int cout, int relu, void* stream) {
  return conv3x3_fwd_impl(levels, n_levels, packed, bias, cin, cout, relu, stream, false, 1.f, true);
}
SAD_EXPORT int sad_conv3x3_fwd_f16(const sad_conv_level* levels, int n_levels, const void* packed_f16, const float* bias, int cin,
                                   int cout, int relu, float nchw_scale, void* stream) {
  return conv3x3_fwd_impl(levels, n_levels, static_cast<const float*>(packed_f16), bias, cin, cout, relu, stream, true, nchw_scale);
}

}  // extern "C"
