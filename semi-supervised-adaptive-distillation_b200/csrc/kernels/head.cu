// RetinaNet FPN head, forward and backward, as ONE object over the tensor-core convolution kernels.
//
// Replaces the per-level operator chains that add_fpn_retinanet_outputs emits
// (detectron/lib/modeling/retinanet_heads.py:63-245): for each FPN level 4 x (Conv 3x3 dim->dim + in-place
// Relu) + Conv 3x3 dim->A*C for classification, the same tower + Conv 3x3 dim->A*4 for box regression,
// weights shared by all levels (level k_min owns them, the others use ConvShared), and what Caffe2
// autograd appends for them (ConvGradient per conv and level, ReluGradient, Sum of the per-level dW/db
// of a shared weight, Sum of the two towers' gradients into fpn_L: caffe2/caffe2/python/core.py:695,706-842,
// caffe2/caffe2/operators/conv_gradient_op.cc:35-77).
//
// The reference runs ~10 cuDNN calls + 8 Relu launches per level and direction, each followed by a
// stream synchronisation (operator.h:369-382).  Here one direction of the head is
//   forward : 1 layout pass + 1 weight-pack launch + 10 convolution launches (all levels per launch,
//             bias + ReLU fused, activations kept channels-last for the next layer and for backward)
//   backward: 2 layout passes + per conv {1 weight-gradient launch + bias partials + finish, 1 data-gradient
//             launch with ReluGradient fused (from 1-bit sign planes the forward pass leaves)}; the two towers
//             run on two streams, and each tower's weight gradients on a further stream behind its data-gradient chain.
// Boundary tensors keep the operator contract: fpn_L, logits, box deltas and their gradients are NCHW
// fp32; weights (Cout, Cin, 3, 3); gradients in the weights' layouts.
#include <cuda_runtime.h>

#include <new>
#include <string>
#include <vector>

#include "sad_b200.h"
#include "sad_internal.h"

using namespace sad;

struct sad_head {
  sad_head_config cfg{};
  int device = 0;
  size_t pixels[SAD_MAX_LEVELS] = {};
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  // channels-last activations: x0 = fpn input, act[t][i] = output of tower t's conv i (post-ReLU)
  float* x0[SAD_MAX_LEVELS] = {};
  float* act[2][SAD_HEAD_MAX_CONVS][SAD_MAX_LEVELS] = {};
  // sign bits of act (1 bit per element): what the fused ReluGradient of the backward pass reads
  uint32_t* bits[2][SAD_HEAD_MAX_CONVS][SAD_MAX_LEVELS] = {};
  // channels-last gradients: gpred[t] = d(prediction) (Cout = pred_out[t]); g[t][i] = d(output of tower conv i) (dim).
  // One buffer per layer (no ping-pong): the weight-gradient stream may still read g[t][i] while the data-gradient
  // stream is two layers further down
  float* gpred[2][SAD_MAX_LEVELS] = {};
  float* g[2][SAD_HEAD_MAX_CONVS][SAD_MAX_LEVELS] = {};
  // packed weights [mode][tower][conv], conv index num_convs = prediction conv
  float* packed[2][2][SAD_HEAD_MAX_CONVS + 1] = {};
  void* wg_ws[2] = {};
  size_t wg_ws_bytes = 0;
  cudaStream_t s_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_main = nullptr;
  // backward: per tower the weight gradients (wgrad + bias partials + finish) run on their own stream behind the
  // data-gradient chain, so their small tail kernels and the ragged last wave of the data-gradient kernels fill each other
  cudaStream_t s_wg[2] = {};
  cudaEvent_t ev_dy[2][SAD_HEAD_MAX_CONVS + 1] = {};   // "the gradient of conv i's output is ready"
  cudaEvent_t ev_wg_done[2] = {};
  bool packed_bwd_valid = false;
};

namespace {

size_t align256(size_t b) { return (b + 255) / 256 * 256; }
int split32(int c) { return (c + 31) & ~31; }   // 3xTF32 head: offset of the lo half of a split row (rows are 2 * split32(C) floats)
int pad8(int c) { return (c + 7) & ~7; }   // fp16 gradient tensors: channel counts padded to 16-byte rows (36 -> 40)
float grad_scale(const sad_head* h) { return h->cfg.f16_grad_scale > 0.f ? h->cfg.f16_grad_scale : 4096.f; }

int pred_out(const sad_head_config& c, int tower) { return tower == 0 ? c.cls_out : c.bbox_out; }

const float* tower_w(const sad_head_weights* w, int t, int i) { return t == 0 ? w->cls_tower_w[i] : w->bbox_tower_w[i]; }
const float* tower_b(const sad_head_weights* w, int t, int i) { return t == 0 ? w->cls_tower_b[i] : w->bbox_tower_b[i]; }
const float* pred_w(const sad_head_weights* w, int t) { return t == 0 ? w->cls_pred_w : w->bbox_pred_w; }
const float* pred_b(const sad_head_weights* w, int t) { return t == 0 ? w->cls_pred_b : w->bbox_pred_b; }

int validate_weights(const sad_head* h, const sad_head_weights* w, const char* who) {
  if (!w) return set_error(SAD_ERR_INVALID, std::string(who) + ": null weights");
  for (int t = 0; t < 2; ++t) {
    for (int i = 0; i < h->cfg.num_convs; ++i)
      if (!tower_w(w, t, i)) return set_error(SAD_ERR_INVALID, std::string(who) + ": null tower weight");
    if (!pred_w(w, t)) return set_error(SAD_ERR_INVALID, std::string(who) + ": null prediction weight");
  }
  return SAD_OK;
}

// one launch packs every weight of the head for the forward (mode 0) and, when training, the data-gradient (mode 1) pass
int pack_all(sad_head* h, const sad_head_weights* w, bool with_bwd, cudaStream_t st) {
  sad_pack_item items[SAD_MAX_PACK_ITEMS];
  int n = 0;
  const int dim = h->cfg.dim, nc = h->cfg.num_convs;
  for (int mode = 0; mode < (with_bwd ? 2 : 1); ++mode)
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i <= nc; ++i) {
        sad_pack_item& it = items[n++];
        it.weight = i < nc ? tower_w(w, t, i) : pred_w(w, t);
        it.packed = h->packed[mode][t][i];
        it.cin = dim;
        it.cout = i < nc ? dim : pred_out(h->cfg, t);
        it.mode = mode;
      }
  h->packed_bwd_valid = with_bwd;
  if (h->cfg.compute_f32x3) return sad_conv3x3_pack_weights_multi_f32x3(items, n, st);
  return h->cfg.compute_f16 ? sad_conv3x3_pack_weights_multi_f16(items, n, st) : sad_conv3x3_pack_weights_multi_f32(items, n, st);
}

int conv_levels(const sad_head* h, float* const* x, float* const* y_nchw, float* const* y_nhwc, uint32_t* const* bits_out,
                uint32_t* const* bits_in, int accumulate, const float* packed, const float* bias, int cin, int cout, int relu,
                cudaStream_t st, float nchw_scale = 1.f) {
  sad_conv_level lv[SAD_MAX_LEVELS];
  for (int l = 0; l < h->cfg.n_levels; ++l) {
    lv[l].x_nhwc = x[l];
    lv[l].y_nchw = y_nchw ? y_nchw[l] : nullptr;
    lv[l].y_nhwc = y_nhwc ? y_nhwc[l] : nullptr;
    lv[l].N = h->cfg.N;
    lv[l].H = h->cfg.H[l];
    lv[l].W = h->cfg.W[l];
    lv[l].relu_mask_nhwc = nullptr;
    lv[l].relu_bits_out = bits_out ? bits_out[l] : nullptr;
    lv[l].relu_bits_in = bits_in ? bits_in[l] : nullptr;
    lv[l].accumulate_nchw = accumulate;
  }
  // fp16 head: the arena's channels-last / packed buffers hold fp16 elements (half of each fp32-sized slot is used)
  if (h->cfg.compute_f16) return sad_conv3x3_fwd_f16(lv, h->cfg.n_levels, packed, bias, pad8(cin), cout, relu, nchw_scale, st);
  // 3xTF32 head: every channels-last tensor of the arena is a split [hi | lo] tensor
  if (h->cfg.compute_f32x3) return sad_conv3x3_fwd_f32x3(lv, h->cfg.n_levels, packed, bias, cin, cout, relu, st);
  return sad_conv3x3_fwd_f32(lv, h->cfg.n_levels, packed, bias, cin, cout, relu, st);
}

// scale: fp16 head only (the loss scale of gradient tensors; 1 for activations)
int layout_levels(const sad_head* h, const float* const* src, float* const* dst, int channels, cudaStream_t st, float scale = 1.f) {
  sad_layout_level lv[SAD_MAX_LEVELS];
  for (int l = 0; l < h->cfg.n_levels; ++l) {
    lv[l].src_nchw = src[l];
    lv[l].dst_nhwc = dst[l];
    lv[l].N = h->cfg.N;
    lv[l].H = h->cfg.H[l];
    lv[l].W = h->cfg.W[l];
  }
  if (h->cfg.compute_f32x3) return sad_nchw_to_nhwc_f32x3(lv, h->cfg.n_levels, channels, st);
  return h->cfg.compute_f16 ? sad_nchw_to_nhwc_f16(lv, h->cfg.n_levels, channels, pad8(channels), scale, st)
                            : sad_nchw_to_nhwc_f32(lv, h->cfg.n_levels, channels, st);
}

int wgrad_levels(const sad_head* h, float* const* x, float* const* dy, int cin, int cout, float* dw, float* db, int accumulate, void* ws,
                 cudaStream_t st) {
  sad_wgrad_level lv[SAD_MAX_LEVELS];
  for (int l = 0; l < h->cfg.n_levels; ++l) {
    lv[l].x_nhwc = x[l];
    lv[l].dy_nhwc = dy[l];
    lv[l].N = h->cfg.N;
    lv[l].H = h->cfg.H[l];
    lv[l].W = h->cfg.W[l];
  }
  if (h->cfg.compute_f16)
    return sad_conv3x3_wgrad_f16(lv, h->cfg.n_levels, cin, pad8(cout), cout, 1.f / grad_scale(h), dw, db, accumulate, ws, h->wg_ws_bytes, st);
  if (h->cfg.compute_f32x3) return sad_conv3x3_wgrad_f32x3(lv, h->cfg.n_levels, cin, cout, dw, db, accumulate, ws, h->wg_ws_bytes, st);
  return sad_conv3x3_wgrad_f32(lv, h->cfg.n_levels, cin, cout, dw, db, accumulate, ws, h->wg_ws_bytes, st);
}

}  // namespace

extern "C" {

SAD_EXPORT void sad_head_default_config(sad_head_config* c) {
  if (!c) return;
  *c = sad_head_config{};
  c->dim = 256;        // cfg.FPN.DIM (config.py:701)
  c->num_convs = 4;    // cfg.RETINANET.NUM_CONVS (config.py:503-566)
  c->cls_out = 9 * 80; // A * (NUM_CLASSES - 1), retinanet_heads.py:72,79-81
  c->bbox_out = 9 * 4; // A * 4, retinanet_heads.py:83-85
}

SAD_EXPORT int sad_head_create(const sad_head_config* cfg, sad_head** out) {
  if (!cfg || !out) return set_error(SAD_ERR_INVALID, "sad_head_create: null argument");
  *out = nullptr;
  if (cfg->n_levels < 1 || cfg->n_levels > SAD_MAX_LEVELS) return set_error(SAD_ERR_INVALID, "sad_head_create: n_levels must be in [1, 8]");
  if (cfg->N < 1 || cfg->dim < 1 || cfg->cls_out < 1 || cfg->bbox_out < 1)
    return set_error(SAD_ERR_INVALID, "sad_head_create: N, dim, cls_out, bbox_out must be positive");
  if (cfg->num_convs < 0 || cfg->num_convs > SAD_HEAD_MAX_CONVS) return set_error(SAD_ERR_INVALID, "sad_head_create: num_convs must be in [0, 8]");
  if (2 * 2 * (cfg->num_convs + 1) > SAD_MAX_PACK_ITEMS) return set_error(SAD_ERR_INVALID, "sad_head_create: too many convolutions");
  if (cfg->compute_f16 && cfg->compute_f32x3) return set_error(SAD_ERR_INVALID, "sad_head_create: compute_f16 and compute_f32x3 exclude each other");
  sad_head* h = new (std::nothrow) sad_head();
  if (!h) return set_error(SAD_ERR_CUDA, "sad_head_create: out of host memory");
  h->cfg = *cfg;
  int rc = check_cuda(cudaGetDevice(&h->device), "cudaGetDevice");
  if (rc != SAD_OK) {
    delete h;
    return rc;
  }
  const int L = cfg->n_levels, nc = cfg->num_convs, dim = cfg->dim;
  size_t total_pixels = 0;
  for (int l = 0; l < L; ++l) {
    if (cfg->H[l] < 0 || cfg->W[l] < 0) {
      delete h;
      return set_error(SAD_ERR_INVALID, "sad_head_create: negative level size");
    }
    h->pixels[l] = (size_t)cfg->N * cfg->H[l] * cfg->W[l];
    total_pixels += h->pixels[l];
  }
  // workspace for the weight gradients: the largest of the three shapes
  sad_wgrad_level wl[SAD_MAX_LEVELS];
  for (int l = 0; l < L; ++l) {
    wl[l].x_nhwc = wl[l].dy_nhwc = nullptr;
    wl[l].N = cfg->N;
    wl[l].H = cfg->H[l];
    wl[l].W = cfg->W[l];
  }
  size_t wsb = sad_conv3x3_wgrad_workspace_bytes(wl, L, dim, dim);
  const size_t wsb_cls = sad_conv3x3_wgrad_workspace_bytes(wl, L, dim, pad8(cfg->cls_out));
  const size_t wsb_box = sad_conv3x3_wgrad_workspace_bytes(wl, L, dim, pad8(cfg->bbox_out));   // the fp16 path pads dY's channels
  if (wsb_cls > wsb) wsb = wsb_cls;
  if (wsb_box > wsb) wsb = wsb_box;
  h->wg_ws_bytes = align256(wsb);

  // arena layout
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += align256(bytes);
    return o;
  };
  std::vector<std::pair<float**, size_t>> slots;
  auto slot = [&](float** p, size_t floats) { slots.emplace_back(p, take(floats * sizeof(float))); };
  const bool x3 = cfg->compute_f32x3 != 0;
  const size_t dim_row = x3 ? 2 * (size_t)split32(dim) : (size_t)dim;   // floats per pixel of a dim-channel channels-last tensor
  for (int l = 0; l < L; ++l) {
    slot(&h->x0[l], h->pixels[l] * dim_row);
    for (int t = 0; t < 2; ++t) {
      for (int i = 0; i < nc; ++i) {
        slot(&h->act[t][i][l], h->pixels[l] * dim_row);
        slot(reinterpret_cast<float**>(&h->bits[t][i][l]), sad_conv3x3_sign_bits_bytes(cfg->N, dim, cfg->H[l], cfg->W[l]) / sizeof(float));
      }
      // sized for the fp16 form too: channels padded to a multiple of 8 (2-byte elements; pad8(c) * 2 <= c * 4 only for c >= 4)
      slot(&h->gpred[t][l], h->pixels[l] * (x3 ? 2 * (size_t)split32(pred_out(*cfg, t)) : (size_t)pad8(pred_out(*cfg, t))));
      for (int k = 0; k < nc; ++k) slot(&h->g[t][k][l], h->pixels[l] * dim_row);
    }
  }
  for (int mode = 0; mode < 2; ++mode)
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i <= nc; ++i) {
        const int co = i < nc ? dim : pred_out(*cfg, t);
        const int M = mode == 0 ? co : dim, K = mode == 0 ? dim : co;   // packed [tap][M][K]
        slot(&h->packed[mode][t][i], x3 ? (size_t)9 * M * 2 * split32(K) : (size_t)9 * pad8(dim) * pad8(co));
      }
  const size_t ws_off0 = take(h->wg_ws_bytes), ws_off1 = take(h->wg_ws_bytes);
  h->arena_bytes = off ? off : 256;
  if ((rc = check_cuda(cudaMalloc(reinterpret_cast<void**>(&h->arena), h->arena_bytes), "sad_head_create: cudaMalloc")) != SAD_OK) {
    delete h;
    return rc;
  }
  // pad channels of split (3xTF32) and padded (fp16) rows are never written by the kernels and must read as zero
  if ((rc = check_cuda(cudaMemset(h->arena, 0, h->arena_bytes), "sad_head_create: cudaMemset")) != SAD_OK) {
    cudaFree(h->arena);
    delete h;
    return rc;
  }
  for (auto& s : slots) *s.first = reinterpret_cast<float*>(h->arena + s.second);
  h->wg_ws[0] = h->arena + ws_off0;
  h->wg_ws[1] = h->arena + ws_off1;
  if ((rc = check_cuda(cudaStreamCreateWithFlags(&h->s_side, cudaStreamNonBlocking), "cudaStreamCreate")) != SAD_OK ||
      (rc = check_cuda(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming), "cudaEventCreate")) != SAD_OK ||
      (rc = check_cuda(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming), "cudaEventCreate")) != SAD_OK ||
      (rc = check_cuda(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming), "cudaEventCreate")) != SAD_OK) {
    sad_head_destroy(h);
    return rc;
  }
  for (int t = 0; t < 2; ++t) {
    rc = check_cuda(cudaStreamCreateWithFlags(&h->s_wg[t], cudaStreamNonBlocking), "cudaStreamCreate");
    if (rc == SAD_OK) rc = check_cuda(cudaEventCreateWithFlags(&h->ev_wg_done[t], cudaEventDisableTiming), "cudaEventCreate");
    for (int i = 0; i <= nc && rc == SAD_OK; ++i)
      rc = check_cuda(cudaEventCreateWithFlags(&h->ev_dy[t][i], cudaEventDisableTiming), "cudaEventCreate");
    if (rc != SAD_OK) {
      sad_head_destroy(h);
      return rc;
    }
  }
  *out = h;
  return SAD_OK;
}

SAD_EXPORT void sad_head_destroy(sad_head* h) {
  if (!h) return;
  if (h->s_side) cudaStreamDestroy(h->s_side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  for (int t = 0; t < 2; ++t) {
    if (h->s_wg[t]) cudaStreamDestroy(h->s_wg[t]);
    if (h->ev_wg_done[t]) cudaEventDestroy(h->ev_wg_done[t]);
    for (int i = 0; i <= SAD_HEAD_MAX_CONVS; ++i)
      if (h->ev_dy[t][i]) cudaEventDestroy(h->ev_dy[t][i]);
  }
  if (h->arena) cudaFree(h->arena);
  delete h;
}

SAD_EXPORT size_t sad_head_device_bytes(const sad_head* h) { return h ? h->arena_bytes : 0; }

SAD_EXPORT int sad_head_set_f16_grad_scale(sad_head* h, float scale) {
  if (!h || !(scale > 0.f) || !(scale < 3.0e38f)) return set_error(SAD_ERR_INVALID, "sad_head_set_f16_grad_scale: the scale must be positive and finite");
  h->cfg.f16_grad_scale = scale;
  return SAD_OK;
}
SAD_EXPORT float sad_head_f16_grad_scale(const sad_head* h) { return h ? grad_scale(h) : 0.f; }

SAD_EXPORT int sad_head_copy_activation(const sad_head* h, int tower, int conv, int level, float* dst_nhwc, void* stream) {
  if (!h || !dst_nhwc || tower < 0 || tower > 1 || level < 0 || level >= h->cfg.n_levels || conv < -1 || conv >= h->cfg.num_convs)
    return set_error(SAD_ERR_INVALID, "sad_head_copy_activation: bad argument");
  const float* src = conv < 0 ? h->x0[level] : h->act[tower][conv][level];
  // an fp16 head keeps fp16 activations: dst then receives pixels * dim fp16 elements
  // a 3xTF32 head keeps split rows [hi | lo]: dst then receives pixels * 2 * round_up(dim, 32) floats
  const size_t bytes = h->cfg.compute_f32x3 ? h->pixels[level] * 2 * (size_t)split32(h->cfg.dim) * sizeof(float)
                                            : h->pixels[level] * h->cfg.dim * (h->cfg.compute_f16 ? 2 : sizeof(float));
  if (bytes == 0) return SAD_OK;
  return check_cuda(cudaMemcpyAsync(dst_nhwc, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)),
                    "sad_head_copy_activation");
}

SAD_EXPORT int sad_head_forward(sad_head* h, const sad_head_weights* w, const float* const* fpn_nchw, float* const* cls_logits_nchw,
                                float* const* bbox_pred_nchw, int training, void* stream) {
  if (!h || !fpn_nchw || !cls_logits_nchw || !bbox_pred_nchw) return set_error(SAD_ERR_INVALID, "sad_head_forward: null argument");
  int rc;
  if ((rc = validate_weights(h, w, "sad_head_forward")) != SAD_OK) return rc;
  if (h->cfg.compute_f16 && (h->cfg.dim % 8))
    return set_error(SAD_ERR_UNSUPPORTED, "sad_head_forward: an fp16 head needs dim % 8 == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = h->cfg.num_convs, dim = h->cfg.dim;
  if ((rc = layout_levels(h, fpn_nchw, h->x0, dim, st)) != SAD_OK) return rc;
  if ((rc = pack_all(h, w, training != 0, st)) != SAD_OK) return rc;
  // fork: box tower on the side stream
  if ((rc = check_cuda(cudaEventRecord(h->ev_fork, st), "cudaEventRecord")) != SAD_OK) return rc;
  if ((rc = check_cuda(cudaStreamWaitEvent(h->s_side, h->ev_fork, 0), "cudaStreamWaitEvent")) != SAD_OK) return rc;
  for (int t = 0; t < 2; ++t) {
    cudaStream_t s = t == 0 ? st : h->s_side;
    float* const* in = h->x0;
    for (int i = 0; i < nc; ++i) {
      // Conv + in-place Relu (retinanet_heads.py:101-124, 188-209); only the channels-last copy is materialised
      if ((rc = conv_levels(h, in, nullptr, h->act[t][i], training ? h->bits[t][i] : nullptr, nullptr, 0, h->packed[0][t][i],
                            tower_b(w, t, i), dim, dim, 1, s)) != SAD_OK)
        return rc;
      in = h->act[t][i];
    }
    float* const* out = t == 0 ? cls_logits_nchw : bbox_pred_nchw;
    const int act = (t == 0 && h->cfg.cls_output_sigmoid) ? 2 : 0;
    if ((rc = conv_levels(h, in, out, nullptr, nullptr, nullptr, 0, h->packed[0][t][nc], pred_b(w, t), dim, pred_out(h->cfg, t), act, s)) != SAD_OK)
      return rc;
  }
  if ((rc = check_cuda(cudaEventRecord(h->ev_join, h->s_side), "cudaEventRecord")) != SAD_OK) return rc;
  return check_cuda(cudaStreamWaitEvent(st, h->ev_join, 0), "cudaStreamWaitEvent");
}

SAD_EXPORT int sad_head_backward(sad_head* h, const sad_head_weights* w, const float* const* d_cls_logits_nchw,
                                 const float* const* d_bbox_pred_nchw, const sad_head_grads* grads, float* const* d_fpn_nchw,
                                 int accumulate, void* stream) {
  if (!h || !grads) return set_error(SAD_ERR_INVALID, "sad_head_backward: null argument");
  if (!d_cls_logits_nchw && !d_bbox_pred_nchw) return set_error(SAD_ERR_INVALID, "sad_head_backward: no output gradient given");
  if (h->cfg.cls_output_sigmoid)
    return set_error(SAD_ERR_UNSUPPORTED, "sad_head_backward: a head whose classification output is Sigmoid(logits) is forward-only (the teacher)");
  if (!h->packed_bwd_valid)
    return set_error(SAD_ERR_INVALID, "sad_head_backward: call sad_head_forward(training = 1) first (it keeps the activations and packs the weights)");
  int rc;
  if ((rc = validate_weights(h, w, "sad_head_backward")) != SAD_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = h->cfg.num_convs, dim = h->cfg.dim;
  const bool two = d_cls_logits_nchw && d_bbox_pred_nchw;
  if (two) {
    if ((rc = check_cuda(cudaEventRecord(h->ev_fork, st), "cudaEventRecord")) != SAD_OK) return rc;
    if ((rc = check_cuda(cudaStreamWaitEvent(h->s_side, h->ev_fork, 0), "cudaStreamWaitEvent")) != SAD_OK) return rc;
  }
  bool fpn_written = false;
  for (int t = 0; t < 2; ++t) {
    const float* const* dpred = t == 0 ? d_cls_logits_nchw : d_bbox_pred_nchw;
    if (!dpred) continue;
    cudaStream_t s = (t == 1 && two) ? h->s_side : st;
    void* ws = h->wg_ws[t];
    const int po = pred_out(h->cfg, t);
    float* const* pred_in = nc > 0 ? h->act[t][nc - 1] : h->x0;
    float* dwp = t == 0 ? grads->cls_pred_w : grads->bbox_pred_w;
    float* dbp = t == 0 ? grads->cls_pred_b : grads->bbox_pred_b;
    if (!dwp) return set_error(SAD_ERR_INVALID, "sad_head_backward: null prediction weight gradient");
    cudaStream_t sw = h->s_wg[t];   // weight gradients of this tower
    // hand a finished output gradient to the weight-gradient stream
    auto dy_ready = [&](int i) -> int {
      int r = check_cuda(cudaEventRecord(h->ev_dy[t][i], s), "cudaEventRecord");
      if (r != SAD_OK) return r;
      return check_cuda(cudaStreamWaitEvent(sw, h->ev_dy[t][i], 0), "cudaStreamWaitEvent");
    };
    // fp16 head: the output gradients are multiplied by the loss scale as they are rounded to fp16 (d_logits ~ 1e-6 would
    // otherwise sit in fp16's subnormal range); every gradient that leaves the head is divided by it again
    if ((rc = layout_levels(h, dpred, h->gpred[t], po, s, grad_scale(h))) != SAD_OK) return rc;
    if ((rc = dy_ready(nc)) != SAD_OK) return rc;
    if ((rc = wgrad_levels(h, pred_in, h->gpred[t], dim, po, dwp, dbp, accumulate, ws, sw)) != SAD_OK) return rc;
    // data gradient of the prediction conv; its input is the last tower activation (post-ReLU) -> ReluGradient fused
    float* const* dy = h->gpred[t];
    int dy_c = po;
    for (int i = nc; i >= 0; --i) {
      const bool last = i == 0;  // this pass produces d(fpn_L)
      if (last && !d_fpn_nchw) break;
      float* const* out_cl = last ? nullptr : h->g[t][i - 1];
      uint32_t* const* mask = last ? nullptr : h->bits[t][i - 1];
      int acc = 0;
      if (last) {
        if (t == 1 && two) {  // the other tower writes d_fpn first; this one adds to it
          if ((rc = check_cuda(cudaStreamWaitEvent(s, h->ev_main, 0), "cudaStreamWaitEvent")) != SAD_OK) return rc;
        }
        acc = fpn_written ? 1 : 0;
      }
      if ((rc = conv_levels(h, dy, last ? d_fpn_nchw : nullptr, out_cl, nullptr, mask, acc, h->packed[1][t][i], nullptr, dy_c, dim, 0, s,
                            1.f / grad_scale(h))) != SAD_OK)
        return rc;
      if (last) {
        fpn_written = true;
        if (t == 0 && two && (rc = check_cuda(cudaEventRecord(h->ev_main, s), "cudaEventRecord")) != SAD_OK) return rc;
        break;
      }
      // tower conv i-1: weight gradient from its input and the gradient just produced, on the weight-gradient stream
      float* const* in = i - 1 > 0 ? h->act[t][i - 2] : h->x0;
      float* dwt = t == 0 ? grads->cls_tower_w[i - 1] : grads->bbox_tower_w[i - 1];
      float* dbt = t == 0 ? grads->cls_tower_b[i - 1] : grads->bbox_tower_b[i - 1];
      if (!dwt) return set_error(SAD_ERR_INVALID, "sad_head_backward: null tower weight gradient");
      if ((rc = dy_ready(i - 1)) != SAD_OK) return rc;
      if ((rc = wgrad_levels(h, in, out_cl, dim, dim, dwt, dbt, accumulate, ws, sw)) != SAD_OK) return rc;
      dy = out_cl;
      dy_c = dim;
    }
    // join the weight-gradient stream into the caller's stream
    if ((rc = check_cuda(cudaEventRecord(h->ev_wg_done[t], sw), "cudaEventRecord")) != SAD_OK) return rc;
    if ((rc = check_cuda(cudaStreamWaitEvent(st, h->ev_wg_done[t], 0), "cudaStreamWaitEvent")) != SAD_OK) return rc;
  }
  if (two) {
    if ((rc = check_cuda(cudaEventRecord(h->ev_join, h->s_side), "cudaEventRecord")) != SAD_OK) return rc;
    if ((rc = check_cuda(cudaStreamWaitEvent(st, h->ev_join, 0), "cudaStreamWaitEvent")) != SAD_OK) return rc;
  }
  return SAD_OK;
}

}  // extern "C"
