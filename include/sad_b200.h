/* sad_b200.h — C ABI of the B200-native adaptive-distillation hot path (libsad_b200.so).
 *
 * Plain pointers and sizes only; no C++ or torch types.  All data pointers are DEVICE pointers on
 * the current CUDA device unless a function name ends in _host.  `stream` is a cudaStream_t passed
 * as void* (NULL = the legacy default stream).  Every function only ENQUEUES work on `stream` and
 * never synchronises it (the reference contract: RunOnDevice must not synchronise,
 * caffe2/caffe2/core/operator.h:373-413), except the *_host entry points, which return when the
 * host output buffers are valid.
 *
 * Return value: 0 on success, SAD_ERR_* (<0) otherwise; sad_last_error() describes the failure for
 * the calling thread.  There is no CPU fallback: without a CUDA device every compute entry point
 * fails with SAD_ERR_CUDA.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * repository root).  The reference exposes these only as C++ operator classes registered with
 * REGISTER_CUDA_OPERATOR; the operator library libcaffe2_detectron_ops_gpu.so in this repo keeps
 * that C++ surface and forwards to this ABI.
 */
#ifndef SAD_B200_H_
#define SAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAD_OK 0
#define SAD_ERR_INVALID (-1)   /* bad argument (null pointer, shape, D % num_classes, scale < 0 ...) */
#define SAD_ERR_CUDA (-2)      /* CUDA runtime/driver error, including "no device" */
#define SAD_ERR_WORKSPACE (-3) /* workspace too small or misaligned */
#define SAD_ERR_UNSUPPORTED (-4)

#define SAD_MAX_LEVELS 8   /* FPN levels per fused launch (RetinaNet uses 5: P3..P7) */
#define SAD_MAX_INPUTS 16  /* PowSum inputs per launch */

const char* sad_last_error(void);
/* "sad_b200 <version> sm_100a" — also proves which library was loaded */
const char* sad_version(void);
/* number of kernels this library has launched in the calling process (bench.py: gpu_launches) */
uint64_t sad_launch_count(void);

/* Workspaces.  The kernels below reduce in two stages and need caller-owned scratch (the role of
 * the reference ops' member tensors losses_ / _buff, loss_op.h:54, pow_sum_op.h:38-40): 256-byte
 * aligned device memory of at least sad_*_workspace_bytes().  Call sad_workspace_init ONCE after
 * allocating it; afterwards it can be reused launch after launch by one caller at a time. */
int sad_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * PowSum — replaces PowSumOp<float, CUDAContext>::RunOnDevice
 *   (caffe2/modules/detectron/pow_sum_op.cu:25-43; argument `power`, pow_sum_op.h:24-41).
 * out[0] = sum_k sum_j powf(inputs[k][j], power), one fp32 scalar.  One launch for all inputs,
 * 4 B/element of HBM traffic (the reference makes 1 + 3*n_inputs launches and 12 B/element).
 * `inputs`/`sizes` are HOST arrays of n_inputs device pointers / element counts.
 * ------------------------------------------------------------------------------------------ */
size_t sad_pow_sum_workspace_bytes(const int64_t* sizes, int n_inputs);
int sad_pow_sum_f32(const float* const* inputs, const int64_t* sizes, int n_inputs, float power,
                    float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * SigmoidAdaptiveDistillLoss (+Gradient) — replaces
 *   SigmoidAdaptiveDistillLossOp<float, CUDAContext>::RunOnDevice          (...loss_op.cu:108-141)
 *   SigmoidAdaptiveDistillLossGradientOp<float, CUDAContext>::RunOnDevice  (...loss_op.cu:144-171)
 *   and their kernels (...loss_op.cu:28-67, 69-105), file
 *   caffe2/modules/detectron/sigmoid_adaptive_distillation_loss_op.cu.
 * One launch covers up to SAD_MAX_LEVELS FPN levels.  Per level, which outputs are non-NULL
 * selects the work: loss only (the forward op), d_logits only (the gradient op), or both (fused,
 * 12.05 B/element).  All levels of one call must request the same outputs.
 * ------------------------------------------------------------------------------------------ */
typedef struct sad_distill_level {
  const float* logits;       /* X: (N, D = A*num_classes, H, W) fp32, NCHW        [Input(0)] */
  const float* teacher_prob; /* T: same shape, teacher sigmoid probabilities      [Input(1)] */
  const int32_t* labels;     /* G: (N, A, H, W) int32; only `!= ignored_label` matters [Input(2)] */
  float* d_logits;           /* out: dX, same shape as X, or NULL                 [Gradient Output(0)] */
  float* loss;               /* out: fp32 scalar, or NULL                         [Output(0)] */
  const float* d_loss;       /* upstream gradient of `loss` (fp32 scalar), NULL = 1.0 [Gradient Input(4)] */
  int32_t N, D, H, W;
} sad_distill_level;

typedef struct sad_distill_params {
  float gamma;            /* arg "gamma", default 1.0 */
  float alpha;            /* arg "alpha", default 0.25 */
  float beta;             /* arg "beta", default 0.0 */
  float scale;            /* arg "scale", default 1.0, must be >= 0 (loss_op.h:38) */
  int32_t num_classes;    /* arg "num_classes", default 80 */
  int32_t ignored_label;  /* arg "ignored_label", default -1 */
} sad_distill_params;

void sad_distill_default_params(sad_distill_params* p);
size_t sad_distill_workspace_bytes(const sad_distill_level* levels, int n_levels);
/* normalizer: device fp32, element 0 is read (Input(3): PowSum's output or retnet_fg_num). */
int sad_distill_f32(const sad_distill_level* levels, int n_levels, const float* normalizer,
                    const sad_distill_params* params, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * SigmoidFocalLoss (+Gradient) — replaces
 *   SigmoidFocalLossOp<float, CUDAContext>::RunOnDevice          (caffe2/modules/detectron/sigmoid_focal_loss_op.cu:112-144)
 *   SigmoidFocalLossGradientOp<float, CUDAContext>::RunOnDevice  (caffe2/modules/detectron/sigmoid_focal_loss_op.cu:147-173)
 * and their kernels (:26-66, :68-109): the classification loss on the SAME logits and labels as the distillation loss
 * (retinanet_heads.py:277-293).  loss and/or d_logits in one pass; accumulate_grad != 0 adds the gradient into d_logits
 * (the autograd Sum of the two consumers of the logits, core.py:695,792-842).  The ignore label is the literal -1 as in
 * the reference (:42).  fg_num: device fp32, element 0 (retnet_fg_num).
 * ------------------------------------------------------------------------------------------ */
typedef struct sad_focal_params {
  float gamma;         /* arg "gamma", default 1.0 */
  float alpha;         /* arg "alpha", default 0.25 */
  float scale;         /* arg "scale", default 1.0, must be >= 0 */
  int32_t num_classes; /* arg "num_classes", default 80 */
} sad_focal_params;
void sad_focal_default_params(sad_focal_params* p);
size_t sad_focal_workspace_bytes(void); /* needed when loss != NULL; 256-byte aligned, sad_workspace_init once */
int sad_sigmoid_focal_loss_f32(const float* logits, const int32_t* labels, const float* fg_num, int N, int D, int H, int W,
                               const sad_focal_params* params, float* loss /* or NULL */, const float* d_loss /* NULL = 1.0 */,
                               float* d_logits /* or NULL */, int accumulate_grad, void* workspace, size_t workspace_bytes, void* stream);

/* SelectSmoothL1Loss (+Gradient) — replaces SelectSmoothL1LossOp / SelectSmoothL1LossGradientOp<float, CUDAContext>::RunOnDevice
 * (caffe2/modules/detectron/select_smooth_l1_loss_op.cu:90-143, 145-181; kernels :23-54, :57-86): smooth-L1 over the 4 deltas of the
 * M foreground anchors listed in `locs` (M x 4 FLOAT rows {n, first channel, y, x}) against `y` (M x 4), / max(fg_num, 1), x scale.
 * loss and/or d_y_hat (zero-filled, then the 4*M entries scattered, scaled by d_loss (NULL = 1) and scale).  M = 0: loss 0. */
size_t sad_smooth_l1_workspace_bytes(void); /* needed when loss != NULL */
int sad_select_smooth_l1_loss_f32(const float* y_hat, const float* y, const float* locs, const float* fg_num, int N, int D, int H, int W,
                                  int M, float beta, float scale, float* loss /* or NULL */, const float* d_loss /* NULL = 1.0 */,
                                  float* d_y_hat /* or NULL */, void* workspace, size_t workspace_bytes, void* stream);

/* Momentum SGD over one flat parameter buffer — replaces, per parameter blob, the Scale(2.0) of bias gradients, the
 * WeightedSum weight decay and MomentumSGDUpdate (detectron/lib/modeling/optimizer.py:95-130;
 * caffe2/caffe2/sgd/momentum_sgd_op_gpu.cu:23-54) with ONE launch.  param, grad, momentum_buf: flat fp32 buffers of
 * sum(count) elements, 16-byte aligned, updated in place exactly like the op's outputs {grad, momentum, param}.
 * Consecutive `segments` tile the buffer; per element  g' = grad_multiplier * g + weight_decay * p,
 * then  adjusted = lr * g' + momentum * m;  m = g = adjusted;  p -= adjusted   (nesterov != 0: the :44-51 form).
 * lr: device fp32 scalar (the blob `lr`). */
#define SAD_MAX_SGD_SEGMENTS 8
typedef struct sad_sgd_segment {
  int64_t count;         /* elements */
  float grad_multiplier; /* 1 for weights, 2 for biases (optimizer.py:115-121) */
  float weight_decay;    /* SOLVER.WEIGHT_DECAY for weights, 0 for biases */
} sad_sgd_segment;
int sad_momentum_sgd_f32(float* param, float* grad, float* momentum_buf, const sad_sgd_segment* segments, int n_segments,
                         const float* lr, float momentum, int nesterov, void* stream);

/* Overflow guard of mixed-precision training (BASELINE.json configs[4]: fp16 compute).  The fp16 head scales its gradient tensors by
 * a loss scale (sad_head_config.f16_grad_scale); when a gradient leaves fp16's range the parameter gradients come out inf / NaN.
 *   sad_nonfinite_flag_f32        *flag |= 1 if any of x[0 .. n) is inf or NaN (one pass, 4 B/element; run it on the REDUCED gradient
 *                                 so that every rank of a data-parallel job takes the same decision)
 *   sad_momentum_sgd_guarded_f32  sad_momentum_sgd_f32 that leaves parameters, gradients and the update history untouched when
 *                                 *skip_if_nonzero != 0: the step is skipped on the device, no host round trip
 * The caller lowers the scale (sad_head_set_f16_grad_scale) and clears the flag when it next looks at it (solver.LossScaler). */
int sad_nonfinite_flag_f32(const float* x, int64_t n, uint32_t* flag, void* stream);
int sad_momentum_sgd_guarded_f32(float* param, float* grad, float* momentum_buf, const sad_sgd_segment* segments, int n_segments,
                                 const float* lr, float momentum, int nesterov, const uint32_t* skip_if_nonzero, void* stream);

/* The same two optimiser steps as the single-blob operators the reference graph names (operator classes MomentumSGDUpdate,
 * MomentumSGD and WeightedSum in libcaffe2_detectron_ops_gpu.so forward here):
 *   MomentumSGDUpdateOp<float, CUDAContext>::RunOnDevice   caffe2/caffe2/sgd/momentum_sgd_op.h:90-127, kernel momentum_sgd_op_gpu.cu:23-54
 *     adjusted = lr[0] * grad + momentum * mom;  mom_out = grad_out = adjusted;  param_out = param - adjusted   (nesterov: :44-51)
 *     param == param_out == NULL is MomentumSGDOp (momentum_sgd_op.h:53-88: no parameter update).  In place allowed.
 *   WeightedSumOp<CUDAContext>::RunOnDevice                 caffe2/caffe2/operators/utility_ops.h:333-378
 *     out = ws[0][0] * xs[0] + ws[1][0] * xs[1] + ...  (weights are device fp32 scalars; in place only with xs[0])
 * Element expressions are the reference's, so the results are bit-identical to its operators. */
int sad_momentum_sgd_update_f32(const float* grad, const float* mom, const float* lr, const float* param /* or NULL */, float* grad_out,
                                float* mom_out, float* param_out /* or NULL */, int64_t n, float momentum, int nesterov, void* stream);
int sad_weighted_sum_f32(const float* const* xs, const float* const* ws, int n_inputs /* <= SAD_MAX_INPUTS */, float* out, int64_t n,
                         void* stream);

/* The whole loss step of add_distill_loss (detectron/lib/modeling/retinanet_heads.py:313-352) in ONE launch:
 *   normalizer_out[0] = PowSum(levels[0..n).teacher_prob, power)           (pow_sum_op.cu:25-43)
 *   levels[l].loss, levels[l].d_logits = SigmoidAdaptiveDistillLoss(+Gradient)(..., normalizer_out)  for every level
 * i.e. exactly sad_pow_sum_f32 over the levels' teacher probabilities followed by sad_distill_f32 with that
 * normaliser.  When gamma == 2, beta == 0, both outputs are requested and the tensors are 16-byte aligned with
 * H*W % 4 == 0 this is one cooperative kernel (grid-wide barrier between the two phases, teacher probabilities
 * re-read through L2, work units handed out dynamically); otherwise it runs as the two launches.  The workspace
 * (sad_distill_fused_workspace_bytes, 256-byte aligned, sad_workspace_init once) is private to this entry point. */
size_t sad_distill_fused_workspace_bytes(const sad_distill_level* levels, int n_levels, int num_classes);
int sad_distill_fused_f32(const sad_distill_level* levels, int n_levels, float power, float* normalizer_out,
                          const sad_distill_params* params, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole distillation-loss step on HOST buffers (what a host-memory caller of the reference's
 * operators pays end to end): H2D of teacher probs / logits / labels, PowSum -> normaliser,
 * fused loss + gradient for every level, D2H of the per-level losses, the normaliser and
 * (if requested) the gradients.  Copies and kernels are pipelined over internal streams.
 * The context owns device buffers, streams and events and is reused across calls.
 * ------------------------------------------------------------------------------------------ */
typedef struct sad_host_level {
  const float* logits;       /* host, (N, D, H, W) */
  const float* teacher_prob; /* host, (N, D, H, W) */
  const int32_t* labels;     /* host, (N, A, H, W) */
  float* d_logits;           /* host out, may be NULL (gradient stays on the device) */
  int32_t N, D, H, W;
} sad_host_level;

typedef struct sad_ctx sad_ctx;
int sad_ctx_create(int device, sad_ctx** out);
void sad_ctx_destroy(sad_ctx* ctx);
/* losses_out: host, n_levels floats.  normalizer_out: host, 1 float (may be NULL). */
int sad_distill_step_host(sad_ctx* ctx, const sad_host_level* levels, int n_levels, float power,
                          const sad_distill_params* params, float* losses_out, float* normalizer_out);
/* Pipeline granularity of sad_distill_step_host: a chunk is a run of whole anchors of one image of one level, capped at `bytes`
 * of logits (0 = default: environment SAD_HOST_CHUNK_BYTES, else 16 MB).  Smaller chunks shorten the un-overlapped head of the
 * pipeline (the first logits copy after the teacher probabilities) at the price of more launches; measured sweep in
 * profiles/r01l_e2e_chunk_sweep.json. */
int sad_ctx_set_host_chunk_bytes(sad_ctx* ctx, size_t bytes);
/* device address of level i's gradient from the last sad_distill_step_host call */
float* sad_ctx_device_d_logits(sad_ctx* ctx, int level);

/* ------------------------------------------------------------------------------------------
 * RetinaNet head convolution: 3x3, stride 1, pad 1, NCHW fp32, weights shared by all FPN levels —
 * replaces, for these shapes, the cuDNN calls of
 *   CudnnConvOp::RunOnDevice / DoRunWithType          (caffe2/caffe2/operators/conv_op_cudnn.cc:293-643)
 *   CudnnConvGradientOp::DoRunWithType                (caffe2/caffe2/operators/conv_op_cudnn.cc:645-1100)
 * with the semantics of ConvOp<T>::RunOnDeviceWithOrderNCHW (caffe2/caffe2/operators/conv_op_impl.h:31-180)
 * and the in-place Relu / ReluGradient that follow each tower conv (relu_op.cu:22-62,
 * retinanet_heads.py:124,209).  tcgen05 kind::tf32 tensor-core tiles fed by TMA; one launch covers
 * all levels.  fp32 tensors, 16-byte aligned.
 * ------------------------------------------------------------------------------------------ */
typedef struct sad_conv_level {
  const float* x_nhwc; /* input, channels-last (N, H, W, Cin): from sad_nchw_to_nhwc_f32 or a previous y_nhwc */
  float* y_nchw;       /* output (N, Cout, H, W) — the operator's output blob — or NULL */
  float* y_nhwc;       /* output channels-last (N, H, W, Cout), tf32-rounded, or NULL */
  int32_t N, H, W;
  const float* relu_mask_nhwc; /* NULL, or channels-last (N, H, W, Cout): outputs are zeroed where mask <= 0.  Fuses the
                                  tower's in-place ReluGradient (relu_op.cu:29-35, dX = Y > 0 ? dY : 0) into the data-gradient
                                  pass of the NEXT convolution: mask = forward output Y of the layer whose dY is produced */
  int32_t accumulate_nchw;     /* != 0: y_nchw += result instead of = (the autograd Sum when a blob has two consumers,
                                  caffe2/caffe2/python/core.py:695,792-842: fpn_L feeds both towers) */
  uint32_t* relu_bits_out;       /* NULL, or sad_conv3x3_sign_bits_bytes() bytes: receives one bit per output element,
                                    set where the (bias-added, ReLU-ed) output is > 0 — 1/32 of the traffic of a float mask */
  const uint32_t* relu_bits_in;  /* NULL, or such a bit plane (of a tensor shaped like THIS pass's output): outputs are
                                    zeroed where the bit is clear — the compact form of relu_mask_nhwc */
} sad_conv_level;
/* bytes of a sign-bit plane for an (N, channels, H, W) output: layout [N][H][ceil(W/32)][channels] uint32,
 * bit i of a word <-> x = 32 * segment + i */
size_t sad_conv3x3_sign_bits_bytes(int N, int channels, int H, int W);

/* The tensor-core kernels read activations channels-last (TMA cannot shift the innermost NCHW
 * coordinate by one element).  One launch converts every level. */
typedef struct sad_layout_level {
  const float* src_nchw; /* (N, C, H, W) */
  float* dst_nhwc;       /* (N, H, W, C), values rounded to tf32 */
  int32_t N, H, W;
} sad_layout_level;
int sad_nchw_to_nhwc_f32(const sad_layout_level* levels, int n_levels, int channels, void* stream);

/* Weights (Cout, Cin, 3, 3) are repacked once per step into [tap][M][K] (tf32-rounded):
 *   mode 0: forward operator       M = Cout, K = Cin
 *   mode 1: data gradient operator M = Cin,  K = Cout, taps flipped (dX = conv(dY, W^T flipped))
 * sad_conv3x3_packed_bytes(cin, cout) bytes of 16-byte aligned device memory. */
size_t sad_conv3x3_packed_bytes(int cin, int cout);
int sad_conv3x3_pack_weights_f32(const float* weight, int cin, int cout, int mode, float* packed, void* stream);
/* The same for up to SAD_MAX_PACK_ITEMS weight tensors in ONE launch (all 10 head weights, both modes). */
#define SAD_MAX_PACK_ITEMS 32
typedef struct sad_pack_item {
  const float* weight; /* (cout, cin, 3, 3) */
  float* packed;       /* sad_conv3x3_packed_bytes(cin, cout) bytes */
  int32_t cin, cout, mode;
} sad_pack_item;
int sad_conv3x3_pack_weights_multi_f32(const sad_pack_item* items, int n_items, void* stream);
/* y = conv3x3(x, packed) + bias (bias may be NULL), followed by the activation `relu`: 0 none, 1 ReLU (relu_op.cu:22-28),
 * 2 Sigmoid (caffe2/caffe2/operators/sigmoid_op.cu:24-29 — the teacher's retnet_cls_prob_fpnL, retinanet_heads.py:153-163).
 * `cin`/`cout` are the K/M of `packed` (for the data gradient pass cin = Cout_of_forward, cout = Cin_of_forward). */
int sad_conv3x3_fwd_f32(const sad_conv_level* levels, int n_levels, const float* packed, const float* bias, int cin,
                        int cout, int relu, void* stream);

/* The convolution with fp16 operands (tcgen05 kind::f16, fp32 accumulation, bias / activation / NCHW output in fp32) —
 * BASELINE.json configs[4]'s "mixed fp16 compute / fp32 loss accumulate": forward and, on mode-1 packed weights, data gradient
 * (the weight gradient is sad_conv3x3_wgrad_f16); the reference's counterpart is the fp16 branch of CudnnConvOp /
 * CudnnConvGradientOp (caffe2/caffe2/operators/conv_op_cudnn.cc:623-643, 1082-1100, unused by its configs).
 * The *_f16 entry points take the SAME structs as the fp32 ones; the channels-last tensors (sad_layout_level.dst_nhwc,
 * sad_conv_level.x_nhwc / y_nhwc) and the packed weights (sad_pack_item.packed: 9 * M * pad8(K) fp16 elements) then hold
 * IEEE fp16 elements behind the float-typed pointers.  relu_mask_nhwc must be NULL (ReluGradient is taken from relu_bits_in).
 * Needs Cin % 8 == 0 and 16-byte aligned tensors (SAD_ERR_UNSUPPORTED otherwise: there is no SIMT fp16 path). */
/* channels_dst: channel count of the destination rows (0 = channels; larger = zero padding, e.g. 36 -> 40 for the box-regression
 * gradient); scale: multiplied in before rounding to fp16 (1 for activations, the loss scale for gradient tensors). */
int sad_nchw_to_nhwc_f16(const sad_layout_level* levels, int n_levels, int channels, int channels_dst, float scale, void* stream);
int sad_conv3x3_pack_weights_multi_f16(const sad_pack_item* items, int n_items, void* stream);
/* nchw_scale multiplies what is stored to y_nchw (1 for the forward operator; 1 / loss scale on the data-gradient pass that
 * produces d(fpn_L)); the fp16 channels-last output is not rescaled.  The data gradient is this entry point on mode-1 packed
 * weights: cin = the padded K of the pack (sad_conv3x3_pack_weights_multi_f16 pads K with zeros to a multiple of 8). */
int sad_conv3x3_fwd_f16(const sad_conv_level* levels, int n_levels, const void* packed_f16, const float* bias, int cin, int cout,
                        int relu, float nchw_scale, void* stream);

/* The convolution in "3xTF32" arithmetic — the fp32-accurate mode.  The reference's head convolution is fp32 (cuDNN without tensor-op
 * math: caffe2/caffe2/operators/conv_op_cudnn.cc:494-498 enables it for fp16 only); tf32 operands keep 10 mantissa bits and sit at
 * ~5e-4 of max|ref| per convolution, which cannot meet a 1e-4 gate.  Here every fp32 operand v is carried as two tf32 numbers,
 * hi = rna_tf32(v) and lo = rna_tf32(v - hi), and a product is accumulated (fp32, TMEM) as W_hi X_hi + W_hi X_lo + W_lo X_hi:
 * three tensor-core passes, error ~1e-6 of max|ref| (measured: tests/test_conv_gpu.py), 3x the tensor work of the tf32 mode.
 * The *_f32x3 entry points take the SAME structs as the fp32 ones; channels-last tensors (sad_layout_level.dst_nhwc,
 * sad_conv_level.x_nhwc / y_nhwc, sad_wgrad_level.x_nhwc / dy_nhwc) are then SPLIT tensors: per pixel a row of
 * 2 * sad_conv3x3_split_channels(C) floats, [hi(0..C) zero pad | lo(0..C) zero pad]; packed weights (sad_pack_item.packed) are
 * [tap][M][2 * sad_conv3x3_split_channels(K)] floats laid out the same way (sad_conv3x3_packed_bytes_f32x3).  Pad channels of a
 * y_nhwc the caller allocates must be zero-initialised once (the kernels never write them).  NCHW outputs, weights, gradients
 * and biases are plain fp32.  relu_mask_nhwc must be NULL (ReluGradient is taken from relu_bits_in).  16-byte aligned tensors
 * (SAD_ERR_UNSUPPORTED otherwise: there is no SIMT split path). */
int sad_conv3x3_split_channels(int channels); /* round_up(channels, 32) */
size_t sad_conv3x3_packed_bytes_f32x3(int cin, int cout, int mode);
int sad_nchw_to_nhwc_f32x3(const sad_layout_level* levels, int n_levels, int channels, void* stream);
int sad_conv3x3_pack_weights_multi_f32x3(const sad_pack_item* items, int n_items, void* stream);
int sad_conv3x3_fwd_f32x3(const sad_conv_level* levels, int n_levels, const float* packed, const float* bias, int cin, int cout,
                          int relu, void* stream);

/* Relu / ReluGradient as stand-alone operators — replace ReluOp / ReluGradientOp<float, CUDAContext>::RunOnDevice
 * (caffe2/caffe2/operators/relu_op.cu:22-62): y = x > 0 ? x : 0;  dx = y > 0 ? dy : 0.  In place allowed (y == x,
 * dx == dy), as the towers use them (retinanet_heads.py:124,209). */
int sad_relu_f32(const float* x, float* y, int64_t n, void* stream);
int sad_relu_grad_f32(const float* y, const float* dy, float* dx, int64_t n, void* stream);
/* Sigmoid (caffe2/caffe2/operators/sigmoid_op.cu:24-29): y = 1 / (1 + exp(-x)), in place allowed.  The teacher graph's
 * retnet_cls_pred_fpnL -> retnet_cls_prob_fpnL (retinanet_heads.py:153-163) as a stand-alone operator. */
int sad_sigmoid_f32(const float* x, float* y, int64_t n, void* stream);

/* Scale (caffe2/caffe2/operators/scale_op.h:31-50, math::Scale caffe2/caffe2/utils/math_gpu.cu:1293-1302): y = x * alpha, in place
 * allowed.  The momentum correction of a learning-rate change (detectron/lib/modeling/detector.py:628-648) over the flat
 * momentum buffer in one launch. */
int sad_scale_f32(const float* x, float* y, int64_t n, float alpha, void* stream);

/* AffineChannel / AffineChannelGradient — replace AffineChannelOp / AffineChannelGradientOp<float, CUDAContext>::RunOnDevice
 * (caffe2/modules/detectron/affine_channel_op.cu:52-98; kernels :22-48): the frozen batch-norm of every ResNet / FPN body
 * convolution (detectron/lib/modeling/ResNet.py:219-278).  x: (N, C, H, W) viewed as N*C rows of HW elements;
 * y = x * scale[c] + bias[c] (one fused multiply-add, as the reference's expression compiles).  bias == NULL is the gradient
 * form dx = dy * scale[c] (x = dy, y = dx).  In place allowed (schema AllowInplace, affine_channel_op.cc:29,56). */
int sad_affine_channel_f32(const float* x, const float* scale, const float* bias /* or NULL */, float* y, int N, int C, int64_t HW,
                           void* stream);
/* UpsampleNearest / UpsampleNearestGradient — replace UpsampleNearestOp / UpsampleNearestGradientOp<float, CUDAContext>::
 * RunOnDevice (caffe2/modules/detectron/upsample_nearest_op.cu:116-158, 162-211; kernels :62-113): FPN's top-down path
 * (detectron/lib/modeling/FPN.py:230-249).  x: (outer, H, W) with outer = product of the leading dimensions (N*C for 4-D,
 * C for 3-D tensors); y: (outer, H*scale, W*scale), y[o][Y][X] = x[o][Y/scale][X/scale].  The gradient sums each
 * scale x scale block of dy in the reference's order (x offset outer, y offset inner) and is therefore bit-identical to it. */
int sad_upsample_nearest_f32(const float* x, float* y, int64_t outer, int H, int W, int scale, void* stream);
int sad_upsample_nearest_grad_f32(const float* dy, float* dx, int64_t outer, int H, int W, int scale, void* stream);

/* FPN top-down merge in one pass — replaces the operator pair UpsampleNearest(scale 2) + Sum that builds every merged level
 * (detectron/lib/modeling/FPN.py:230-249: td = UpsampleNearest(fpn_top); fpn_bottom = Sum([lateral, td])):
 *   out[o][Y][X][c] = lateral[o][Y][X][c] + top[o][Y/2][X/2][c]   over tensors viewed as (outer, H_out, W_out, inner)
 * inner = 1: NCHW blobs (outer = N*C); inner = C: channels-last (outer = N).  top is (outer, H_out/2, W_out/2, inner).  In place
 * with the lateral allowed (out == lateral).  9 instead of 17 B per output element; exact (one fp32 add per element).  The
 * gradient needs no kernel of its own: d(lateral) = d(out), d(top) = sad_upsample_nearest_grad_f32(d(out)). */
int sad_upsample_nearest_add_f32(const float* top, const float* lateral, float* out, int64_t outer, int H_out, int W_out, int64_t inner,
                                 void* stream);

/* Weight and bias gradient — replaces the filter/bias half of CudnnConvGradientOp::DoRunWithType
 *   (caffe2/caffe2/operators/conv_op_cudnn.cc:1011-1040: cudnnConvolutionBackwardBias / BackwardFilter)
 * and the autograd Sum over the FPN levels that share the weight (caffe2/caffe2/python/core.py:695,706-842):
 *   d_weight[co][ci][ky][kx] = sum over levels, images, pixels of dY[co][y][x] * X[ci][y+ky-1][x+kx-1]
 *   d_bias[co]               = sum over levels, images, pixels of dY[co][y][x]
 * Both operands are the channels-last tensors the forward / data-gradient passes already hold.  One
 * tensor-core launch over all levels (pixels are the reduction axis, split across CTAs) + a fixed-order
 * finish pass: deterministic.  accumulate != 0 adds into d_weight / d_bias instead of overwriting. */
typedef struct sad_wgrad_level {
  const float* x_nhwc;  /* forward input, channels-last (N, H, W, Cin) */
  const float* dy_nhwc; /* output gradient, channels-last (N, H, W, Cout) */
  int32_t N, H, W;
} sad_wgrad_level;
size_t sad_conv3x3_wgrad_workspace_bytes(const sad_wgrad_level* levels, int n_levels, int cin, int cout);
/* d_weight: (Cout, Cin, 3, 3) fp32; d_bias: (Cout) fp32 or NULL; workspace: 256-byte aligned device memory */
int sad_conv3x3_wgrad_f32(const sad_wgrad_level* levels, int n_levels, int cin, int cout, float* d_weight, float* d_bias,
                          int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* Weight (+ bias) gradient with fp16 operands: x_nhwc (N, H, W, cin) and dy_nhwc (N, H, W, dy_channels) hold fp16 elements behind
 * the float-typed pointers of sad_wgrad_level; both are MN-major UMMA operands under the plain 128-byte swizzle (16-bit types,
 * unlike tf32, have that layout).  dy_channels >= cout is the channel count of the dY tensors: gradient tensors whose channel
 * count is not a multiple of 8 (the 36 box-regression channels) are stored padded, and the pad channels never leave the partial
 * buffers.  out_scale multiplies the finished dW / db: 1 / (the loss scale the caller applied to dY to keep fp16 gradients out
 * of the subnormal range).  Workspace: sad_conv3x3_wgrad_workspace_bytes(levels, n_levels, cin, dy_channels). */
int sad_conv3x3_wgrad_f16(const sad_wgrad_level* levels, int n_levels, int cin, int dy_channels, int cout, float out_scale,
                          float* d_weight, float* d_bias, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Weight (+ bias) gradient in 3xTF32 arithmetic: x_nhwc and dy_nhwc are split tensors (see sad_conv3x3_fwd_f32x3); every pixel
 * block is accumulated as dY_hi X_hi + dY_hi X_lo + dY_lo X_hi; db sums hi + lo.  Same workspace as sad_conv3x3_wgrad_f32. */
int sad_conv3x3_wgrad_f32x3(const sad_wgrad_level* levels, int n_levels, int cin, int cout, float* d_weight, float* d_bias,
                            int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * The whole RetinaNet FPN head, forward and backward — replaces the operator chains emitted by
 *   add_fpn_retinanet_outputs              (detectron/lib/modeling/retinanet_heads.py:63-245)
 * and the gradient operators Caffe2 autograd appends for them (ConvGradient, ReluGradient, Sum of the
 * per-level gradients of shared weights, Sum of both towers' gradients into fpn_L;
 * caffe2/caffe2/python/core.py:695,706-842, caffe2/caffe2/operators/conv_gradient_op.cc:35-77).
 * Per FPN level: num_convs x (Conv 3x3 dim->dim + Relu) + Conv 3x3 dim->cls_out, and the same tower +
 * Conv 3x3 dim->bbox_out; weights shared by all levels (blob names retnet_{cls,bbox}_conv_n{i}_fpn{k_min}_{w,b},
 * retnet_{cls,bbox}_pred_fpn{k_min}_{w,b}).  All tensors fp32; fpn_L, predictions and their gradients NCHW;
 * weights (Cout, Cin, 3, 3).  The object owns the channels-last activations kept between forward and
 * backward, the packed weights and the reduction scratch (sad_head_device_bytes()).
 * ------------------------------------------------------------------------------------------ */
#define SAD_HEAD_MAX_CONVS 8
typedef struct sad_head_config {
  int32_t n_levels;             /* k_max - k_min + 1 (5) */
  int32_t N;                    /* images per GPU (TRAIN.IMS_PER_BATCH) */
  int32_t H[SAD_MAX_LEVELS];    /* level 0 = finest (fpn3) */
  int32_t W[SAD_MAX_LEVELS];
  int32_t dim;                  /* FPN.DIM = 256 */
  int32_t num_convs;            /* RETINANET.NUM_CONVS = 4 */
  int32_t cls_out;              /* A * (NUM_CLASSES - 1) = 720 */
  int32_t bbox_out;             /* A * 4 = 36 */
  int32_t cls_output_sigmoid;   /* 0: cls output = logits (retnet_cls_pred_fpnL, the student).  1: = Sigmoid(logits)
                                   (retnet_cls_prob_fpnL: what the graph adds when model.train is False, i.e. for the
                                   teacher, retinanet_heads.py:153-163) fused into the prediction convolution's epilogue */
  int32_t compute_f16;          /* 1: fp16 operands (tcgen05 kind::f16, fp32 accumulation) for every convolution of the head, forward
                                   and backward: activations, packed weights and the channels-last gradient tensors are fp16;
                                   parameters, parameter gradients and the NCHW boundary tensors stay fp32 */
  float f16_grad_scale;         /* compute_f16 only: loss scale applied to d(logits) / d(box deltas) when they are rounded to fp16
                                   and divided out of every gradient the head returns; a power of two (0 = default 4096) */
  int32_t compute_f32x3;        /* 1: 3xTF32, the fp32-accurate mode (sad_conv3x3_fwd_f32x3 / sad_conv3x3_wgrad_f32x3) for every
                                   convolution of the head, forward and backward: what matches the reference's fp32 cuDNN convolution
                                   to 1e-4; the kept activations and gradient tensors are split [hi | lo] tensors (2x the memory,
                                   3x the tensor-core work of the default tf32 mode).  Excludes compute_f16. */
} sad_head_config;
typedef struct sad_head_weights {
  const float* cls_tower_w[SAD_HEAD_MAX_CONVS];  /* (dim, dim, 3, 3) */
  const float* cls_tower_b[SAD_HEAD_MAX_CONVS];  /* (dim) or NULL */
  const float* bbox_tower_w[SAD_HEAD_MAX_CONVS];
  const float* bbox_tower_b[SAD_HEAD_MAX_CONVS];
  const float* cls_pred_w;   /* (cls_out, dim, 3, 3) */
  const float* cls_pred_b;
  const float* bbox_pred_w;  /* (bbox_out, dim, 3, 3) */
  const float* bbox_pred_b;
} sad_head_weights;
typedef struct sad_head_grads {  /* same shapes as the weights; bias gradients may be NULL */
  float* cls_tower_w[SAD_HEAD_MAX_CONVS];
  float* cls_tower_b[SAD_HEAD_MAX_CONVS];
  float* bbox_tower_w[SAD_HEAD_MAX_CONVS];
  float* bbox_tower_b[SAD_HEAD_MAX_CONVS];
  float* cls_pred_w;
  float* cls_pred_b;
  float* bbox_pred_w;
  float* bbox_pred_b;
} sad_head_grads;
typedef struct sad_head sad_head;
void sad_head_default_config(sad_head_config* cfg); /* dim 256, num_convs 4, cls_out 720, bbox_out 36; levels unset */
int sad_head_create(const sad_head_config* cfg, sad_head** out); /* on the current device */
void sad_head_destroy(sad_head* head);
size_t sad_head_device_bytes(const sad_head* head);
/* compute_f16 heads: change the loss scale of the gradient tensors (a power of two) for the following backward passes.  The scale is
 * passed to the kernels by value: a step captured in a CUDA graph has to be captured again after a change. */
int sad_head_set_f16_grad_scale(sad_head* head, float scale);
float sad_head_f16_grad_scale(const sad_head* head);
/* fpn_nchw[l]: (N, dim, H_l, W_l); cls_logits_nchw[l]: (N, cls_out, H_l, W_l); bbox_pred_nchw[l]: (N, bbox_out, H_l, W_l).
 * training != 0 also prepares the data-gradient weights (required before sad_head_backward). */
int sad_head_forward(sad_head* head, const sad_head_weights* weights, const float* const* fpn_nchw,
                     float* const* cls_logits_nchw, float* const* bbox_pred_nchw, int training, void* stream);
/* Introspection: copies a kept activation of the last forward into dst_nhwc (N, H_l, W_l, dim), channels-last,
 * tf32-rounded as stored.  tower 0 = cls, 1 = bbox; conv = -1: the head's input fpn_L, 0..num_convs-1: output of
 * that tower conv after ReLU (the blob retnet_{cls,bbox}_conv_n{conv}_fpn{L}).  For a compute_f16 head dst_nhwc receives
 * fp16 elements (N * H_l * W_l * dim * 2 bytes), fp16-rounded as stored. */
int sad_head_copy_activation(const sad_head* head, int tower, int conv, int level, float* dst_nhwc, void* stream);
/* d_cls_logits_nchw / d_bbox_pred_nchw: gradients of the two predictions (either may be NULL: that branch is
 * skipped and its weight gradients are left untouched).  d_fpn_nchw: NULL, or (N, dim, H_l, W_l) receiving the sum
 * of both towers' input gradients.  accumulate != 0 adds into the weight/bias gradients instead of overwriting. */
int sad_head_backward(sad_head* head, const sad_head_weights* weights, const float* const* d_cls_logits_nchw,
                      const float* const* d_bbox_pred_nchw, const sad_head_grads* grads, float* const* d_fpn_nchw,
                      int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAD_B200_H_ */
