// Operator classes of the drop-in library libcaffe2_detectron_ops_gpu.so: same operator names,
// inputs, outputs, arguments, defaults and error behaviour as the reference's
// caffe2/modules/detectron/{pow_sum_op, sigmoid_adaptive_distillation_loss_op}.{h,cc,cu}, with the
// device work forwarded to the sm_100a kernels behind include/sad_b200.h.
//
//   PowSum                                 (reference pow_sum_op.cc:22-38, pow_sum_op.cu:25-46)
//   SigmoidAdaptiveDistillLoss             (reference ...loss_op.cc:21-69, ...loss_op.cu:108-141)
//   SigmoidAdaptiveDistillLossGradient     (reference ...loss_op.cc:71-112, ...loss_op.cu:144-171)
//   SigmoidAdaptiveDistillLossMultiLevel   NEW: every FPN level's loss AND gradient in one launch;
//                                          produced from the three above by FuseAdaptiveDistillOps()
//
// Like the reference, the CPU registrations exist but throw "Not Implemented."
// (pow_sum_op.h:33-36, ...loss_op.h:42-45,73-76): there is no CPU fallback.
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "ops/distill_ops.h"
#include "sad_b200.h"

namespace caffe2 {

namespace {

void EnforceSad(int rc, const char* what) {
  CAFFE_ENFORCE(rc == SAD_OK, what, " failed: ", sad_last_error());
}

// Per-op scratch with the role of the reference's member tensors (_buff/_buff_sum, losses_):
// sized on first use, kept across runs, initialised once.
class KernelWorkspace {
 public:
  void* Ensure(size_t bytes, CUDAContext* context) {
    const TIndex floats = (TIndex)((bytes + 255) / 256 * 64);
    if (t_.size() < floats) {
      t_.Resize(floats);
      void* p = t_.mutable_data<float>();
      EnforceSad(sad_workspace_init(p, (size_t)floats * 4, context->cuda_stream()), "sad_workspace_init");
    }
    return t_.mutable_data<float>();
  }
  size_t bytes() const { return (size_t)t_.size() * 4; }

 private:
  Tensor<CUDAContext> t_;
};

sad_distill_params ReadDistillParams(OperatorBase* op) {
  sad_distill_params p;
  p.scale = op->GetSingleArgument<float>("scale", 1.f);
  p.num_classes = op->GetSingleArgument<int>("num_classes", 80);
  p.gamma = op->GetSingleArgument<float>("gamma", 1.f);
  p.alpha = op->GetSingleArgument<float>("alpha", 0.25f);
  p.beta = op->GetSingleArgument<float>("beta", 0.f);
  p.ignored_label = op->GetSingleArgument<int>("ignored_label", -1);
  CAFFE_ENFORCE(p.scale >= 0);
  return p;
}

void FillLevel(sad_distill_level* L, const Tensor<CUDAContext>& X, const Tensor<CUDAContext>& T,
               const Tensor<CUDAContext>& G, int num_classes) {
  CAFFE_ENFORCE_EQ(X.ndim(), 4, "logits must be (N, A*num_classes, H, W)");
  L->N = X.dim32(0);
  L->D = X.dim32(1);
  L->H = X.dim32(2);
  L->W = X.dim32(3);
  CAFFE_ENFORCE_EQ(T.size(), X.size(), "teacher probabilities must have the shape of the logits");
  CAFFE_ENFORCE(L->D % num_classes == 0, "channel dim ", L->D, " is not a multiple of num_classes ", num_classes);
  CAFFE_ENFORCE_EQ(G.size(), X.size() / num_classes, "labels must be (N, A, H, W)");
  L->logits = X.data<float>();
  L->teacher_prob = T.data<float>();
  L->labels = G.data<int>();
  L->d_logits = nullptr;
  L->loss = nullptr;
  L->d_loss = nullptr;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
template <typename T, class Context>
class PowSumOp final : public Operator<Context> {
 public:
  PowSumOp(const OperatorDef& operator_def, Workspace* ws)
      : Operator<Context>(operator_def, ws), power_(OperatorBase::GetSingleArgument<float>("power", 1.0f)) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  float power_;
  KernelWorkspace scratch_;
};

template <>
bool PowSumOp<float, CUDAContext>::RunOnDevice() {
  const int n = InputSize();
  CAFFE_ENFORCE(n <= SAD_MAX_INPUTS, "PowSum: at most ", SAD_MAX_INPUTS, " inputs per op, got ", n);
  const float* ptrs[SAD_MAX_INPUTS];
  int64_t sizes[SAD_MAX_INPUTS];
  for (int i = 0; i < n; ++i) {
    const auto& in = Input(i);
    sizes[i] = in.size();
    ptrs[i] = in.data<float>();
  }
  auto* res = Output(0);
  res->Resize(vector<TIndex>());
  void* ws = scratch_.Ensure(sad_pow_sum_workspace_bytes(sizes, n), &context_);
  EnforceSad(sad_pow_sum_f32(ptrs, sizes, n, power_, res->mutable_data<float>(), ws, scratch_.bytes(), context_.cuda_stream()),
             "sad_pow_sum_f32");
  return true;
}

// ---------------------------------------------------------------------------------------------
template <typename T, class Context>
class SigmoidAdaptiveDistillLossOp final : public Operator<Context> {
 public:
  SigmoidAdaptiveDistillLossOp(const OperatorDef& operator_def, Workspace* ws)
      : Operator<Context>(operator_def, ws), params_(ReadDistillParams(this)) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  sad_distill_params params_;
  KernelWorkspace scratch_;
};

template <>
bool SigmoidAdaptiveDistillLossOp<float, CUDAContext>::RunOnDevice() {
  sad_distill_level L;
  FillLevel(&L, Input(0), Input(1), Input(2), params_.num_classes);
  const float* wp = Input(3).data<float>();
  auto* avg_loss = Output(0);
  avg_loss->Resize(vector<TIndex>());
  L.loss = avg_loss->mutable_data<float>();
  void* ws = scratch_.Ensure(sad_distill_workspace_bytes(&L, 1), &context_);
  EnforceSad(sad_distill_f32(&L, 1, wp, &params_, ws, scratch_.bytes(), context_.cuda_stream()), "sad_distill_f32");
  return true;
}

template <typename T, class Context>
class SigmoidAdaptiveDistillLossGradientOp final : public Operator<Context> {
 public:
  SigmoidAdaptiveDistillLossGradientOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), params_(ReadDistillParams(this)) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  sad_distill_params params_;
};

template <>
bool SigmoidAdaptiveDistillLossGradientOp<float, CUDAContext>::RunOnDevice() {
  sad_distill_level L;
  const auto& X = Input(0);
  FillLevel(&L, X, Input(1), Input(2), params_.num_classes);
  const float* wp = Input(3).data<float>();
  L.d_loss = Input(InputSize() - 1).data<float>();
  auto* dX = Output(0);
  dX->ResizeLike(X);
  L.d_logits = dX->mutable_data<float>();
  EnforceSad(sad_distill_f32(&L, 1, wp, &params_, nullptr, 0, context_.cuda_stream()), "sad_distill_f32");
  return true;
}

// ---------------------------------------------------------------------------------------------
// Inputs : X_0, T_0, G_0, ..., X_{L-1}, T_{L-1}, G_{L-1}, normalizer            (3L + 1)
// Outputs: loss_0 .. loss_{L-1}, dX_0 .. dX_{L-1}                               (2L)
// Args   : those of SigmoidAdaptiveDistillLoss + "d_loss" (float, default 1.0): the constant every
//          loss gradient is filled with (Detectron: ConstantFill(value=1.0), utils/blob.py:166-172).
template <typename T, class Context>
class SigmoidAdaptiveDistillLossMultiLevelOp final : public Operator<Context> {
 public:
  SigmoidAdaptiveDistillLossMultiLevelOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), params_(ReadDistillParams(this)),
        d_loss_(OperatorBase::GetSingleArgument<float>("d_loss", 1.f)) {
    CAFFE_ENFORCE((InputSize() - 1) % 3 == 0 && InputSize() >= 4, "expected 3*L + 1 inputs");
    levels_ = (InputSize() - 1) / 3;
    CAFFE_ENFORCE(levels_ <= SAD_MAX_LEVELS, "at most ", SAD_MAX_LEVELS, " levels");
    CAFFE_ENFORCE_EQ(OutputSize(), 2 * levels_, "expected L losses followed by L gradients");
  }
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  sad_distill_params params_;
  float d_loss_;
  int levels_;
  KernelWorkspace scratch_;
  Tensor<CUDAContext> d_loss_dev_;
};

template <>
bool SigmoidAdaptiveDistillLossMultiLevelOp<float, CUDAContext>::RunOnDevice() {
  sad_distill_level L[SAD_MAX_LEVELS];
  if (d_loss_ != 1.f && d_loss_dev_.size() != 1) {
    d_loss_dev_.Resize(vector<TIndex>());
    CUDA_ENFORCE(cudaMemcpyAsync(d_loss_dev_.mutable_data<float>(), &d_loss_, sizeof(float), cudaMemcpyHostToDevice,
                                 context_.cuda_stream()));
  }
  for (int l = 0; l < levels_; ++l) {
    const auto& X = Input(3 * l);
    FillLevel(&L[l], X, Input(3 * l + 1), Input(3 * l + 2), params_.num_classes);
    auto* loss = Output(l);
    loss->Resize(vector<TIndex>());
    L[l].loss = loss->mutable_data<float>();
    auto* dX = Output(levels_ + l);
    dX->ResizeLike(X);
    L[l].d_logits = dX->mutable_data<float>();
    L[l].d_loss = d_loss_ != 1.f ? d_loss_dev_.data<float>() : nullptr;
  }
  const float* wp = Input(3 * levels_).data<float>();
  void* ws = scratch_.Ensure(sad_distill_workspace_bytes(L, levels_), &context_);
  EnforceSad(sad_distill_f32(L, levels_, wp, &params_, ws, scratch_.bytes(), context_.cuda_stream()), "sad_distill_f32");
  return true;
}

// ---------------------------------------------------------------------------------------------
// SigmoidAdaptiveDistillStep: the PowSum that produces the normaliser folded into the multi-level op — the whole
// add_distill_loss sub-graph (retinanet_heads.py:313-352) plus its gradient ops as ONE cooperative launch
// (sad_distill_fused_f32).  Produced by FuseAdaptiveDistillOps when the group's normaliser blob is the output of a
// PowSum over exactly the group's teacher-probability blobs, in the same order.
// Inputs : X_0, T_0, G_0, ..., X_{L-1}, T_{L-1}, G_{L-1}                        (3L)
// Outputs: loss_0 .. loss_{L-1}, dX_0 .. dX_{L-1}, normalizer                   (2L + 1)
// Args   : those of SigmoidAdaptiveDistillLoss + "d_loss" + "power" (PowSum's argument, default 1.0)
template <typename T, class Context>
class SigmoidAdaptiveDistillStepOp final : public Operator<Context> {
 public:
  SigmoidAdaptiveDistillStepOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), params_(ReadDistillParams(this)),
        d_loss_(OperatorBase::GetSingleArgument<float>("d_loss", 1.f)),
        power_(OperatorBase::GetSingleArgument<float>("power", 1.f)) {
    CAFFE_ENFORCE(InputSize() % 3 == 0 && InputSize() >= 3, "expected 3*L inputs");
    levels_ = InputSize() / 3;
    CAFFE_ENFORCE(levels_ <= SAD_MAX_LEVELS, "at most ", SAD_MAX_LEVELS, " levels");
    CAFFE_ENFORCE_EQ(OutputSize(), 2 * levels_ + 1, "expected L losses, L gradients and the normaliser");
  }
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  sad_distill_params params_;
  float d_loss_, power_;
  int levels_;
  KernelWorkspace scratch_;
  Tensor<CUDAContext> d_loss_dev_;
};

template <>
bool SigmoidAdaptiveDistillStepOp<float, CUDAContext>::RunOnDevice() {
  sad_distill_level L[SAD_MAX_LEVELS];
  if (d_loss_ != 1.f && d_loss_dev_.size() != 1) {
    d_loss_dev_.Resize(vector<TIndex>());
    CUDA_ENFORCE(cudaMemcpyAsync(d_loss_dev_.mutable_data<float>(), &d_loss_, sizeof(float), cudaMemcpyHostToDevice,
                                 context_.cuda_stream()));
  }
  for (int l = 0; l < levels_; ++l) {
    const auto& X = Input(3 * l);
    FillLevel(&L[l], X, Input(3 * l + 1), Input(3 * l + 2), params_.num_classes);
    auto* loss = Output(l);
    loss->Resize(vector<TIndex>());
    L[l].loss = loss->mutable_data<float>();
    auto* dX = Output(levels_ + l);
    dX->ResizeLike(X);
    L[l].d_logits = dX->mutable_data<float>();
    L[l].d_loss = d_loss_ != 1.f ? d_loss_dev_.data<float>() : nullptr;
  }
  auto* norm = Output(2 * levels_);
  norm->Resize(vector<TIndex>());   // PowSum's output is a scalar (pow_sum_op.cu:31)
  void* ws = scratch_.Ensure(sad_distill_fused_workspace_bytes(L, levels_, params_.num_classes), &context_);
  EnforceSad(sad_distill_fused_f32(L, levels_, power_, norm->mutable_data<float>(), &params_, ws, scratch_.bytes(),
                                   context_.cuda_stream()),
             "sad_distill_fused_f32");
  return true;
}

// ---------------------------------------------------------------------------------------------
// ConstantFill (subset): output shaped like input 0 (or arg "shape"), filled with float "value".
// Only here so NetDefs dumped by Detectron (loss-gradient seeds, utils/blob.py:166-172) run
// unmodified through the executor; reference: caffe2/caffe2/operators/filler_op.h.
template <class Context>
class ConstantFillOp final : public Operator<Context> {
 public:
  ConstantFillOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), value_(OperatorBase::GetSingleArgument<float>("value", 0.f)),
        shape_(OperatorBase::GetRepeatedArgument<int64_t>("shape")) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override;

 private:
  float value_;
  vector<int64_t> shape_;
  vector<float> host_;
};

template <>
bool ConstantFillOp<CUDAContext>::RunOnDevice() {
  auto* out = Output(0);
  if (InputSize()) out->ResizeLike(Input(0));
  else out->Resize(shape_);
  float* p = out->mutable_data<float>();
  if ((TIndex)host_.size() != out->size()) host_.assign(out->size(), value_);
  if (out->size())
    CUDA_ENFORCE(cudaMemcpyAsync(p, host_.data(), out->size() * sizeof(float), cudaMemcpyHostToDevice, context_.cuda_stream()));
  return true;
}

// ---------------------------------------------------------------------------------------------
// registration: CPU (unimplemented, as in the reference) + CUDA, schemas, gradient maker
// ---------------------------------------------------------------------------------------------
REGISTER_CPU_OPERATOR(PowSum, PowSumOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(PowSum, PowSumOp<float, CUDAContext>);
OPERATOR_SCHEMA(PowSum)
    .NumInputs(1, INT_MAX)
    .NumOutputs(1)
    .SetDoc("Sum over every element of every input of element^power: the adaptive normaliser "
            "of the distillation loss when fed the teacher's per-level class probabilities.")
    .Arg("power", "(float) default 1.0; exponent applied to each element before summing.")
    .Input(0, "X_0 .. X_{k-1}", "float tensors of any shape")
    .Output(0, "sum", "float scalar");

REGISTER_CPU_OPERATOR(SigmoidAdaptiveDistillLoss, SigmoidAdaptiveDistillLossOp<float, CPUContext>);
REGISTER_CPU_OPERATOR(SigmoidAdaptiveDistillLossGradient, SigmoidAdaptiveDistillLossGradientOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(SigmoidAdaptiveDistillLoss, SigmoidAdaptiveDistillLossOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(SigmoidAdaptiveDistillLossGradient, SigmoidAdaptiveDistillLossGradientOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(SigmoidAdaptiveDistillLossMultiLevel, SigmoidAdaptiveDistillLossMultiLevelOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(SigmoidAdaptiveDistillStep, SigmoidAdaptiveDistillStepOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(ConstantFill, ConstantFillOp<CUDAContext>);

OPERATOR_SCHEMA(SigmoidAdaptiveDistillLoss)
    .NumInputs(4)
    .NumOutputs(1)
    .SetDoc("Adaptive (focal-style) distillation loss between student logits and teacher sigmoid "
            "probabilities, summed over all anchors and classes and multiplied by `scale`.")
    .Arg("scale", "(float) default 1.0; multiplies the summed loss (must be >= 0).")
    .Arg("alpha", "(float) default 0.25; weight of the positive (teacher-probability) branch.")
    .Arg("gamma", "(float) default 1.0; exponent of the adaptive weight 1 - exp(-KL-like distance).")
    .Arg("beta", "(float) default 0.0; weight of the teacher-entropy term inside the distance.")
    .Arg("num_classes", "(int) default 80; classes per anchor (no background).")
    .Arg("ignored_label", "(int) default -1; anchors carrying this label contribute nothing.")
    .Input(0, "logits", "(N, A*num_classes, H, W) float student logits")
    .Input(1, "teacher_prob", "same shape: teacher sigmoid probabilities in (0, 1)")
    .Input(2, "labels", "(N, A, H, W) int32 anchor labels; only `!= ignored_label` is used")
    .Input(3, "normalizer", "float, element 0 used: loss is divided by max(normalizer, 1)")
    .Output(0, "loss", "float scalar");

OPERATOR_SCHEMA(SigmoidAdaptiveDistillLossGradient)
    .NumInputs(5)
    .NumOutputs(1)
    .Input(0, "logits", "as the forward op")
    .Input(1, "teacher_prob", "as the forward op")
    .Input(2, "labels", "as the forward op")
    .Input(3, "normalizer", "as the forward op")
    .Input(4, "d_loss", "float scalar: gradient of the forward output")
    .Output(0, "d_logits", "gradient with respect to logits; the teacher receives none");

OPERATOR_SCHEMA(SigmoidAdaptiveDistillLossMultiLevel)
    .NumInputs(4, 3 * SAD_MAX_LEVELS + 1)
    .NumOutputs(2, 2 * SAD_MAX_LEVELS)
    .SetDoc("All FPN levels of SigmoidAdaptiveDistillLoss and its gradient in one kernel launch.")
    .Arg("d_loss", "(float) default 1.0; constant upstream gradient of every level's loss.");

OPERATOR_SCHEMA(SigmoidAdaptiveDistillStep)
    .NumInputs(3, 3 * SAD_MAX_LEVELS)
    .NumOutputs(3, 2 * SAD_MAX_LEVELS + 1)
    .SetDoc("PowSum over the levels' teacher probabilities, then every level's SigmoidAdaptiveDistillLoss and gradient, "
            "in one cooperative kernel launch.")
    .Arg("power", "(float) default 1.0; PowSum's exponent.")
    .Arg("d_loss", "(float) default 1.0; constant upstream gradient of every level's loss.");

OPERATOR_SCHEMA(ConstantFill).NumInputs(0, 1).NumOutputs(1).AllowInplace({{0, 0}});

class GetSigmoidAdaptiveDistillLossGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    // gradient flows to the student logits only (reference ...loss_op.cc:99-110)
    return SingleGradientDef("SigmoidAdaptiveDistillLossGradient", "",
                             vector<string>{I(0), I(1), I(2), I(3), GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(SigmoidAdaptiveDistillLoss, GetSigmoidAdaptiveDistillLossGradient);

// ---------------------------------------------------------------------------------------------
// graph pass
// ---------------------------------------------------------------------------------------------
namespace {
bool SameArgs(const OperatorDef& a, const OperatorDef& b) {
  return OperatorDefToText([&] { OperatorDef c; for (const auto& x : a.arg()) *c.add_arg() = x; return c; }()) ==
         OperatorDefToText([&] { OperatorDef c; for (const auto& x : b.arg()) *c.add_arg() = x; return c; }());
}
}  // namespace

int FuseAdaptiveDistillOps(NetDef* net) {
  int fused_groups = 0;
  vector<OperatorDef>& ops = *net->mutable_op();
  vector<bool> consumed(ops.size(), false);
  vector<OperatorDef> result;
  for (size_t i = 0; i < ops.size(); ++i) {
    if (consumed[i]) continue;
    const OperatorDef& first = ops[i];
    if (first.type() != "SigmoidAdaptiveDistillLoss" || first.device_option().device_type() != CUDA) {
      result.push_back(first);
      continue;
    }
    // group: later forward ops with the same normaliser blob, device and arguments
    vector<size_t> fwd{i};
    for (size_t j = i + 1; j < ops.size() && (int)fwd.size() < SAD_MAX_LEVELS; ++j)
      if (!consumed[j] && ops[j].type() == first.type() && ops[j].input(3) == first.input(3) &&
          ops[j].device_option().cuda_gpu_id() == first.device_option().cuda_gpu_id() && SameArgs(ops[j], first))
        fwd.push_back(j);
    // every member's inputs must already exist where the fused op will sit (position i)
    bool ok = true;
    for (size_t f : fwd)
      for (size_t j = i; j < f && ok; ++j)
        for (const auto& out : ops[j].output())
          for (const auto& in : ops[f].input())
            if (out == in) ok = false;
    // every member needs its gradient op whose d_loss comes from a ConstantFill of one common value
    vector<size_t> grad, fill;
    float d_loss_value = 0.f;
    for (size_t f : fwd) {
      if (!ok) break;
      size_t g = ops.size(), c = ops.size();
      for (size_t j = f + 1; j < ops.size(); ++j)
        if (!consumed[j] && ops[j].type() == "SigmoidAdaptiveDistillLossGradient" && ops[j].input(0) == ops[f].input(0) &&
            ops[j].input(1) == ops[f].input(1) && ops[j].input(2) == ops[f].input(2) && ops[j].input(3) == ops[f].input(3)) {
          g = j;
          break;
        }
      if (g == ops.size()) { ok = false; break; }
      for (size_t j = 0; j < g; ++j)
        if (ops[j].type() == "ConstantFill" && ops[j].output_size() == 1 && ops[j].output(0) == ops[g].input(4)) c = j;
      if (c == ops.size()) { ok = false; break; }
      const float v = ArgumentHelper::GetSingleArgument<OperatorDef, float>(ops[c], "value", 0.f);
      if (grad.empty()) d_loss_value = v;
      else if (v != d_loss_value) { ok = false; break; }
      // the baked-in d_loss is only the ConstantFill's value if nothing rewrites that blob between the fill and the gradient op
      // (e.g. a Scale on the loss gradient), and moving the gradient up to the forward op's position is only legal if no op in
      // between reads or writes the gradient's output blob
      for (size_t j = c + 1; j < g && ok; ++j)
        for (const auto& out : ops[j].output())
          if (out == ops[g].input(4)) ok = false;
      for (size_t j = i; j < g && ok; ++j) {
        if (j == f) continue;
        for (const auto& out : ops[j].output())
          if (out == ops[g].output(0)) ok = false;
        for (const auto& in : ops[j].input())
          if (in == ops[g].output(0)) ok = false;
      }
      if (!ok) break;
      // the logits must not be rewritten between the forward and the gradient op
      for (size_t j = f + 1; j < g && ok; ++j)
        for (const auto& out : ops[j].output())
          if (out == ops[f].input(0) || out == ops[f].input(1) || out == ops[f].input(2) || out == ops[f].input(3)) ok = false;
      if (!ok) break;
      grad.push_back(g);
      fill.push_back(c);
    }
    if (!ok || fwd.size() < 2) {
      result.push_back(first);
      continue;
    }
    // the normaliser: if it is the output of a PowSum (already emitted, same device) over exactly this group's
    // teacher-probability blobs in this order, and nothing between that PowSum and here rewrites them, the PowSum is
    // folded in too (retinanet_heads.py:320-328 builds it right before the loss ops)
    size_t pow_at = result.size();
    for (size_t r = 0; r < result.size(); ++r) {
      const OperatorDef& c = result[r];
      if (c.type() == "PowSum" && c.output_size() == 1 && c.output(0) == first.input(3) &&
          c.device_option().device_type() == CUDA && c.device_option().cuda_gpu_id() == first.device_option().cuda_gpu_id() &&
          c.input_size() == (int)fwd.size()) {
        bool same = true;
        for (size_t k = 0; k < fwd.size(); ++k) same = same && c.input((int)k) == ops[fwd[k]].input(1);
        if (same) pow_at = r;
      }
    }
    if (pow_at != result.size()) {
      const string& nb = first.input(3);
      for (size_t r = pow_at + 1; r < result.size(); ++r) {       // ops emitted after the PowSum
        for (const auto& in : result[r].input())
          if (in == nb) pow_at = result.size();                    // someone else already reads the normaliser
        for (const auto& out : result[r].output())
          for (size_t f : fwd)
            if (out == ops[f].input(1) || out == nb) pow_at = result.size();
        if (pow_at == result.size()) break;
      }
    }
    OperatorDef fused;
    fused.set_name(first.name());
    for (size_t f : fwd)
      for (int k = 0; k < 3; ++k) fused.add_input(ops[f].input(k));
    for (const auto& a : first.arg()) *fused.add_arg() = a;
    Argument* dl = fused.add_arg();
    dl->set_name("d_loss");
    dl->set_f(d_loss_value);
    if (pow_at != result.size()) {
      fused.set_type("SigmoidAdaptiveDistillStep");
      Argument* pw = fused.add_arg();
      pw->set_name("power");
      pw->set_f(ArgumentHelper::GetSingleArgument<OperatorDef, float>(result[pow_at], "power", 1.f));
      for (size_t f : fwd) fused.add_output(ops[f].output(0));
      for (size_t g : grad) fused.add_output(ops[g].output(0));
      fused.add_output(first.input(3));
      result.erase(result.begin() + pow_at);
    } else {
      fused.set_type("SigmoidAdaptiveDistillLossMultiLevel");
      fused.add_input(first.input(3));
      for (size_t f : fwd) fused.add_output(ops[f].output(0));
      for (size_t g : grad) fused.add_output(ops[g].output(0));
    }
    *fused.mutable_device_option() = first.device_option();
    result.push_back(fused);
    for (size_t f : fwd) consumed[f] = true;
    for (size_t g : grad) consumed[g] = true;
    ++fused_groups;
  }
  ops.swap(result);
  return fused_groups;
}

}  // namespace caffe2

extern "C" __attribute__((visibility("default"))) const char* c2_fuse_adaptive_distill_ops(const char* net_text, int* n_fused) {
  static thread_local std::string out;
  try {
    caffe2::NetDef net;
    caffe2::ParseNetDefText(net_text, &net);
    int n = caffe2::FuseAdaptiveDistillOps(&net);
    if (n_fused) *n_fused = n;
    out = caffe2::NetDefToText(net);
    return out.c_str();
  } catch (const std::exception&) {
    return nullptr;
  }
}
