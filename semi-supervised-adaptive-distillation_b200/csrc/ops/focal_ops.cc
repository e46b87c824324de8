// SigmoidFocalLoss / SigmoidFocalLossGradient under the reference's names, arguments, defaults and gradient maker
// (caffe2/modules/detectron/sigmoid_focal_loss_op.{h,cc,cu}: class + args sigmoid_focal_loss_op.h:27-43,
// schema .cc:26-101, maker .cc:103-116), forwarding to sad_sigmoid_focal_loss_f32 (include/sad_b200.h).
// As in the reference the CPU registrations exist but throw "Not Implemented." (sigmoid_focal_loss_op.h:45-48).
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "sad_b200.h"

namespace caffe2 {

namespace {
sad_focal_params ReadFocalParams(OperatorBase* op) {
  sad_focal_params p;
  p.scale = op->GetSingleArgument<float>("scale", 1.f);
  p.num_classes = op->GetSingleArgument<int>("num_classes", 80);
  p.gamma = op->GetSingleArgument<float>("gamma", 1.f);
  p.alpha = op->GetSingleArgument<float>("alpha", 0.25f);
  CAFFE_ENFORCE(p.scale >= 0);
  return p;
}
void CheckFocalInputs(const Tensor<CUDAContext>& X, const Tensor<CUDAContext>& T, int num_classes) {
  CAFFE_ENFORCE_EQ(X.ndim(), 4, "logits must be (N, A*num_classes, H, W)");
  CAFFE_ENFORCE(X.dim32(1) % num_classes == 0, "channel dim ", X.dim32(1), " is not a multiple of num_classes ", num_classes);
  CAFFE_ENFORCE_EQ(T.size(), X.size() / num_classes, "labels must be (N, A, H, W)");
}
}  // namespace

template <typename T, class Context>
class SigmoidFocalLossOp final : public Operator<Context> {
 public:
  SigmoidFocalLossOp(const OperatorDef& def, Workspace* ws) : Operator<Context>(def, ws), params_(ReadFocalParams(this)) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  sad_focal_params params_;
  Tensor<CUDAContext> scratch_;
};

template <>
bool SigmoidFocalLossOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  const auto& T = Input(1);
  const auto& wp = Input(2);
  CheckFocalInputs(X, T, params_.num_classes);
  auto* avg_loss = Output(0);
  avg_loss->Resize(vector<TIndex>());
  const size_t need = sad_focal_workspace_bytes();
  if (scratch_.ndim() == 0) {  // sized and initialised once, kept across runs (the role of losses_, sigmoid_focal_loss_op.h:41)
    scratch_.Resize((TIndex)(need / sizeof(float) + 64));
    void* raw = scratch_.mutable_data<float>();
    void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(raw) + 255) & ~(uintptr_t)255);
    CAFFE_ENFORCE(sad_workspace_init(ws, need, context_.cuda_stream()) == SAD_OK, sad_last_error());
  }
  void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(scratch_.mutable_data<float>()) + 255) & ~(uintptr_t)255);
  const int rc = sad_sigmoid_focal_loss_f32(X.data<float>(), T.data<int>(), wp.data<float>(), X.dim32(0), X.dim32(1), X.dim32(2), X.dim32(3),
                                            &params_, avg_loss->mutable_data<float>(), nullptr, nullptr, 0, ws, need, context_.cuda_stream());
  CAFFE_ENFORCE(rc == SAD_OK, "sad_sigmoid_focal_loss_f32 failed: ", sad_last_error());
  return true;
}

template <typename T, class Context>
class SigmoidFocalLossGradientOp final : public Operator<Context> {
 public:
  SigmoidFocalLossGradientOp(const OperatorDef& def, Workspace* ws) : Operator<Context>(def, ws), params_(ReadFocalParams(this)) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  sad_focal_params params_;
};

template <>
bool SigmoidFocalLossGradientOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  const auto& T = Input(1);
  const auto& wp = Input(2);
  const auto& d_avg_loss = Input(InputSize() - 1);
  CheckFocalInputs(X, T, params_.num_classes);
  auto* dX = Output(0);
  dX->ResizeLike(X);
  const int rc = sad_sigmoid_focal_loss_f32(X.data<float>(), T.data<int>(), wp.data<float>(), X.dim32(0), X.dim32(1), X.dim32(2), X.dim32(3),
                                            &params_, nullptr, d_avg_loss.data<float>(), dX->mutable_data<float>(), 0, nullptr, 0,
                                            context_.cuda_stream());
  CAFFE_ENFORCE(rc == SAD_OK, "sad_sigmoid_focal_loss_f32 failed: ", sad_last_error());
  return true;
}

REGISTER_CPU_OPERATOR(SigmoidFocalLoss, SigmoidFocalLossOp<float, CPUContext>);
REGISTER_CPU_OPERATOR(SigmoidFocalLossGradient, SigmoidFocalLossGradientOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(SigmoidFocalLoss, SigmoidFocalLossOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(SigmoidFocalLossGradient, SigmoidFocalLossGradientOp<float, CUDAContext>);

OPERATOR_SCHEMA(SigmoidFocalLoss)
    .NumInputs(3)
    .NumOutputs(1)
    .SetDoc("RetinaNet's sigmoid focal loss over all anchors and classes of one FPN level, divided by max(fg_num, 1) and "
            "multiplied by `scale`.")
    .Arg("scale", "(float) default 1.0; multiplies the loss (must be >= 0).")
    .Arg("alpha", "(float) default 0.25; weight of the positive class.")
    .Arg("gamma", "(float) default 1.0; focusing exponent.")
    .Arg("num_classes", "(int) default 80; classes per anchor (no background).")
    .Input(0, "logits", "(N, A*num_classes, H, W) float")
    .Input(1, "labels", "(N, A, H, W) int32: -1 ignore, 0 background, 1..num_classes foreground class")
    .Input(2, "normalizer", "float, element 0: number of foreground anchors")
    .Output(0, "loss", "float scalar");
OPERATOR_SCHEMA(SigmoidFocalLossGradient)
    .NumInputs(4)
    .NumOutputs(1)
    .Input(0, "logits", "as the forward op")
    .Input(1, "labels", "as the forward op")
    .Input(2, "normalizer", "as the forward op")
    .Input(3, "d_loss", "float scalar: gradient of the forward output")
    .Output(0, "d_logits", "gradient with respect to the logits");

class GetSigmoidFocalLossGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    return SingleGradientDef("SigmoidFocalLossGradient", "", vector<string>{I(0), I(1), I(2), GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(SigmoidFocalLoss, GetSigmoidFocalLossGradient);

}  // namespace caffe2
