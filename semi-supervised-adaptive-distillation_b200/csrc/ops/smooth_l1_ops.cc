// SelectSmoothL1Loss / SelectSmoothL1LossGradient under the reference's names, arguments (beta, scale; use_gt_flag is read
// and unused there too), defaults, enforces and gradient maker (caffe2/modules/detectron/select_smooth_l1_loss_op.{h,cc,cu}:
// class select_smooth_l1_loss_op.h:27-52, schema .cc:27-92, maker .cc:94-104), forwarding to sad_select_smooth_l1_loss_f32.
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "sad_b200.h"

namespace caffe2 {

namespace {
struct SmoothL1Params {
  float beta, scale;
  explicit SmoothL1Params(OperatorBase* op)
      : beta(op->GetSingleArgument<float>("beta", 1.f)), scale(op->GetSingleArgument<float>("scale", 1.f)) {
    CAFFE_ENFORCE(beta > 0);
    CAFFE_ENFORCE(scale >= 0);
  }
};
int BoxCount(const Tensor<CUDAContext>& Y, const Tensor<CUDAContext>& L) {
  if (Y.size() == 0) return 0;
  CAFFE_ENFORCE(Y.ndim() == 2 && Y.dim32(1) == 4, "targets must be (M, 4)");
  CAFFE_ENFORCE_EQ(L.size(), Y.size(), "locations must be (M, 4)");
  return Y.dim32(0);
}
}  // namespace

template <typename T, class Context>
class SelectSmoothL1LossOp final : public Operator<Context> {
 public:
  SelectSmoothL1LossOp(const OperatorDef& def, Workspace* ws) : Operator<Context>(def, ws), p_(this) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  SmoothL1Params p_;
  Tensor<CUDAContext> scratch_;
};

template <>
bool SelectSmoothL1LossOp<float, CUDAContext>::RunOnDevice() {
  const auto& Y_hat = Input(0);
  const auto& Y = Input(1);
  const auto& L = Input(2);
  const auto& S = Input(3);
  auto* avg_loss = Output(0);
  avg_loss->Resize(vector<TIndex>());
  CAFFE_ENFORCE_EQ(Y_hat.ndim(), 4, "predictions must be (N, A*4, H, W)");
  const int M = BoxCount(Y, L);
  const size_t need = sad_smooth_l1_workspace_bytes();
  if (scratch_.ndim() == 0) {
    scratch_.Resize((TIndex)(need / sizeof(float) + 64));
    void* ws0 = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(scratch_.mutable_data<float>()) + 255) & ~(uintptr_t)255);
    CAFFE_ENFORCE(sad_workspace_init(ws0, need, context_.cuda_stream()) == SAD_OK, sad_last_error());
  }
  void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(scratch_.mutable_data<float>()) + 255) & ~(uintptr_t)255);
  const int rc = sad_select_smooth_l1_loss_f32(Y_hat.data<float>(), M ? Y.data<float>() : nullptr, M ? L.data<float>() : nullptr, S.data<float>(),
                                               Y_hat.dim32(0), Y_hat.dim32(1), Y_hat.dim32(2), Y_hat.dim32(3), M, p_.beta, p_.scale,
                                               avg_loss->mutable_data<float>(), nullptr, nullptr, ws, need, context_.cuda_stream());
  CAFFE_ENFORCE(rc == SAD_OK, "sad_select_smooth_l1_loss_f32 failed: ", sad_last_error());
  return true;
}

template <typename T, class Context>
class SelectSmoothL1LossGradientOp final : public Operator<Context> {
 public:
  SelectSmoothL1LossGradientOp(const OperatorDef& def, Workspace* ws) : Operator<Context>(def, ws), p_(this) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  SmoothL1Params p_;
};

template <>
bool SelectSmoothL1LossGradientOp<float, CUDAContext>::RunOnDevice() {
  const auto& Y_hat = Input(0);
  const auto& Y = Input(1);
  const auto& L = Input(2);
  const auto& S = Input(3);
  const auto& d_avg_loss = Input(4);
  auto* d_Y_hat = Output(0);
  d_Y_hat->ResizeLike(Y_hat);
  CAFFE_ENFORCE_EQ(Y_hat.ndim(), 4, "predictions must be (N, A*4, H, W)");
  const int M = BoxCount(Y, L);
  const int rc = sad_select_smooth_l1_loss_f32(Y_hat.data<float>(), M ? Y.data<float>() : nullptr, M ? L.data<float>() : nullptr, S.data<float>(),
                                               Y_hat.dim32(0), Y_hat.dim32(1), Y_hat.dim32(2), Y_hat.dim32(3), M, p_.beta, p_.scale, nullptr,
                                               d_avg_loss.data<float>(), d_Y_hat->mutable_data<float>(), nullptr, 0, context_.cuda_stream());
  CAFFE_ENFORCE(rc == SAD_OK, "sad_select_smooth_l1_loss_f32 failed: ", sad_last_error());
  return true;
}

REGISTER_CPU_OPERATOR(SelectSmoothL1Loss, SelectSmoothL1LossOp<float, CPUContext>);
REGISTER_CPU_OPERATOR(SelectSmoothL1LossGradient, SelectSmoothL1LossGradientOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(SelectSmoothL1Loss, SelectSmoothL1LossOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(SelectSmoothL1LossGradient, SelectSmoothL1LossGradientOp<float, CUDAContext>);

OPERATOR_SCHEMA(SelectSmoothL1Loss)
    .NumInputs(4)
    .NumOutputs(1)
    .SetDoc("RetinaNet's smooth-L1 box-regression loss over the foreground anchors of one FPN level, divided by "
            "max(fg_num, 1) and multiplied by `scale`.")
    .Arg("beta", "(float) default 1.0; transition point between the quadratic and the linear branch (must be > 0).")
    .Arg("scale", "(float) default 1.0; multiplies the loss (must be >= 0).")
    .Input(0, "Y_hat", "(N, A*4, H, W) predicted box deltas")
    .Input(1, "Y", "(M, 4) regression targets of the M foreground anchors")
    .Input(2, "locations", "(M, 4) float: image index, first channel, y, x of each foreground anchor")
    .Input(3, "normalizer", "float, element 0: number of foreground anchors over all levels")
    .Output(0, "loss", "float scalar");
OPERATOR_SCHEMA(SelectSmoothL1LossGradient).NumInputs(5).NumOutputs(1);

class GetSelectSmoothL1LossGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    return SingleGradientDef("SelectSmoothL1LossGradient", "", vector<string>{I(0), I(1), I(2), I(3), GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(SelectSmoothL1Loss, GetSelectSmoothL1LossGradient);

}  // namespace caffe2
