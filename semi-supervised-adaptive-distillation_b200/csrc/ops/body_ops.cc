// AffineChannel(+Gradient) and UpsampleNearest(+Gradient) under the reference's names, arguments, schema arity, in-place
// permissions and gradient makers (caffe2/modules/detectron/affine_channel_op.{h,cc,cu}: classes .h:27-50, schema .cc:26-69,
// maker .cc:71-80; caffe2/modules/detectron/upsample_nearest_op.{h,cc,cu}: classes .h:27-62, schema .cc:26-59, maker
// .cc:61-72), forwarding to sad_affine_channel_f32 / sad_upsample_nearest(_grad)_f32.  CPU registrations keep the reference's
// CAFFE_NOT_IMPLEMENTED.
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "sad_b200.h"

namespace caffe2 {

namespace {
void EnforceSadBody(int rc, const char* what) { CAFFE_ENFORCE(rc == SAD_OK, what, " failed: ", sad_last_error()); }

// (outer, H, W) of a 3-D or 4-D tensor as upsample_nearest_op.cu:129-138 reads them
void OuterHW(const Tensor<CUDAContext>& X, int64_t* outer, int* H, int* W) {
  CAFFE_ENFORCE(X.ndim() == 3 || X.ndim() == 4, "UpsampleNearest takes a 3-D or 4-D tensor");
  const int nd = X.ndim();
  *W = X.dim32(nd - 1);
  *H = X.dim32(nd - 2);
  *outer = 1;
  for (int i = 0; i < nd - 2; ++i) *outer *= X.dim32(i);
}
}  // namespace

template <typename T, class Context>
class AffineChannelOp final : public Operator<Context> {
 public:
  USE_SIMPLE_CTOR_DTOR(AffineChannelOp);
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }
};

template <>
bool AffineChannelOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  const auto& scale = Input(1);
  const auto& bias = Input(2);
  auto* Y = Output(0);
  Y->ResizeLike(X);
  CAFFE_ENFORCE_EQ(X.ndim(), 4, "X must be (N, C, H, W)");
  CAFFE_ENFORCE_EQ(scale.size(), X.dim32(1), "scale must have C elements");
  CAFFE_ENFORCE_EQ(bias.size(), X.dim32(1), "bias must have C elements");
  EnforceSadBody(sad_affine_channel_f32(X.data<float>(), scale.data<float>(), bias.data<float>(), Y->mutable_data<float>(), X.dim32(0),
                                        X.dim32(1), (int64_t)X.dim32(2) * X.dim32(3), context_.cuda_stream()),
                 "sad_affine_channel_f32");
  return true;
}

template <typename T, class Context>
class AffineChannelGradientOp final : public Operator<Context> {
 public:
  USE_SIMPLE_CTOR_DTOR(AffineChannelGradientOp);
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }
};

template <>
bool AffineChannelGradientOp<float, CUDAContext>::RunOnDevice() {
  const auto& scale = Input(0);
  const auto& dY = Input(1);
  auto* dX = Output(0);
  dX->ResizeLike(dY);
  CAFFE_ENFORCE_EQ(dY.ndim(), 4, "dY must be (N, C, H, W)");
  CAFFE_ENFORCE_EQ(scale.size(), dY.dim32(1), "scale must have C elements");
  EnforceSadBody(sad_affine_channel_f32(dY.data<float>(), scale.data<float>(), nullptr, dX->mutable_data<float>(), dY.dim32(0), dY.dim32(1),
                                        (int64_t)dY.dim32(2) * dY.dim32(3), context_.cuda_stream()),
                 "sad_affine_channel_f32 (gradient)");
  return true;
}

template <typename T, class Context>
class UpsampleNearestOp final : public Operator<Context> {
 public:
  UpsampleNearestOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), scale_(OperatorBase::GetSingleArgument<int>("scale", 2)) {
    CAFFE_ENFORCE_GE(scale_, 1);  // the reference only DCHECKs (upsample_nearest_op.h:32); a scale < 1 divides by zero there
  }
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  int scale_;
};

template <>
bool UpsampleNearestOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  auto* Y = Output(0);
  int64_t outer;
  int H, W;
  OuterHW(X, &outer, &H, &W);
  vector<TIndex> out_shape;
  for (int i = 0; i < X.ndim(); ++i) out_shape.push_back(X.dim32(i));
  out_shape[X.ndim() - 1] *= scale_;
  out_shape[X.ndim() - 2] *= scale_;
  Y->Resize(out_shape);
  EnforceSadBody(sad_upsample_nearest_f32(X.data<float>(), Y->mutable_data<float>(), outer, H, W, scale_, context_.cuda_stream()),
                 "sad_upsample_nearest_f32");
  return true;
}

template <typename T, class Context>
class UpsampleNearestGradientOp final : public Operator<Context> {
 public:
  UpsampleNearestGradientOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), scale_(OperatorBase::GetSingleArgument<int>("scale", 2)) {
    CAFFE_ENFORCE_GE(scale_, 1);
  }
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  int scale_;
};

template <>
bool UpsampleNearestGradientOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  const auto& dY = Input(1);
  auto* dX = Output(0);
  dX->ResizeLike(X);
  int64_t outer;
  int H, W;
  OuterHW(X, &outer, &H, &W);
  CAFFE_ENFORCE_EQ(dY.size(), X.size() * scale_ * scale_, "dY must be the upsampled shape of X");
  EnforceSadBody(sad_upsample_nearest_grad_f32(dY.data<float>(), dX->mutable_data<float>(), outer, H, W, scale_, context_.cuda_stream()),
                 "sad_upsample_nearest_grad_f32");
  return true;
}

REGISTER_CPU_OPERATOR(AffineChannel, AffineChannelOp<float, CPUContext>);
REGISTER_CPU_OPERATOR(AffineChannelGradient, AffineChannelGradientOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(AffineChannel, AffineChannelOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(AffineChannelGradient, AffineChannelGradientOp<float, CUDAContext>);
REGISTER_CPU_OPERATOR(UpsampleNearest, UpsampleNearestOp<float, CPUContext>);
REGISTER_CPU_OPERATOR(UpsampleNearestGradient, UpsampleNearestGradientOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(UpsampleNearest, UpsampleNearestOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(UpsampleNearestGradient, UpsampleNearestGradientOp<float, CUDAContext>);

OPERATOR_SCHEMA(AffineChannel)
    .NumInputs(3)
    .NumOutputs(1)
    .AllowInplace({{0, 0}})
    .SetDoc("Per-channel affine transformation Y = X * scale[c] + bias[c]: batch normalisation frozen into its fixed form.")
    .Input(0, "X", "(N, C, H, W)")
    .Input(1, "scale", "(C)")
    .Input(2, "bias", "(C)")
    .Output(0, "Y", "(N, C, H, W)");
OPERATOR_SCHEMA(AffineChannelGradient)
    .NumInputs(2)
    .NumOutputs(1)
    .AllowInplace({{1, 0}})
    .Input(0, "scale", "(C)")
    .Input(1, "dY", "(N, C, H, W)")
    .Output(0, "dX", "(N, C, H, W)");
OPERATOR_SCHEMA(UpsampleNearest)
    .NumInputs(1)
    .NumOutputs(1)
    .SetDoc("Nearest-neighbour upsampling by an integer factor over the last two dimensions.")
    .Arg("scale", "(int) default 2; integer upsampling factor.")
    .Input(0, "X", "(N, C, H, W)")
    .Output(0, "Y", "(N, C, scale * H, scale * W)");
OPERATOR_SCHEMA(UpsampleNearestGradient)
    .NumInputs(2)
    .NumOutputs(1)
    .Input(0, "X", "forward input")
    .Input(1, "dY", "gradient of the forward output")
    .Output(0, "dX", "gradient of the forward input");

class GetAffineChannelGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    return SingleGradientDef("AffineChannelGradient", "", vector<string>{I(1), GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(AffineChannel, GetAffineChannelGradient);

class GetUpsampleNearestGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    return SingleGradientDef("UpsampleNearestGradient", "", vector<string>{I(0), GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(UpsampleNearest, GetUpsampleNearestGradient);

}  // namespace caffe2
