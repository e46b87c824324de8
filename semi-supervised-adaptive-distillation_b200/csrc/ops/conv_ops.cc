// Operator classes for the RetinaNet head's convolutions under the reference's operator names, so the head
// sub-graph Detectron emits (detectron/lib/modeling/retinanet_heads.py:63-245 through DetectionModelHelper.Conv /
// ConvShared, detector.py:449-482, engine CUDNN via detector.py:56-60) instantiates from this library:
//
//   Conv          (+ engine CUDNN)   reference CudnnConvOp          caffe2/caffe2/operators/conv_op_cudnn.cc:293-643,1130-1131
//   ConvGradient  (+ engine CUDNN)   reference CudnnConvGradientOp  caffe2/caffe2/operators/conv_op_cudnn.cc:645-1100
//                                    gradient maker                 caffe2/caffe2/operators/conv_gradient_op.cc:35-77
//   Relu / ReluGradient              reference relu_op.cu:22-62, relu_op.cc:98-108
//
// Scope: the head's shape class only — 2-D, NCHW, kernel 3, stride 1, pad 1, dilation 1, group 1, float.
// Anything else throws (there is no cuDNN / CPU fallback behind these names in this library).
// Arguments are parsed like ConvPoolOpBase (conv_pool_op_base.h:53-123): kernel | kernel_h+kernel_w, stride(_h/_w),
// pad | pad_t/l/b/r, dilation, group, order; ConvGradient also reads no_bias (conv_op_cudnn.cc:655-663).
// Each operator instance keeps its channels-last staging buffers, packed weights and reduction scratch across runs
// (the role of the cuDNN workspace in the reference, cudnn_wrappers.h:49-93).
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "sad_b200.h"

namespace caffe2 {

namespace {

void EnforceSadConv(int rc, const char* what) { CAFFE_ENFORCE(rc == SAD_OK, what, " failed: ", sad_last_error()); }

// ConvPoolOpBase argument parsing restricted to what the tensor-core kernels implement
struct HeadConvArgs {
  explicit HeadConvArgs(OperatorBase* op) {
    auto pair = [&](const char* both, const char* h, const char* w, int dflt, int* oh, int* ow) {
      if (op->HasArgument(both)) {
        *oh = *ow = op->GetSingleArgument<int>(both, dflt);
      } else if (op->HasArgument(h) && op->HasArgument(w)) {
        *oh = op->GetSingleArgument<int>(h, dflt);
        *ow = op->GetSingleArgument<int>(w, dflt);
      } else {
        *oh = *ow = dflt;
      }
    };
    int kh, kw, sh, sw, dh, dw;
    pair("kernel", "kernel_h", "kernel_w", 0, &kh, &kw);
    pair("stride", "stride_h", "stride_w", 1, &sh, &sw);
    pair("dilation", "dilation_h", "dilation_w", 1, &dh, &dw);
    int pt, pl, pb, pr;
    if (op->HasArgument("pad")) {
      pt = pl = pb = pr = op->GetSingleArgument<int>("pad", 0);
    } else {
      pt = op->GetSingleArgument<int>("pad_t", 0);
      pl = op->GetSingleArgument<int>("pad_l", 0);
      pb = op->GetSingleArgument<int>("pad_b", 0);
      pr = op->GetSingleArgument<int>("pad_r", 0);
    }
    const string order = op->GetSingleArgument<string>("order", "NCHW");
    const int group = op->GetSingleArgument<int>("group", 1);
    CAFFE_ENFORCE(order == "NCHW", "B200 head convolution: only order NCHW is implemented, got ", order);
    CAFFE_ENFORCE(kh == 3 && kw == 3, "B200 head convolution: only 3x3 kernels are implemented, got ", kh, "x", kw);
    CAFFE_ENFORCE(sh == 1 && sw == 1, "B200 head convolution: only stride 1 is implemented");
    CAFFE_ENFORCE(pt == 1 && pl == 1 && pb == 1 && pr == 1, "B200 head convolution: only pad 1 is implemented");
    CAFFE_ENFORCE(dh == 1 && dw == 1, "B200 head convolution: dilation is not implemented");
    CAFFE_ENFORCE(group == 1, "B200 head convolution: groups are not implemented");
    // conv_op_cudnn.cc:77-86,494-498: the reference only lets cuDNN use reduced-precision tensor-core math when the operator
    // carries enable_tensor_core = 1 (default 0: plain fp32).  Here every mode runs on the tensor cores; the flag selects the
    // arithmetic: 0 (default) = 3xTF32, fp32-accurate (matches the reference's fp32 convolution to ~1e-5), 1 = single-pass tf32
    // (10-bit operand mantissas, 3x faster).
    tensor_core_math = op->GetSingleArgument<int>("enable_tensor_core", 0) != 0;
  }
  bool tensor_core_math = false;
};

// grows a member tensor to at least `floats` elements and returns its storage (a default-constructed tensor has
// no shape yet: ndim() == 0 and size() would throw, tensor.h:82)
float* Ensure(Tensor<CUDAContext>* t, size_t floats) {
  if (t->ndim() == 0 || (size_t)t->size() < floats) t->Resize((TIndex)(floats ? floats : 1));
  return t->mutable_data<float>();
}

void CheckConvInputs(const Tensor<CUDAContext>& X, const Tensor<CUDAContext>& W) {
  CAFFE_ENFORCE_EQ(X.ndim(), 4, "Conv input must be (N, C, H, W)");
  CAFFE_ENFORCE_EQ(W.ndim(), 4, "Conv filter must be (M, C, 3, 3)");
  CAFFE_ENFORCE(W.dim32(2) == 3 && W.dim32(3) == 3, "Conv filter must be (M, C, 3, 3)");
  CAFFE_ENFORCE_EQ(X.dim32(1), W.dim32(1), "Conv: input channels ", X.dim32(1), " do not match the filter's ", W.dim32(1));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
template <typename T, class Context>
class HeadConvOp final : public Operator<Context> {
 public:
  HeadConvOp(const OperatorDef& def, Workspace* ws) : Operator<Context>(def, ws), args_(this) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  HeadConvArgs args_;
  Tensor<CUDAContext> x_nhwc_, packed_;
};

template <>
bool HeadConvOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  const auto& W = Input(1);
  CheckConvInputs(X, W);
  const int N = X.dim32(0), C = X.dim32(1), H = X.dim32(2), Wd = X.dim32(3), M = W.dim32(0);
  const float* bias = nullptr;
  if (InputSize() == 3) {
    const auto& b = Input(2);
    CAFFE_ENFORCE_EQ(b.size(), (TIndex)M, "Conv bias must have one element per output channel");
    bias = b.data<float>();
  }
  auto* Y = Output(0);
  Y->Resize(vector<TIndex>{N, M, H, Wd});  // stride 1, pad 1, kernel 3: same spatial size (conv_pool_op_base.h:245-270)
  float* y = Y->mutable_data<float>();
  if (Y->size() == 0) return true;
  void* st = context_.cuda_stream();
  const bool x3 = !args_.tensor_core_math;
  const size_t pixels = (size_t)N * H * Wd;
  float* xt = Ensure(&x_nhwc_, x3 ? pixels * 2 * sad_conv3x3_split_channels(C) : (size_t)X.size());
  float* pk = Ensure(&packed_, (x3 ? sad_conv3x3_packed_bytes_f32x3(C, M, 0) : sad_conv3x3_packed_bytes(C, M)) / sizeof(float));
  sad_layout_level ll{X.data<float>(), xt, N, H, Wd};
  sad_pack_item pi{W.data<float>(), pk, C, M, 0};
  sad_conv_level cl{};
  cl.x_nhwc = xt;
  cl.y_nchw = y;
  cl.N = N;
  cl.H = H;
  cl.W = Wd;
  if (x3) {
    EnforceSadConv(sad_nchw_to_nhwc_f32x3(&ll, 1, C, st), "sad_nchw_to_nhwc_f32x3");
    EnforceSadConv(sad_conv3x3_pack_weights_multi_f32x3(&pi, 1, st), "sad_conv3x3_pack_weights_multi_f32x3");
    EnforceSadConv(sad_conv3x3_fwd_f32x3(&cl, 1, pk, bias, C, M, 0, st), "sad_conv3x3_fwd_f32x3");
  } else {
    EnforceSadConv(sad_nchw_to_nhwc_f32(&ll, 1, C, st), "sad_nchw_to_nhwc_f32");
    EnforceSadConv(sad_conv3x3_pack_weights_f32(W.data<float>(), C, M, 0, pk, st), "sad_conv3x3_pack_weights_f32");
    EnforceSadConv(sad_conv3x3_fwd_f32(&cl, 1, pk, bias, C, M, 0, st), "sad_conv3x3_fwd_f32");
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Inputs X, filter, dY; outputs dfilter, [dbias unless no_bias], [dX]   (conv_op_cudnn.cc:645-700, INPUT_TAGS/OUTPUT_TAGS)
template <typename T, class Context>
class HeadConvGradientOp final : public Operator<Context> {
 public:
  HeadConvGradientOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), args_(this), no_bias_(OperatorBase::GetSingleArgument<int>("no_bias", 0)) {
    CAFFE_ENFORCE(!(no_bias_ && OutputSize() == 3), "If bias is not present, you should not have 3 grad output.");
  }
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  HeadConvArgs args_;
  bool no_bias_;
  Tensor<CUDAContext> x_nhwc_, dy_nhwc_, packed_, scratch_;
};

template <>
bool HeadConvGradientOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  const auto& W = Input(1);
  const auto& dY = Input(2);
  CheckConvInputs(X, W);
  const int N = X.dim32(0), C = X.dim32(1), H = X.dim32(2), Wd = X.dim32(3), M = W.dim32(0);
  CAFFE_ENFORCE(dY.ndim() == 4 && dY.dim32(0) == N && dY.dim32(1) == M && dY.dim32(2) == H && dY.dim32(3) == Wd,
                "ConvGradient: dY must be (N, M, H, W) of the forward output");
  auto* dW = Output(0);
  dW->ResizeLike(W);
  float* dw = dW->mutable_data<float>();
  float* db = nullptr;
  if (!no_bias_) {
    auto* dB = Output(1);
    dB->Resize(vector<TIndex>{M});
    db = dB->mutable_data<float>();
  }
  const bool want_dx = OutputSize() == 3 || (no_bias_ && OutputSize() == 2);
  void* st = context_.cuda_stream();
  const bool x3 = !args_.tensor_core_math;
  const size_t pixels = (size_t)N * H * Wd;
  float* xt = Ensure(&x_nhwc_, x3 ? pixels * 2 * sad_conv3x3_split_channels(C) : (size_t)X.size());
  float* dyt = Ensure(&dy_nhwc_, x3 ? pixels * 2 * sad_conv3x3_split_channels(M) : (size_t)dY.size());
  if (X.size()) {
    sad_layout_level lx{X.data<float>(), xt, N, H, Wd};
    sad_layout_level ld{dY.data<float>(), dyt, N, H, Wd};
    if (x3) {
      EnforceSadConv(sad_nchw_to_nhwc_f32x3(&lx, 1, C, st), "sad_nchw_to_nhwc_f32x3(X)");
      EnforceSadConv(sad_nchw_to_nhwc_f32x3(&ld, 1, M, st), "sad_nchw_to_nhwc_f32x3(dY)");
    } else {
      EnforceSadConv(sad_nchw_to_nhwc_f32(&lx, 1, C, st), "sad_nchw_to_nhwc_f32(X)");
      EnforceSadConv(sad_nchw_to_nhwc_f32(&ld, 1, M, st), "sad_nchw_to_nhwc_f32(dY)");
    }
  }
  sad_wgrad_level wl{xt, dyt, N, H, Wd};
  const size_t wsb = sad_conv3x3_wgrad_workspace_bytes(&wl, 1, C, M);
  // 256-byte aligned scratch inside a float tensor
  float* raw = Ensure(&scratch_, wsb / sizeof(float) + 64);
  void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(raw) + 255) & ~(uintptr_t)255);
  if (x3) EnforceSadConv(sad_conv3x3_wgrad_f32x3(&wl, 1, C, M, dw, db, 0, ws, wsb, st), "sad_conv3x3_wgrad_f32x3");
  else EnforceSadConv(sad_conv3x3_wgrad_f32(&wl, 1, C, M, dw, db, 0, ws, wsb, st), "sad_conv3x3_wgrad_f32");
  if (want_dx) {
    auto* dX = Output(no_bias_ ? 1 : 2);
    dX->ResizeLike(X);
    float* dx = dX->mutable_data<float>();
    if (X.size()) {
      float* pk = Ensure(&packed_, (x3 ? sad_conv3x3_packed_bytes_f32x3(C, M, 1) : sad_conv3x3_packed_bytes(C, M)) / sizeof(float));
      sad_conv_level cl{};
      cl.x_nhwc = dyt;
      cl.y_nchw = dx;
      cl.N = N;
      cl.H = H;
      cl.W = Wd;
      if (x3) {
        sad_pack_item pi{W.data<float>(), pk, C, M, 1};
        EnforceSadConv(sad_conv3x3_pack_weights_multi_f32x3(&pi, 1, st), "sad_conv3x3_pack_weights_multi_f32x3");
        EnforceSadConv(sad_conv3x3_fwd_f32x3(&cl, 1, pk, nullptr, M, C, 0, st), "sad_conv3x3_fwd_f32x3(dgrad)");
      } else {
        EnforceSadConv(sad_conv3x3_pack_weights_f32(W.data<float>(), C, M, 1, pk, st), "sad_conv3x3_pack_weights_f32");
        EnforceSadConv(sad_conv3x3_fwd_f32(&cl, 1, pk, nullptr, M, C, 0, st), "sad_conv3x3_fwd_f32(dgrad)");
      }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
template <typename T, class Context>
class HeadReluOp final : public Operator<Context> {
 public:
  USE_SIMPLE_CTOR_DTOR(HeadReluOp);
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }
};
template <>
bool HeadReluOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  auto* Y = Output(0);
  Y->ResizeLike(X);
  EnforceSadConv(sad_relu_f32(X.data<float>(), Y->mutable_data<float>(), X.size(), context_.cuda_stream()), "sad_relu_f32");
  return true;
}

// Sigmoid (sigmoid_op.cu:24-29) for the teacher graph's class probabilities (retinanet_heads.py:153-163)
template <typename T, class Context>
class HeadSigmoidOp final : public Operator<Context> {
 public:
  USE_SIMPLE_CTOR_DTOR(HeadSigmoidOp);
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }
};
template <>
bool HeadSigmoidOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  auto* Y = Output(0);
  Y->ResizeLike(X);
  EnforceSadConv(sad_sigmoid_f32(X.data<float>(), Y->mutable_data<float>(), X.size(), context_.cuda_stream()), "sad_sigmoid_f32");
  return true;
}

template <typename T, class Context>
class HeadReluGradientOp final : public Operator<Context> {
 public:
  USE_SIMPLE_CTOR_DTOR(HeadReluGradientOp);
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }
};
template <>
bool HeadReluGradientOp<float, CUDAContext>::RunOnDevice() {
  const auto& Y = Input(0);
  const auto& dY = Input(1);
  auto* dX = Output(0);
  CAFFE_ENFORCE_EQ(dY.size(), Y.size());  // relu_op.cu:53
  dX->ResizeLike(Y);
  EnforceSadConv(sad_relu_grad_f32(Y.data<float>(), dY.data<float>(), dX->mutable_data<float>(), Y.size(), context_.cuda_stream()),
                 "sad_relu_grad_f32");
  return true;
}

// Scale (caffe2/caffe2/operators/scale_op.h:31-50): the operator _CorrectMomentum creates per momentum blob (detector.py:643-647)
template <typename T, class Context>
class HeadScaleOp final : public Operator<Context> {
 public:
  HeadScaleOp(const OperatorDef& def, Workspace* ws)
      : Operator<Context>(def, ws), scale_(OperatorBase::GetSingleArgument<float>("scale", 1.0f)) {}
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  float scale_;
};
template <>
bool HeadScaleOp<float, CUDAContext>::RunOnDevice() {
  const auto& X = Input(0);
  auto* Y = Output(0);
  Y->ResizeLike(X);
  EnforceSadConv(sad_scale_f32(X.data<float>(), Y->mutable_data<float>(), X.size(), scale_, context_.cuda_stream()), "sad_scale_f32");
  return true;
}

// ---------------------------------------------------------------------------------------------
REGISTER_CUDA_OPERATOR(Scale, HeadScaleOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(Conv, HeadConvOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(ConvGradient, HeadConvGradientOp<float, CUDAContext>);
REGISTER_CUDNN_OPERATOR(Conv, HeadConvOp<float, CUDAContext>);                  // conv_op_cudnn.cc:1130-1131
REGISTER_CUDNN_OPERATOR(ConvGradient, HeadConvGradientOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(Relu, HeadReluOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(ReluGradient, HeadReluGradientOp<float, CUDAContext>);
REGISTER_CUDA_OPERATOR(Sigmoid, HeadSigmoidOp<float, CUDAContext>);

OPERATOR_SCHEMA(Conv)
    .NumInputs(2, 3)
    .NumOutputs(1)
    .SetDoc("2-D convolution (cross-correlation) Y = conv(X, filter) + bias, NCHW.  This library implements the RetinaNet "
            "head's shape class on the Blackwell tensor cores: kernel 3, stride 1, pad 1, group 1, float.")
    .Input(0, "X", "(N, C, H, W)")
    .Input(1, "filter", "(M, C, 3, 3)")
    .Input(2, "bias", "(M), optional")
    .Output(0, "Y", "(N, M, H, W)");
OPERATOR_SCHEMA(ConvGradient).NumInputs(2, 3).NumOutputs(1, 3);
OPERATOR_SCHEMA(Relu).NumInputs(1).NumOutputs(1).AllowInplace({{0, 0}}).IdenticalTypeAndShape();
OPERATOR_SCHEMA(ReluGradient).NumInputs(2).NumOutputs(1).AllowInplace({{1, 0}});
OPERATOR_SCHEMA(Scale).NumInputs(1).NumOutputs(1).AllowInplace({{0, 0}}).IdenticalTypeAndShape().Arg("scale", "(float, default 1.0)");  // scale_op.cc
OPERATOR_SCHEMA(Sigmoid).NumInputs(1).NumOutputs(1).AllowInplace({{0, 0}}).IdenticalTypeAndShape();  // sigmoid_op.cc

// caffe2/caffe2/operators/conv_gradient_op.cc:35-77
class GetConvGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    CAFFE_ENFORCE(def_.input_size() == 3 || def_.input_size() == 2);
    const bool compute_dX = !ArgumentHelper::GetSingleArgument<OperatorDef, bool>(def_, "no_gradient_to_input", false);
    if (def_.input_size() == 3) {
      if (compute_dX)
        return SingleGradientDef(def_.type() + "Gradient", "", vector<string>{I(0), I(1), GO(0)}, vector<string>{GI(1), GI(2), GI(0)});
      return SingleGradientDef(def_.type() + "Gradient", "", vector<string>{I(0), I(1), GO(0)}, vector<string>{GI(1), GI(2)});
    }
    Argument no_bias;
    no_bias.set_name("no_bias");
    no_bias.set_i(1);
    if (compute_dX)
      return SingleGradientDef(def_.type() + "Gradient", "", vector<string>{I(0), I(1), GO(0)}, vector<string>{GI(1), GI(0)},
                               vector<Argument>{no_bias});
    return SingleGradientDef(def_.type() + "Gradient", "", vector<string>{I(0), I(1), GO(0)}, vector<string>{GI(1)},
                             vector<Argument>{no_bias});
  }
};
REGISTER_GRADIENT(Conv, GetConvGradient);

// caffe2/caffe2/operators/relu_op.cc:98-108
class GetReluGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    return SingleGradientDef(def_.type() + "Gradient", "", vector<string>{O(0), GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(Relu, GetReluGradient);

// caffe2/caffe2/operators/scale_op.cc:33-41: the gradient of Scale is Scale with the copied argument
class GetScaleGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override {
    return SingleGradientDef("Scale", "", vector<string>{GO(0)}, vector<string>{GI(0)});
  }
};
REGISTER_GRADIENT(Scale, GetScaleGradient);

}  // namespace caffe2
