// The optimiser operators of Detectron's parameter-update graph under the reference's names, arguments, schema arity and in-place
// permissions (detectron/lib/modeling/optimizer.py:95-130 emits, per parameter blob, Scale / WeightedSum / MomentumSGDUpdate):
//   MomentumSGDUpdate, MomentumSGD   caffe2/caffe2/sgd/momentum_sgd_op.h:53-127 (classes), momentum_sgd_op.cc:21-86 (schemas),
//                                    momentum_sgd_op_gpu.cu:23-77,138-139 (kernel + CUDA registration)
//   WeightedSum                      caffe2/caffe2/operators/utility_ops.h:333-378, utility_ops.cc (schema: NumInputs even, 1 output,
//                                    in-place with input 0), utility_ops.cu (CUDA registration)
// forwarding to sad_momentum_sgd_update_f32 / sad_weighted_sum_f32.  (`Scale` is registered in conv_ops.cc.)  CPU registrations keep
// CAFFE_NOT_IMPLEMENTED: there is no CPU path in this library.
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "sad_b200.h"

namespace caffe2 {

namespace {
void EnforceSadSgd(int rc, const char* what) { CAFFE_ENFORCE(rc == SAD_OK, what, " failed: ", sad_last_error()); }
}  // namespace

template <typename T, class Context>
class MomentumSGDOp final : public Operator<Context> {
 public:
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  MomentumSGDOp(const OperatorDef& operator_def, Workspace* ws)
      : Operator<Context>(operator_def, ws),
        momentum_(OperatorBase::GetSingleArgument<T>("momentum", 0.0)),
        nesterov_(OperatorBase::GetSingleArgument<int>("nesterov", 0)) {}
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  T momentum_{0.9};
  bool nesterov_;
  INPUT_TAGS(GRAD, MOMENTUM, LR);
  OUTPUT_TAGS(OUTPUT_GRAD, OUTPUT_MOMENTUM);
};

template <>
bool MomentumSGDOp<float, CUDAContext>::RunOnDevice() {
  CAFFE_ENFORCE(OperatorBase::InputIsType<Tensor<CUDAContext>>(GRAD));
  CAFFE_ENFORCE(OperatorBase::InputIsType<Tensor<CUDAContext>>(MOMENTUM));
  CAFFE_ENFORCE(Input(LR).size() == 1);
  CAFFE_ENFORCE(Input(GRAD).size() == Input(MOMENTUM).size());
  Output(OUTPUT_GRAD)->ResizeLike(Input(GRAD));
  Output(OUTPUT_MOMENTUM)->ResizeLike(Input(MOMENTUM));
  EnforceSadSgd(sad_momentum_sgd_update_f32(Input(GRAD).data<float>(), Input(MOMENTUM).data<float>(), Input(LR).data<float>(), nullptr,
                                            Output(OUTPUT_GRAD)->mutable_data<float>(), Output(OUTPUT_MOMENTUM)->mutable_data<float>(), nullptr,
                                            Input(GRAD).size(), momentum_, nesterov_ ? 1 : 0, context_.cuda_stream()),
                "sad_momentum_sgd_update_f32");
  return true;
}

template <typename T, class Context>
class MomentumSGDUpdateOp final : public Operator<Context> {
 public:
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  MomentumSGDUpdateOp(const OperatorDef& operator_def, Workspace* ws)
      : Operator<Context>(operator_def, ws),
        momentum_(OperatorBase::GetSingleArgument<T>("momentum", 0.0)),
        nesterov_(OperatorBase::GetSingleArgument<int>("nesterov", 0)) {}
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }

 protected:
  T momentum_{0.9};
  bool nesterov_;
  INPUT_TAGS(GRAD, MOMENTUM, LR, PARAM);
  OUTPUT_TAGS(OUTPUT_GRAD, OUTPUT_MOMENTUM, OUTPUT_PARAM);
};

template <>
bool MomentumSGDUpdateOp<float, CUDAContext>::RunOnDevice() {
  CAFFE_ENFORCE(OperatorBase::InputIsType<Tensor<CUDAContext>>(GRAD));
  CAFFE_ENFORCE(OperatorBase::InputIsType<Tensor<CUDAContext>>(MOMENTUM));
  CAFFE_ENFORCE_EQ(Input(LR).size(), 1);
  CAFFE_ENFORCE_EQ(Input(GRAD).size(), Input(MOMENTUM).size());
  Output(OUTPUT_GRAD)->ResizeLike(Input(GRAD));
  Output(OUTPUT_MOMENTUM)->ResizeLike(Input(MOMENTUM));
  // the reference writes through Output(OUTPUT_PARAM)->mutable_data without reading Input(PARAM): it only works in place
  // (optimizer.py:125-130 passes the same blob).  Out of place the parameter is read from the input here.
  auto* P = Output(OUTPUT_PARAM);
  P->ResizeLike(Input(PARAM));
  CAFFE_ENFORCE_EQ(Input(PARAM).size(), Input(GRAD).size());
  EnforceSadSgd(sad_momentum_sgd_update_f32(Input(GRAD).data<float>(), Input(MOMENTUM).data<float>(), Input(LR).data<float>(),
                                            Input(PARAM).data<float>(), Output(OUTPUT_GRAD)->mutable_data<float>(),
                                            Output(OUTPUT_MOMENTUM)->mutable_data<float>(), P->mutable_data<float>(), Input(GRAD).size(),
                                            momentum_, nesterov_ ? 1 : 0, context_.cuda_stream()),
                "sad_momentum_sgd_update_f32");
  return true;
}

template <class Context>
class WeightedSumOp final : public Operator<Context> {
 public:
  USE_OPERATOR_CONTEXT_FUNCTIONS;
  USE_SIMPLE_CTOR_DTOR(WeightedSumOp);
  bool RunOnDevice() override { CAFFE_NOT_IMPLEMENTED; }
};

template <>
bool WeightedSumOp<CUDAContext>::RunOnDevice() {
  CAFFE_ENFORCE_EQ(InputSize() % 2, 0);
  const auto& X0 = Input(0);
  CAFFE_ENFORCE_GT(X0.size(), 0);
  CAFFE_ENFORCE(X0.IsType<float>(), "WeightedSum: only float tensors are implemented");
  const int pairs = InputSize() / 2;
  CAFFE_ENFORCE(pairs <= SAD_MAX_INPUTS, "WeightedSum: at most ", SAD_MAX_INPUTS, " (tensor, weight) pairs");
  auto* output = Output(0);
  output->ResizeLike(X0);
  const float* xs[SAD_MAX_INPUTS];
  const float* ws[SAD_MAX_INPUTS];
  for (int k = 0; k < pairs; ++k) {
    const auto& X = Input(2 * k);
    const auto& w = Input(2 * k + 1);
    if (k > 0 && &X == output) return false;   // utility_ops.h:357-364: in-place only with input 0
    CAFFE_ENFORCE_EQ(X.size(), X0.size());
    CAFFE_ENFORCE_EQ(w.size(), 1);
    xs[k] = X.data<float>();
    ws[k] = w.data<float>();
  }
  EnforceSadSgd(sad_weighted_sum_f32(xs, ws, pairs, output->mutable_data<float>(), X0.size(), context_.cuda_stream()), "sad_weighted_sum_f32");
  return true;
}

REGISTER_CPU_OPERATOR(MomentumSGD, MomentumSGDOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(MomentumSGD, MomentumSGDOp<float, CUDAContext>);
OPERATOR_SCHEMA(MomentumSGD)
    .NumInputs(3)
    .NumOutputs(2)
    .AllowInplace({{0, 0}, {1, 1}})
    .SetDoc("(grad, momentum, lr) -> (adjusted grad, momentum): adjusted = lr * grad + momentum * m; nesterov: the :44-51 form.")
    .Arg("momentum", "(float, default 0) momentum coefficient")
    .Arg("nesterov", "(int, default 0) Nesterov form");
SHOULD_NOT_DO_GRADIENT(MomentumSGD);

REGISTER_CPU_OPERATOR(MomentumSGDUpdate, MomentumSGDUpdateOp<float, CPUContext>);
REGISTER_CUDA_OPERATOR(MomentumSGDUpdate, MomentumSGDUpdateOp<float, CUDAContext>);
OPERATOR_SCHEMA(MomentumSGDUpdate)
    .NumInputs(4)
    .NumOutputs(3)
    .AllowInplace({{0, 0}, {1, 1}, {3, 2}})
    .SetDoc("(grad, momentum, lr, param) -> (adjusted grad, momentum, param - adjusted grad).")
    .Arg("momentum", "(float, default 0) momentum coefficient")
    .Arg("nesterov", "(int, default 0) Nesterov form");
SHOULD_NOT_DO_GRADIENT(MomentumSGDUpdate);

REGISTER_CPU_OPERATOR(WeightedSum, WeightedSumOp<CPUContext>);
REGISTER_CUDA_OPERATOR(WeightedSum, WeightedSumOp<CUDAContext>);
OPERATOR_SCHEMA(WeightedSum)
    .NumInputs([](int n) { return (n > 0 && n % 2 == 0); })
    .NumOutputs(1)
    .AllowInplace({{0, 0}})
    .SetDoc("Element-wise weighted sum of (tensor, scalar weight) pairs: X_0 * w_0 + X_1 * w_1 + ...; in place only with X_0.")
    .Input(0, "data_0", "first tensor")
    .Input(1, "weight_0", "its scalar weight (float, 1 element)")
    .Output(0, "output", "the weighted sum");

}  // namespace caffe2
