// TEST INFRASTRUCTURE — part of the GPU oracle (oracle/_ref), never linked into the product.
//
// Restatement of the five generic primitives the reference ops call, keeping the reference's
// launch shapes and arithmetic order so the unmodified reference .cu files behave here as they do
// inside Caffe2:
//   Set    caffe2/caffe2/utils/math_gpu.cu:835-854     grid-stride fill
//   Powx   caffe2/caffe2/utils/math_gpu.cu:1257-1262,1279-1291   y = powf(x, b)
//   Sum    caffe2/caffe2/utils/math_gpu.cu:1021-1058,1101-1114   no scratch tensor is passed by the
//          ops, so the <<<1,128>>> single-block kernel runs: 128 strided float partials,
//          lane j<32 adds partials j+32, j+64, j+96, then lane 0 adds the 32 in order
//   Add    caffe2/caffe2/utils/math_gpu.cu:85-103      y = a + b
//   Scale  caffe2/caffe2/utils/math_gpu.cu:1241-1247,1293-1302   y = x * alpha
#include "caffe2/core/context_gpu.h"
#include "caffe2/core/operator.h"
#include "caffe2/utils/math.h"

namespace caffe2 {
namespace math {
namespace {

__global__ void fill_kernel(const int n, const float v, float* y) {
  CUDA_1D_KERNEL_LOOP(i, n) { y[i] = v; }
}
__global__ void pow_kernel(const int n, const float* x, const float e, float* y) {
  CUDA_1D_KERNEL_LOOP(i, n) { y[i] = powf(x[i], e); }
}
__global__ void add_kernel(const int n, const float* a, const float* b, float* y) {
  CUDA_1D_KERNEL_LOOP(i, n) {
    float r = a[i] + b[i];
    y[i] = r;
  }
}
__global__ void scale_kernel(const int n, const float alpha, const float* x, float* y) {
  CUDA_1D_KERNEL_LOOP(i, n) { y[i] = x[i] * alpha; }
}
constexpr int kSumLanes = 128;
__global__ void single_block_sum_kernel(const int n, const float* x, float* y) {
  __shared__ float part[kSumLanes];
  const int lane = threadIdx.x;
  part[lane] = 0;
  for (int i = lane; i < n; i += kSumLanes) part[lane] += x[i];
  __syncthreads();
  if (lane < 32) part[lane] += part[lane + 32] + part[lane + 64] + part[lane + 96];
  __syncthreads();
  if (lane == 0) {
    float total = 0;
    for (int j = 0; j < 32; ++j) total += part[j];
    *y = total;
  }
}

}  // namespace

template <>
void Set<float, CUDAContext>(const size_t N, const float alpha, float* Y, CUDAContext* context) {
  fill_kernel<<<CAFFE_GET_BLOCKS((int)N), CAFFE_CUDA_NUM_THREADS, 0, context->cuda_stream()>>>((int)N, alpha, Y);
}
template <>
void Powx<float, CUDAContext>(const int N, const float* a, const float b, float* y, CUDAContext* context) {
  pow_kernel<<<CAFFE_GET_BLOCKS(N), CAFFE_CUDA_NUM_THREADS, 0, context->cuda_stream()>>>(N, a, b, y);
}
template <>
void Sum<float, CUDAContext>(const int N, const float* x, float* y, CUDAContext* context,
                             Tensor<CUDAContext>* scratch_ptr) {
  // the ops on this path never pass scratch (pow_sum_op.cu:38, sigmoid_adaptive_distillation_loss_op.cu:135-136)
  CAFFE_ENFORCE(scratch_ptr == nullptr, "GPU oracle: only the scratch-less Sum path is restated");
  single_block_sum_kernel<<<1, kSumLanes, 0, context->cuda_stream()>>>(N, x, y);
}
template <>
void Add<float, CUDAContext>(const int N, const float* a, const float* b, float* y, CUDAContext* context) {
  add_kernel<<<CAFFE_GET_BLOCKS(N), CAFFE_CUDA_NUM_THREADS, 0, context->cuda_stream()>>>(N, a, b, y);
}
template <>
void Scale<float, CUDAContext>(const int N, const float alpha, const float* x, float* y, CUDAContext* context) {
  scale_kernel<<<CAFFE_GET_BLOCKS(N), CAFFE_CUDA_NUM_THREADS, 0, context->cuda_stream()>>>(N, alpha, x, y);
}

}  // namespace math

// ConstantFill (subset: float value, output shaped like input 0) so the NetDef text Detectron emits
// for the loss-gradient seeds (detectron/lib/utils/blob.py:166-172; caffe2/caffe2/operators/filler_op.h)
// also runs through the GPU oracle's executor.  Uses the restated Set primitive above.
class RefConstantFillOp final : public Operator<CUDAContext> {
 public:
  RefConstantFillOp(const OperatorDef& def, Workspace* ws)
      : Operator<CUDAContext>(def, ws), value_(OperatorBase::GetSingleArgument<float>("value", 0.f)) {}
  bool RunOnDevice() override {
    auto* out = Output(0);
    CAFFE_ENFORCE(InputSize() == 1, "GPU oracle ConstantFill: only the shape-from-input form is restated");
    out->ResizeLike(Input(0));
    math::Set<float, CUDAContext>(out->size(), value_, out->mutable_data<float>(), &context_);
    return true;
  }

 private:
  float value_;
};
REGISTER_CUDA_OPERATOR(ConstantFill, RefConstantFillOp);
OPERATOR_SCHEMA(ConstantFill).NumInputs(0, 1).NumOutputs(1);

}  // namespace caffe2
