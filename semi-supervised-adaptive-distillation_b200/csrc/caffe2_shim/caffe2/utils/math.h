// Shim of caffe2/caffe2/utils/math.h:102,286,315-316,334 — declarations of the five generic
// primitives the reference ops call.  The product ops do NOT use them (their work is fused into
// the sm_100a kernels behind include/sad_b200.h); definitions restating
// caffe2/caffe2/utils/math_gpu.cu exist only in oracle/ref_math.cu for the GPU oracle build.
#ifndef SAD_SHIM_MATH_H_
#define SAD_SHIM_MATH_H_

#include "caffe2/core/common.h"
#include "caffe2/core/tensor.h"

namespace caffe2 {
namespace math {

template <typename T, class Context>
void Set(const size_t N, const T alpha, T* X, Context* context);
template <typename T, class Context>
void Powx(const int N, const T* a, const T b, T* y, Context* context);
template <typename T, class Context>
void Sum(const int N, const T* x, T* y, Context* context, Tensor<Context>* scratch_ptr = nullptr);
template <typename T, class Context>
void Add(const int N, const T* a, const T* b, T* y, Context* context);
template <typename T, class Context>
void Scale(const int N, const float alpha, const T* x, T* y, Context* context);

}  // namespace math
}  // namespace caffe2
#endif
