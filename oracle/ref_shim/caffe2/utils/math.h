// TEST INFRASTRUCTURE (oracle/_ref build only).  Extends the product shim's caffe2/utils/math.h with the
// declarations the reference's CPU convolution (caffe2/caffe2/operators/conv_op_impl.h:31-180, 346-700)
// calls: caffe2/caffe2/utils/math.h:216-229 (Gemm), :272-283 (Gemv), :360-389 (Im2colNd / Col2imNd),
// :391-427 (Im2col / Col2im), plus StorageOrder (caffe2/caffe2/core/types.h:32-47), the CBLAS transpose tags
// (caffe2/caffe2/utils/cblas.h) and the TensorShape / cost helpers conv_pool_op_base.h:375-520 names.
// Definitions: oracle/ref_math_cpu.cc.
#ifndef SAD_REF_SHIM_MATH_H_
#define SAD_REF_SHIM_MATH_H_

#include_next "caffe2/utils/math.h"

#include <cmath>
#include <iostream>

#include "caffe2/core/context.h"
#include "caffe2/core/flags.h"
#include "caffe2/core/logging.h"
#include "caffe2/proto/caffe2.pb.h"

extern "C" {
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
}

// glog-style LOG(severity) << ...: messages go to stderr, FATAL aborts (logging_is_not_google_glog.h:41-70)
#ifndef LOG
namespace caffe2 {
struct RefShimLogLine {
  explicit RefShimLogLine(bool fatal) : fatal_(fatal) {}
  ~RefShimLogLine() { std::cerr << std::endl; if (fatal_) abort(); }
  template <typename T> RefShimLogLine& operator<<(const T& v) { std::cerr << v; return *this; }
  bool fatal_;
};
}
#define SAD_REF_LOG_INFO false
#define SAD_REF_LOG_WARNING false
#define SAD_REF_LOG_ERROR false
#define SAD_REF_LOG_FATAL true
#define LOG(severity) ::caffe2::RefShimLogLine(SAD_REF_LOG_##severity)
#define VLOG(n) if (false) ::caffe2::RefShimLogLine(false)
#endif

namespace caffe2 {

enum StorageOrder { UNKNOWN = 0, NHWC = 1, NCHW = 2 };
inline StorageOrder StringToStorageOrder(const string& str) {
  if (str == "NHWC" || str == "nhwc") return StorageOrder::NHWC;
  if (str == "NCHW" || str == "nchw") return StorageOrder::NCHW;
  return StorageOrder::UNKNOWN;
}

class DefaultEngine {};

namespace math {

template <typename T, class Context, class Engine = DefaultEngine>
void Gemm(const CBLAS_TRANSPOSE TransA, const CBLAS_TRANSPOSE TransB, const int M, const int N, const int K,
          const float alpha, const T* A, const T* B, const float beta, T* C, Context* context,
          int math_type = 1);
template <typename T, class Context, class Engine = DefaultEngine>
void Gemv(const CBLAS_TRANSPOSE TransA, const int M, const int N, const float alpha, const T* A, const T* x,
          const float beta, T* y, Context* context, int math_type = 1);
template <typename T, class Context, int order>
void Im2colNd(const T* data_img, const int* im_shape, const int* col_shape, const int img_size, const int col_size,
              const int* kernel_shape, const int* stride, const int* dilation, const int* pad, const int N,
              T* data_col, Context* context, bool accumulate_output = false);
template <typename T, class Context, int order>
void Col2imNd(const T* data_col, const int* img_shape, const int* col_shape, const int img_size, const int col_size,
              const int* kernel_shape, const int* stride, const int* dilation, const int* pad, const int N,
              T* data_img, Context* context);
template <typename T, class Context, int order>
void Im2col(const T* data_im, const int channels, const int height, const int width, const int kernel_h,
            const int kernel_w, const int dilation_h, const int dilation_w, const int pad_t, const int pad_l,
            const int pad_b, const int pad_r, const int stride_h, const int stride_w, T* data_col, Context* context);
template <typename T, class Context, int order>
void Col2im(const T* data_col, const int channels, const int height, const int width, const int kernel_h,
            const int kernel_w, const int dilation_h, const int dilation_w, const int pad_t, const int pad_l,
            const int pad_b, const int pad_r, const int stride_h, const int stride_w, T* data_im, Context* context);

}  // namespace math
}  // namespace caffe2
#endif
