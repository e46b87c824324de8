// Shim of caffe2/caffe2/core/common_gpu.h:246-248,274-288 — launch macros of the generic kernels
// (used by the unmodified reference .cu files when they are built as the GPU oracle) and the
// CUDA error-check convention.
#ifndef SAD_SHIM_COMMON_GPU_H_
#define SAD_SHIM_COMMON_GPU_H_

#include <cuda_runtime.h>

#include <algorithm>

#include "caffe2/core/common.h"
#include "caffe2/core/logging.h"

namespace caffe2 {

#define CUDA_ENFORCE(condition, ...)                                                          \
  do {                                                                                        \
    cudaError_t error = condition;                                                            \
    CAFFE_ENFORCE_EQ(error, cudaSuccess, "Error at: ", __FILE__, ":", __LINE__, ": ",         \
                     cudaGetErrorString(error), ##__VA_ARGS__);                               \
  } while (0)
#define CUDA_CHECK(condition) CUDA_ENFORCE(condition)

#define CUDA_1D_KERNEL_LOOP(i, n) \
  for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += blockDim.x * gridDim.x)

constexpr int CAFFE_CUDA_NUM_THREADS = 512;
constexpr int CAFFE_MAXIMUM_NUM_BLOCKS = 4096;
inline int CAFFE_GET_BLOCKS(const int N) {
  return std::min((N + CAFFE_CUDA_NUM_THREADS - 1) / CAFFE_CUDA_NUM_THREADS, CAFFE_MAXIMUM_NUM_BLOCKS);
}

int NumCudaDevices();
int CaffeCudaGetDevice();
void CaffeCudaSetDevice(const int id);

class DeviceGuard {
 public:
  explicit DeviceGuard(int newDevice) : previous_(CaffeCudaGetDevice()) {
    if (previous_ != newDevice) CaffeCudaSetDevice(newDevice);
  }
  ~DeviceGuard() noexcept { cudaSetDevice(previous_); }

 private:
  int previous_;
};

}  // namespace caffe2
#endif
