// Shim of caffe2/caffe2/core/context_gpu.h:54-134,136-272 — CUDAContext.
//
// Kept: one non-blocking stream per (thread, gpu, stream_id); work is enqueued on cuda_stream();
// FinishDeviceComputation() synchronises and surfaces asynchronous CUDA errors; New() allocates
// device memory.  Added: AdoptExternalStream() so a host framework (PyTorch here, the DAG
// executor's worker in real Caffe2) can make the ops enqueue on a stream it owns.
#ifndef SAD_SHIM_CONTEXT_GPU_H_
#define SAD_SHIM_CONTEXT_GPU_H_

#include "caffe2/core/common_gpu.h"
#include "caffe2/core/context.h"
#include "caffe2/core/tensor.h"

namespace caffe2 {

class CUDAContext final {
 public:
  explicit CUDAContext(const int gpu_id = -1);
  explicit CUDAContext(const DeviceOption& option);
  ~CUDAContext() {}

  void SwitchToDevice(int stream_id = 0) {
    set_stream_id(stream_id);
    CaffeCudaSetDevice(gpu_id_);
  }
  bool FinishDeviceComputation() {
    cudaStreamSynchronize(cuda_stream());
    cudaError_t error = cudaGetLastError();
    if (error == cudaSuccess) return true;
    fprintf(stderr, "Encountered CUDA error: %s\n", cudaGetErrorString(error));
    return false;
  }
  int cuda_gpu_id() const { return gpu_id_; }
  cudaStream_t cuda_stream() const { return cuda_stream(gpu_id_, stream_id_); }
  cudaStream_t cuda_stream() { return cuda_stream(gpu_id_, stream_id_); }
  static cudaStream_t cuda_stream(int gpu_id, int stream_id);
  // Route (this thread, gpu_id, stream_id) to a caller-owned stream; nullptr restores the default.
  static void AdoptExternalStream(int gpu_id, int stream_id, cudaStream_t stream);

  static std::pair<void*, std::function<void(void*)>> New(size_t nbytes);

  template <class SrcContext, class DstContext>
  void CopyBytes(size_t nbytes, const void* src, void* dst) {
    if (nbytes) CUDA_ENFORCE(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, cuda_stream()));
  }
  template <typename T, class SrcContext, class DstContext>
  void Copy(size_t n, const T* src, T* dst) {
    CopyBytes<SrcContext, DstContext>(n * sizeof(T), src, dst);
  }
  static bool HasAsyncPartDefault() { return true; }

 protected:
  void set_stream_id(int stream_id) { stream_id_ = stream_id; }
  int gpu_id_;
  int stream_id_ = 0;
};

typedef Tensor<CUDAContext> TensorCUDA;

}  // namespace caffe2
#endif
