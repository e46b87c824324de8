// Shim of caffe2/caffe2/core/context.h — CPUContext (host allocation + copies).
#ifndef SAD_SHIM_CONTEXT_H_
#define SAD_SHIM_CONTEXT_H_

#include <cstdlib>
#include <cstring>

#include "caffe2/core/logging.h"
#include "caffe2/core/typeid.h"

namespace caffe2 {

class CPUContext final {
 public:
  CPUContext() {}
  explicit CPUContext(const DeviceOption& option) {
    CAFFE_ENFORCE_EQ(option.device_type(), (int)CPU);
  }
  void SwitchToDevice(int /*stream_id*/ = 0) {}
  bool FinishDeviceComputation() { return true; }
  static std::pair<void*, std::function<void(void*)>> New(size_t nbytes) {
    void* p = nullptr;
    if (nbytes) {
      CAFFE_ENFORCE(posix_memalign(&p, 64, nbytes) == 0, "host allocation of ", nbytes, " bytes failed");
      memset(p, 0, nbytes);
    }
    return {p, [](void* q) { free(q); }};
  }
  template <class SrcContext, class DstContext>
  void CopyBytes(size_t nbytes, const void* src, void* dst) {
    if (nbytes) memcpy(dst, src, nbytes);
  }
  template <typename T, class SrcContext, class DstContext>
  void Copy(size_t n, const T* src, T* dst) {
    CopyBytes<SrcContext, DstContext>(n * sizeof(T), src, dst);
  }
  static bool HasAsyncPartDefault() { return false; }
};

}  // namespace caffe2
#endif
