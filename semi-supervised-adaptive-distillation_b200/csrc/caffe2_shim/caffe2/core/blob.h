// Shim of caffe2/caffe2/core/blob.h — a type-erased owner of one object (Tensor<Context> here).
#ifndef SAD_SHIM_BLOB_H_
#define SAD_SHIM_BLOB_H_

#include "caffe2/core/logging.h"
#include "caffe2/core/typeid.h"

namespace caffe2 {

class Blob {
 public:
  Blob() {}
  ~Blob() { Reset(); }
  template <class T>
  bool IsType() const { return meta_.Match<T>() && pointer_; }
  template <class T>
  const T& Get() const {
    CAFFE_ENFORCE(IsType<T>(), "wrong type for the Blob instance. Blob contains ", meta_.name());
    return *static_cast<const T*>(pointer_);
  }
  template <class T>
  T* GetMutable() {
    if (IsType<T>()) return static_cast<T*>(pointer_);
    Reset();
    pointer_ = new T();
    meta_ = TypeMeta::Make<T>();
    destroy_ = [](void* p) { delete static_cast<T*>(p); };
    return static_cast<T*>(pointer_);
  }
  void Reset() {
    if (pointer_ && destroy_) destroy_(pointer_);
    pointer_ = nullptr;
    destroy_ = nullptr;
    meta_ = TypeMeta();
  }
  const TypeMeta& meta() const { return meta_; }
  DISABLE_COPY_AND_ASSIGN(Blob);

 private:
  TypeMeta meta_;
  void* pointer_ = nullptr;
  void (*destroy_)(void*) = nullptr;
};

}  // namespace caffe2
#endif
