// Shim of caffe2/caffe2/core/tensor.h:109,290-333,500-506,594-598,609,673 — Tensor<Context>.
// Semantics kept: Resize(vector<TIndex>()) is a 1-element scalar; storage is allocated lazily by
// mutable_data<T>() through Context::New and kept when the tensor shrinks; data<T>() throws on a
// dtype mismatch or on unallocated storage; ShareExternalPointer borrows caller memory.
#ifndef SAD_SHIM_TENSOR_H_
#define SAD_SHIM_TENSOR_H_

#include <numeric>

#include "caffe2/core/context.h"
#include "caffe2/core/logging.h"
#include "caffe2/core/typeid.h"

namespace caffe2 {

template <class Context>
class Tensor {
 public:
  Tensor() {}
  explicit Tensor(const vector<TIndex>& dims) { Resize(dims); }
  explicit Tensor(const vector<int>& dims) { Resize(dims); }
  virtual ~Tensor() noexcept {}

  template <typename... Ts>
  void Resize(Ts... dim_source) {
    bool size_changed = SetDims(dim_source...);
    if (size_changed && size_ * meta_.itemsize() > capacity_) FreeMemory();
  }
  template <class OtherContext>
  void ResizeLike(const Tensor<OtherContext>& src) {
    Resize(src.dims());
  }
  void FreeMemory() {
    data_.reset();
    capacity_ = 0;
  }

  const vector<TIndex>& dims() const { return dims_; }
  int ndim() const { return (int)dims_.size(); }
  TIndex size() const { return size_; }
  size_t itemsize() const { return meta_.itemsize(); }
  size_t nbytes() const { return size_ * meta_.itemsize(); }
  const TypeMeta& meta() const { return meta_; }
  TIndex dim(int i) const {
    CAFFE_ENFORCE_LT_WITH_IDX(i);
    return dims_[i];
  }
  int dim32(int i) const {
    CAFFE_ENFORCE_LT_WITH_IDX(i);
    CAFFE_ENFORCE_LT(dims_[i], (TIndex)INT32_MAX);
    return (int)dims_[i];
  }
  TIndex size_from_dim(int k) const {
    TIndex r = 1;
    for (int i = k; i < (int)dims_.size(); ++i) r *= dims_[i];
    return r;
  }

  template <typename T>
  bool IsType() const { return meta_.Match<T>(); }

  const void* raw_data() const {
    CAFFE_ENFORCE(data_.get() || size_ == 0, "tensor has no storage yet");
    return data_.get();
  }
  template <typename T>
  const T* data() const {
    CAFFE_ENFORCE_WITH_CALLER(
        data_.get() || size_ == 0,
        "The tensor is of non-zero shape, but its data is not allocated yet. "
        "Caffe2 uses a lazy allocation, so you will need to call mutable_data() or "
        "raw_mutable_data() to actually allocate memory.");
    CAFFE_ENFORCE_WITH_CALLER(
        IsType<T>(), "Tensor type mismatch, caller expects elements to be ",
        TypeMeta::Make<T>().name(), " while tensor contains ", meta_.name());
    return static_cast<const T*>(data_.get());
  }
  void* raw_mutable_data(const TypeMeta& meta) {
    if (meta_ == meta && (data_.get() || size_ == 0)) return data_.get();
    bool had = data_.get() != nullptr;
    meta_ = meta;
    CAFFE_ENFORCE_WITH_CALLER(size_ >= 0, "Tensor is not initialized. You probably need to call Resize() first.");
    if (size_ == 0) return data_.get();
    if (!had || size_ * meta_.itemsize() > capacity_) {
      auto ptr_and_deleter = Context::New(size_ * meta_.itemsize());
      data_.reset(ptr_and_deleter.first, ptr_and_deleter.second);
      capacity_ = size_ * meta_.itemsize();
    }
    return data_.get();
  }
  template <typename T>
  T* mutable_data() {
    if ((size_ == 0 || data_.get()) && IsType<T>()) return static_cast<T*>(data_.get());
    return static_cast<T*>(raw_mutable_data(TypeMeta::Make<T>()));
  }

  // Borrow caller-owned memory (real Caffe2: tensor.h ShareExternalPointer).
  template <typename T>
  void ShareExternalPointer(T* src, size_t capacity = 0) {
    meta_ = TypeMeta::Make<T>();
    CAFFE_ENFORCE_WITH_CALLER(size_ >= 0, "To share data with a raw pointer, you need to set shape first.");
    data_.reset(static_cast<void*>(src), [](void*) {});
    capacity_ = capacity ? capacity : size_ * meta_.itemsize();
  }

  template <class SrcContext, class ContextForCopy>
  void CopyFrom(const Tensor<SrcContext>& src, ContextForCopy* context) {
    if ((void*)&src == (void*)this) return;
    meta_ = src.meta();
    Resize(src.dims());
    if (size() > 0) {
      context->template CopyBytes<SrcContext, Context>(nbytes(), src.raw_data(), raw_mutable_data(meta_));
    }
  }

 protected:
  vector<TIndex> dims_;
  TIndex size_ = -1;
  TypeMeta meta_;
  std::shared_ptr<void> data_;
  size_t capacity_ = 0;

  void CAFFE_ENFORCE_LT_WITH_IDX(int i) const {
    CAFFE_ENFORCE_WITH_CALLER(i >= 0 && i < (int)dims_.size(), "Exceeding ndim limit: ", i, " vs ", dims_.size());
  }
  template <typename T, typename = typename std::enable_if<std::is_integral<T>::value>::type>
  bool SetDims(const vector<T>& src) {
    auto old_size = size_;
    dims_.resize(src.size());
    TIndex new_size = 1;
    for (size_t i = 0; i < src.size(); ++i) {
      new_size *= src[i];
      dims_[i] = src[i];
    }
    size_ = new_size;
    return size_ != old_size;
  }
  bool SetDims() {
    auto old_size = size_;
    dims_.resize(0);
    size_ = 1;
    return size_ != old_size;
  }
  bool SetDims(const TIndex d0) { return SetDims(vector<TIndex>{d0}); }
  bool SetDims(const TIndex d0, const TIndex d1) { return SetDims(vector<TIndex>{d0, d1}); }
  bool SetDims(const TIndex d0, const TIndex d1, const TIndex d2) { return SetDims(vector<TIndex>{d0, d1, d2}); }
  bool SetDims(const TIndex d0, const TIndex d1, const TIndex d2, const TIndex d3) {
    return SetDims(vector<TIndex>{d0, d1, d2, d3});
  }
};

typedef Tensor<CPUContext> TensorCPU;

}  // namespace caffe2
#endif
