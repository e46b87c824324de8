// Shim of caffe2/caffe2/core/typeid.h — a tiny TypeMeta: identity, item size, name.
#ifndef SAD_SHIM_TYPEID_H_
#define SAD_SHIM_TYPEID_H_

#include <cstddef>
#include <cstdint>

namespace caffe2 {

class TypeMeta {
 public:
  TypeMeta() : id_(0), itemsize_(0), name_("nullptr (uninitialized)") {}
  template <typename T>
  static TypeMeta Make() {
    return TypeMeta(Id<T>(), sizeof(T), Name<T>());
  }
  template <typename T>
  bool Match() const { return id_ == Id<T>(); }
  size_t itemsize() const { return itemsize_; }
  const char* name() const { return name_; }
  intptr_t id() const { return id_; }
  bool operator==(const TypeMeta& o) const { return id_ == o.id_; }
  bool operator!=(const TypeMeta& o) const { return id_ != o.id_; }

 private:
  TypeMeta(intptr_t id, size_t sz, const char* name) : id_(id), itemsize_(sz), name_(name) {}
  // ids are fixed small integers for the POD types that cross the C boundary; any other type
  // gets the address of a per-type tag (unique within one shared object, which is all we need).
  template <typename T>
  static intptr_t Id() {
    static const char tag = 0;
    return reinterpret_cast<intptr_t>(&tag);
  }
  template <typename T>
  static const char* Name() { return "non-POD"; }
  intptr_t id_;
  size_t itemsize_;
  const char* name_;
};

#define SAD_SHIM_POD_TYPE(T, ID)                                         \
  template <> inline intptr_t TypeMeta::Id<T>() { return ID; }           \
  template <> inline const char* TypeMeta::Name<T>() { return #T; }
SAD_SHIM_POD_TYPE(float, 1)
SAD_SHIM_POD_TYPE(int, 2)
SAD_SHIM_POD_TYPE(int64_t, 3)
SAD_SHIM_POD_TYPE(double, 4)
SAD_SHIM_POD_TYPE(uint8_t, 5)
SAD_SHIM_POD_TYPE(bool, 6)
SAD_SHIM_POD_TYPE(uint16_t, 7)  // raw bf16 / fp16 storage
#undef SAD_SHIM_POD_TYPE

}  // namespace caffe2
#endif
