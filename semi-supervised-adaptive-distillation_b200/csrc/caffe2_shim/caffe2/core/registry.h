// Shim of caffe2/caffe2/core/registry.h:217-220 — string-keyed creator registry filled by static
// initialisers when the operator library is dlopen'ed (reference: detectron/lib/utils/env.py:59-73).
#ifndef SAD_SHIM_REGISTRY_H_
#define SAD_SHIM_REGISTRY_H_

#include <mutex>

#include "caffe2/core/common.h"
#include "caffe2/core/logging.h"

namespace caffe2 {

template <class SrcType, class ObjectType, class... Args>
class Registry {
 public:
  typedef std::function<std::unique_ptr<ObjectType>(Args...)> Creator;
  Registry() {}
  void Register(const SrcType& key, Creator creator) {
    std::lock_guard<std::mutex> lock(register_mutex_);
    if (registry_.count(key) != 0) {
      fprintf(stderr, "Key already registered: %s\n", std::string(key).c_str());
      std::abort();
    }
    registry_[key] = creator;
  }
  inline bool Has(const SrcType& key) { return registry_.count(key) != 0; }
  unique_ptr<ObjectType> Create(const SrcType& key, Args... args) {
    if (registry_.count(key) == 0) return nullptr;
    return registry_[key](args...);
  }
  vector<SrcType> Keys() {
    vector<SrcType> keys;
    for (const auto& it : registry_) keys.push_back(it.first);
    return keys;
  }
  DISABLE_COPY_AND_ASSIGN(Registry);

 private:
  CaffeMap<SrcType, Creator> registry_;
  std::mutex register_mutex_;
};

template <class SrcType, class ObjectType, class... Args>
class Registerer {
 public:
  Registerer(const SrcType& key, Registry<SrcType, ObjectType, Args...>* registry,
             typename Registry<SrcType, ObjectType, Args...>::Creator creator) {
    registry->Register(key, creator);
  }
  template <class DerivedType>
  static unique_ptr<ObjectType> DefaultCreator(Args... args) {
    return unique_ptr<ObjectType>(new DerivedType(args...));
  }
};

}  // namespace caffe2
#endif
