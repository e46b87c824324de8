// Shim of caffe2/caffe2/core/common.h (reference :48 TIndex) — only what the hot path uses.
#ifndef SAD_SHIM_COMMON_H_
#define SAD_SHIM_COMMON_H_

#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace caffe2 {
using std::map;
using std::set;
using std::string;
using std::unique_ptr;
using std::vector;
typedef int64_t TIndex;
template <typename K, typename V>
using CaffeMap = std::map<K, V>;

#ifndef DISABLE_COPY_AND_ASSIGN
#define DISABLE_COPY_AND_ASSIGN(classname) \
  classname(const classname&) = delete;    \
  classname& operator=(const classname&) = delete
#endif

#define CAFFE_CONCATENATE_IMPL(s1, s2) s1##s2
#define CAFFE_CONCATENATE(s1, s2) CAFFE_CONCATENATE_IMPL(s1, s2)
#define CAFFE_ANONYMOUS_VARIABLE(str) CAFFE_CONCATENATE(str, __LINE__)

}  // namespace caffe2
#endif
