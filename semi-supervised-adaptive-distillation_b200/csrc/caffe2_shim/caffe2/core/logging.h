// Shim of caffe2/caffe2/core/logging.h:122-163,293 — the error convention of the operator
// boundary: soft failure = RunOnDevice returns false, hard failure = EnforceNotMet thrown.
#ifndef SAD_SHIM_LOGGING_H_
#define SAD_SHIM_LOGGING_H_

#include <cstdio>
#include <exception>
#include <sstream>
#include <cstring>
#include <string>
#include <vector>

#include "caffe2/core/common.h"
#include "caffe2/proto/caffe2.pb.h"

namespace caffe2 {

inline void MakeStringInternal(std::stringstream&) {}
template <typename T, typename... Args>
inline void MakeStringInternal(std::stringstream& ss, const T& t, const Args&... args) {
  ss << t;
  MakeStringInternal(ss, args...);
}
template <typename... Args>
inline std::string MakeString(const Args&... args) {
  std::stringstream ss;
  MakeStringInternal(ss, args...);
  return ss.str();
}

class EnforceNotMet : public std::exception {
 public:
  EnforceNotMet(const char* file, int line, const char* condition, const std::string& msg,
                const void* caller = nullptr)
      : caller_(caller) {
    msg_stack_.push_back(MakeString("[enforce fail at ", file, ":", line, "] ", condition, ". ", msg));
    full_msg_ = msg_stack_[0];
  }
  void AppendMessage(const std::string& msg) {
    msg_stack_.push_back(msg);
    full_msg_ += " " + msg;
  }
  std::string msg() const { return full_msg_; }
  const char* what() const noexcept override { return full_msg_.c_str(); }
  const void* caller() const noexcept { return caller_; }

 private:
  std::vector<std::string> msg_stack_;
  std::string full_msg_;
  const void* caller_;
};

#define CAFFE_ENFORCE(condition, ...)                                                      \
  do {                                                                                     \
    if (!(condition)) {                                                                    \
      throw ::caffe2::EnforceNotMet(__FILE__, __LINE__, #condition,                        \
                                    ::caffe2::MakeString(__VA_ARGS__));                    \
    }                                                                                      \
  } while (false)

#define CAFFE_ENFORCE_WITH_CALLER(condition, ...)                                          \
  do {                                                                                     \
    if (!(condition)) {                                                                    \
      throw ::caffe2::EnforceNotMet(__FILE__, __LINE__, #condition,                        \
                                    ::caffe2::MakeString(__VA_ARGS__), this);              \
    }                                                                                      \
  } while (false)

#define CAFFE_THROW(...) \
  throw ::caffe2::EnforceNotMet(__FILE__, __LINE__, "", ::caffe2::MakeString(__VA_ARGS__))

#define CAFFE_ENFORCE_BINARY_OP_(op, x, y, ...)                                            \
  do {                                                                                     \
    const auto& _x = (x);                                                                  \
    const auto& _y = (y);                                                                  \
    if (!(_x op _y)) {                                                                     \
      throw ::caffe2::EnforceNotMet(                                                       \
          __FILE__, __LINE__, #x " " #op " " #y,                                           \
          ::caffe2::MakeString(_x, " vs ", _y, ". ", ::caffe2::MakeString(__VA_ARGS__)));  \
    }                                                                                      \
  } while (false)
#define CAFFE_ENFORCE_EQ(x, y, ...) CAFFE_ENFORCE_BINARY_OP_(==, x, y, __VA_ARGS__)
#define CAFFE_ENFORCE_NE(x, y, ...) CAFFE_ENFORCE_BINARY_OP_(!=, x, y, __VA_ARGS__)
#define CAFFE_ENFORCE_LE(x, y, ...) CAFFE_ENFORCE_BINARY_OP_(<=, x, y, __VA_ARGS__)
#define CAFFE_ENFORCE_LT(x, y, ...) CAFFE_ENFORCE_BINARY_OP_(<, x, y, __VA_ARGS__)
#define CAFFE_ENFORCE_GE(x, y, ...) CAFFE_ENFORCE_BINARY_OP_(>=, x, y, __VA_ARGS__)
#define CAFFE_ENFORCE_GT(x, y, ...) CAFFE_ENFORCE_BINARY_OP_(>, x, y, __VA_ARGS__)

// glog's debug-only checks (caffe2/caffe2/core/logging_is_not_google_glog.h:121-146): compiled out in release builds,
// which is how Detectron's operators are built; the operands are not evaluated.
#define SAD_SHIM_DCHECK_(x, y) \
  while (false) (void)((x), (y))
#define DCHECK_EQ(x, y) SAD_SHIM_DCHECK_(x, y)
#define DCHECK_NE(x, y) SAD_SHIM_DCHECK_(x, y)
#define DCHECK_LE(x, y) SAD_SHIM_DCHECK_(x, y)
#define DCHECK_LT(x, y) SAD_SHIM_DCHECK_(x, y)
#define DCHECK_GE(x, y) SAD_SHIM_DCHECK_(x, y)
#define DCHECK_GT(x, y) SAD_SHIM_DCHECK_(x, y)

// reference logging.h:115 / logging.cc:41-58 — in-place substring replacement used by schema doc generators
inline size_t ReplaceAll(std::string& s, const char* from, const char* to) {
  size_t n = 0, pos = 0;
  const size_t lf = strlen(from), lt = strlen(to);
  if (!lf) return 0;
  while ((pos = s.find(from, pos)) != std::string::npos) {
    s.replace(pos, lf, to);
    pos += lt;
    ++n;
  }
  return n;
}

// CUDA_ENFORCE lives in common_gpu.h
}  // namespace caffe2
#endif
