// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for caffe2/caffe2/core/flags.h:96-171 without the
// gflags / registry machinery: a flag is a plain global in namespace scope named FLAGS_<name>.
#ifndef SAD_REF_SHIM_FLAGS_H_
#define SAD_REF_SHIM_FLAGS_H_
#include <string>
#define CAFFE2_DECLARE_bool(name) extern bool FLAGS_##name
#define CAFFE2_DECLARE_int(name) extern int FLAGS_##name
#define CAFFE2_DEFINE_bool(name, default_value, help_str) bool FLAGS_##name = default_value
#define CAFFE2_DEFINE_int(name, default_value, help_str) int FLAGS_##name = default_value
#endif
