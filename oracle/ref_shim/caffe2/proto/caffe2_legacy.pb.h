// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for the protoc-generated header of
// caffe2/caffe2/proto/caffe2_legacy.proto:5-37 (no protoc in this image): the one enum
// conv_pool_op_base.h:46-50 reads.
#ifndef SAD_REF_SHIM_CAFFE2_LEGACY_PB_H_
#define SAD_REF_SHIM_CAFFE2_LEGACY_PB_H_
namespace caffe2 {
enum LegacyPadding { NOTSET = 0, VALID = 1, SAME = 2, CAFFE_LEGACY_POOLING = 3 };
}
#endif
