// Operator-boundary shim: plain-struct stand-ins for the protobuf messages the hot path touches.
//
// The reference generates these from caffe2/caffe2/proto/caffe2.proto:97-176 with protoc
// (absent from this image).  Only the fields the distillation ops, Conv/Relu and the
// in-order executor read are kept; accessor names follow protobuf's generated C++ API so
// operator sources written against real Caffe2 compile unchanged.
#ifndef SAD_SHIM_CAFFE2_PB_H_
#define SAD_SHIM_CAFFE2_PB_H_

#include <cstdint>
#include <string>
#include <vector>

namespace caffe2 {

enum DeviceType { CPU = 0, CUDA = 1 };

// caffe2.proto:97-107
class Argument {
 public:
  const std::string& name() const { return name_; }
  void set_name(const std::string& n) { name_ = n; }
  bool has_f() const { return has_f_; }
  bool has_i() const { return has_i_; }
  bool has_s() const { return has_s_; }
  float f() const { return f_; }
  int64_t i() const { return i_; }
  const std::string& s() const { return s_; }
  void set_f(float v) { f_ = v; has_f_ = true; }
  void set_i(int64_t v) { i_ = v; has_i_ = true; }
  void set_s(const std::string& v) { s_ = v; has_s_ = true; }
  const std::vector<float>& floats() const { return floats_; }
  const std::vector<int64_t>& ints() const { return ints_; }
  const std::vector<std::string>& strings() const { return strings_; }
  void add_floats(float v) { floats_.push_back(v); }
  void add_ints(int64_t v) { ints_.push_back(v); }
  void add_strings(const std::string& v) { strings_.push_back(v); }
  int floats_size() const { return (int)floats_.size(); }
  int ints_size() const { return (int)ints_.size(); }
  int strings_size() const { return (int)strings_.size(); }

 private:
  std::string name_;
  float f_ = 0.f;
  int64_t i_ = 0;
  std::string s_;
  bool has_f_ = false, has_i_ = false, has_s_ = false;
  std::vector<float> floats_;
  std::vector<int64_t> ints_;
  std::vector<std::string> strings_;
};

// caffe2.proto:16-68 — the shape-only descriptors schema inference functions exchange
// (operator_schema.h:150-185; conv_pool_op_base.h:375-520 builds them for Conv / pooling).
class TensorProto {
 public:
  enum DataType { UNDEFINED = 0, FLOAT = 1, INT32 = 2, BYTE = 3, STRING = 4, BOOL = 5, UINT8 = 6, INT8 = 7,
                  UINT16 = 8, INT16 = 9, INT64 = 10, FLOAT16 = 12, DOUBLE = 13 };
};
const TensorProto::DataType TensorProto_DataType_FLOAT = TensorProto::FLOAT;
class TensorShape {
 public:
  const std::vector<int64_t>& dims() const { return dims_; }
  int64_t dims(int i) const { return dims_.at(i); }
  int dims_size() const { return (int)dims_.size(); }
  void add_dims(int64_t d) { dims_.push_back(d); }
  void clear_dims() { dims_.clear(); }
  TensorProto::DataType data_type() const { return data_type_; }
  void set_data_type(TensorProto::DataType t) { data_type_ = t; }
  bool unknown_shape() const { return unknown_shape_; }
  void set_unknown_shape(bool b) { unknown_shape_ = b; }

 private:
  std::vector<int64_t> dims_;
  TensorProto::DataType data_type_ = TensorProto::FLOAT;
  bool unknown_shape_ = false;
};

// caffe2.proto:122-137
class DeviceOption {
 public:
  int device_type() const { return device_type_; }
  void set_device_type(int t) { device_type_ = t; has_device_type_ = true; }
  bool has_device_type() const { return has_device_type_; }
  int cuda_gpu_id() const { return cuda_gpu_id_; }
  void set_cuda_gpu_id(int g) { cuda_gpu_id_ = g; has_cuda_gpu_id_ = true; }
  bool has_cuda_gpu_id() const { return has_cuda_gpu_id_; }

 private:
  int device_type_ = CPU;
  int cuda_gpu_id_ = 0;
  bool has_device_type_ = false, has_cuda_gpu_id_ = false;
};

// caffe2.proto:142-163
class OperatorDef {
 public:
  const std::vector<std::string>& input() const { return input_; }
  const std::vector<std::string>& output() const { return output_; }
  const std::string& input(int i) const { return input_.at(i); }
  const std::string& output(int i) const { return output_.at(i); }
  int input_size() const { return (int)input_.size(); }
  int output_size() const { return (int)output_.size(); }
  void add_input(const std::string& s) { input_.push_back(s); }
  void add_output(const std::string& s) { output_.push_back(s); }
  void set_input(int i, const std::string& s) { input_.at(i) = s; }
  void set_output(int i, const std::string& s) { output_.at(i) = s; }
  void clear_input() { input_.clear(); }
  void clear_output() { output_.clear(); }
  const std::string& name() const { return name_; }
  void set_name(const std::string& s) { name_ = s; }
  const std::string& type() const { return type_; }
  void set_type(const std::string& s) { type_ = s; }
  const std::string& engine() const { return engine_; }
  void set_engine(const std::string& s) { engine_ = s; }
  bool has_engine() const { return !engine_.empty(); }
  const std::vector<Argument>& arg() const { return arg_; }
  const Argument& arg(int i) const { return arg_.at(i); }
  int arg_size() const { return (int)arg_.size(); }
  Argument* add_arg() { arg_.emplace_back(); return &arg_.back(); }
  std::vector<Argument>* mutable_arg() { return &arg_; }
  const DeviceOption& device_option() const { return device_option_; }
  DeviceOption* mutable_device_option() { has_device_option_ = true; return &device_option_; }
  bool has_device_option() const { return has_device_option_; }
  bool is_gradient_op() const { return is_gradient_op_; }
  void set_is_gradient_op(bool b) { is_gradient_op_ = b; }

 private:
  std::vector<std::string> input_, output_;
  std::string name_, type_, engine_;
  std::vector<Argument> arg_;
  DeviceOption device_option_;
  bool has_device_option_ = false;
  bool is_gradient_op_ = false;
};

// caffe2.proto:166-176
class NetDef {
 public:
  const std::string& name() const { return name_; }
  void set_name(const std::string& s) { name_ = s; }
  const std::string& type() const { return type_; }
  void set_type(const std::string& s) { type_ = s; }
  int num_workers() const { return num_workers_; }
  void set_num_workers(int n) { num_workers_ = n; }
  const std::vector<OperatorDef>& op() const { return op_; }
  const OperatorDef& op(int i) const { return op_.at(i); }
  int op_size() const { return (int)op_.size(); }
  OperatorDef* add_op() { op_.emplace_back(); return &op_.back(); }
  std::vector<OperatorDef>* mutable_op() { return &op_; }
  const DeviceOption& device_option() const { return device_option_; }
  DeviceOption* mutable_device_option() { has_device_option_ = true; return &device_option_; }
  bool has_device_option() const { return has_device_option_; }
  const std::vector<std::string>& external_input() const { return external_input_; }
  const std::vector<std::string>& external_output() const { return external_output_; }
  void add_external_input(const std::string& s) { external_input_.push_back(s); }
  void add_external_output(const std::string& s) { external_output_.push_back(s); }

 private:
  std::string name_, type_;
  int num_workers_ = 0;
  std::vector<OperatorDef> op_;
  DeviceOption device_option_;
  bool has_device_option_ = false;
  std::vector<std::string> external_input_, external_output_;
};

// protobuf text format (what Detectron dumps as net.pbtxt, tools/train_net.py:306-312).
// Throws EnforceNotMet on malformed input.  Implemented in shim_runtime.cc.
bool ParseNetDefText(const std::string& text, NetDef* out);
bool ParseOperatorDefText(const std::string& text, OperatorDef* out);
std::string OperatorDefToText(const OperatorDef& def);
std::string NetDefToText(const NetDef& def);
inline std::string ProtoDebugString(const OperatorDef& def) { return OperatorDefToText(def); }

}  // namespace caffe2

#endif  // SAD_SHIM_CAFFE2_PB_H_
