// Shim of caffe2/caffe2/core/operator_schema.h:496-499 — OPERATOR_SCHEMA(Name) with the arity
// checks CreateOperator enforces, plus the doc-only builders the reference .cc files chain.
#ifndef SAD_SHIM_OPERATOR_SCHEMA_H_
#define SAD_SHIM_OPERATOR_SCHEMA_H_

#include <climits>

#include "caffe2/core/common.h"
#include "caffe2/core/logging.h"
#include "caffe2/proto/caffe2.pb.h"

namespace caffe2 {

class OpSchema {
 public:
  OpSchema() : file_("unknown"), line_(0) {}
  OpSchema(const string& file, const int line) : file_(file), line_(line) {}
  const string& file() const { return file_; }
  int line() const { return line_; }
  const char* doc() const { return doc_.empty() ? nullptr : doc_.c_str(); }
  bool Verify(const OperatorDef& def) const;

  OpSchema& NumInputs(int n) { return NumInputs(n, n); }
  OpSchema& NumInputs(int min, int max) { min_input_ = min; max_input_ = max; return *this; }
  // reference operator_schema.h:70-80: an arbitrary predicate on the input count (WeightedSum: even and positive)
  OpSchema& NumInputs(std::function<bool(int)> allowed) { num_inputs_allowed_ = allowed; return *this; }
  bool num_inputs_allowed(int n) const { return !num_inputs_allowed_ || num_inputs_allowed_(n); }
  OpSchema& NumOutputs(int n) { return NumOutputs(n, n); }
  OpSchema& NumOutputs(int min, int max) { min_output_ = min; max_output_ = max; return *this; }
  OpSchema& AllowInplace(std::function<bool(int, int)> inplace) { inplace_allowed_ = inplace; return *this; }
  OpSchema& AllowInplace(std::set<std::pair<int, int>> inplace) {
    return AllowInplace([inplace](int in, int out) { return inplace.count(std::make_pair(in, out)) != 0; });
  }
  OpSchema& EnforceInplace(std::set<std::pair<int, int>> /*inplace*/) { return *this; }
  OpSchema& AllowOneToOneInplace() { return AllowInplace([](int in, int out) { return in == out; }); }
  OpSchema& IdenticalTypeAndShape() { return *this; }
  OpSchema& IdenticalTypeAndShapeOfInput(int) { return *this; }
  OpSchema& SetDoc(const string& doc) { doc_ = doc; return *this; }
  // reference operator_schema.h:150-214 — shape / cost inference hooks and FillUsing; kept so reference .cc files
  // that chain them compile and so callers can query them
  struct Cost {
    uint64_t flops = 0;
    uint64_t bytes_moved = 0;
    uint64_t params_bytes = 0;
  };
  typedef std::function<vector<TensorShape>(const OperatorDef&, const vector<TensorShape>&)> TensorInferenceFunctionType;
  typedef std::function<struct Cost(const OperatorDef&, const vector<TensorShape>&)> CostInferenceFunctionType;
  OpSchema& TensorInferenceFunction(TensorInferenceFunctionType f) { tensor_inference_function_ = f; return *this; }
  OpSchema& CostInferenceFunction(CostInferenceFunctionType f) { cost_inference_function_ = f; return *this; }
  bool HasCostInferenceFunction() const { return !!cost_inference_function_; }
  vector<TensorShape> InferTensor(const OperatorDef& def, const vector<TensorShape>& in) const {
    return tensor_inference_function_(def, in);
  }
  struct Cost InferCost(const OperatorDef& def, const vector<TensorShape>& in) const {
    return cost_inference_function_(def, in);
  }
  OpSchema& FillUsing(std::function<void(OpSchema&)> populator) { if (populator) populator(*this); return *this; }
  OpSchema& Arg(const char* name, const char* description) { args_.emplace_back(name, description); return *this; }
  OpSchema& Input(const int n, const char* name, const char* description) {
    if ((int)input_desc_.size() <= n) input_desc_.resize(n + 1);
    input_desc_[n] = std::make_pair(name, description);
    return *this;
  }
  OpSchema& Output(const int n, const char* name, const char* description) {
    if ((int)output_desc_.size() <= n) output_desc_.resize(n + 1);
    output_desc_[n] = std::make_pair(name, description);
    return *this;
  }
  int min_input() const { return min_input_; }
  int max_input() const { return max_input_; }
  int min_output() const { return min_output_; }
  int max_output() const { return max_output_; }
  const std::vector<std::pair<const char*, const char*>>& args() const { return args_; }

 private:
  string file_, doc_;
  int line_;
  int min_input_ = 0, max_input_ = INT_MAX, min_output_ = 0, max_output_ = INT_MAX;
  std::function<bool(int, int)> inplace_allowed_ = [](int, int) { return false; };
  std::function<bool(int)> num_inputs_allowed_;
  std::vector<std::pair<const char*, const char*>> args_, input_desc_, output_desc_;
  TensorInferenceFunctionType tensor_inference_function_;
  CostInferenceFunctionType cost_inference_function_;
};

// reference proto_utils.h:48-60 / proto_utils.cc — shape descriptor helpers used by inference functions
inline TensorShape CreateTensorShape(vector<int> dims, TensorProto::DataType dt) {
  TensorShape ts;
  for (int d : dims) ts.add_dims(d);
  ts.set_data_type(dt);
  ts.set_unknown_shape(false);
  return ts;
}
inline vector<TIndex> GetDimsVector(const TensorShape& shape) {
  vector<TIndex> dims;
  for (auto d : shape.dims()) dims.push_back(d);
  return dims;
}

class OpSchemaRegistry {
 public:
  static OpSchema& NewSchema(const string& key, const string& file, const int line);
  static const OpSchema* Schema(const string& key);

 private:
  OpSchemaRegistry() = delete;
  static CaffeMap<string, OpSchema>& map();
};

#define OPERATOR_SCHEMA(name)                                   \
  void CAFFE2_PLEASE_ADD_OPERATOR_SCHEMA_FOR_##name() {};       \
  static OpSchema* CAFFE_ANONYMOUS_VARIABLE(name) = &OpSchemaRegistry::NewSchema(#name, __FILE__, __LINE__)

}  // namespace caffe2
#endif
