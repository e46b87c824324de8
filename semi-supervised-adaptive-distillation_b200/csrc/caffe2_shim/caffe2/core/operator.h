// Shim of caffe2/caffe2/core/operator.h:40-149,329-499,698-789 — the drop-in boundary.
//
// An operator is a class deriving Operator<Context>, constructed from (OperatorDef, Workspace*),
// that overrides RunOnDevice() and enqueues on context_.cuda_stream() without synchronising.
// It is created by name through a per-device registry filled at library-load time by
// REGISTER_{CPU,CUDA}_OPERATOR, after its OpSchema (arity) is verified.
#ifndef SAD_SHIM_OPERATOR_H_
#define SAD_SHIM_OPERATOR_H_

#include "caffe2/core/blob.h"
#include "caffe2/core/common.h"
#include "caffe2/core/logging.h"
#include "caffe2/core/operator_schema.h"
#include "caffe2/core/registry.h"
#include "caffe2/core/tensor.h"
#include "caffe2/core/workspace.h"
#include "caffe2/proto/caffe2.pb.h"

namespace caffe2 {

// Reference: caffe2/caffe2/utils/proto_utils.cc:253-300.  Type-strict: a float argument must be
// stored in Argument.f, an int/bool argument in Argument.i, a string in Argument.s.
class ArgumentHelper {
 public:
  // instance form (proto_utils.h:227-251): ArgumentHelper helper(def); helper.GetSingleArgument<T>(name, default)
  explicit ArgumentHelper(const OperatorDef& def) : def_(&def) {}
  bool HasArgument(const string& name) const { return HasArgument(*def_, name); }
  template <typename T>
  T GetSingleArgument(const string& name, const T& default_value) const {
    return GetSingleArgument<OperatorDef, T>(*def_, name, default_value);
  }
  template <typename T>
  bool HasSingleArgumentOfType(const string& name) const { return HasSingleArgumentOfType<OperatorDef, T>(*def_, name); }
  template <typename T>
  vector<T> GetRepeatedArgument(const string& name, const vector<T>& default_value = vector<T>()) const {
    return GetRepeatedArgument<OperatorDef, T>(*def_, name, default_value);
  }

  template <typename Def>
  static bool HasArgument(const Def& def, const string& name) {
    for (const auto& a : def.arg()) if (a.name() == name) return true;
    return false;
  }
  template <typename Def, typename T>
  static T GetSingleArgument(const Def& def, const string& name, const T& default_value);
  template <typename Def, typename T>
  static bool HasSingleArgumentOfType(const Def& def, const string& name);
  template <typename Def, typename T>
  static vector<T> GetRepeatedArgument(const Def& def, const string& name,
                                       const vector<T>& default_value = vector<T>());

 private:
  template <typename Def>
  static const Argument* Find(const Def& def, const string& name) {
    for (const auto& a : def.arg()) if (a.name() == name) return &a;
    return nullptr;
  }
  const OperatorDef* def_ = nullptr;
};

// reference proto_utils.h:253-262 / proto_utils.cc:302-330
template <typename T>
Argument MakeArgument(const string& name, const T& value);
template <>
inline Argument MakeArgument(const string& name, const int& value) { Argument a; a.set_name(name); a.set_i(value); return a; }
template <>
inline Argument MakeArgument(const string& name, const int64_t& value) { Argument a; a.set_name(name); a.set_i(value); return a; }
template <>
inline Argument MakeArgument(const string& name, const bool& value) { Argument a; a.set_name(name); a.set_i(value); return a; }
template <>
inline Argument MakeArgument(const string& name, const float& value) { Argument a; a.set_name(name); a.set_f(value); return a; }
template <>
inline Argument MakeArgument(const string& name, const string& value) { Argument a; a.set_name(name); a.set_s(value); return a; }

#define SAD_SHIM_SINGLE_ARG(T, fieldname)                                                            \
  template <>                                                                                        \
  inline T ArgumentHelper::GetSingleArgument<OperatorDef, T>(const OperatorDef& def, const string& name, \
                                                             const T& default_value) {              \
    const Argument* a = Find(def, name);                                                             \
    if (!a) return default_value;                                                                    \
    CAFFE_ENFORCE(a->has_##fieldname(), "Argument ", name, " does not have the right field: expected field " #fieldname); \
    return static_cast<T>(a->fieldname());                                                           \
  }                                                                                                  \
  template <>                                                                                        \
  inline bool ArgumentHelper::HasSingleArgumentOfType<OperatorDef, T>(const OperatorDef& def, const string& name) { \
    const Argument* a = Find(def, name);                                                             \
    return a && a->has_##fieldname();                                                                \
  }
SAD_SHIM_SINGLE_ARG(float, f)
SAD_SHIM_SINGLE_ARG(double, f)
SAD_SHIM_SINGLE_ARG(bool, i)
SAD_SHIM_SINGLE_ARG(int, i)
SAD_SHIM_SINGLE_ARG(int64_t, i)
SAD_SHIM_SINGLE_ARG(size_t, i)
SAD_SHIM_SINGLE_ARG(string, s)
#undef SAD_SHIM_SINGLE_ARG

#define SAD_SHIM_REPEATED_ARG(T, fieldname)                                                          \
  template <>                                                                                        \
  inline vector<T> ArgumentHelper::GetRepeatedArgument<OperatorDef, T>(                              \
      const OperatorDef& def, const string& name, const vector<T>& default_value) {                  \
    const Argument* a = Find(def, name);                                                             \
    if (!a) return default_value;                                                                    \
    vector<T> values;                                                                                \
    for (const auto& v : a->fieldname()) values.push_back(static_cast<T>(v));                        \
    return values;                                                                                   \
  }
SAD_SHIM_REPEATED_ARG(float, floats)
SAD_SHIM_REPEATED_ARG(int, ints)
SAD_SHIM_REPEATED_ARG(int64_t, ints)
SAD_SHIM_REPEATED_ARG(string, strings)
#undef SAD_SHIM_REPEATED_ARG

class OperatorBase {
 public:
  explicit OperatorBase(const OperatorDef& operator_def, Workspace* ws);
  virtual ~OperatorBase() noexcept {}

  inline bool HasArgument(const string& name) const { return ArgumentHelper::HasArgument(operator_def_, name); }
  template <typename T>
  inline T GetSingleArgument(const string& name, const T& default_value) const {
    return ArgumentHelper::GetSingleArgument<OperatorDef, T>(operator_def_, name, default_value);
  }
  template <typename T>
  inline bool HasSingleArgumentOfType(const string& name) const {
    return ArgumentHelper::HasSingleArgumentOfType<OperatorDef, T>(operator_def_, name);
  }
  template <typename T>
  inline vector<T> GetRepeatedArgument(const string& name, const vector<T>& default_value = vector<T>()) const {
    return ArgumentHelper::GetRepeatedArgument<OperatorDef, T>(operator_def_, name, default_value);
  }

  template <typename T>
  inline const T& Input(int idx) {
    try {
      return inputs_.at(idx)->template Get<T>();
    } catch (EnforceNotMet& enf) {
      enf.AppendMessage(".\nOffending Blob name: " + operator_def_.input(idx) + ".\n");
      throw;
    }
  }
  template <typename T>
  inline T* Output(int idx) { return outputs_.at(idx)->template GetMutable<T>(); }
  inline const Blob& InputBlob(int idx) { return *inputs_.at(idx); }
  inline Blob* OutputBlob(int idx) { return outputs_.at(idx); }
  template <typename T>
  inline bool InputIsType(int idx) { return inputs_.at(idx)->template IsType<T>(); }
  inline int InputSize() const { return (int)inputs_.size(); }
  inline int OutputSize() const { return (int)outputs_.size(); }
  inline const vector<const Blob*>& Inputs() const { return inputs_; }
  inline const vector<Blob*>& Outputs() { return outputs_; }

  virtual bool Run(int /*stream_id*/ = 0) { CAFFE_THROW("Not implemented"); }
  virtual bool RunAsync(int stream_id = 0) { return Run(stream_id); }
  inline const OperatorDef& debug_def() const { return operator_def_; }
  inline const OperatorDef& def() const { return operator_def_; }
  const string& type() const { return operator_def_.type(); }

 protected:
  void AddRelatedBlobInfo(EnforceNotMet* err);
  OperatorDef operator_def_;
  vector<const Blob*> inputs_;
  vector<Blob*> outputs_;
  DISABLE_COPY_AND_ASSIGN(OperatorBase);
};

#define OP_SINGLE_ARG(type, name, variable, default) variable(OperatorBase::GetSingleArgument<type>(name, (default)))
#define INPUT_TAGS(first_input, ...) enum _InputTags { first_input = 0, __VA_ARGS__ }
#define OUTPUT_TAGS(first_input, ...) enum _OutputTags { first_input = 0, __VA_ARGS__ }

template <class Context>
class Operator : public OperatorBase {
 public:
  explicit Operator(const OperatorDef& operator_def, Workspace* ws)
      : OperatorBase(operator_def, ws), context_(operator_def.device_option()) {
    context_.SwitchToDevice(0);
  }
  ~Operator() noexcept override {}

  inline const Tensor<Context>& Input(int idx) { return OperatorBase::template Input<Tensor<Context>>(idx); }
  inline Tensor<Context>* Output(int idx) { return OperatorBase::template Output<Tensor<Context>>(idx); }

  // reference operator.h:369-394: run, then synchronise the stream (host-synchronous per op)
  bool Run(int stream_id = 0) final {
    try {
      context_.SwitchToDevice(stream_id);
      bool started = RunOnDevice();
      bool finished = context_.FinishDeviceComputation();
      if (!finished) throw EnforceNotMet(__FILE__, __LINE__, "finished", "Error from operator: \n" + ProtoDebugString(operator_def_));
      return started && finished;
    } catch (EnforceNotMet& err) {
      err.AppendMessage("Error from operator: \n" + ProtoDebugString(operator_def_));
      AddRelatedBlobInfo(&err);
      throw;
    }
  }
  // reference operator.h:397-413: enqueue only; the caller fences
  bool RunAsync(int stream_id = 0) final {
    try {
      context_.SwitchToDevice(stream_id);
      return RunOnDevice();
    } catch (EnforceNotMet& err) {
      err.AppendMessage("Error from operator: \n" + ProtoDebugString(operator_def_));
      AddRelatedBlobInfo(&err);
      throw;
    }
  }
  virtual bool RunOnDevice() = 0;

 protected:
  Context context_;
};

// reference operator.h:540-645 — run-time dispatch on a tensor's element type: DispatchHelper<TensorTypes<int32_t, int64_t>>::
// call(this, Input(INDICES)) invokes op->DoRunWithType<T>() for the first listed T the tensor holds
template <typename... Types>
struct TensorTypes {};
template <typename Sizes, typename... ExtraArgs>
struct DispatchHelper;
template <typename FirstType, typename... Types, typename... ExtraArgs>
struct DispatchHelper<TensorTypes<FirstType, Types...>, ExtraArgs...> {
  template <typename Op, typename Context>
  static bool call(Op* op, const Tensor<Context>& tensor) {
    if (tensor.template IsType<FirstType>()) return op->template DoRunWithType<ExtraArgs..., FirstType>();
    return DispatchHelper<TensorTypes<Types...>, ExtraArgs...>::template call<Op, Context>(op, tensor);
  }
};
template <typename... ExtraArgs>
struct DispatchHelper<TensorTypes<>, ExtraArgs...> {
  template <typename Op, typename Context>
  static bool call(Op* /*op*/, const Tensor<Context>& /*tensor*/) {
    CAFFE_THROW("Unsupported type of tensor");
  }
};

#define USE_OPERATOR_BASE_FUNCTIONS                  \
  /* using override */ using OperatorBase::HasArgument; \
  /* using override */ using OperatorBase::GetSingleArgument; \
  /* using override */ using OperatorBase::HasSingleArgumentOfType; \
  /* using override */ using OperatorBase::GetRepeatedArgument; \
  /* using override */ using OperatorBase::InputIsType; \
  /* using override */ using OperatorBase::InputSize; \
  /* using override */ using OperatorBase::OutputSize

#define USE_OPERATOR_FUNCTIONS(context)                    \
  USE_OPERATOR_BASE_FUNCTIONS;                             \
  /* using override */ using Operator<context>::context_;  \
  /* using override */ using Operator<context>::Input;     \
  /* using override */ using Operator<context>::Output

#define USE_OPERATOR_CONTEXT_FUNCTIONS USE_OPERATOR_FUNCTIONS(Context)

#define USE_SIMPLE_CTOR_DTOR(name)                                              \
  name(const OperatorDef& operator_def, Workspace* ws) : Operator<Context>(operator_def, ws) {} \
  virtual ~name() noexcept {}

#define CAFFE_NOT_IMPLEMENTED CAFFE_THROW("Not Implemented.")

typedef Registry<std::string, OperatorBase, const OperatorDef&, Workspace*> OperatorRegistry;
typedef Registerer<std::string, OperatorBase, const OperatorDef&, Workspace*> OperatorRegisterer;
OperatorRegistry* CPUOperatorRegistry();
OperatorRegistry* CUDAOperatorRegistry();

#define SAD_SHIM_REGISTER_CLASS(RegistryFn, key, keyvar, ...)                                   \
  namespace {                                                                                   \
  static OperatorRegisterer CAFFE_ANONYMOUS_VARIABLE(g_##keyvar)(                               \
      key, RegistryFn(), OperatorRegisterer::DefaultCreator<__VA_ARGS__>);                      \
  }

#define REGISTER_CPU_OPERATOR(name, ...)                           \
  extern void CAFFE2_PLEASE_ADD_OPERATOR_SCHEMA_FOR_##name();      \
  static void __attribute__((unused)) CAFFE_ANONYMOUS_VARIABLE_CPU##name() { CAFFE2_PLEASE_ADD_OPERATOR_SCHEMA_FOR_##name(); } \
  SAD_SHIM_REGISTER_CLASS(CPUOperatorRegistry, #name, cpu_##name, __VA_ARGS__)
#define REGISTER_CUDA_OPERATOR(name, ...)                          \
  extern void CAFFE2_PLEASE_ADD_OPERATOR_SCHEMA_FOR_##name();      \
  static void __attribute__((unused)) CAFFE_ANONYMOUS_VARIABLE_CUDA##name() { CAFFE2_PLEASE_ADD_OPERATOR_SCHEMA_FOR_##name(); } \
  SAD_SHIM_REGISTER_CLASS(CUDAOperatorRegistry, #name, cuda_##name, __VA_ARGS__)
// reference operator.h:726-732 — engine-qualified key "<Name>_ENGINE_<ENGINE>"
#define REGISTER_CPU_OPERATOR_WITH_ENGINE(name, engine, ...) \
  SAD_SHIM_REGISTER_CLASS(CPUOperatorRegistry, #name "_ENGINE_" #engine, cpu_##name##_##engine, __VA_ARGS__)
#define REGISTER_CUDA_OPERATOR_WITH_ENGINE(name, engine, ...) \
  SAD_SHIM_REGISTER_CLASS(CUDAOperatorRegistry, #name "_ENGINE_" #engine, cuda_##name##_##engine, __VA_ARGS__)
#define REGISTER_CUDNN_OPERATOR(name, ...) REGISTER_CUDA_OPERATOR_WITH_ENGINE(name, CUDNN, __VA_ARGS__)

// reference operator.h:786-789
unique_ptr<OperatorBase> CreateOperator(const OperatorDef& operator_def, Workspace* ws, int net_position = -1);

}  // namespace caffe2

#include "caffe2/core/operator_gradient.h"

#endif
