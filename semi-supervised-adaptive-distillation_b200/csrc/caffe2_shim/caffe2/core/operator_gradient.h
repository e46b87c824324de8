// Shim of caffe2/caffe2/core/operator_gradient.h:216,320-321 — gradient makers.  The reference's
// Python autograd (caffe2/caffe2/python/core.py:1818) calls these through pybind to obtain the
// "<Op>Gradient" OperatorDefs; the shim exposes the same through GetGradientForOp().
#ifndef SAD_SHIM_OPERATOR_GRADIENT_H_
#define SAD_SHIM_OPERATOR_GRADIENT_H_

#include "caffe2/core/registry.h"
#include "caffe2/proto/caffe2.pb.h"

namespace caffe2 {

struct GradientWrapper {
  string dense_;
  string indices_;
  string values_;
  inline bool IsDense() const { return dense_.size() != 0; }
  inline bool IsSparse() const { return indices_.size() != 0 || values_.size() != 0; }
  inline bool IsEmpty() const { return !IsDense() && !IsSparse(); }
};

struct GradientOpsMeta {
  vector<OperatorDef> ops_;
  vector<GradientWrapper> g_input_;
  GradientOpsMeta() {}
  GradientOpsMeta(const vector<OperatorDef>& ops, const vector<GradientWrapper>& v) : ops_(ops), g_input_(v) {}
};

class GradientMakerBase {
 public:
  GradientMakerBase(const OperatorDef& def, const vector<GradientWrapper>& g_output)
      : def_(def), g_output_(g_output), g_input_(def.input_size()) {}
  virtual ~GradientMakerBase() {}
  virtual bool CopyDeviceOption() const { return true; }
  virtual bool CopyEngine() const { return true; }
  virtual bool CopyArguments() const { return true; }

  virtual GradientOpsMeta Get() {
    vector<OperatorDef> new_defs = GetGradientDefs();
    for (auto& opdef : new_defs) opdef.set_is_gradient_op(true);
    return GradientOpsMeta(new_defs, g_input_);
  }
  const OperatorDef& Def() const { return def_; }

 protected:
  virtual vector<OperatorDef> GetGradientDefs() { CAFFE_THROW("Not Implemented."); }

  string I(const int i) { return def_.input(i); }
  string O(const int i) { return def_.output(i); }
  string GI(const int i) {
    CAFFE_ENFORCE(!g_input_.at(i).IsSparse(), "Input ", def_.input(i), " already set to sparse.");
    g_input_.at(i).dense_ = GradientName(def_.input(i));
    return GradientName(def_.input(i));
  }
  string GO(const int i) {
    CAFFE_ENFORCE(g_output_.at(i).IsDense(), "Gradient of output ", def_.output(i), " is not dense or not provided.");
    return g_output_.at(i).dense_;
  }
  bool GradOut(int i) { return !g_output_.at(i).IsEmpty(); }
  void SetDense(const int i, const string& name) { g_input_.at(i).dense_ = name; }

  // reference operator_gradient.h:216 — one gradient op inheriting device option, engine and
  // (this is what carries gamma/alpha/beta/scale to the backward op) every argument of the forward.
  inline vector<OperatorDef> SingleGradientDef(const string& type, const string& name,
                                               const vector<string>& inputs, const vector<string>& outputs) {
    OperatorDef g;
    g.set_type(type);
    g.set_name(name);
    for (const auto& in : inputs) g.add_input(in);
    for (const auto& out : outputs) g.add_output(out);
    if (CopyDeviceOption() && def_.has_device_option()) *g.mutable_device_option() = def_.device_option();
    if (CopyEngine() && def_.has_engine()) g.set_engine(def_.engine());
    if (CopyArguments()) for (const auto& a : def_.arg()) *g.add_arg() = a;
    return vector<OperatorDef>{g};
  }
  // the overload with extra arguments (operator_gradient.h:216-234), used by the Conv gradient maker for no_bias
  inline vector<OperatorDef> SingleGradientDef(const string& type, const string& name, const vector<string>& inputs,
                                               const vector<string>& outputs, const vector<Argument>& extra_args) {
    vector<OperatorDef> v = SingleGradientDef(type, name, inputs, outputs);
    for (const auto& a : extra_args) *v[0].add_arg() = a;
    return v;
  }

 public:
  static string GradientName(const string& name) { return name + "_grad"; }
  static bool IsGradientBlob(const string& name) {
    return name.length() > 5 && name.find("_grad") == name.length() - 5;
  }

 protected:
  const OperatorDef& def_;
  const vector<GradientWrapper>& g_output_;
  vector<GradientWrapper> g_input_;
};

class NoGradient : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  vector<OperatorDef> GetGradientDefs() override { return vector<OperatorDef>(); }
};

typedef Registry<std::string, GradientMakerBase, const OperatorDef&, const vector<GradientWrapper>&> GradientRegistryT;
typedef Registerer<std::string, GradientMakerBase, const OperatorDef&, const vector<GradientWrapper>&> GradientRegisterer;
GradientRegistryT* GradientRegistry();

#define REGISTER_GRADIENT(name, ...)                                             \
  namespace {                                                                    \
  static GradientRegisterer CAFFE_ANONYMOUS_VARIABLE(g_grad_##name)(             \
      #name, GradientRegistry(), GradientRegisterer::DefaultCreator<__VA_ARGS__>); \
  }
#define NO_GRADIENT(name) REGISTER_GRADIENT(name, NoGradient)
// reference operator_gradient.h:281-293,328-332: asking for the gradient of such an operator is an error
class ThrowInTheTowelIfGradientIsCalled : public GradientMakerBase {
  using GradientMakerBase::GradientMakerBase;
  GradientOpsMeta Get() override { CAFFE_THROW("One should not call gradient for operator ", def_.type(), "."); }
};
#define SHOULD_NOT_DO_GRADIENT(name) REGISTER_GRADIENT(name, ThrowInTheTowelIfGradientIsCalled)

GradientOpsMeta GetGradientForOp(const OperatorDef& def, const vector<GradientWrapper>& g_output);

}  // namespace caffe2
#endif
