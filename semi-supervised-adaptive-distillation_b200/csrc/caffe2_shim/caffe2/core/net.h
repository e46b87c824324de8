// Shim of caffe2/caffe2/core/net.h — NetBase plus ONE executor: ops run in definition order,
// enqueue-only (RunAsync) on the context stream, with a single fence at the end of Run().
// The reference's DAGNet synchronises after every operator (operator.h:369-382); on this path
// that host round trip costs more than the kernels, so the fence is per net, not per op.
#ifndef SAD_SHIM_NET_H_
#define SAD_SHIM_NET_H_

#include "caffe2/core/operator.h"

namespace caffe2 {

class NetBase {
 public:
  NetBase(const NetDef& net_def, Workspace* ws);
  virtual ~NetBase() noexcept {}
  virtual bool Run();
  // enqueue without the trailing fence (for callers that own the stream, e.g. CUDA-graph capture)
  virtual bool RunAsync();
  const string& Name() const { return name_; }
  const vector<unique_ptr<OperatorBase>>& GetOperators() const { return operators_; }

 protected:
  string name_;
  vector<unique_ptr<OperatorBase>> operators_;
  DISABLE_COPY_AND_ASSIGN(NetBase);
};

unique_ptr<NetBase> CreateNet(const NetDef& net_def, Workspace* ws);

}  // namespace caffe2
#endif
