// Shim of caffe2/caffe2/core/workspace.h — name → Blob map plus the nets created in it.
#ifndef SAD_SHIM_WORKSPACE_H_
#define SAD_SHIM_WORKSPACE_H_

#include "caffe2/core/blob.h"
#include "caffe2/core/common.h"
#include "caffe2/proto/caffe2.pb.h"

namespace caffe2 {

class NetBase;

class Workspace {
 public:
  Workspace();
  ~Workspace();
  bool HasBlob(const string& name) const { return blob_map_.count(name) != 0; }
  Blob* CreateBlob(const string& name);
  const Blob* GetBlob(const string& name) const;
  Blob* GetBlob(const string& name);
  bool RemoveBlob(const string& name);
  vector<string> Blobs() const;

  NetBase* CreateNet(const NetDef& net_def, bool overwrite = false);
  NetBase* GetNet(const string& net_name);
  bool RunNet(const string& net_name);
  bool RunOperatorOnce(const OperatorDef& op_def);
  bool RunNetOnce(const NetDef& net_def);
  DISABLE_COPY_AND_ASSIGN(Workspace);

 private:
  std::map<string, unique_ptr<Blob>> blob_map_;
  std::map<string, unique_ptr<NetBase>> net_map_;
};

}  // namespace caffe2
#endif
