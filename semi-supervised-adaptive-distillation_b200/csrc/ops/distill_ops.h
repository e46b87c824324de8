// Public C++ surface of the operator library beyond the registered operators themselves.
#ifndef SAD_DISTILL_OPS_H_
#define SAD_DISTILL_OPS_H_

#include "caffe2/proto/caffe2.pb.h"

namespace caffe2 {

// Graph pass over a NetDef as Detectron builds it (retinanet_heads.py:313-352 + autograd): groups
// the per-level SigmoidAdaptiveDistillLoss ops that share a normaliser blob, device and arguments,
// together with their SigmoidAdaptiveDistillLossGradient ops whose d_loss is a ConstantFill of one
// common value, into ONE SigmoidAdaptiveDistillLossMultiLevel op placed where the first forward op
// was.  Blob names are unchanged, so the rest of the net (ConvGradient consumers, loss logging)
// is untouched.  Returns the number of groups fused; nets without the pattern are left as is.
int FuseAdaptiveDistillOps(NetDef* net);

}  // namespace caffe2
#endif
