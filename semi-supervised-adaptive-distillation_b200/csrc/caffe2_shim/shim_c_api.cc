// extern "C" handle API over the operator-boundary shim so a non-C++ host (the Python tests and
// bench here; pybind in real Caffe2: caffe2/caffe2/python/pybind_state.cc) can feed blobs, create
// operators/nets from NetDef text and fetch results.  Declared in include/c2_shim_api.h.
#include <cstring>
#include <string>

#include "caffe2/core/context_gpu.h"
#include "caffe2/core/net.h"
#include "caffe2/core/operator.h"
#include "caffe2/core/workspace.h"

#define C2_API extern "C" __attribute__((visibility("default")))

using namespace caffe2;

namespace {
thread_local std::string g_last_error;
thread_local std::string g_text_result;

template <typename F>
int Guard(F&& f) {
  try {
    g_last_error.clear();
    return f();
  } catch (const EnforceNotMet& e) {
    g_last_error = e.msg();
    return -1;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

enum DType { kFloat = 1, kInt32 = 2 };

template <class Context>
void ShareExternal(Tensor<Context>* t, int dtype, const vector<TIndex>& dims, void* ptr) {
  t->Resize(dims);
  if (dtype == kFloat) t->ShareExternalPointer(static_cast<float*>(ptr));
  else if (dtype == kInt32) t->ShareExternalPointer(static_cast<int*>(ptr));
  else CAFFE_THROW("unsupported dtype code ", dtype);
}
}  // namespace

C2_API const char* c2_last_error() { return g_last_error.c_str(); }

C2_API void* c2_workspace_create() { return new Workspace(); }
C2_API void c2_workspace_destroy(void* ws) { delete static_cast<Workspace*>(ws); }

// Make blob `name` a tensor that BORROWS `ptr` (host memory for device_type 0, device memory on
// `gpu_id` for device_type 1).  The caller keeps ownership and must keep it alive.
C2_API int c2_feed_external(void* ws, const char* name, int device_type, int dtype, const int64_t* dims,
                            int ndim, void* ptr) {
  return Guard([&] {
    vector<TIndex> d(dims, dims + ndim);
    Blob* blob = static_cast<Workspace*>(ws)->CreateBlob(name);
    if (device_type == CUDA) ShareExternal(blob->GetMutable<Tensor<CUDAContext>>(), dtype, d, ptr);
    else ShareExternal(blob->GetMutable<Tensor<CPUContext>>(), dtype, d, ptr);
    return 0;
  });
}

// Describe a tensor blob. dims must have room for 8 entries. ptr may be null when unallocated.
C2_API int c2_tensor_info(void* ws, const char* name, int* device_type, int* dtype, int64_t* dims, int* ndim,
                          void** ptr) {
  return Guard([&] {
    const Blob* blob = static_cast<Workspace*>(ws)->GetBlob(name);
    CAFFE_ENFORCE(blob, "no blob named ", name);
    auto fill = [&](const auto& t, int dev) {
      *device_type = dev;
      *dtype = t.template IsType<float>() ? kFloat : (t.template IsType<int>() ? kInt32 : 0);
      CAFFE_ENFORCE(t.ndim() <= 8, "too many dims");
      *ndim = t.ndim();
      for (int i = 0; i < t.ndim(); ++i) dims[i] = t.dim(i);
      *ptr = t.size() > 0 ? const_cast<void*>(t.raw_data()) : nullptr;
    };
    if (blob->IsType<Tensor<CUDAContext>>()) fill(blob->Get<Tensor<CUDAContext>>(), (int)CUDA);
    else if (blob->IsType<Tensor<CPUContext>>()) fill(blob->Get<Tensor<CPUContext>>(), (int)CPU);
    else CAFFE_THROW("blob ", name, " is not a tensor");
    return 0;
  });
}

// Copy a tensor blob to host memory (synchronises the device for CUDA tensors).
C2_API int c2_fetch(void* ws, const char* name, void* host_dst, size_t nbytes) {
  return Guard([&] {
    const Blob* blob = static_cast<Workspace*>(ws)->GetBlob(name);
    CAFFE_ENFORCE(blob, "no blob named ", name);
    if (blob->IsType<Tensor<CUDAContext>>()) {
      const auto& t = blob->Get<Tensor<CUDAContext>>();
      CAFFE_ENFORCE_EQ(t.nbytes(), nbytes);
      CUDA_ENFORCE(cudaDeviceSynchronize());
      if (nbytes) CUDA_ENFORCE(cudaMemcpy(host_dst, t.raw_data(), nbytes, cudaMemcpyDeviceToHost));
    } else {
      const auto& t = blob->Get<Tensor<CPUContext>>();
      CAFFE_ENFORCE_EQ(t.nbytes(), nbytes);
      if (nbytes) memcpy(host_dst, t.raw_data(), nbytes);
    }
    return 0;
  });
}

C2_API int c2_has_blob(void* ws, const char* name) { return static_cast<Workspace*>(ws)->HasBlob(name) ? 1 : 0; }

C2_API int c2_has_operator(const char* type, int device_type) {
  OperatorRegistry* r = device_type == CUDA ? CUDAOperatorRegistry() : CPUOperatorRegistry();
  return r->Has(type) ? 1 : 0;
}
C2_API int c2_has_schema(const char* type) { return OpSchemaRegistry::Schema(type) ? 1 : 0; }
C2_API int c2_schema_arity(const char* type, int* min_in, int* max_in, int* min_out, int* max_out) {
  const OpSchema* s = OpSchemaRegistry::Schema(type);
  if (!s) return -1;
  *min_in = s->min_input(); *max_in = s->max_input(); *min_out = s->min_output(); *max_out = s->max_output();
  return 0;
}

// comma-separated operator keys registered for a device type
C2_API const char* c2_registered_operators(int device_type) {
  OperatorRegistry* r = device_type == CUDA ? CUDAOperatorRegistry() : CPUOperatorRegistry();
  g_text_result.clear();
  for (const auto& k : r->Keys()) { if (!g_text_result.empty()) g_text_result += ","; g_text_result += k; }
  return g_text_result.c_str();
}

// Instantiate (schema-checked) and Run() one operator given as OperatorDef text: the reference's
// host-synchronous Operator::Run semantics (operator.h:369-382).
C2_API int c2_run_operator_once(void* ws, const char* op_text) {
  return Guard([&] {
    OperatorDef def;
    ParseOperatorDefText(op_text, &def);
    return static_cast<Workspace*>(ws)->RunOperatorOnce(def) ? 0 : 1;
  });
}

C2_API int c2_create_net(void* ws, const char* net_text, int overwrite) {
  return Guard([&] {
    NetDef def;
    ParseNetDefText(net_text, &def);
    static_cast<Workspace*>(ws)->CreateNet(def, overwrite != 0);
    return 0;
  });
}
C2_API int c2_run_net(void* ws, const char* name) {
  return Guard([&] { return static_cast<Workspace*>(ws)->RunNet(name) ? 0 : 1; });
}
// enqueue only (no fence): for timing loops and CUDA-graph capture on an adopted stream
C2_API int c2_run_net_async(void* ws, const char* name) {
  return Guard([&] {
    NetBase* net = static_cast<Workspace*>(ws)->GetNet(name);
    CAFFE_ENFORCE(net, "Network ", name, " does not exist yet.");
    return net->RunAsync() ? 0 : 1;
  });
}

// Gradient OperatorDefs for a forward op (what caffe2/python/core.py:1818 obtains through pybind).
// g_outputs[i] = name of the gradient blob of output i ("" = none).  Returns NetDef text whose ops
// are the gradient ops, or nullptr on error.  g_inputs_out (optional, comma-joined) receives the
// gradient blob name produced for each forward input ("" = none).
C2_API const char* c2_gradient_defs(const char* op_text, const char* const* g_outputs, int n_outputs) {
  int rc = Guard([&] {
    OperatorDef def;
    ParseOperatorDefText(op_text, &def);
    vector<GradientWrapper> go(def.output_size());
    for (int i = 0; i < n_outputs && i < def.output_size(); ++i) go[i].dense_ = g_outputs[i] ? g_outputs[i] : "";
    GradientOpsMeta meta = GetGradientForOp(def, go);
    NetDef net;
    for (const auto& op : meta.ops_) *net.add_op() = op;
    for (const auto& gi : meta.g_input_) net.add_external_output(gi.dense_);
    g_text_result = NetDefToText(net);
    return 0;
  });
  return rc == 0 ? g_text_result.c_str() : nullptr;
}

// Route this thread's (gpu, stream_id) context stream to a caller-owned cudaStream_t
// (nullptr = the legacy default stream).
C2_API int c2_adopt_stream(int gpu_id, int stream_id, void* stream) {
  return Guard([&] {
    CUDAContext::AdoptExternalStream(gpu_id, stream_id, static_cast<cudaStream_t>(stream));
    return 0;
  });
}

// round-trip a NetDef through the text reader/writer (used by host-logic tests)
C2_API const char* c2_normalize_net_text(const char* net_text) {
  int rc = Guard([&] {
    NetDef def;
    ParseNetDefText(net_text, &def);
    g_text_result = NetDefToText(def);
    return 0;
  });
  return rc == 0 ? g_text_result.c_str() : nullptr;
}
