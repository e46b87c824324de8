// Runtime behind the operator-boundary shim: text-format NetDef reader/writer, registries,
// CreateOperator, Workspace, the in-order executor and CUDAContext's stream table.
//
// Reference behaviour followed (not code): caffe2/caffe2/core/operator.cc (CreateOperator: schema
// verification, engine-qualified lookup, device dispatch), workspace.cc, net_simple.cc,
// context_gpu.h:54-134 (thread-local stream per gpu/stream_id).
#include <cctype>
#include <cstring>
#include <mutex>
#include <sstream>

#include "caffe2/core/context_gpu.h"
#include "caffe2/core/net.h"
#include "caffe2/core/operator.h"
#include "caffe2/core/workspace.h"

namespace caffe2 {

// ---------------------------------------------------------------------------------------------
// protobuf text format: a small recursive-descent reader producing a field tree.
// ---------------------------------------------------------------------------------------------
namespace {

struct TextNode {
  // scalar fields: name -> raw token (string tokens are unescaped); message fields: children
  std::vector<std::pair<std::string, std::string>> scalars;
  std::vector<std::pair<std::string, TextNode>> messages;
};

class TextReader {
 public:
  explicit TextReader(const std::string& s) : s_(s) {}
  TextNode ParseTop() {
    TextNode n = ParseFields(/*until_brace=*/false);
    SkipWs();
    CAFFE_ENFORCE(pos_ == s_.size(), "text format: trailing characters at offset ", pos_);
    return n;
  }

 private:
  const std::string& s_;
  size_t pos_ = 0;

  void SkipWs() {
    while (pos_ < s_.size()) {
      char c = s_[pos_];
      if (c == '#') {
        while (pos_ < s_.size() && s_[pos_] != '\n') ++pos_;
      } else if (isspace((unsigned char)c) || c == ',' || c == ';') {
        ++pos_;
      } else {
        break;
      }
    }
  }
  std::string Ident() {
    SkipWs();
    size_t b = pos_;
    while (pos_ < s_.size() && (isalnum((unsigned char)s_[pos_]) || s_[pos_] == '_')) ++pos_;
    CAFFE_ENFORCE(pos_ > b, "text format: expected a field name at offset ", b);
    return s_.substr(b, pos_ - b);
  }
  std::string QuotedString() {
    char q = s_[pos_++];
    std::string out;
    while (true) {
      CAFFE_ENFORCE(pos_ < s_.size(), "text format: unterminated string");
      char c = s_[pos_++];
      if (c == q) break;
      if (c == '\\') {
        CAFFE_ENFORCE(pos_ < s_.size(), "text format: dangling escape");
        char e = s_[pos_++];
        switch (e) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case '\\': case '\'': case '"': out.push_back(e); break;
          default:
            if (e >= '0' && e <= '7') {  // octal
              int v = e - '0';
              for (int k = 0; k < 2 && pos_ < s_.size() && s_[pos_] >= '0' && s_[pos_] <= '7'; ++k) v = v * 8 + (s_[pos_++] - '0');
              out.push_back((char)v);
            } else {
              out.push_back(e);
            }
        }
      } else {
        out.push_back(c);
      }
    }
    return out;
  }
  std::string ScalarToken() {
    SkipWs();
    CAFFE_ENFORCE(pos_ < s_.size(), "text format: expected a value");
    if (s_[pos_] == '"' || s_[pos_] == '\'') {
      std::string out = QuotedString();
      // adjacent string literals concatenate
      while (true) {
        SkipWs();
        if (pos_ < s_.size() && (s_[pos_] == '"' || s_[pos_] == '\'')) out += QuotedString();
        else break;
      }
      return out;
    }
    size_t b = pos_;
    while (pos_ < s_.size() && !isspace((unsigned char)s_[pos_]) && s_[pos_] != '}' && s_[pos_] != '{' &&
           s_[pos_] != ',' && s_[pos_] != ';' && s_[pos_] != '#')
      ++pos_;
    CAFFE_ENFORCE(pos_ > b, "text format: empty value at offset ", b);
    return s_.substr(b, pos_ - b);
  }
  TextNode ParseFields(bool until_brace) {
    TextNode node;
    while (true) {
      SkipWs();
      if (pos_ >= s_.size()) {
        CAFFE_ENFORCE(!until_brace, "text format: missing '}'");
        break;
      }
      if (s_[pos_] == '}' || s_[pos_] == '>') {
        CAFFE_ENFORCE(until_brace, "text format: unexpected '}' at offset ", pos_);
        ++pos_;
        break;
      }
      std::string name = Ident();
      SkipWs();
      bool colon = false;
      if (pos_ < s_.size() && s_[pos_] == ':') { colon = true; ++pos_; SkipWs(); }
      if (pos_ < s_.size() && (s_[pos_] == '{' || s_[pos_] == '<')) {
        ++pos_;
        node.messages.emplace_back(name, ParseFields(true));
      } else {
        CAFFE_ENFORCE(colon, "text format: expected ':' after field ", name);
        if (pos_ < s_.size() && s_[pos_] == '[') {  // short repeated form: f: [1, 2]
          ++pos_;
          while (true) {
            SkipWs();
            CAFFE_ENFORCE(pos_ < s_.size(), "text format: missing ']'");
            if (s_[pos_] == ']') { ++pos_; break; }
            node.scalars.emplace_back(name, ScalarToken());
          }
        } else {
          node.scalars.emplace_back(name, ScalarToken());
        }
      }
    }
    return node;
  }
};

int64_t ToInt(const std::string& t) {
  if (t == "true") return 1;
  if (t == "false") return 0;
  try {
    size_t used = 0;
    long long v = std::stoll(t, &used, 0);
    CAFFE_ENFORCE(used == t.size(), "text format: bad integer '", t, "'");
    return v;
  } catch (const std::logic_error&) {
    CAFFE_THROW("text format: bad integer '", t, "'");
  }
}
float ToFloat(const std::string& t) {
  std::string u = t;
  if (!u.empty() && (u.back() == 'f' || u.back() == 'F') && u != "inf" && u != "-inf") u.pop_back();
  try {
    size_t used = 0;
    float v = std::stof(u, &used);
    CAFFE_ENFORCE(used == u.size(), "text format: bad float '", t, "'");
    return v;
  } catch (const std::logic_error&) {
    CAFFE_THROW("text format: bad float '", t, "'");
  }
}

void FillDeviceOption(const TextNode& n, DeviceOption* d) {
  for (const auto& kv : n.scalars) {
    if (kv.first == "device_type") {
      if (kv.second == "CPU") d->set_device_type(CPU);
      else if (kv.second == "CUDA") d->set_device_type(CUDA);
      else d->set_device_type((int)ToInt(kv.second));
    } else if (kv.first == "cuda_gpu_id") {
      d->set_cuda_gpu_id((int)ToInt(kv.second));
    }  // random_seed, node_name: not on the path
  }
}
void FillArgument(const TextNode& n, Argument* a) {
  for (const auto& kv : n.scalars) {
    if (kv.first == "name") a->set_name(kv.second);
    else if (kv.first == "f") a->set_f(ToFloat(kv.second));
    else if (kv.first == "i") a->set_i(ToInt(kv.second));
    else if (kv.first == "s") a->set_s(kv.second);
    else if (kv.first == "floats") a->add_floats(ToFloat(kv.second));
    else if (kv.first == "ints") a->add_ints(ToInt(kv.second));
    else if (kv.first == "strings") a->add_strings(kv.second);
    else CAFFE_THROW("text format: unknown Argument field '", kv.first, "'");
  }
}
void FillOperatorDef(const TextNode& n, OperatorDef* op) {
  for (const auto& kv : n.scalars) {
    if (kv.first == "input") op->add_input(kv.second);
    else if (kv.first == "output") op->add_output(kv.second);
    else if (kv.first == "name") op->set_name(kv.second);
    else if (kv.first == "type") op->set_type(kv.second);
    else if (kv.first == "engine") op->set_engine(kv.second);
    else if (kv.first == "is_gradient_op") op->set_is_gradient_op(ToInt(kv.second) != 0);
    else if (kv.first == "control_input" || kv.first == "debug_info") {}
    else CAFFE_THROW("text format: unknown OperatorDef field '", kv.first, "'");
  }
  for (const auto& kv : n.messages) {
    if (kv.first == "arg") FillArgument(kv.second, op->add_arg());
    else if (kv.first == "device_option") FillDeviceOption(kv.second, op->mutable_device_option());
    else CAFFE_THROW("text format: unknown OperatorDef message '", kv.first, "'");
  }
}

std::string Quote(const std::string& s) {
  std::string o = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
    else if (c == '\n') o += "\\n";
    else o.push_back(c);
  }
  return o + "\"";
}
void EmitOp(std::ostringstream& os, const OperatorDef& d, const char* ind) {
  for (const auto& s : d.input()) os << ind << "input: " << Quote(s) << "\n";
  for (const auto& s : d.output()) os << ind << "output: " << Quote(s) << "\n";
  if (!d.name().empty()) os << ind << "name: " << Quote(d.name()) << "\n";
  os << ind << "type: " << Quote(d.type()) << "\n";
  for (const auto& a : d.arg()) {
    os << ind << "arg {\n" << ind << "  name: " << Quote(a.name()) << "\n";
    char buf[64];
    if (a.has_f()) { snprintf(buf, sizeof buf, "%.9g", a.f()); os << ind << "  f: " << buf << "\n"; }
    if (a.has_i()) os << ind << "  i: " << a.i() << "\n";
    if (a.has_s()) os << ind << "  s: " << Quote(a.s()) << "\n";
    for (float f : a.floats()) { snprintf(buf, sizeof buf, "%.9g", f); os << ind << "  floats: " << buf << "\n"; }
    for (int64_t i : a.ints()) os << ind << "  ints: " << i << "\n";
    for (const auto& s : a.strings()) os << ind << "  strings: " << Quote(s) << "\n";
    os << ind << "}\n";
  }
  if (d.has_device_option()) {
    os << ind << "device_option {\n" << ind << "  device_type: " << d.device_option().device_type() << "\n";
    if (d.device_option().has_cuda_gpu_id()) os << ind << "  cuda_gpu_id: " << d.device_option().cuda_gpu_id() << "\n";
    os << ind << "}\n";
  }
  if (d.has_engine()) os << ind << "engine: " << Quote(d.engine()) << "\n";
  if (d.is_gradient_op()) os << ind << "is_gradient_op: true\n";
}

}  // namespace

bool ParseOperatorDefText(const std::string& text, OperatorDef* out) {
  TextReader r(text);
  TextNode n = r.ParseTop();
  // accept both a bare OperatorDef body and one wrapped as `op { ... }`
  if (n.scalars.empty() && n.messages.size() == 1 && n.messages[0].first == "op") {
    FillOperatorDef(n.messages[0].second, out);
  } else {
    FillOperatorDef(n, out);
  }
  return true;
}

bool ParseNetDefText(const std::string& text, NetDef* out) {
  TextReader r(text);
  TextNode n = r.ParseTop();
  for (const auto& kv : n.scalars) {
    if (kv.first == "name") out->set_name(kv.second);
    else if (kv.first == "type") out->set_type(kv.second);
    else if (kv.first == "num_workers") out->set_num_workers((int)ToInt(kv.second));
    else if (kv.first == "external_input") out->add_external_input(kv.second);
    else if (kv.first == "external_output") out->add_external_output(kv.second);
    else CAFFE_THROW("text format: unknown NetDef field '", kv.first, "'");
  }
  for (const auto& kv : n.messages) {
    if (kv.first == "op") FillOperatorDef(kv.second, out->add_op());
    else if (kv.first == "device_option") FillDeviceOption(kv.second, out->mutable_device_option());
    else if (kv.first == "arg") {}
    else CAFFE_THROW("text format: unknown NetDef message '", kv.first, "'");
  }
  return true;
}

std::string OperatorDefToText(const OperatorDef& def) {
  std::ostringstream os;
  EmitOp(os, def, "");
  return os.str();
}
std::string NetDefToText(const NetDef& def) {
  std::ostringstream os;
  if (!def.name().empty()) os << "name: " << Quote(def.name()) << "\n";
  for (const auto& op : def.op()) {
    os << "op {\n";
    EmitOp(os, op, "  ");
    os << "}\n";
  }
  if (!def.type().empty()) os << "type: " << Quote(def.type()) << "\n";
  if (def.num_workers()) os << "num_workers: " << def.num_workers() << "\n";
  for (const auto& s : def.external_input()) os << "external_input: " << Quote(s) << "\n";
  for (const auto& s : def.external_output()) os << "external_output: " << Quote(s) << "\n";
  return os.str();
}

// ---------------------------------------------------------------------------------------------
// registries
// ---------------------------------------------------------------------------------------------
OperatorRegistry* CPUOperatorRegistry() {
  static OperatorRegistry* r = new OperatorRegistry();
  return r;
}
OperatorRegistry* CUDAOperatorRegistry() {
  static OperatorRegistry* r = new OperatorRegistry();
  return r;
}
GradientRegistryT* GradientRegistry() {
  static GradientRegistryT* r = new GradientRegistryT();
  return r;
}
CaffeMap<string, OpSchema>& OpSchemaRegistry::map() {
  static CaffeMap<string, OpSchema>* m = new CaffeMap<string, OpSchema>();
  return *m;
}
OpSchema& OpSchemaRegistry::NewSchema(const string& key, const string& file, const int line) {
  auto& m = map();
  if (m.count(key)) {
    fprintf(stderr, "Trying to register schema with name %s from file %s line %d, but it is already registered from file %s line %d\n",
            key.c_str(), file.c_str(), line, m[key].file().c_str(), m[key].line());
    abort();
  }
  m.emplace(std::make_pair(key, OpSchema(file, line)));
  return m[key];
}
const OpSchema* OpSchemaRegistry::Schema(const string& key) {
  auto& m = map();
  return m.count(key) ? &m[key] : nullptr;
}

bool OpSchema::Verify(const OperatorDef& def) const {
  if (def.input_size() < min_input_ || def.input_size() > max_input_) {
    fprintf(stderr, "Input size %d not in range [min=%d, max=%d].\n", def.input_size(), min_input_, max_input_);
    return false;
  }
  if (!num_inputs_allowed(def.input_size())) {
    fprintf(stderr, "Input size %d is not allowed by the schema of %s.\n", def.input_size(), def.type().c_str());
    return false;
  }
  if (def.output_size() < min_output_ || def.output_size() > max_output_) {
    fprintf(stderr, "Output size %d not in range [min=%d, max=%d].\n", def.output_size(), min_output_, max_output_);
    return false;
  }
  for (int in = 0; in < def.input_size(); ++in)
    for (int out = 0; out < def.output_size(); ++out)
      if (def.input(in) == def.output(out) && !inplace_allowed_(in, out)) {
        fprintf(stderr, "Input index %d and output idx %d (%s) are set to be in-place but this is actually not supported by op %s\n",
                in, out, def.input(in).c_str(), def.type().c_str());
        return false;
      }
  return true;
}

// ---------------------------------------------------------------------------------------------
// OperatorBase / CreateOperator
// ---------------------------------------------------------------------------------------------
OperatorBase::OperatorBase(const OperatorDef& operator_def, Workspace* ws) : operator_def_(operator_def) {
  for (const string& input_str : operator_def.input()) {
    auto* blob = ws->GetBlob(input_str);
    CAFFE_ENFORCE(blob != nullptr, "op ", operator_def.type(), ": Encountered a non-existing input blob: ", input_str);
    inputs_.push_back(blob);
  }
  for (const string& output_str : operator_def.output()) outputs_.push_back(ws->CreateBlob(output_str));
}

void OperatorBase::AddRelatedBlobInfo(EnforceNotMet* err) {
  if (!err->caller()) return;
  for (size_t i = 0; i < inputs_.size(); ++i) {
    // the offending tensor identifies itself through the exception's caller pointer
    const void* obj = nullptr;
    if (inputs_[i]->IsType<Tensor<CPUContext>>()) obj = &inputs_[i]->Get<Tensor<CPUContext>>();
    else if (inputs_[i]->IsType<Tensor<CUDAContext>>()) obj = &inputs_[i]->Get<Tensor<CUDAContext>>();
    if (obj && obj == err->caller()) {
      err->AppendMessage("Offending Blob name: " + operator_def_.input((int)i) + ".");
      return;
    }
  }
}

unique_ptr<OperatorBase> CreateOperator(const OperatorDef& operator_def, Workspace* ws, int /*net_position*/) {
  const string& type = operator_def.type();
  const OpSchema* schema = OpSchemaRegistry::Schema(type);
  if (schema) {
    CAFFE_ENFORCE(schema->Verify(operator_def), "Operator def did not pass schema checking: ",
                  ProtoDebugString(operator_def));
  }
  OperatorRegistry* registry = nullptr;
  switch (operator_def.device_option().device_type()) {
    case CPU: registry = CPUOperatorRegistry(); break;
    case CUDA: registry = CUDAOperatorRegistry(); break;
    default: CAFFE_THROW("Unknown device type: ", operator_def.device_option().device_type());
  }
  if (operator_def.has_engine()) {
    // comma-separated preference list; first registered engine wins, then the plain op
    std::stringstream ss(operator_def.engine());
    string engine;
    while (std::getline(ss, engine, ',')) {
      const string key = type + "_ENGINE_" + engine;
      if (registry->Has(key)) return registry->Create(key, operator_def, ws);
    }
  }
  auto op = registry->Create(type, operator_def, ws);
  CAFFE_ENFORCE(op, "Cannot create operator of type '", type, "' on the device '",
                operator_def.device_option().device_type() == CUDA ? "CUDA" : "CPU",
                "'. Verify that implementation for the corresponding device exist. Operator def: ",
                ProtoDebugString(operator_def));
  return op;
}

GradientOpsMeta GetGradientForOp(const OperatorDef& def, const vector<GradientWrapper>& g_output) {
  unique_ptr<GradientMakerBase> maker(GradientRegistry()->Create(def.type(), def, g_output));
  CAFFE_ENFORCE(maker, "Gradient maker for operator ", def.type(), " not implemented.");
  GradientOpsMeta meta = maker->Get();
  // copy device option / engine when the maker did not go through SingleGradientDef
  for (OperatorDef& grad_def : meta.ops_) {
    if (maker->CopyDeviceOption() && def.has_device_option() && !grad_def.has_device_option())
      *grad_def.mutable_device_option() = def.device_option();
    if (maker->CopyEngine() && def.has_engine() && !grad_def.has_engine()) grad_def.set_engine(def.engine());
  }
  return meta;
}

// ---------------------------------------------------------------------------------------------
// Workspace / Net
// ---------------------------------------------------------------------------------------------
Workspace::Workspace() {}
Workspace::~Workspace() {
  net_map_.clear();  // nets hold raw Blob pointers: drop them before the blobs
}
Blob* Workspace::CreateBlob(const string& name) {
  auto it = blob_map_.find(name);
  if (it != blob_map_.end()) return it->second.get();
  blob_map_[name] = unique_ptr<Blob>(new Blob());
  return blob_map_[name].get();
}
const Blob* Workspace::GetBlob(const string& name) const {
  auto it = blob_map_.find(name);
  return it == blob_map_.end() ? nullptr : it->second.get();
}
Blob* Workspace::GetBlob(const string& name) {
  return const_cast<Blob*>(static_cast<const Workspace*>(this)->GetBlob(name));
}
bool Workspace::RemoveBlob(const string& name) { return blob_map_.erase(name) != 0; }
vector<string> Workspace::Blobs() const {
  vector<string> names;
  for (const auto& kv : blob_map_) names.push_back(kv.first);
  return names;
}
NetBase* Workspace::CreateNet(const NetDef& net_def, bool overwrite) {
  CAFFE_ENFORCE(!net_def.name().empty(), "NetDef.name is required to create a net in a workspace");
  if (net_map_.count(net_def.name())) {
    CAFFE_ENFORCE(overwrite, "net ", net_def.name(), " already exists; pass overwrite=true to replace it");
    net_map_.erase(net_def.name());
  }
  net_map_[net_def.name()] = caffe2::CreateNet(net_def, this);
  return net_map_[net_def.name()].get();
}
NetBase* Workspace::GetNet(const string& name) {
  auto it = net_map_.find(name);
  return it == net_map_.end() ? nullptr : it->second.get();
}
bool Workspace::RunNet(const string& name) {
  NetBase* net = GetNet(name);
  CAFFE_ENFORCE(net, "Network ", name, " does not exist yet.");
  return net->Run();
}
bool Workspace::RunOperatorOnce(const OperatorDef& op_def) {
  unique_ptr<OperatorBase> op(CreateOperator(op_def, this));
  return op->Run();
}
bool Workspace::RunNetOnce(const NetDef& net_def) {
  unique_ptr<NetBase> net(caffe2::CreateNet(net_def, this));
  return net->Run();
}

NetBase::NetBase(const NetDef& net_def, Workspace* ws) : name_(net_def.name()) {
  // external inputs must already be fed
  for (const string& in : net_def.external_input())
    CAFFE_ENFORCE(ws->HasBlob(in), "net ", net_def.name(), ": external input blob ", in, " is missing");
  for (int idx = 0; idx < net_def.op_size(); ++idx) {
    OperatorDef op_def = net_def.op(idx);
    if (!op_def.has_device_option() && net_def.has_device_option()) *op_def.mutable_device_option() = net_def.device_option();
    operators_.emplace_back(CreateOperator(op_def, ws, idx));
  }
}
bool NetBase::RunAsync() {
  for (auto& op : operators_)
    if (!op->RunAsync()) return false;
  return true;
}
bool NetBase::Run() {
  if (!RunAsync()) return false;
  // one fence for the whole net on every device it touched
  bool ok = true;
  std::set<int> gpus;
  for (auto& op : operators_)
    if (op->def().device_option().device_type() == CUDA) gpus.insert(op->def().device_option().cuda_gpu_id());
  for (int g : gpus) {
    CUDAContext ctx(g);
    ctx.SwitchToDevice(0);
    ok = ctx.FinishDeviceComputation() && ok;
  }
  return ok;
}
unique_ptr<NetBase> CreateNet(const NetDef& net_def, Workspace* ws) {
  return unique_ptr<NetBase>(new NetBase(net_def, ws));
}

// ---------------------------------------------------------------------------------------------
// CUDA context
// ---------------------------------------------------------------------------------------------
int NumCudaDevices() {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}
int CaffeCudaGetDevice() {
  int id = 0;
  CUDA_ENFORCE(cudaGetDevice(&id));
  return id;
}
void CaffeCudaSetDevice(const int id) { CUDA_ENFORCE(cudaSetDevice(id)); }

namespace {
constexpr int kMaxGpus = 16;
constexpr int kMaxStreams = 8;
struct ThreadStreams {
  cudaStream_t own[kMaxGpus][kMaxStreams] = {};
  cudaStream_t adopted[kMaxGpus][kMaxStreams] = {};
  bool has_adopted[kMaxGpus][kMaxStreams] = {};
  ~ThreadStreams() {
    for (int g = 0; g < kMaxGpus; ++g)
      for (int s = 0; s < kMaxStreams; ++s)
        if (own[g][s]) cudaStreamDestroy(own[g][s]);
  }
};
thread_local ThreadStreams tls_streams;
}  // namespace

CUDAContext::CUDAContext(const int gpu_id) : gpu_id_(gpu_id == -1 ? CaffeCudaGetDevice() : gpu_id) {}
CUDAContext::CUDAContext(const DeviceOption& option)
    : gpu_id_(option.has_cuda_gpu_id() ? option.cuda_gpu_id() : CaffeCudaGetDevice()) {
  CAFFE_ENFORCE_EQ(option.device_type(), (int)CUDA);
}
cudaStream_t CUDAContext::cuda_stream(int gpu_id, int stream_id) {
  CAFFE_ENFORCE(gpu_id >= 0 && gpu_id < kMaxGpus && stream_id >= 0 && stream_id < kMaxStreams,
                "gpu/stream id out of range: ", gpu_id, "/", stream_id);
  if (tls_streams.has_adopted[gpu_id][stream_id]) return tls_streams.adopted[gpu_id][stream_id];
  cudaStream_t& s = tls_streams.own[gpu_id][stream_id];
  if (!s) {
    DeviceGuard guard(gpu_id);
    CUDA_ENFORCE(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  return s;
}
void CUDAContext::AdoptExternalStream(int gpu_id, int stream_id, cudaStream_t stream) {
  CAFFE_ENFORCE(gpu_id >= 0 && gpu_id < kMaxGpus && stream_id >= 0 && stream_id < kMaxStreams);
  tls_streams.adopted[gpu_id][stream_id] = stream;
  tls_streams.has_adopted[gpu_id][stream_id] = true;
}
std::pair<void*, std::function<void(void*)>> CUDAContext::New(size_t nbytes) {
  void* p = nullptr;
  if (nbytes) CUDA_ENFORCE(cudaMalloc(&p, nbytes));
  return {p, [](void* q) { if (q) cudaFree(q); }};
}

}  // namespace caffe2
