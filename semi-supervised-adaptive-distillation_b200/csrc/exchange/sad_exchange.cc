// Gradient exchange of the data-parallel step: host C++ over NCCL (include/sad_exchange.h).
//
// Reference (single process, all GPUs): detectron/lib/modeling/optimizer.py:72-92 adds one NCCLAllreduce operator per gradient
// blob; caffe2/caffe2/contrib/nccl/cuda_nccl_gpu.cc:139-225 (runNCCL) records an event on every participating context's stream,
// makes the NCCL stream wait for it, calls ncclAllReduce and makes the context streams wait for the result.  The same
// event plumbing, one process per GPU: record on the producer stream -> the communication stream waits -> ncclAllReduce ->
// join: record on the communication stream -> the consumer (optimiser) stream waits.  The host never blocks.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "sad_exchange.h"

#define SAD_EXPORT __attribute__((visibility("default")))

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& m) {
  g_err = m;
  return code;
}
int cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SAD_EXCHANGE_OK;
  return fail(SAD_EXCHANGE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ---- NCCL, resolved at run time (nccl.h: ncclUniqueId is 128 opaque bytes passed by value; ncclFloat32 = 7, ncclSum = 0) ----
struct NcclUniqueId {
  char internal[SAD_EXCHANGE_UNIQUE_ID_BYTES];
};
typedef void* NcclComm;
// ncclConfig_t as of NCCL 2.27 (nccl.h:70-88).  NCCL reads `size` / `version` first and accepts configuration structs of older
// releases, so this restatement stays valid for newer libraries; fields left at "undefined" keep the library's defaults.
struct NcclConfig {
  size_t size;
  unsigned int magic;
  unsigned int version;
  int blocking, cgaClusterSize, minCTAs, maxCTAs;
  const char* netName;
  int splitShare, trafficClass;
  const char* commName;
  int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
  int nChannelsPerNetPeer, nvlinkCentricSched;   // NCCL 2.28 (nccl.h:83-101); only written when the library is >= 2.28
};
constexpr size_t kNcclConfigBytesV22703 = offsetof(NcclConfig, nChannelsPerNetPeer);
constexpr int kNcclUndefInt = -2147483647 - 1;   // NCCL_CONFIG_UNDEF_INT = INT_MIN
constexpr int kNcclCtaPolicyZero = 2;            // NCCL_CTA_POLICY_ZERO (nccl.h:66, NCCL >= 2.28): collectives on the copy engines where possible
constexpr int kNcclWinCollSymmetric = 1;         // NCCL_WIN_COLL_SYMMETRIC (nccl.h:59)
typedef void* NcclWindow;
struct Nccl {
  int (*CommInitRankConfig)(NcclComm*, int, NcclUniqueId, int, NcclConfig*) = nullptr;   // optional (NCCL >= 2.14)
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  // optional: the copy-engine form (NCCL >= 2.28: symmetric windows + zero-CTA all-gather)
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*MemAlloc)(void**, size_t) = nullptr;
  int (*MemFree)(void*) = nullptr;
  int (*CommWindowRegister)(NcclComm, void*, size_t, NcclWindow*, int) = nullptr;
  int (*CommWindowDeregister)(NcclComm, NcclWindow) = nullptr;
  std::string why;   // non-empty: NCCL is not usable
  bool ok = false;
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

const Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    std::string tried;
    if (const char* e = getenv("SAD_NCCL_LIBRARY")) {
      h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
      tried += std::string(e) + " ";
    }
    // the copy already mapped into this process (PyTorch's bundled libnccl.so.2) comes first: one NCCL per process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      n.why = "NCCL not found (tried " + tried + "libnccl.so.2, libnccl.so): " + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (!p && n.why.empty()) n.why = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    n.ok = n.why.empty();
    n.CommInitRankConfig = reinterpret_cast<decltype(n.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
    n.MemAlloc = reinterpret_cast<decltype(n.MemAlloc)>(dlsym(h, "ncclMemAlloc"));
    n.MemFree = reinterpret_cast<decltype(n.MemFree)>(dlsym(h, "ncclMemFree"));
    n.CommWindowRegister = reinterpret_cast<decltype(n.CommWindowRegister)>(dlsym(h, "ncclCommWindowRegister"));
    n.CommWindowDeregister = reinterpret_cast<decltype(n.CommWindowDeregister)>(dlsym(h, "ncclCommWindowDeregister"));
  });
  return n;
}
int nccl_check(int rc, const char* what) {
  if (rc == 0) return SAD_EXCHANGE_OK;
  const Nccl& n = nccl();
  return fail(SAD_EXCHANGE_ERR_NCCL, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

}  // namespace

struct PlannedBucket {
  float* buf;
  size_t count;
  cudaEvent_t ready;   // recorded by an EXTERNAL event-record node of the captured graph
};

struct sad_exchange {
  int rank = 0, world = 1, device = 0;
  int max_ctas = 0;                  // CTA bound of this communicator (0 = NCCL's default)
  // copy-engine form (sad_exchange_create_gather): a symmetric NCCL window of world x capacity floats; the buckets of one step take
  // consecutive regions of it (cursor, reset by the join)
  float* stage = nullptr;
  NcclWindow window = nullptr;
  size_t capacity = 0, cursor = 0;
  uint64_t gathered = 0;             // buckets that went through the copy-engine form
  std::vector<PlannedBucket> plan;   // buckets announced while the producer stream was being captured (see sad_exchange_flush)
  NcclComm comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  std::vector<cudaEvent_t> ready;   // one "bucket is ready" event per bucket in flight (re-used after a join)
  size_t in_flight = 0;             // buckets enqueued since the last join
  cudaEvent_t done = nullptr;
  uint64_t buckets = 0, bytes = 0;
};

namespace {
constexpr size_t kSlotAlign = 128;   // floats: every rank's slot of a bucket starts 512-byte aligned

// The exchange of one bucket, enqueued on the communication stream (which already waits for the bucket's producer).
int enqueue_bucket(sad_exchange* ex, float* buf, size_t count) {
  int rc;
  if (ex->world > 1) {
    const size_t padded = (count + kSlotAlign - 1) / kSlotAlign * kSlotAlign;
    if (ex->stage && ex->cursor + padded <= ex->capacity) {
      // Copy-engine form: this rank's bucket goes into its slot of the symmetric window (a local copy), ncclAllGather (in place;
      // zero-CTA policy: NVLink copies issued by the copy engines, no SMs) fills the other ranks' slots, and one HBM-bound kernel
      // sums the slots in rank order back into the bucket — the same bits on every rank.
      float* region = ex->stage + (size_t)ex->world * ex->cursor;
      float* mine = region + (size_t)ex->rank * padded;
      if ((rc = cuda_check(cudaMemcpyAsync(mine, buf, count * sizeof(float), cudaMemcpyDeviceToDevice, ex->comm_stream), "cudaMemcpyAsync(bucket -> slot)")) != SAD_EXCHANGE_OK) return rc;
      if ((rc = nccl_check(nccl().AllGather(mine, region, padded, kNcclFloat32, ex->comm, ex->comm_stream), "ncclAllGather")) != SAD_EXCHANGE_OK) return rc;
      if ((rc = sad_exchange_slot_sum_f32(region, padded, ex->world, buf, count, ex->comm_stream)) != 0)
        return fail(SAD_EXCHANGE_ERR_CUDA, std::string("slot sum launch: ") + cudaGetErrorString((cudaError_t)rc));
      ex->cursor += padded;
      ex->gathered += 1;
    } else {
      if ((rc = nccl_check(nccl().AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, ex->comm, ex->comm_stream), "ncclAllReduce")) != SAD_EXCHANGE_OK) return rc;
    }
  }
  ex->buckets += 1;
  ex->bytes += (uint64_t)count * sizeof(float);
  return SAD_EXCHANGE_OK;
}
}  // namespace

extern "C" {

SAD_EXPORT const char* sad_exchange_last_error(void) { return g_err.c_str(); }

SAD_EXPORT int sad_exchange_nccl_version(void) {
  const Nccl& n = nccl();
  if (!n.ok) return fail(SAD_EXCHANGE_ERR_NCCL, n.why);
  int v = 0;
  int rc = nccl_check(n.GetVersion(&v), "ncclGetVersion");
  return rc != SAD_EXCHANGE_OK ? rc : v;
}

SAD_EXPORT int sad_exchange_unique_id(void* id_out) {
  if (!id_out) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_unique_id: null id");
  const Nccl& n = nccl();
  if (!n.ok) return fail(SAD_EXCHANGE_ERR_NCCL, n.why);
  return nccl_check(n.GetUniqueId(static_cast<NcclUniqueId*>(id_out)), "ncclGetUniqueId");
}

static int exchange_create_impl(const void* id, int rank, int world, int max_ctas, size_t gather_capacity, sad_exchange** out);
SAD_EXPORT int sad_exchange_create(const void* id, int rank, int world, sad_exchange** out) {
  return exchange_create_impl(id, rank, world, 0, 0, out);
}
SAD_EXPORT int sad_exchange_create_config(const void* id, int rank, int world, int max_ctas, sad_exchange** out) {
  return exchange_create_impl(id, rank, world, max_ctas, 0, out);
}
SAD_EXPORT int sad_exchange_create_gather(const void* id, int rank, int world, size_t capacity_floats, sad_exchange** out) {
  if (capacity_floats == 0) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_create_gather: capacity must be > 0");
  return exchange_create_impl(id, rank, world, 0, capacity_floats, out);
}
SAD_EXPORT int sad_exchange_gather_supported(void) {
  const Nccl& n = nccl();
  if (!n.ok) return 0;
  int v = 0;
  if (n.GetVersion(&v) != 0 || v < 22800) return 0;
  return n.CommInitRankConfig && n.AllGather && n.MemAlloc && n.MemFree && n.CommWindowRegister && n.CommWindowDeregister;
}
SAD_EXPORT int sad_exchange_max_ctas(const sad_exchange* ex) { return ex ? ex->max_ctas : 0; }
SAD_EXPORT size_t sad_exchange_gather_capacity(const sad_exchange* ex) { return ex && ex->stage ? ex->capacity : 0; }
SAD_EXPORT uint64_t sad_exchange_gathered(const sad_exchange* ex) { return ex ? ex->gathered : 0; }
static int exchange_create_impl(const void* id, int rank, int world, int max_ctas, size_t gather_capacity, sad_exchange** out) {
  if (!out) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_create: null out");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_create: rank must be in [0, world)");
  sad_exchange* ex = new (std::nothrow) sad_exchange();
  if (!ex) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_create: out of host memory");
  ex->rank = rank;
  ex->world = world;
  ex->max_ctas = max_ctas;
  int rc = cuda_check(cudaGetDevice(&ex->device), "cudaGetDevice");
  if (rc == SAD_EXCHANGE_OK) {
    int lo = 0, hi = 0;   // the exchange should win the SMs it needs as soon as a bucket is ready: highest stream priority
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    rc = cuda_check(cudaStreamCreateWithPriority(&ex->comm_stream, cudaStreamNonBlocking, hi), "cudaStreamCreateWithPriority");
  }
  if (rc == SAD_EXCHANGE_OK) rc = cuda_check(cudaEventCreateWithFlags(&ex->done, cudaEventDisableTiming), "cudaEventCreate");
  if (rc == SAD_EXCHANGE_OK && world > 1) {
    const Nccl& n = nccl();
    if (!n.ok) rc = fail(SAD_EXCHANGE_ERR_NCCL, n.why);
    else if (!id) rc = fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_create: world > 1 needs the unique id of rank 0");
    else {
      // The exchange has to run BESIDE the backward pass.  An NCCL kernel's CTAs want (almost) a whole SM each; with the library's
      // default CTA count (enough to saturate NVLink on an idle GPU) they do not find room while compute kernels keep every SM
      // partly occupied, and the "overlapped" exchange ends up running after the backward pass (measured on 8 B200s: 9.03 ms with
      // and 9.06 ms without the overlap; with 8 CTAs 0.74 ms of a then 1.40 ms exchange is hidden).  max_ctas bounds the CTAs
      // of THIS communicator only (ncclConfig_t.maxCTAs); 0 keeps NCCL's default.
      int max_ctas = ex->max_ctas;
      if (const char* e = getenv("SAD_EXCHANGE_MAX_CTAS")) max_ctas = atoi(e);
      if (gather_capacity && !sad_exchange_gather_supported())
        rc = fail(SAD_EXCHANGE_ERR_UNSUPPORTED, "sad_exchange_create_gather: needs NCCL >= 2.28 (ncclCommWindowRegister, ncclMemAlloc, zero-CTA policy)");
      else if ((max_ctas > 0 || gather_capacity) && n.CommInitRankConfig) {
        NcclConfig cfg{};
        cfg.magic = 0xcafebeefu;
        cfg.blocking = cfg.cgaClusterSize = cfg.minCTAs = cfg.maxCTAs = cfg.splitShare = cfg.trafficClass = kNcclUndefInt;
        cfg.collnetEnable = cfg.CTAPolicy = cfg.shrinkShare = cfg.nvlsCTAs = kNcclUndefInt;
        cfg.nChannelsPerNetPeer = cfg.nvlinkCentricSched = kNcclUndefInt;
        cfg.netName = cfg.commName = nullptr;
        if (gather_capacity) {   // NCCL 2.28 layout; collectives that can run on the copy engines take no CTAs
          cfg.size = sizeof(NcclConfig);
          cfg.version = 22800;
          cfg.CTAPolicy = kNcclCtaPolicyZero;
        } else {
          cfg.size = kNcclConfigBytesV22703;
          cfg.version = 22703;
        }
        if (max_ctas > 0) {
          cfg.maxCTAs = max_ctas;
          cfg.minCTAs = max_ctas < 4 ? max_ctas : 4;
        }
        rc = nccl_check(n.CommInitRankConfig(&ex->comm, world, *static_cast<const NcclUniqueId*>(id), rank, &cfg), "ncclCommInitRankConfig");
        ex->max_ctas = max_ctas > 0 ? max_ctas : 0;
        if (rc == SAD_EXCHANGE_OK && gather_capacity) {
          // the window: world slots of `capacity` floats, allocated by NCCL (cuMem, shareable over NVLink) and registered
          // symmetrically — every rank registers the same size at the same point, which is what lets the copy engines address it
          ex->capacity = (gather_capacity + kSlotAlign - 1) / kSlotAlign * kSlotAlign;
          size_t bytes = (size_t)world * ex->capacity * sizeof(float);
          bytes = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
          void* p = nullptr;
          rc = nccl_check(n.MemAlloc(&p, bytes), "ncclMemAlloc");
          if (rc == SAD_EXCHANGE_OK) {
            ex->stage = static_cast<float*>(p);
            rc = nccl_check(n.CommWindowRegister(ex->comm, p, bytes, &ex->window, kNcclWinCollSymmetric), "ncclCommWindowRegister");
          }
        }
      } else {
        rc = nccl_check(n.CommInitRank(&ex->comm, world, *static_cast<const NcclUniqueId*>(id), rank), "ncclCommInitRank");
        ex->max_ctas = 0;
      }
    }
  }
  if (rc != SAD_EXCHANGE_OK) {
    const std::string keep = g_err;
    sad_exchange_destroy(ex);
    g_err = keep;
    return rc;
  }
  *out = ex;
  return SAD_EXCHANGE_OK;
}

SAD_EXPORT void sad_exchange_destroy(sad_exchange* ex) {
  if (!ex) return;
  if (ex->comm_stream) cudaStreamSynchronize(ex->comm_stream);
  if (ex->window && ex->comm && nccl().CommWindowDeregister) nccl().CommWindowDeregister(ex->comm, ex->window);
  if (ex->stage && nccl().MemFree) nccl().MemFree(ex->stage);
  if (ex->comm && nccl().CommDestroy) nccl().CommDestroy(ex->comm);
  for (cudaEvent_t e : ex->ready) cudaEventDestroy(e);
  for (auto& b : ex->plan) cudaEventDestroy(b.ready);
  if (ex->done) cudaEventDestroy(ex->done);
  if (ex->comm_stream) cudaStreamDestroy(ex->comm_stream);
  delete ex;
}

SAD_EXPORT int sad_exchange_world(const sad_exchange* ex) { return ex ? ex->world : 0; }
SAD_EXPORT int sad_exchange_rank(const sad_exchange* ex) { return ex ? ex->rank : -1; }
SAD_EXPORT uint64_t sad_exchange_buckets(const sad_exchange* ex) { return ex ? ex->buckets : 0; }
SAD_EXPORT uint64_t sad_exchange_bytes(const sad_exchange* ex) { return ex ? ex->bytes : 0; }

SAD_EXPORT int sad_exchange_allreduce_async_f32(sad_exchange* ex, float* buf, size_t count, void* producer_stream) {
  if (!ex || (!buf && count)) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_allreduce_async_f32: null argument");
  if (count == 0) return SAD_EXCHANGE_OK;
  int rc;
  cudaStream_t ps = static_cast<cudaStream_t>(producer_stream);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if ((rc = cuda_check(cudaStreamIsCapturing(ps, &cap), "cudaStreamIsCapturing")) != SAD_EXCHANGE_OK) return rc;
  if (cap == cudaStreamCaptureStatusActive) {
    // The producer stream is being captured into a CUDA graph.  The collective itself stays OUT of the graph: the graph only gets
    // an external event-record node ("this bucket is final"), and sad_exchange_flush — called after every launch of the graph —
    // makes the communication stream wait for that event and enqueues the allreduce eagerly.  (Capturing the NCCL kernels as a
    // parallel branch of the graph works — scripts/exchange_check.py — but measured on 8 B200s the branch did not overlap the
    // backward pass at all: 9.08 ms with and without it.  An eager, high-priority stream beside the graph does.)
    cudaEvent_t e = nullptr;
    if ((rc = cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) != SAD_EXCHANGE_OK) return rc;
    if ((rc = cuda_check(cudaEventRecordWithFlags(e, ps, cudaEventRecordExternal), "cudaEventRecordWithFlags(external)")) != SAD_EXCHANGE_OK) {
      cudaEventDestroy(e);
      return rc;
    }
    ex->plan.push_back(PlannedBucket{buf, count, e});
    return SAD_EXCHANGE_OK;
  }
  if (ex->in_flight == ex->ready.size()) {
    cudaEvent_t e = nullptr;
    if ((rc = cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) != SAD_EXCHANGE_OK) return rc;
    ex->ready.push_back(e);
  }
  cudaEvent_t ev = ex->ready[ex->in_flight++];
  // the bucket is complete once everything enqueued so far on the producer stream has run (cuda_nccl_gpu.cc:157-166)
  if ((rc = cuda_check(cudaEventRecord(ev, ps), "cudaEventRecord(bucket ready)")) != SAD_EXCHANGE_OK) return rc;
  if ((rc = cuda_check(cudaStreamWaitEvent(ex->comm_stream, ev, 0), "cudaStreamWaitEvent(comm stream)")) != SAD_EXCHANGE_OK) return rc;
  return enqueue_bucket(ex, buf, count);
}

SAD_EXPORT int sad_exchange_plan_reset(sad_exchange* ex) {
  if (!ex) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_plan_reset: null exchange");
  for (auto& b : ex->plan) cudaEventDestroy(b.ready);
  ex->plan.clear();
  return SAD_EXCHANGE_OK;
}

SAD_EXPORT int sad_exchange_planned(const sad_exchange* ex) { return ex ? (int)ex->plan.size() : 0; }

SAD_EXPORT int sad_exchange_flush(sad_exchange* ex) {
  if (!ex) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_flush: null exchange");
  int rc;
  for (auto& b : ex->plan) {
    // waits for the record node of the graph launch that precedes this call
    if ((rc = cuda_check(cudaStreamWaitEvent(ex->comm_stream, b.ready, 0), "cudaStreamWaitEvent(planned bucket)")) != SAD_EXCHANGE_OK) return rc;
    if ((rc = enqueue_bucket(ex, b.buf, b.count)) != SAD_EXCHANGE_OK) return rc;
  }
  if (!ex->plan.empty()) ex->in_flight += 1;   // something for the next join to wait for
  return SAD_EXCHANGE_OK;
}

SAD_EXPORT int sad_exchange_join(sad_exchange* ex, void* consumer_stream) {
  if (!ex) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_join: null exchange");
  if (ex->in_flight == 0) return SAD_EXCHANGE_OK;
  int rc;
  // the consumer waits for the communication stream (cuda_nccl_gpu.cc:196-205)
  if ((rc = cuda_check(cudaEventRecord(ex->done, ex->comm_stream), "cudaEventRecord(exchange done)")) != SAD_EXCHANGE_OK) return rc;
  if ((rc = cuda_check(cudaStreamWaitEvent(static_cast<cudaStream_t>(consumer_stream), ex->done, 0), "cudaStreamWaitEvent(consumer)")) != SAD_EXCHANGE_OK)
    return rc;
  ex->in_flight = 0;
  ex->cursor = 0;   // the next step's buckets reuse the window from its start (ordered behind this step's sums on the communication stream)
  return SAD_EXCHANGE_OK;
}

SAD_EXPORT int sad_exchange_allreduce_f32(sad_exchange* ex, float* buf, size_t count, void* stream) {
  if (!ex || (!buf && count)) return fail(SAD_EXCHANGE_ERR_INVALID, "sad_exchange_allreduce_f32: null argument");
  if (count == 0 || ex->world == 1) return SAD_EXCHANGE_OK;
  int rc = nccl_check(nccl().AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, ex->comm, static_cast<cudaStream_t>(stream)), "ncclAllReduce");
  if (rc == SAD_EXCHANGE_OK) {
    ex->buckets += 1;
    ex->bytes += (uint64_t)count * sizeof(float);
  }
  return rc;
}

}  // extern "C"
