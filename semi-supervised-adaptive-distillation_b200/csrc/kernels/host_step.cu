// sad_distill_step_host — the whole distillation-loss step for a caller whose tensors live in HOST
// memory (the end-to-end number bench.py reports).  Copies and kernels are pipelined on three
// streams so PCIe runs in both directions while the kernels execute:
//
//   copy-in  stream : T(all levels) | X,G chunk 0 | X,G chunk 1 | ...
//   compute  stream :               PowSum        | loss+grad 0 | loss+grad 1 | ...
//   copy-out stream :                                           | dX chunk 0  | dX chunk 1 ... | losses
//
// A chunk is a run of whole ANCHORS of one image of one level: the (A*C, H, W) block of an image is contiguous and the
// kernel's label indexing (loss_op.cu:35-42: t = gt[n*H*W*A + a*H*W + y*W + x]) only needs the label base moved by
// a0*H*W, so anchors [a0, a0 + k) of image n are a valid (N = 1, D = k*C, H, W) level of their own.  Chunks are capped at
// 16 MB (SAD_HOST_CHUNK_BYTES / sad_ctx_set_host_chunk_bytes) and ordered largest first so the un-overlapped D2H tail is the
// smallest chunk.  The normaliser needs every teacher probability (PowSum runs over all levels, reference
// retinanet_heads.py:320-328), hence T (and the small label tensors) go first and the critical path is
//   H2D(T) + H2D(first X chunk) + kernel + D2H(all dX).
// Measured at configs[1] (profiles/r01l_e2e_chunk_sweep.json): whole images 3.64 ms, 16 MB 3.56 ms, 8 MB 3.67, 4 MB 3.96, 1 MB 4.21 —
// the copies of both directions contend on this host (236.7 MB in 3.56 ms = 66 GB/s aggregate) and every extra chunk costs
// launches, so the cap stays coarse.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "sad_b200.h"
#include "sad_internal.h"

using namespace sad;

struct sad_ctx {
  int device = 0;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t ev_T = nullptr, ev_done = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_k;
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  Buf X[SAD_MAX_LEVELS], T[SAD_MAX_LEVELS], G[SAD_MAX_LEVELS], dX[SAD_MAX_LEVELS];
  Buf ws_pow, ws_dist, scalars;  // scalars: [0] normaliser, [1..] per-chunk losses
  float* h_scalars = nullptr;    // pinned mirror of `scalars`
  size_t h_scalars_cap = 0;
  size_t chunk_bytes = 0;        // 0 = default (SAD_HOST_CHUNK_BYTES or 4 MB)
};

namespace {
int grow(sad_ctx::Buf& b, size_t bytes, bool is_workspace, cudaStream_t st) {
  if (bytes <= b.cap) return SAD_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  int rc = check_cuda(cudaMalloc(&b.p, bytes), "cudaMalloc");
  if (rc != SAD_OK) return rc;
  b.cap = bytes;
  if (is_workspace) return sad_workspace_init(b.p, bytes, st);
  return SAD_OK;
}
struct Chunk {
  int level, n, a0, k;   // anchors [a0, a0 + k) of image n of `level`
  size_t elems;          // k * num_classes * H * W
};
size_t host_chunk_bytes() {
  static const size_t v = [] {
    const char* e = std::getenv("SAD_HOST_CHUNK_BYTES");
    const long long x = e ? std::atoll(e) : 0;
    return x > 0 ? (size_t)x : (size_t)16 << 20;
  }();
  return v;
}
}  // namespace

extern "C" {

SAD_EXPORT int sad_ctx_create(int device, sad_ctx** out) {
  if (!out) return set_error(SAD_ERR_INVALID, "sad_ctx_create: null out");
  int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice");
  if (rc != SAD_OK) return rc;
  sad_ctx* c = new (std::nothrow) sad_ctx();
  if (!c) return set_error(SAD_ERR_INVALID, "out of host memory");
  c->device = device;
  if ((rc = check_cuda(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking), "stream")) != SAD_OK ||
      (rc = check_cuda(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking), "stream")) != SAD_OK ||
      (rc = check_cuda(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking), "stream")) != SAD_OK ||
      (rc = check_cuda(cudaEventCreateWithFlags(&c->ev_T, cudaEventDisableTiming), "event")) != SAD_OK ||
      (rc = check_cuda(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming), "event")) != SAD_OK) {
    sad_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return SAD_OK;
}

SAD_EXPORT void sad_ctx_destroy(sad_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int l = 0; l < SAD_MAX_LEVELS; ++l) {
    cudaFree(c->X[l].p); cudaFree(c->T[l].p); cudaFree(c->G[l].p); cudaFree(c->dX[l].p);
  }
  cudaFree(c->ws_pow.p); cudaFree(c->ws_dist.p); cudaFree(c->scalars.p);
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  for (auto e : c->ev_in) cudaEventDestroy(e);
  for (auto e : c->ev_k) cudaEventDestroy(e);
  if (c->ev_T) cudaEventDestroy(c->ev_T);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_k) cudaStreamDestroy(c->s_k);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
}

SAD_EXPORT int sad_ctx_set_host_chunk_bytes(sad_ctx* c, size_t bytes) {
  if (!c) return set_error(SAD_ERR_INVALID, "sad_ctx_set_host_chunk_bytes: null context");
  c->chunk_bytes = bytes;
  return SAD_OK;
}

SAD_EXPORT float* sad_ctx_device_d_logits(sad_ctx* c, int level) {
  if (!c || level < 0 || level >= SAD_MAX_LEVELS) return nullptr;
  return static_cast<float*>(c->dX[level].p);
}

SAD_EXPORT int sad_distill_step_host(sad_ctx* c, const sad_host_level* levels, int n_levels, float power,
                                     const sad_distill_params* params, float* losses_out, float* normalizer_out) {
  if (!c || !levels || !params || !losses_out) return set_error(SAD_ERR_INVALID, "sad_distill_step_host: null argument");
  if (n_levels < 1 || n_levels > SAD_MAX_LEVELS) return set_error(SAD_ERR_INVALID, "sad_distill_step_host: bad n_levels");
  if (params->num_classes < 1) return set_error(SAD_ERR_INVALID, "num_classes must be >= 1");
  int rc = check_cuda(cudaSetDevice(c->device), "cudaSetDevice");
  if (rc != SAD_OK) return rc;

  // ---- plan: buffers and chunks -------------------------------------------------------------
  std::vector<Chunk> chunks;
  int64_t sizes[SAD_MAX_LEVELS];
  const float* t_dev[SAD_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) {
    const sad_host_level& L = levels[l];
    if (L.N < 0 || L.D < 0 || L.H < 0 || L.W < 0 || L.D % params->num_classes)
      return set_error(SAD_ERR_INVALID, "sad_distill_step_host: bad level shape");
    const size_t per_img = (size_t)L.D * L.H * L.W;
    const size_t elems = per_img * L.N;
    const size_t lab = (size_t)L.N * (L.D / params->num_classes) * L.H * L.W;
    if (elems && (!L.logits || !L.teacher_prob || !L.labels)) return set_error(SAD_ERR_INVALID, "null host input");
    if ((rc = grow(c->X[l], elems * 4, false, c->s_k)) != SAD_OK || (rc = grow(c->T[l], elems * 4, false, c->s_k)) != SAD_OK ||
        (rc = grow(c->G[l], lab * 4, false, c->s_k)) != SAD_OK || (rc = grow(c->dX[l], elems * 4, false, c->s_k)) != SAD_OK)
      return rc;
    sizes[l] = (int64_t)elems;
    t_dev[l] = static_cast<const float*>(c->T[l].p);
    const int A = L.D / params->num_classes;
    const size_t per_anchor = (size_t)params->num_classes * L.H * L.W;
    if (per_img) {
      // anchors per chunk: as many as fit the byte cap, spread evenly (9 anchors, cap 2 -> 2,2,2,2,1 becomes 5 chunks of <= 2)
      int k_cap = (int)std::max<size_t>(1, (c->chunk_bytes ? c->chunk_bytes : host_chunk_bytes()) / (per_anchor * 4));
      if (k_cap > A) k_cap = A;
      const int pieces = (A + k_cap - 1) / k_cap;
      for (int n = 0; n < L.N; ++n)
        for (int pc = 0; pc < pieces; ++pc) {
          const int a0 = (int)(((int64_t)A * pc) / pieces), a1 = (int)(((int64_t)A * (pc + 1)) / pieces);
          if (a1 > a0) chunks.push_back({l, n, a0, a1 - a0, (size_t)(a1 - a0) * per_anchor});
        }
    }
  }
  std::stable_sort(chunks.begin(), chunks.end(), [](const Chunk& a, const Chunk& b) { return a.elems > b.elems; });
  const size_t n_chunks = chunks.size();
  while (c->ev_in.size() < n_chunks) {
    cudaEvent_t e1, e2;
    if ((rc = check_cuda(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming), "event")) != SAD_OK) return rc;
    c->ev_in.push_back(e1);
    if ((rc = check_cuda(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming), "event")) != SAD_OK) return rc;
    c->ev_k.push_back(e2);
  }
  if ((rc = grow(c->scalars, (1 + n_chunks) * sizeof(float) + 64, false, c->s_k)) != SAD_OK) return rc;
  if (c->h_scalars_cap < 1 + n_chunks) {
    if (c->h_scalars) cudaFreeHost(c->h_scalars);
    c->h_scalars = nullptr;
    if ((rc = check_cuda(cudaMallocHost(&c->h_scalars, (1 + n_chunks + 16) * sizeof(float)), "cudaMallocHost")) != SAD_OK) return rc;
    c->h_scalars_cap = 1 + n_chunks + 16;
  }
  float* d_norm = static_cast<float*>(c->scalars.p);
  float* d_loss = d_norm + 1;
  if ((rc = grow(c->ws_pow, sad_pow_sum_workspace_bytes(sizes, n_levels), true, c->s_k)) != SAD_OK) return rc;
  {
    // the largest chunk bounds the distill workspace
    size_t need = 256;
    for (const Chunk& ch : chunks) {
      sad_distill_level one{};
      one.N = 1; one.D = ch.k * params->num_classes; one.H = levels[ch.level].H; one.W = levels[ch.level].W;
      need = std::max(need, sad_distill_workspace_bytes(&one, 1));
    }
    if ((rc = grow(c->ws_dist, need, true, c->s_k)) != SAD_OK) return rc;
  }

  // ---- copy-in: teacher probabilities first, PowSum as soon as they have landed -------------
  for (int l = 0; l < n_levels; ++l)
    if (sizes[l] && (rc = check_cuda(cudaMemcpyAsync(c->T[l].p, levels[l].teacher_prob, (size_t)sizes[l] * 4, cudaMemcpyHostToDevice, c->s_in), "H2D T")) != SAD_OK)
      return rc;
  cudaEventRecord(c->ev_T, c->s_in);
  cudaStreamWaitEvent(c->s_k, c->ev_T, 0);
  // labels: 1/80 of the logits' bytes; one copy per level here instead of one small copy per chunk
  for (int l = 0; l < n_levels; ++l) {
    const sad_host_level& L = levels[l];
    const size_t lab = (size_t)L.N * (L.D / params->num_classes) * L.H * L.W;
    if (lab && (rc = check_cuda(cudaMemcpyAsync(c->G[l].p, L.labels, lab * 4, cudaMemcpyHostToDevice, c->s_in), "H2D G")) != SAD_OK) return rc;
  }
  if ((rc = sad_pow_sum_f32(t_dev, sizes, n_levels, power, d_norm, c->ws_pow.p, c->ws_pow.cap, c->s_k)) != SAD_OK) return rc;

  // ---- per chunk: H2D X,G -> loss+grad -> D2H dX --------------------------------------------
  for (size_t i = 0; i < n_chunks; ++i) {
    const Chunk& ch = chunks[i];
    const sad_host_level& L = levels[ch.level];
    const size_t hw = (size_t)L.H * L.W;
    const size_t xoff = ((size_t)ch.n * L.D + (size_t)ch.a0 * params->num_classes) * hw;
    const size_t lab_per_img = (size_t)(L.D / params->num_classes) * hw;
    const size_t goff = (size_t)ch.n * lab_per_img + (size_t)ch.a0 * hw;
    float* dXd = static_cast<float*>(c->dX[ch.level].p) + xoff;
    if ((rc = check_cuda(cudaMemcpyAsync(static_cast<float*>(c->X[ch.level].p) + xoff, L.logits + xoff, ch.elems * 4, cudaMemcpyHostToDevice, c->s_in), "H2D X")) != SAD_OK)
      return rc;
    cudaEventRecord(c->ev_in[i], c->s_in);
    cudaStreamWaitEvent(c->s_k, c->ev_in[i], 0);
    sad_distill_level one{};
    one.logits = static_cast<const float*>(c->X[ch.level].p) + xoff;
    one.teacher_prob = static_cast<const float*>(c->T[ch.level].p) + xoff;
    one.labels = static_cast<const int32_t*>(c->G[ch.level].p) + goff;
    one.d_logits = dXd;
    one.loss = d_loss + i;
    one.d_loss = nullptr;
    one.N = 1; one.D = ch.k * params->num_classes; one.H = L.H; one.W = L.W;
    if ((rc = sad_distill_f32(&one, 1, d_norm, params, c->ws_dist.p, c->ws_dist.cap, c->s_k)) != SAD_OK) return rc;
    cudaEventRecord(c->ev_k[i], c->s_k);
    if (L.d_logits) {
      cudaStreamWaitEvent(c->s_out, c->ev_k[i], 0);
      if ((rc = check_cuda(cudaMemcpyAsync(L.d_logits + xoff, dXd, ch.elems * 4, cudaMemcpyDeviceToHost, c->s_out), "D2H dX")) != SAD_OK) return rc;
    }
  }
  // ---- scalars back, then wait for everything ----------------------------------------------
  cudaEventRecord(c->ev_done, c->s_k);
  cudaStreamWaitEvent(c->s_out, c->ev_done, 0);
  if ((rc = check_cuda(cudaMemcpyAsync(c->h_scalars, c->scalars.p, (1 + n_chunks) * sizeof(float), cudaMemcpyDeviceToHost, c->s_out), "D2H scalars")) != SAD_OK) return rc;
  if ((rc = check_cuda(cudaStreamSynchronize(c->s_out), "sync")) != SAD_OK) return rc;
  if ((rc = check_cuda(cudaStreamSynchronize(c->s_in), "sync")) != SAD_OK) return rc;

  for (int l = 0; l < n_levels; ++l) losses_out[l] = 0.f;
  for (size_t i = 0; i < n_chunks; ++i) losses_out[chunks[i].level] += c->h_scalars[1 + i];  // fixed order: (image, anchor run) ascending per level
  if (normalizer_out) *normalizer_out = c->h_scalars[0];
  return SAD_OK;
}

}  // extern "C"
