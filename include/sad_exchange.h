/* sad_exchange.h — C ABI of the gradient exchange of the data-parallel distillation step (libsad_exchange.so).
 *
 * Host C++ over NCCL (NVLink 5 / NVSwitch inside one B200 box): one communicator per process (one process per GPU), a
 * dedicated communication stream, and BUCKETED SUM-allreduces that are ordered by CUDA events against the stream that
 * produces the gradients, so a bucket's exchange runs while the backward pass is still producing the next bucket and the
 * optimiser step waits only for the join.  Capturable into a CUDA graph together with the step.
 *
 * Replaces, in the reference,
 *   detectron/lib/modeling/optimizer.py:72-92          _add_allreduce_graph: one NCCLAllreduce operator per parameter-gradient blob
 *   caffe2/caffe2/contrib/nccl/cuda_nccl_op_gpu.cc:80-121   NCCLAllreduceOp<T>::RunOnDevice
 *   caffe2/caffe2/contrib/nccl/cuda_nccl_gpu.cc:139-225     runNCCL / NCCL<T>::AllReduce: stream + event plumbing around ncclAllReduce
 * (single process driving all GPUs there; one process per GPU here, the same in-place SUM over the same blobs, issued per
 * contiguous bucket of the flat gradient buffer instead of per blob).
 *
 * NCCL itself is resolved at run time (dlopen of the libnccl.so.2 already in the process, e.g. PyTorch's, else the system's;
 * SAD_NCCL_LIBRARY overrides) so the library loads on hosts without NCCL or a GPU; every call then fails with a clear error.
 * All functions return 0 on success, a negative SAD_EXCHANGE_ERR_* otherwise; sad_exchange_last_error() describes it.
 */
#ifndef SAD_EXCHANGE_H_
#define SAD_EXCHANGE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAD_EXCHANGE_OK 0
#define SAD_EXCHANGE_ERR_INVALID (-1)
#define SAD_EXCHANGE_ERR_CUDA (-2)
#define SAD_EXCHANGE_ERR_NCCL (-3)      /* NCCL missing or an NCCL call failed */
#define SAD_EXCHANGE_ERR_UNSUPPORTED (-4) /* the resolved NCCL lacks what the requested form needs */
#define SAD_EXCHANGE_UNIQUE_ID_BYTES 128 /* sizeof(ncclUniqueId) */

typedef struct sad_exchange sad_exchange;

const char* sad_exchange_last_error(void);
/* NCCL_VERSION_CODE of the library that was resolved (e.g. 22809), or a negative error */
int sad_exchange_nccl_version(void);
/* rank 0 creates the rendezvous id (ncclGetUniqueId) and hands it to every rank by any host channel */
int sad_exchange_unique_id(void* id_out /* SAD_EXCHANGE_UNIQUE_ID_BYTES */);
/* collective over all ranks: joins the communicator on the CURRENT CUDA device.  world == 1 needs neither NCCL nor an id. */
int sad_exchange_create(const void* id, int rank, int world, sad_exchange** out);
/* The same with a bound on the CTAs this communicator's kernels may use (ncclConfig_t.maxCTAs; 0 = NCCL's default; the environment
 * variable SAD_EXCHANGE_MAX_CTAS overrides).  An exchange that has to overlap a backward pass needs kernels small enough to find room
 * beside the compute kernels. */
int sad_exchange_create_config(const void* id, int rank, int world, int max_ctas, sad_exchange** out);
int sad_exchange_max_ctas(const sad_exchange* ex);
/* The COPY-ENGINE form of the same exchange (NCCL >= 2.28; sad_exchange_gather_supported() says whether the resolved library has it).
 * Why: an NCCL allreduce kernel holds whole SMs for its whole duration, and beside a backward pass whose kernels are sized for all 148
 * SMs every held SM turns one wave into two — measured on 8 B200s the overlapped allreduce hid nothing.  Here the
 * communicator is created with the zero-CTA policy and owns a symmetric NCCL window of world x capacity_floats; a bucket is copied into
 * this rank's slot, ncclAllGather moves the slots over NVLink with the copy engines (no SMs), and ONE short HBM-bound kernel
 * (sad_exchange_slot_sum_f32) adds the world slots in rank order back into the bucket: the result is the rank-ordered fp32 sum, bit-identical
 * on every rank.  Same calls afterwards (allreduce_async / flush / join); a bucket that does not fit what is left of the window in
 * the current step falls back to ncclAllReduce.  capacity_floats = the most floats one step exchanges (+ 128 per bucket of padding).
 * MEASURED (profiles/r02q_exchange_copy_engine.json): correct at 2 and 8 GPUs but slower than the ncclAllReduce
 * buckets (an all-gather moves (world - 1) x the bytes through every rank's HBM) — opt-in, not the default. */
int sad_exchange_gather_supported(void);
int sad_exchange_create_gather(const void* id, int rank, int world, size_t capacity_floats, sad_exchange** out);
size_t sad_exchange_gather_capacity(const sad_exchange* ex);   /* 0 = this exchange runs the ncclAllReduce form */
uint64_t sad_exchange_gathered(const sad_exchange* ex);         /* buckets that took the copy-engine form since creation */
/* out[i] = slots[i] + slots[stride + i] + ... + slots[(world-1) * stride + i], fp32, added in that order; returns a cudaError_t */
int sad_exchange_slot_sum_f32(const float* slots, size_t stride, int world, float* out, size_t count, void* stream);
void sad_exchange_destroy(sad_exchange* ex);
int sad_exchange_world(const sad_exchange* ex);
int sad_exchange_rank(const sad_exchange* ex);

/* In-place SUM-allreduce of buf[0 .. count) (fp32) on the exchange's communication stream, ordered AFTER everything enqueued
 * so far on producer_stream (cudaStream_t as void*; NULL = legacy default stream).  Returns at once; neither the host nor
 * producer_stream waits.  Several buckets may be in flight; NCCL runs them in issue order. */
int sad_exchange_allreduce_async_f32(sad_exchange* ex, float* buf, size_t count, void* producer_stream);
/* CUDA graphs.  When producer_stream is being CAPTURED, sad_exchange_allreduce_async_f32 does not enqueue the collective: it adds
 * an external event-record node to the graph ("this bucket is final") and remembers the bucket.  After EVERY launch of that graph
 * call sad_exchange_flush: the communication stream then waits for each bucket's event of that launch and runs the allreduce
 * eagerly, beside the graph.  sad_exchange_plan_reset forgets the remembered buckets (call it before capturing again). */
int sad_exchange_flush(sad_exchange* ex);
int sad_exchange_plan_reset(sad_exchange* ex);
int sad_exchange_planned(const sad_exchange* ex);
/* consumer_stream waits (device side) for every bucket enqueued since the previous join */
int sad_exchange_join(sad_exchange* ex, void* consumer_stream);
/* the un-overlapped form: the allreduce is enqueued on `stream` itself */
int sad_exchange_allreduce_f32(sad_exchange* ex, float* buf, size_t count, void* stream);
/* buckets enqueued / bytes reduced since creation (bench.py reports them) */
uint64_t sad_exchange_buckets(const sad_exchange* ex);
uint64_t sad_exchange_bytes(const sad_exchange* ex);

#ifdef __cplusplus
}
#endif
#endif /* SAD_EXCHANGE_H_ */
