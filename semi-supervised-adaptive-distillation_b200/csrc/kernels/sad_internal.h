// Internal helpers shared by the translation units of libsad_b200.so (not part of the ABI).
#ifndef SAD_INTERNAL_H_
#define SAD_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "sad_b200.h"

#define SAD_EXPORT __attribute__((visibility("default")))

namespace sad {
int set_error(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
void count_launch(uint64_t n);

// persistent bulk-copy-ring kernels (distill_ring.cu): the production path for 16-byte aligned tensors
constexpr int kMaxRingCtas = 1024;  // upper bound of a ring grid (2 CTAs x #SMs); sizes the partial-sum scratch
bool distill_ring_supported(const sad_distill_level* levels, int n_levels, int num_classes);
int launch_distill_ring(const sad_distill_level* levels, int n_levels, const float* normalizer, const sad_distill_params* p,
                        void* workspace, size_t workspace_bytes, cudaStream_t st);
// one cooperative launch for PowSum + loss + gradient (distill_fused.cu)
bool distill_fused_supported(const sad_distill_level* levels, int n_levels, const sad_distill_params* p, float power);
size_t distill_fused_workspace_bytes(const sad_distill_level* levels, int n_levels, int num_classes);
int launch_distill_fused(const sad_distill_level* levels, int n_levels, float power, float* norm_out, const sad_distill_params* p,
                         void* workspace, size_t workspace_bytes, cudaStream_t st);
bool pow_sum_ring_supported(const float* const* inputs, const int64_t* sizes, int n_inputs);
int launch_pow_sum_ring(const float* const* inputs, const int64_t* sizes, int n_inputs, float power, float* out,
                        void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace sad

#endif
