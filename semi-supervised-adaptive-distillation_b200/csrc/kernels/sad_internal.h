// Internal helpers shared by the translation units of libsad_b200.so (not part of the ABI).
#ifndef SAD_INTERNAL_H_
#define SAD_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#define SAD_EXPORT __attribute__((visibility("default")))

namespace sad {
int set_error(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
void count_launch(uint64_t n);
}  // namespace sad

#endif
