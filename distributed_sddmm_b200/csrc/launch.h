// launch.h -- helpers shared by the translation units of libhnh_b200.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cuda_runtime.h>

#define HNH_STR2(x) #x
#define HNH_STR(x) HNH_STR2(x)

namespace hnh {
// Record a failure description for hnh_last_error_string() and return `code`.
int set_error(int code, const char *fmt, ...);
// HNH_OK, or HNH_E_CUDA with the CUDA error text recorded.
int check_cuda(cudaError_t e, const char *what);
// Add to the process-wide kernel-launch counter (hnh_launch_count()).
void count_launch(uint64_t n);
// Number of SMs of the current device (cached).
int device_sm_count(int *out);
}  // namespace hnh
