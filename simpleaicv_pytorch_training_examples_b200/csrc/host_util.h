// Error reporting shared by the translation units of libsaicv_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace saicv {
// Formats into a thread-local buffer returned by saicv_last_error(); always returns 1.
int set_error(const char* fmt, ...);
int check_launch(const char* what);
// Number of kernels this library has launched in this process (bench.py's gpu_launches).
void count_launch(int n);
}  // namespace saicv
