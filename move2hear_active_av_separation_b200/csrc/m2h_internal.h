// Internal helpers shared by the translation units of libm2h_b200.so (not part of the ABI).
#pragma once
#include <stdarg.h>
#include <cuda_runtime.h>

#define M2H_MAX_DEVICES 64

void m2h_set_error(const char* fmt, ...);
int m2h_current_device_checked(int* dev);
int m2h_num_sms(int dev);

#define M2H_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t m2h_e_ = (expr);                                                              \
    if (m2h_e_ != cudaSuccess) {                                                              \
      m2h_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(m2h_e_), __FILE__, __LINE__); \
      return M2H_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)
