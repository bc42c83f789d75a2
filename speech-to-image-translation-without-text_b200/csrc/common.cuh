// Shared host-side error reporting for the C ABI (sg2_last_error()).
#pragma once
#include <cstdio>
#include <cuda_runtime.h>

namespace sg2 {
extern thread_local char g_err[512];
}
#define SG2_FAIL(code, ...)                                 \
  do {                                                      \
    snprintf(::sg2::g_err, sizeof(::sg2::g_err), __VA_ARGS__); \
    return (code);                                          \
  } while (0)
#define SG2_LAUNCH_OK(what)                                                               \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) SG2_FAIL((int)e__, "%s launch: %s", what, cudaGetErrorString(e__)); \
    return 0;                                                                             \
  } while (0)
