// Shared host/device helpers of libideas_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ideas_b200.h"

namespace ideas {

// thread-local last error (ideas_last_error); defined in elementwise.cu
void set_error(const char* fmt, ...);
void count_launch();

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return IDEAS_ERR_CUDA;
}

#define IDEAS_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ideas::set_error(__VA_ARGS__);        \
      return IDEAS_ERR_INVALID;             \
    }                                       \
  } while (0)

// every kernel launch of the library goes through this macro: it counts the launch
// (ideas_launch_count, reported by bench.py as gpu_launches) and checks for a launch error
#define IDEAS_CHECK_LAUNCH(name)                                  \
  do {                                                            \
    ideas::count_launch();                                        \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return ideas::cuda_fail(e__, name);   \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Streaming (read-once / write-once) 128-bit accesses: keep L1 for data that is re-read.
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float lrelu(float v, float alpha) { return v > 0.f ? v : v * alpha; }

}  // namespace ideas
