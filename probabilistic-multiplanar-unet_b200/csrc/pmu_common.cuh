// Shared helpers for the pmu_b200 kernels: error plumbing, launch checks, small PTX wrappers.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <algorithm>

#include "../../include/pmu_b200.h"

namespace pmu {

// thread-local last-error message (pmu_last_error)
void set_error(const char* fmt, ...);

#define PMU_CHECK_ARG(cond, ...)                                                      \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::pmu::set_error(__VA_ARGS__);                                                  \
      return PMU_ERR_INVALID;                                                         \
    }                                                                                 \
  } while (0)

#define PMU_CHECK_SUPPORTED(cond, ...)                                                \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::pmu::set_error(__VA_ARGS__);                                                  \
      return PMU_ERR_UNSUPPORTED;                                                     \
    }                                                                                 \
  } while (0)

#define PMU_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::pmu::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return PMU_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

// after a <<<>>> launch
#define PMU_LAUNCH_CHECK() PMU_CUDA(cudaGetLastError())

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();  // cached multiProcessorCount of the current device

// ---- device helpers --------------------------------------------------------

// Order-preserving float <-> int mapping so atomicMax on ints implements float max
// (handles negatives; -inf maps to the smallest key we ever see).
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return (i >= 0) ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float((i >= 0) ? i : (i ^ 0x7fffffff));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // The buffer holds plain floats (pre-filled with -inf); CAS-free trick: positive floats
  // order like ints, negative floats order reversed like unsigned ints.
  if (v >= 0.f) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  } else {
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// reference normalisation: (float)((double)x / (double)m) if m != 0   (mri_dataset.py:109-110,142).
// x and m are fp32 values, so the fp64 quotient rounded to fp32 equals the correctly rounded
// fp32 quotient: double rounding is innocuous for division when the wide format has at least
// 2p+2 = 50 significand bits (fp64 has 53).  IEEE div.rn.f32 is therefore bit-identical to the
// reference's fp64 divide + .float(), without touching the fp64 pipe.
__device__ __forceinline__ float ref_normalise(float x, float m) {
  return (m != 0.f) ? __fdiv_rn(x, m) : x;
}

// streaming 128-bit global access (read-once / write-once data; keep L1 for reused lines)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace pmu
