// The two 16-bit storage formats of the tensor-core mode.  F16 = true: IEEE half — the INFERENCE format (11 significand
// bits: bf16 operands put the worst pixel of a single view's N-sample mean at 2.0-2.3e-2 on this 22-layer network,
// outside the 2e-2 bound; f16 operands: ~3e-3, tests/tools/emulate_bf16_net.py).  F16 = false: bfloat16 — the TRAINING
// format (gradients need the exponent range).  tcgen05 kind::f16 runs both at the same rate with fp32 accumulation.
// Conversions to f16 saturate (cvt.satfinite: +-65504 instead of inf), so an out-of-range activation cannot poison the
// network with infinities.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pmu {

template <bool F16>
__device__ __forceinline__ uint32_t pack16_rn(float lo, float hi) {
  uint32_t d;
  if constexpr (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi) {
  uint32_t d;
  if constexpr (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// (a0 + b0, a1 + b1) -> [relu] -> packed pair (low half = first element): add.f32x2 + one cvt
template <bool RELU, bool F16>
__device__ __forceinline__ uint32_t add_pack16(float a0, float a1, float b0, float b1) {
  float lo, hi;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(lo), "=f"(hi) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
  return RELU ? pack16_relu<F16>(lo, hi) : pack16_rn<F16>(lo, hi);
}
template <bool F16>
__device__ __forceinline__ float2 unpack16(uint32_t v) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}
template <bool F16>
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
  if constexpr (F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}
template <bool F16>
__device__ __forceinline__ uint16_t cvt16(float v) {
  if constexpr (F16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h); }
  else { const __nv_bfloat16 h = __float2bfloat16(v); return *reinterpret_cast<const uint16_t*>(&h); }
}
template <bool F16>
__device__ __forceinline__ float cvt16_to_f32(uint16_t v) {
  if constexpr (F16) return __half2float(*reinterpret_cast<const __half*>(&v));
  else return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&v));
}

}  // namespace pmu
