// Shared host/device helpers of the wxformer_b200 C-ABI library.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "wxformer_b200.h"

extern thread_local char wxf_err_buf[512];

#define WXF_FAIL(code, ...)                                  \
  do {                                                       \
    snprintf(wxf_err_buf, sizeof(wxf_err_buf), __VA_ARGS__); \
    return (code);                                           \
  } while (0)

#define WXF_CHECK_LAUNCH(what)                                                                     \
  do {                                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) {                                                                      \
      snprintf(wxf_err_buf, sizeof(wxf_err_buf), "%s: %s", (what), cudaGetErrorString(e__));       \
      return (int)e__;                                                                             \
    }                                                                                              \
  } while (0)

__host__ __device__ static inline bool wxf_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float wxf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exact-erf GELU (nn.GELU() default), branch-free: erf(t) = 1 - 2^(-t q(t)) with a degree-8 minimax fit of
// q(t) = -log2(erfc(t))/t on [0, 4] (erfc(4) = 1.5e-8 is below half an ulp of 1).  |erf error| <= 1e-7 absolute, the same
// resolution "1 + erf" has in fp32; measured GELU error vs fp64 4.5e-7 max on [-8, 8] (torch's fp32 GELU: 1.2e-6).
__device__ __forceinline__ float wxf_gelu_erf(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float q = -1.160479314e-05f;
  q = fmaf(q, t, 1.529642177e-04f);
  q = fmaf(q, t, -8.482338744e-04f);
  q = fmaf(q, t, 2.274784725e-03f);
  q = fmaf(q, t, -8.480441466e-05f);
  q = fmaf(q, t, -2.772447653e-02f);
  q = fmaf(q, t, 1.483079046e-01f);
  q = fmaf(q, t, 9.184429049e-01f);
  q = fmaf(q, t, 1.627907276e+00f);
  float e;
  asm("ex2.approx.f32 %0, %1;" : "=f"(e) : "f"(-q * t));
  const float h = 0.5f * x;
  return fmaf(h, copysignf(1.0f - e, x), h);
}
__device__ __forceinline__ float wxf_silu(float x) { return x / (1.0f + expf(-x)); }

// two values at once: packed conversions (cvt.rn.f16x2.f32)
__device__ __forceinline__ void wxf_split2_f16x2(float a, float b, __half2& hi, __half2& lo) {
  const float ca = fminf(fmaxf(a, -65504.f), 65504.f), cb = fminf(fmaxf(b, -65504.f), 65504.f);
  hi = __floats2half2_rn(ca, cb);
  const float2 back = __half22float2(hi);
  lo = __floats2half2_rn(a - back.x, b - back.y);
}

// fp32 -> (hi, lo) fp16 operand planes of the f16x2 tensor-core scheme (22 significant bits; hi saturates)
__device__ __forceinline__ void wxf_split_f16x2(float v, __half& hi, __half& lo) {
  const float c = fminf(fmaxf(v, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = __float2half_rn(v - __half2float(hi));
}
