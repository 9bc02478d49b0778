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

__device__ __forceinline__ float wxf_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float wxf_silu(float x) { return x / (1.0f + expf(-x)); }

// fp32 -> (hi, lo) fp16 operand planes of the f16x2 tensor-core scheme (22 significant bits; hi saturates)
__device__ __forceinline__ void wxf_split_f16x2(float v, __half& hi, __half& lo) {
  const float c = fminf(fmaxf(v, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = __float2half_rn(v - __half2float(hi));
}
