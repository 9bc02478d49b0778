// Shared host/device helpers of the wxformer_b200 C-ABI library.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "wxformer_b200.h"

// One value per CUDA device of the process (function attributes, SM counts and occupancy limits are per device; a
// process-wide `static` would skip the opt-in on the second GPU of a multi-device process).
template <typename T>
struct WxfPerDevice {
  T v[64] = {};
  T& get() {
    int dev = 0;
    cudaGetDevice(&dev);
    return v[dev & 63];
  }
};

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

// ---- programmatic dependent launch (WXF_PDL=1; round-2 candidate, off by default) --------------------------------------
// With the attribute, kernel N+1 may be scheduled while kernel N drains: its prologue (barrier init, TMEM allocation,
// index math) overlaps N's tail, and griddepcontrol.wait holds it before the first access to memory N produced.  Every
// kernel triggers its dependents at entry and waits before touching global memory; without the launch attribute both
// instructions are no-ops.
__device__ __forceinline__ void wxf_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void wxf_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool wxf_pdl_enabled();

template <typename... KArgs, typename... Args>
inline void wxf_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  if (wxf_pdl_enabled()) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
  } else {
    kernel<<<grid, block, smem, st>>>(args...);
  }
}

__host__ __device__ static inline bool wxf_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float wxf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exact-erf GELU (nn.GELU() default), branch-free: erf(t) = 1 - 2^(-t q(t))  [the exponent lies in [-26.6, 0]: the .ftz
// form of ex2 returns the same bits and saves the four-instruction subnormal-range fix-up per element] with a degree-8 minimax fit of
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
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q * t));
  const float h = 0.5f * x;
  return fmaf(h, copysignf(1.0f - e, x), h);
}
__device__ __forceinline__ float wxf_silu(float x) { return x / (1.0f + expf(-x)); }

// Two GELUs at once on Blackwell's packed fp32 pipe (FFMA2/FMUL2): same polynomial as wxf_gelu_erf.
__device__ __forceinline__ float2 wxf_gelu_erf2(float2 x) {
  float2 t;
  t.x = fminf(fabsf(x.x) * 0.70710678118654752440f, 4.0f);
  t.y = fminf(fabsf(x.y) * 0.70710678118654752440f, 4.0f);
  float2 q = make_float2(-1.160479314e-05f, -1.160479314e-05f);
  q = __ffma2_rn(q, t, make_float2(1.529642177e-04f, 1.529642177e-04f));
  q = __ffma2_rn(q, t, make_float2(-8.482338744e-04f, -8.482338744e-04f));
  q = __ffma2_rn(q, t, make_float2(2.274784725e-03f, 2.274784725e-03f));
  q = __ffma2_rn(q, t, make_float2(-8.480441466e-05f, -8.480441466e-05f));
  q = __ffma2_rn(q, t, make_float2(-2.772447653e-02f, -2.772447653e-02f));
  q = __ffma2_rn(q, t, make_float2(1.483079046e-01f, 1.483079046e-01f));
  q = __ffma2_rn(q, t, make_float2(9.184429049e-01f, 9.184429049e-01f));
  q = __ffma2_rn(q, t, make_float2(1.627907276e+00f, 1.627907276e+00f));
  const float2 a = __fmul2_rn(q, t);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-a.y));
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(h, make_float2(copysignf(1.0f - e0, x.x), copysignf(1.0f - e1, x.y)), h);
}

// The same function with three instructions fewer per pair: GELU(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2), and with
// t = min(|x| / sqrt 2, 4): 0.5 |x| = t / sqrt 2 wherever the clamp is inactive (beyond it the term is < 5e-8 either way).
// No copysign, no 1 - e: u = x / sqrt 2 (packed), t = min(|u|, 4), e = 2^(-q(t) t), result = max(x, 0) - (t e) / sqrt 2.
__device__ __forceinline__ float2 wxf_gelu_erf2_relu(float2 x) {
  const float2 u = __fmul2_rn(x, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  float2 t;
  t.x = fminf(fabsf(u.x), 4.0f);
  t.y = fminf(fabsf(u.y), 4.0f);
  float2 q = make_float2(-1.160479314e-05f, -1.160479314e-05f);
  q = __ffma2_rn(q, t, make_float2(1.529642177e-04f, 1.529642177e-04f));
  q = __ffma2_rn(q, t, make_float2(-8.482338744e-04f, -8.482338744e-04f));
  q = __ffma2_rn(q, t, make_float2(2.274784725e-03f, 2.274784725e-03f));
  q = __ffma2_rn(q, t, make_float2(-8.480441466e-05f, -8.480441466e-05f));
  q = __ffma2_rn(q, t, make_float2(-2.772447653e-02f, -2.772447653e-02f));
  q = __ffma2_rn(q, t, make_float2(1.483079046e-01f, 1.483079046e-01f));
  q = __ffma2_rn(q, t, make_float2(9.184429049e-01f, 9.184429049e-01f));
  q = __ffma2_rn(q, t, make_float2(1.627907276e+00f, 1.627907276e+00f));
  const float2 a = __fmul2_rn(q, t);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-a.y));
  const float2 m = __fmul2_rn(t, make_float2(e0, e1));
  return __ffma2_rn(m, make_float2(-0.70710678118654752440f, -0.70710678118654752440f),
                    make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

// two values at once: packed conversions (cvt.rn.f16x2.f32)
__device__ __forceinline__ void wxf_split2_f16x2(float a, float b, __half2& hi, __half2& lo) {
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));  // saturating: hi never becomes inf
  hi = *reinterpret_cast<__half2*>(&h);
  const float2 back = __half22float2(hi);
  const float2 d = __ffma2_rn(back, make_float2(-1.0f, -1.0f), make_float2(a, b));
  lo = __floats2half2_rn(d.x, d.y);
}

// fp32 -> (hi, lo) fp16 operand planes of the f16x2 tensor-core scheme (22 significant bits; hi saturates)
__device__ __forceinline__ void wxf_split_f16x2(float v, __half& hi, __half& lo) {
  const float c = fminf(fmaxf(v, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = __float2half_rn(v - __half2float(hi));
}
