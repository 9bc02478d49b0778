// Noise injection of the ensemble variant (credit/models/wxformer/crossformer_ensemble.py:9-177,
// StochasticDecompositionLayer, credit/models/wxformer/stochastic_decomposition_layer.py:5-42):
//   feature + (noise_factor * eps) * style * modulation,   eps ~ N(0,1) per (b, c, y, x),   style = Linear(latent)[b, c]
// Two kernels per injection site: the per-(batch, channel) coefficient (latent draw + the small Linear), and one pass over
// the feature map that draws eps with Philox, adds, and writes fp32 and / or the fp16 hi/lo operand planes of the consumer.
// Tests feed recorded draws of the reference (eps / latent pointers) instead of the generator.
#include <curand_kernel.h>

#include "wxf_common.cuh"

namespace {

// coef[b, c] = noise_factor * (W[c, :] . latent[b, :] + bias[c]) * modulation[c]
__global__ void __launch_bounds__(128) noise_coef_kernel(const float* __restrict__ latent, const float* __restrict__ W,
                                                         const float* __restrict__ bias, const float* __restrict__ mod,
                                                         const float* __restrict__ factor, float* __restrict__ coef, int C,
                                                         int D, unsigned long long seed, const unsigned long long* step,
                                                         int site) {
  extern __shared__ float lat[];
  const int b = blockIdx.y;
  if (latent) {
    for (int k = threadIdx.x; k < D; k += blockDim.x) lat[k] = latent[(size_t)b * D + k];
  } else {
    // one Philox stream per (step, site, batch): the first D normals of the stream are the latent vector
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, ((unsigned long long)site << 32) + (unsigned long long)b, (*step) * 4096ull + (unsigned long long)k, &st);
      lat[k] = curand_normal(&st);
    }
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < D; ++k) s = fmaf(__ldg(W + (size_t)c * D + k), lat[k], s);
  coef[(size_t)b * C + c] = __ldg(factor) * (s + __ldg(bias + c)) * __ldg(mod + c);
}

// out[p, c] = x[p, c] + eps[p, c] * coef[b, c]   (pixel-major, float4 granularity; p = b*HW + pixel)
__global__ void __launch_bounds__(256) noise_inject_kernel(const float* __restrict__ x, int ldx, float* out, int ldo,
                                                           __half* __restrict__ hi, __half* __restrict__ lo, int ldh,
                                                           const float* __restrict__ coef, const float* __restrict__ eps,
                                                           int64_t HW, int C4, int64_t total, unsigned long long seed,
                                                           const unsigned long long* step, int site) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int C = C4 * 4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / C4;
    const int c = (int)(e - p * C4) * 4;
    const int64_t b = p / HW;
    float4 n;
    if (eps) {
      n = *reinterpret_cast<const float4*>(eps + p * C + c);
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, ((unsigned long long)(site + 64) << 32) + (unsigned long long)(e >> 20), (*step) * (1ull << 22) + (unsigned long long)(e & ((1 << 20) - 1)) * 4ull, &st);
      n = curand_normal4(&st);
    }
    const float4 v = *reinterpret_cast<const float4*>(x + p * ldx + c);
    const float4 k = *reinterpret_cast<const float4*>(coef + b * C + c);
    float4 o;
    o.x = fmaf(n.x, k.x, v.x);
    o.y = fmaf(n.y, k.y, v.y);
    o.z = fmaf(n.z, k.z, v.z);
    o.w = fmaf(n.w, k.w, v.w);
    if (out) *reinterpret_cast<float4*>(out + p * ldo + c) = o;
    if (hi) {
      __align__(8) __half2 h2[2];
      __align__(8) __half2 l2[2];
      wxf_split2_f16x2(o.x, o.y, h2[0], l2[0]);
      wxf_split2_f16x2(o.z, o.w, h2[1], l2[1]);
      *reinterpret_cast<uint2*>(hi + p * ldh + c) = *reinterpret_cast<const uint2*>(h2);
      *reinterpret_cast<uint2*>(lo + p * ldh + c) = *reinterpret_cast<const uint2*>(l2);
    }
  }
}

__global__ void noise_step_kernel(unsigned long long* step) {
  if (threadIdx.x == 0) *step += 1ull;
}

}  // namespace

extern "C" int wxf_noise_coef(const float* latent, const float* W, const float* bias, const float* modulation, const float* factor,
                              float* coef, int B, int C, int D, uint64_t seed, const void* step_counter, int site, void* stream) {
  if (!W || !bias || !modulation || !factor || !coef || B <= 0 || C <= 0 || D <= 0 || D > 4096 || (!latent && !step_counter))
    WXF_FAIL(WXF_EINVAL, "noise_coef: bad arguments");
  dim3 grid((C + 127) / 128, B);
  noise_coef_kernel<<<grid, 128, D * sizeof(float), (cudaStream_t)stream>>>(latent, W, bias, modulation, factor, coef, C, D,
                                                                          (unsigned long long)seed,
                                                                          reinterpret_cast<const unsigned long long*>(step_counter), site);
  WXF_CHECK_LAUNCH("noise_coef");
  return 0;
}

extern "C" int wxf_noise_inject(const float* x, int ldx, float* out, int ldo, void* out_hi, void* out_lo, int ldh, int h_off,
                                const float* coef, const float* eps, int B, int64_t HW, int C, uint64_t seed,
                                const void* step_counter, int site, void* stream) {
  if (!x || !coef || B <= 0 || HW <= 0 || C <= 0 || (C & 3) || (ldx & 3) || ldx < C || !wxf_aligned16(x) || !wxf_aligned16(coef) ||
      (!eps && !step_counter))
    WXF_FAIL(WXF_EINVAL, "noise_inject: bad arguments (C, strides multiples of 4; 16-byte aligned)");
  if (!out && !out_hi) WXF_FAIL(WXF_EINVAL, "noise_inject: no output");
  if ((out_hi == nullptr) != (out_lo == nullptr)) WXF_FAIL(WXF_EINVAL, "noise_inject: out_hi/out_lo come together");
  if (out && ((ldo & 3) || ldo < C || !wxf_aligned16(out))) WXF_FAIL(WXF_EALIGN, "noise_inject: out stride/alignment");
  if (eps && !wxf_aligned16(eps)) WXF_FAIL(WXF_EALIGN, "noise_inject: eps alignment");
  __half* hi = out_hi ? reinterpret_cast<__half*>(out_hi) + h_off : nullptr;
  __half* lo = out_lo ? reinterpret_cast<__half*>(out_lo) + h_off : nullptr;
  if (hi && ((ldh & 3) || (h_off & 3) || ldh < h_off + C || ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7)))
    WXF_FAIL(WXF_EALIGN, "noise_inject: plane stride/alignment");
  const int64_t total = (int64_t)B * HW * (C / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wxf_launch(noise_inject_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, ldx, out, ldo, hi, lo, ldh, coef, eps,
             HW, C / 4, total, (unsigned long long)seed, reinterpret_cast<const unsigned long long*>(step_counter), site);
  WXF_CHECK_LAUNCH("noise_inject");
  return 0;
}

extern "C" int wxf_noise_step_advance(void* step_counter, void* stream) {
  if (!step_counter) WXF_FAIL(WXF_EINVAL, "noise_step_advance: null");
  noise_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(step_counter));
  WXF_CHECK_LAUNCH("noise_step_advance");
  return 0;
}
