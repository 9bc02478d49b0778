// Cross-scale window attention core on CUDA cores (exact fp32): S = (q*scale) k^T + bias, softmax, P v.
// One thread owns one query token (online softmax over the window's keys); K and V of the window/head
// are staged in shared memory and broadcast-read as float4.  Small windows (L < 64) pack several
// (window, head) pairs per CTA.  Reference arithmetic: credit/models/crossformer.py:261-296, 301-314.
#include <math_constants.h>

#include "wxf_common.cuh"

namespace {

constexpr int ATT_THREADS = 128;
constexpr int DH = 32;

__device__ __forceinline__ int64_t token_pixel(int64_t win, int i, int H, int W, int wsz, int nh, int nw, int kind) {
  const int per_img = nh * nw;
  const int b = (int)(win / per_img);
  const int rem = (int)(win - (int64_t)b * per_img);
  const int gh = rem / nw, gw = rem - gh * nw;
  const int ty = i / wsz, tx = i - ty * wsz;
  int y, x;
  if (kind == WXF_ATTN_SHORT) {
    y = gh * wsz + ty;
    x = gw * wsz + tx;
  } else {
    y = ty * nh + gh;
    x = tx * nw + gw;
  }
  return ((int64_t)b * H + y) * W + x;
}

template <bool SPLIT>
__global__ void __launch_bounds__(ATT_THREADS) window_attention_kernel(
    const float* __restrict__ qkv, int ldq, const float* __restrict__ biasT, float* __restrict__ out,
    __half* __restrict__ out_hi, __half* __restrict__ out_lo, int ldo, int H,
    int W, int d, int heads, int wsz, int kind, float scale, int L, int G, int64_t npairs) {
  extern __shared__ float4 smem4[];
  float4* Ks = smem4;
  float4* Vs = smem4 + (size_t)G * L * (DH / 4);
  const int tid = threadIdx.x;
  const int nh = H / wsz, nw = W / wsz;
  const int64_t pair0 = (int64_t)blockIdx.x * G;

  const int ntok = G * L;
  for (int idx = tid; idx < ntok * (DH / 4); idx += ATT_THREADS) {
    const int tt = idx >> 3, q4 = idx & 7;
    const int g = tt / L, i = tt - g * L;
    const int64_t pair = pair0 + g;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (pair < npairs) {
      const int head = (int)(pair % heads);
      const int64_t win = pair / heads;
      const int64_t pix = token_pixel(win, i, H, W, wsz, nh, nw, kind);
      const float* base = qkv + pix * ldq + head * DH + q4 * 4;
      kv = __ldg(reinterpret_cast<const float4*>(base + d));
      vv = __ldg(reinterpret_cast<const float4*>(base + 2 * d));
    }
    Ks[idx] = kv;
    Vs[idx] = vv;
  }
  __syncthreads();

  const int g = tid / L, i = tid - g * L;
  const int64_t pair = pair0 + g;
  if (g >= G || pair >= npairs) return;
  const int head = (int)(pair % heads);
  const int64_t win = pair / heads;
  const int64_t pix = token_pixel(win, i, H, W, wsz, nh, nw, kind);

  float q[DH];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + pix * ldq + head * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 v = __ldg(qp + c);
      q[4 * c + 0] = v.x * scale; q[4 * c + 1] = v.y * scale; q[4 * c + 2] = v.z * scale; q[4 * c + 3] = v.w * scale;
    }
  }
  float o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;
  float mrun = -CUDART_INF_F, lrun = 0.f;
  const float4* Kg = Ks + (size_t)g * L * (DH / 4);
  const float4* Vg = Vs + (size_t)g * L * (DH / 4);

  for (int j0 = 0; j0 < L; j0 += 4) {
    float s[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      if (j < L) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 k4 = Kg[j * (DH / 4) + c];
          acc = fmaf(q[4 * c + 0], k4.x, acc);
          acc = fmaf(q[4 * c + 1], k4.y, acc);
          acc = fmaf(q[4 * c + 2], k4.z, acc);
          acc = fmaf(q[4 * c + 3], k4.w, acc);
        }
        s[jj] = acc + __ldg(biasT + (size_t)j * L + i);
      } else {
        s[jj] = -CUDART_INF_F;
      }
    }
    const float mnew = fmaxf(fmaxf(mrun, fmaxf(s[0], s[1])), fmaxf(s[2], s[3]));
    const float corr = expf(mrun - mnew);
    lrun *= corr;
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] *= corr;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      if (j < L) {
        const float pj = expf(s[jj] - mnew);
        lrun += pj;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 v4 = Vg[j * (DH / 4) + c];
          o[4 * c + 0] = fmaf(pj, v4.x, o[4 * c + 0]);
          o[4 * c + 1] = fmaf(pj, v4.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(pj, v4.z, o[4 * c + 2]);
          o[4 * c + 3] = fmaf(pj, v4.w, o[4 * c + 3]);
        }
      }
    }
    mrun = mnew;
  }
  const float inv = 1.0f / lrun;
  if constexpr (SPLIT) {
    uint4* hp = reinterpret_cast<uint4*>(out_hi + pix * ldo + head * DH);
    uint4* lp = reinterpret_cast<uint4*>(out_lo + pix * ldo + head * DH);
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      __align__(16) __half h8[8];
      __align__(16) __half l8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) wxf_split_f16x2(o[8 * c + e] * inv, h8[e], l8[e]);
      hp[c] = *reinterpret_cast<const uint4*>(h8);
      lp[c] = *reinterpret_cast<const uint4*>(l8);
    }
  } else {
    float4* op = reinterpret_cast<float4*>(out + pix * ldo + head * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c)
      op[c] = make_float4(o[4 * c + 0] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
  }
}

}  // namespace

static int attention_launch(const float* qkv, int ldq, const float* biasT, float* out, void* out_hi, void* out_lo,
                            int ldo, int B, int H, int W, int d, int dh, int wsz, int kind, float scale, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || d <= 0 || wsz <= 0) WXF_FAIL(WXF_EINVAL, "attention: bad dims");
  if (dh != DH) WXF_FAIL(WXF_EUNSUPPORTED, "attention: dim_head must be 32, got %d", dh);
  if (d % dh) WXF_FAIL(WXF_EINVAL, "attention: d %% dh != 0");
  if (H % wsz || W % wsz) WXF_FAIL(WXF_EINVAL, "attention: grid %dx%d not divisible by window %d", H, W, wsz);
  if (kind != WXF_ATTN_SHORT && kind != WXF_ATTN_LONG) WXF_FAIL(WXF_EINVAL, "attention: bad kind");
  const int L = wsz * wsz;
  if (L > ATT_THREADS) WXF_FAIL(WXF_EUNSUPPORTED, "attention: window %d (L=%d) > %d tokens", wsz, L, ATT_THREADS);
  const bool split = out_hi != nullptr;
  if (ldq < 3 * d || ldo < d || (ldq & 3) || (ldo & (split ? 7 : 3)) || !wxf_aligned16(qkv) ||
      !(split ? (wxf_aligned16(out_hi) && wxf_aligned16(out_lo)) : wxf_aligned16(out)))
    WXF_FAIL(WXF_EALIGN, "attention: strides/pointers must be 16-byte aligned");
  const int heads = d / dh;
  const int G = ATT_THREADS / L;
  const int64_t npairs = (int64_t)B * (H / wsz) * (W / wsz) * heads;
  const int64_t blocks = (npairs + G - 1) / G;
  const size_t smem = (size_t)G * L * DH * sizeof(float) * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (split)
    window_attention_kernel<true><<<(unsigned)blocks, ATT_THREADS, smem, st>>>(
        qkv, ldq, biasT, nullptr, (__half*)out_hi, (__half*)out_lo, ldo, H, W, d, heads, wsz, kind, scale, L, G, npairs);
  else
    window_attention_kernel<false><<<(unsigned)blocks, ATT_THREADS, smem, st>>>(
        qkv, ldq, biasT, out, nullptr, nullptr, ldo, H, W, d, heads, wsz, kind, scale, L, G, npairs);
  WXF_CHECK_LAUNCH("window_attention");
  return 0;
}

extern "C" int wxf_window_attention_f32(const float* qkv, int ldq, const float* biasT, float* out, int ldo, int B,
                                        int H, int W, int d, int dh, int wsz, int kind, float scale, void* stream) {
  if (!out) WXF_FAIL(WXF_EINVAL, "attention: null output");
  return attention_launch(qkv, ldq, biasT, out, nullptr, nullptr, ldo, B, H, W, d, dh, wsz, kind, scale, stream);
}

extern "C" int wxf_window_attention_f16x2(const float* qkv, int ldq, const float* biasT, void* out_hi, void* out_lo,
                                          int ldh, int B, int H, int W, int d, int dh, int wsz, int kind, float scale,
                                          void* stream) {
  if (!out_hi || !out_lo) WXF_FAIL(WXF_EINVAL, "attention: null output planes");
  return attention_launch(qkv, ldq, biasT, nullptr, out_hi, out_lo, ldh, B, H, W, d, dh, wsz, kind, scale, stream);
}
