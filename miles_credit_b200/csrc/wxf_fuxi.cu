// FuXi forecast step (credit/models/fuxi.py:454-506): the kernels the WXFormer path does not already provide.
//   * res-post-norm LayerNorm  x <- x + LN(y)          (timm SwinTransformerV2Block: x + norm1(attn(x)), x + norm2(mlp(x)))
//   * Swin-V2 window attention: cyclic shift, L2-normalised q / k, per-head logit scale, per-head position bias,
//     -100 masks between the regions the shift glues together (timm WindowAttention / _calc_attn_mask)
//   * row gather with zero fill (ZeroPad2d to a window multiple, fuxi.py:67-79, 281-283; crop + concat, :288-292)
//   * un-patchify + un-pad + bilinear resize + NHWC -> NCHW of the dense head's output (fuxi.py:484-498)
// The contractions (cube embedding as a k4 s4 implicit GEMM, DownBlock / UpBlock convolutions, qkv / proj / MLP / head
// GEMMs) run on the tcgen05 kernels of wxf_gemm_tc.cu; GroupNorm + SiLU on wxf_pointwise.cu.
#include "wxf_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// out = res + LayerNorm(x) * g + b, one warp per row.  NV4 > 0: d == NV4 * 128, the row lives in registers (float4 per
// lane); NV4 == 0: any d, the row is re-read (it sits in L1 after the first pass).
template <int NV4>
__global__ void __launch_bounds__(256) ln_residual_kernel(const float* __restrict__ x, int ldx, const float* res, int ldr,
                                                          float* out, int ldo, __half* __restrict__ out_hi,
                                                          __half* __restrict__ out_lo, int ldh, const float* __restrict__ g,
                                                          const float* __restrict__ bta, int64_t M, int d, float eps) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * ldx;
  const float* rr = res ? res + row * ldr : nullptr;
  if constexpr (NV4 > 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    float4 v[NV4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      v[k] = x4[lane + 32 * k];
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mean = wxf_warp_sum(s) / (float)d;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
      ss += (a * a + b * b) + (c * c + e * e);
    }
    const float rstd = rsqrtf(wxf_warp_sum(ss) / (float)d + eps);
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c4 = lane + 32 * k;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + c4), bb = __ldg(reinterpret_cast<const float4*>(bta) + c4);
      float4 o;
      o.x = (v[k].x - mean) * rstd * gg.x + bb.x;
      o.y = (v[k].y - mean) * rstd * gg.y + bb.y;
      o.z = (v[k].z - mean) * rstd * gg.z + bb.z;
      o.w = (v[k].w - mean) * rstd * gg.w + bb.w;
      if (rr) {
        const float4 r4 = *reinterpret_cast<const float4*>(rr + 4 * c4);
        o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
      }
      if (out) *reinterpret_cast<float4*>(out + row * ldo + 4 * c4) = o;
      if (out_hi) {
        __align__(8) __half2 h2[2];
        __align__(8) __half2 l2[2];
        wxf_split2_f16x2(o.x, o.y, h2[0], l2[0]);
        wxf_split2_f16x2(o.z, o.w, h2[1], l2[1]);
        *reinterpret_cast<uint2*>(out_hi + row * ldh + 4 * c4) = *reinterpret_cast<const uint2*>(h2);
        *reinterpret_cast<uint2*>(out_lo + row * ldh + 4 * c4) = *reinterpret_cast<const uint2*>(l2);
      }
    }
  } else {
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = wxf_warp_sum(s) / (float)d;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float t = xr[c] - mean;
      ss += t * t;
    }
    const float rstd = rsqrtf(wxf_warp_sum(ss) / (float)d + eps);
    for (int c = lane; c < d; c += 32) {
      float o = (xr[c] - mean) * rstd * __ldg(g + c) + __ldg(bta + c);
      if (rr) o += rr[c];
      if (out) out[row * ldo + c] = o;
      if (out_hi) {
        __half hi, lo;
        wxf_split_f16x2(o, hi, lo);
        out_hi[row * ldh + c] = hi;
        out_lo[row * ldh + c] = lo;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Swin-V2 window attention, exact fp32 on the CUDA cores.  One CTA = one (window, head).
//   qkv  : [B, H, W, ldq] fp32, q = [0, d), k = [d, 2d), v = [2d, 3d), head h = channels h*dh .. h*dh + dh - 1
//   window (wy, wx), token (ty, tx) sits at rolled position (wy*wsh + ty, wx*wsw + tx) = source pixel
//   ((wy*wsh + ty + sh) mod H, (wx*wsw + tx + sw) mod W)   [torch.roll(x, (-sh, -sw))]; the output row goes back to the
//   same source pixel [torch.roll(.., (+sh, +sw)) after window_reverse].
//   S[i][j] = <q_i/|q_i|, k_j/|k_j|> * scale[h] + bias[h][i][j] + (region(i) != region(j) ? -100 : 0)
// Shared memory: q, k, v rows padded to dh + 4 floats (float4 loads of 32 different rows hit 32 different bank groups),
// S padded to L + 1.
constexpr int SW_THREADS = 512;  // 16 warps x 2 CTAs per SM (87 KB of shared memory each): the phases are latency-bound, more warps hide it
constexpr int SW_R = 4;  // query rows per warp pass

__global__ void __launch_bounds__(SW_THREADS, 2) swin_attention_kernel(const float* __restrict__ qkv, int ldq,
                                                                    const float* __restrict__ bias,
                                                                    const float* __restrict__ scale,
                                                                    __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                                    float* __restrict__ out_f32, int ldh, int H, int W,
                                                                    int d, int dh, int wsh, int wsw, int sh, int sw, int nwx,
                                                                    int nwy, int msh) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int L = wsh * wsw, DS = dh + 4, LS = (L + 4) & ~3, dh4 = dh >> 2;  // LS % 4 == 0: float4 loads of P rows
  float* qs = smem;
  float* ks = qs + L * DS;
  float* vs = ks + L * DS;
  float* S = vs + L * DS;
  int* rid = reinterpret_cast<int*>(S + L * LS);
  int64_t* pix = reinterpret_cast<int64_t*>(rid + ((L + 1) & ~1));
  float* inv_n = reinterpret_cast<float*>(pix + L);  // 1 / max(|q_i|, eps) for i < L, then 1 / max(|k_j|, eps)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SW_THREADS >> 5;
  const int head = blockIdx.y;
  int win = blockIdx.x;
  const int wx = win % nwx;
  win /= nwx;
  const int wy = win % nwy, b = win / nwy;

  for (int t = tid; t < L; t += SW_THREADS) {
    const int ry = wy * wsh + t / wsw, rx = wx * wsw + t % wsw;
    int sy = ry + sh, sx = rx + sw;
    if (sy >= H) sy -= H;
    if (sx >= W) sx -= W;
    pix[t] = ((int64_t)b * H + sy) * W + sx;
    int r = 0;
    // msh: the row shift the MASK is built for.  It equals sh except on a latitude band of a decomposed forecast, where the
    // roll over rows is done by the caller (the band buffer starts `shift` rows into the rank's rows: sh = 0) and only the
    // band that holds the wrapped window row still needs the mask (msh = shift there, 0 elsewhere).
    if (msh > 0) r += 3 * (ry < H - wsh ? 0 : (ry < H - msh ? 1 : 2));
    if (sw > 0) r += (rx < W - wsw ? 0 : (rx < W - sw ? 1 : 2));
    rid[t] = r;
  }
  __syncthreads();
  for (int i = tid; i < L * dh4; i += SW_THREADS) {
    const int t = i / dh4, c4 = i - t * dh4;
    const float* src = qkv + pix[t] * ldq + head * dh + 4 * c4;
    *reinterpret_cast<float4*>(qs + t * DS + 4 * c4) = *reinterpret_cast<const float4*>(src);
    *reinterpret_cast<float4*>(ks + t * DS + 4 * c4) = *reinterpret_cast<const float4*>(src + d);
    *reinterpret_cast<float4*>(vs + t * DS + 4 * c4) = *reinterpret_cast<const float4*>(src + 2 * d);
  }
  __syncthreads();
  // F.normalize(q, dim=-1), F.normalize(k, dim=-1): x / max(|x|_2, 1e-12).  The rows stay as loaded; the two reciprocal norms
  // scale the dot product instead (one multiply per score instead of a read-modify-write of both operand tiles).
  for (int r = warp; r < 2 * L; r += nwarps) {
    const float* row = (r < L ? qs + r * DS : ks + (r - L) * DS);
    float ss = 0.f;
    for (int c4 = lane; c4 < dh4; c4 += 32) {
      const float4 t = *reinterpret_cast<const float4*>(row + 4 * c4);
      ss += (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
    }
    ss = wxf_warp_sum(ss);
    if (lane == 0) inv_n[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  const float sc = __ldg(scale + head);
  const float* bh = bias + (size_t)head * L * L;
  // S = cos(q, k) * scale + bias + mask
  for (int i0 = warp * SW_R; i0 < L; i0 += nwarps * SW_R) {
    // packed fp32 pipe (FFMA2): every accumulator is a pair (sum over even, sum over odd channel pairs), so the halves of the
    // float4 loads are the operands as they sit in registers: 2 instructions per 4 multiply-adds
    float2 acc2[SW_R][2];
#pragma unroll
    for (int r = 0; r < SW_R; ++r) acc2[r][0] = acc2[r][1] = make_float2(0.f, 0.f);
    const int j0 = lane, j1 = lane + 32;
    const float* k0 = ks + (j0 < L ? j0 : 0) * DS;
    const float* k1 = ks + (j1 < L ? j1 : 0) * DS;
    const float* qrow[SW_R];
#pragma unroll
    for (int r = 0; r < SW_R; ++r) qrow[r] = qs + ((i0 + r < L) ? i0 + r : L - 1) * DS;
#pragma unroll 2
    for (int c4 = 0; c4 < dh4; ++c4) {
      const float4 a0 = *reinterpret_cast<const float4*>(k0 + 4 * c4);
      const float4 a1 = *reinterpret_cast<const float4*>(k1 + 4 * c4);
#pragma unroll
      for (int r = 0; r < SW_R; ++r) {
        const float4 q4 = *reinterpret_cast<const float4*>(qrow[r] + 4 * c4);
        const float2 qa = make_float2(q4.x, q4.y), qb = make_float2(q4.z, q4.w);
        acc2[r][0] = __ffma2_rn(qa, make_float2(a0.x, a0.y), acc2[r][0]);
        acc2[r][0] = __ffma2_rn(qb, make_float2(a0.z, a0.w), acc2[r][0]);
        acc2[r][1] = __ffma2_rn(qa, make_float2(a1.x, a1.y), acc2[r][1]);
        acc2[r][1] = __ffma2_rn(qb, make_float2(a1.z, a1.w), acc2[r][1]);
      }
    }
    float acc[SW_R][2];
#pragma unroll
    for (int r = 0; r < SW_R; ++r) {
      acc[r][0] = acc2[r][0].x + acc2[r][0].y;
      acc[r][1] = acc2[r][1].x + acc2[r][1].y;
    }
#pragma unroll
    for (int r = 0; r < SW_R; ++r) {
      const int i = i0 + r;
      if (i >= L) break;
      const float qi = inv_n[i] * sc;
      if (j0 < L) S[i * LS + j0] = acc[r][0] * (qi * inv_n[L + j0]) + __ldg(bh + i * L + j0) + (rid[i] != rid[j0] ? -100.f : 0.f);
      if (j1 < L) S[i * LS + j1] = acc[r][1] * (qi * inv_n[L + j1]) + __ldg(bh + i * L + j1) + (rid[i] != rid[j1] ? -100.f : 0.f);
    }
  }
  __syncthreads();
  // softmax over j
  for (int i = warp; i < L; i += nwarps) {
    float* row = S + i * LS;
    const float a = lane < L ? row[lane] : -INFINITY, c = lane + 32 < L ? row[lane + 32] : -INFINITY;
    float m = fmaxf(a, c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float ea = lane < L ? expf(a - m) : 0.f, ec = lane + 32 < L ? expf(c - m) : 0.f;
    const float inv = 1.0f / wxf_warp_sum(ea + ec);
    if (lane < L) row[lane] = ea * inv;
    if (lane + 32 < L) row[lane + 32] = ec * inv;
  }
  __syncthreads();
  // O = P V, written to the token's source pixel as fp16 hi/lo planes (A operand of the proj GEMM) or fp32
  for (int i0 = warp * SW_R; i0 < L; i0 += nwarps * SW_R) {
    for (int e4 = lane; e4 < dh4; e4 += 32) {
      // four keys per step: the probabilities of a row come as one float4 (broadcast load), the products run on the packed
      // fp32 pipe with the probability as the scalar operand: 2 instructions per 4 multiply-adds
      float2 oa[SW_R], ob[SW_R];
#pragma unroll
      for (int r = 0; r < SW_R; ++r) oa[r] = ob[r] = make_float2(0.f, 0.f);
      const float* prow[SW_R];
#pragma unroll
      for (int r = 0; r < SW_R; ++r) prow[r] = S + ((i0 + r < L) ? i0 + r : L - 1) * LS;
      const float* vcol = vs + 4 * e4;
      int j = 0;
      for (; j + 4 <= L; j += 4) {
        float4 v4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v4[u] = *reinterpret_cast<const float4*>(vcol + (j + u) * DS);
#pragma unroll
        for (int r = 0; r < SW_R; ++r) {
          const float4 p4 = *reinterpret_cast<const float4*>(prow[r] + j);
          const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            oa[r] = __ffma2_rn(make_float2(v4[u].x, v4[u].y), make_float2(pp[u], pp[u]), oa[r]);
            ob[r] = __ffma2_rn(make_float2(v4[u].z, v4[u].w), make_float2(pp[u], pp[u]), ob[r]);
          }
        }
      }
      for (; j < L; ++j) {
        const float4 v4 = *reinterpret_cast<const float4*>(vcol + j * DS);
#pragma unroll
        for (int r = 0; r < SW_R; ++r) {
          const float p = prow[r][j];
          oa[r] = __ffma2_rn(make_float2(v4.x, v4.y), make_float2(p, p), oa[r]);
          ob[r] = __ffma2_rn(make_float2(v4.z, v4.w), make_float2(p, p), ob[r]);
        }
      }
      float4 o[SW_R];
#pragma unroll
      for (int r = 0; r < SW_R; ++r) o[r] = make_float4(oa[r].x, oa[r].y, ob[r].x, ob[r].y);
#pragma unroll
      for (int r = 0; r < SW_R; ++r) {
        const int i = i0 + r;
        if (i >= L) break;
        const int64_t off = pix[i] * ldh + head * dh + 4 * e4;
        if (out_hi) {
          __align__(8) __half2 h2[2];
          __align__(8) __half2 l2[2];
          wxf_split2_f16x2(o[r].x, o[r].y, h2[0], l2[0]);
          wxf_split2_f16x2(o[r].z, o[r].w, h2[1], l2[1]);
          *reinterpret_cast<uint2*>(out_hi + off) = *reinterpret_cast<const uint2*>(h2);
          *reinterpret_cast<uint2*>(out_lo + off) = *reinterpret_cast<const uint2*>(l2);
        } else {
          *reinterpret_cast<float4*>(out_f32 + off) = o[r];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dst[i, 0:d] = idx[i] >= 0 ? src[idx[i], 0:d] : 0, as fp32 and / or fp16 hi/lo planes (float4 granularity).
__global__ void __launch_bounds__(256) gather_rows_ex_kernel(const float* __restrict__ src, int ld_src,
                                                             const int32_t* __restrict__ idx, float* __restrict__ dst,
                                                             int ld_dst, __half* __restrict__ hi, __half* __restrict__ lo,
                                                             int ldh, int64_t n, int d4) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int64_t total = n * d4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / d4;
    const int c4 = (int)(e - i * d4);
    const int s = __ldg(idx + i);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0) v = *reinterpret_cast<const float4*>(src + (int64_t)s * ld_src + 4 * c4);
    if (dst) *reinterpret_cast<float4*>(dst + i * ld_dst + 4 * c4) = v;
    if (hi) {
      __align__(8) __half2 h2[2];
      __align__(8) __half2 l2[2];
      wxf_split2_f16x2(v.x, v.y, h2[0], l2[0]);
      wxf_split2_f16x2(v.z, v.w, h2[1], l2[1]);
      *reinterpret_cast<uint2*>(hi + i * ldh + 4 * c4) = *reinterpret_cast<const uint2*>(h2);
      *reinterpret_cast<uint2*>(lo + i * ldh + 4 * c4) = *reinterpret_cast<const uint2*>(l2);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Dense-head output [B, Lat, Lon, ph*pw*cp] (token-major; pixel (py, px) of a patch owns columns (py*pw + px)*cp + c)
// -> un-patchify -> crop [top, top+Hc) x [left, left+Wc) -> bilinear resize -> NCHW.  Same arithmetic as
// unpad_resize_kernel (wxf_pointwise.cu); only the source address of a pixel differs.
__device__ __forceinline__ void bilin_axis_f(int dst, float scale, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float src = fmaf(scale, (float)dst + 0.5f, -0.5f);  // torch's fp32 expression is FMA-contracted (CPU and CUDA)
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + ((i0 < n_in - 1) ? 1 : 0);
  l1 = src - (float)i0;
  if (l1 < 0.f) l1 = 0.f;
  if (l1 > 1.f) l1 = 1.f;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) unpatchify_resize_kernel(const float* __restrict__ y, float* __restrict__ out, int C,
                                                                int cp, int Lat, int Lon, int ph, int pw, int top, int left,
                                                                int Hc, int Wc, int Ho, int Wo, float sh, float sw,
                                                                int cgroups, int o0, int lat0) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  __shared__ float tile[32][33];  // [channel][column]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int b = blockIdx.z / cgroups, cg = blockIdx.z % cgroups;
  const int h = o0 + blockIdx.y, w0 = blockIdx.x * 32;
  int y0, y1;
  float ly0, ly1;
  bilin_axis_f(h, sh, Hc, y0, y1, ly0, ly1);
  const int ch = cg * 32 + tx;
  const int64_t tok = (int64_t)ph * pw * cp;
  // row part of the source address once per CTA, column part once per (thread, column); zero-weight taps are not loaded
  // (W is not resized in the forecast configs, so half of the four taps usually drop out)
  const int Y0 = y0 + top - lat0 * ph, Y1 = y1 + top - lat0 * ph;  // rows of the buffer (it starts at patch row lat0)
  const float* row0 = y + ((int64_t)b * Lat + Y0 / ph) * Lon * tok + (int64_t)(Y0 % ph) * pw * cp + ch;
  const float* row1 = y + ((int64_t)b * Lat + Y1 / ph) * Lon * tok + (int64_t)(Y1 % ph) * pw * cp + ch;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int w = w0 + i;
    float v = 0.f;
    if (w < Wo && ch < C) {
      int x0, x1;
      float lx0, lx1;
      bilin_axis_f(w, sw, Wc, x0, x1, lx0, lx1);
      const int X0 = x0 + left, X1 = x1 + left;
      const int64_t c0 = (int64_t)(X0 / pw) * tok + (X0 % pw) * cp;
      float top_v = lx0 * row0[c0], bot_v = (ly1 != 0.f) ? lx0 * row1[c0] : 0.f;
      if (lx1 != 0.f) {
        const int64_t c1 = (int64_t)(X1 / pw) * tok + (X1 % pw) * cp;
        top_v += lx1 * row0[c1];
        if (ly1 != 0.f) bot_v += lx1 * row1[c1];
      }
      v = ly0 * top_v + ly1 * bot_v;
    }
    tile[tx][i] = v;
  }
  __syncthreads();
  const int w = w0 + tx;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = cg * 32 + i;
    if (w < Wo && c < C) out[((size_t)(b * C + c) * Ho + h) * Wo + w] = tile[i][tx];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Autoregressive state update with a history window (trainers/rollout_utils.py:288-311: drop the oldest time step,
// append the newest; update_x, datasets/gen_2/channel_utils.py:253-291, for the newest step).  x: [B, C, T, plane].
// One thread owns a pixel of one (b, c) and walks t upwards, so the in-place shift never reads a value it has written.
__global__ void __launch_bounds__(256) history_update_kernel(float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ frc, int C, int T, int n_prog,
                                                             int n_dyn, int Cy, int Ty, int64_t plane, int64_t total) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e % plane;
    const int64_t bc = e / plane;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    float* xc = x + (bc * T) * plane + p;
    for (int t = 0; t + 1 < T; ++t) xc[(int64_t)t * plane] = xc[(int64_t)(t + 1) * plane];
    if (c < n_prog)
      xc[(int64_t)(T - 1) * plane] = y[((b * Cy + c) * Ty) * plane + p];
    else if (frc && c < n_prog + n_dyn)
      xc[(int64_t)(T - 1) * plane] = frc[(b * n_dyn + (c - n_prog)) * plane + p];
  }
}

}  // namespace

extern "C" int wxf_history_update(float* x, const float* y, const float* forcing, int B, int C, int T, int n_prog, int n_dyn,
                                  int Cy, int Ty, int64_t plane, void* stream) {
  if (!x || !y || B <= 0 || C <= 0 || T <= 0 || n_prog < 0 || n_dyn < 0 || n_prog + n_dyn > C || n_prog > Cy || Ty <= 0 ||
      plane <= 0)
    WXF_FAIL(WXF_EINVAL, "history_update: bad dims");
  const int64_t total = (int64_t)B * C * plane;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  wxf_launch(history_update_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, y, forcing, C, T, n_prog, n_dyn, Cy, Ty,
             plane, total);
  WXF_CHECK_LAUNCH("history_update");
  return 0;
}

extern "C" int wxf_layernorm_residual(const float* x, int ldx, const float* res, int ldr, float* out, int ldo, void* out_hi,
                                      void* out_lo, int ldh, const float* g, const float* b, int64_t M, int d, float eps,
                                      void* stream) {
  if (!x || !g || !b || M <= 0 || d <= 0 || ldx < d) WXF_FAIL(WXF_EINVAL, "layernorm_residual: bad arguments");
  if (!out && !out_hi) WXF_FAIL(WXF_EINVAL, "layernorm_residual: no output");
  if ((out_hi == nullptr) != (out_lo == nullptr)) WXF_FAIL(WXF_EINVAL, "layernorm_residual: out_hi/out_lo come together");
  if ((res && ldr < d) || (out && ldo < d) || (out_hi && ldh < d)) WXF_FAIL(WXF_EINVAL, "layernorm_residual: bad strides");
  const unsigned blocks = (unsigned)((M + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  __half* hi = (__half*)out_hi;
  __half* lo = (__half*)out_lo;
  const bool vec = d % 128 == 0 && ldx % 4 == 0 && wxf_aligned16(x) && wxf_aligned16(g) && wxf_aligned16(b) &&
                   (!res || (ldr % 4 == 0 && wxf_aligned16(res))) && (!out || (ldo % 4 == 0 && wxf_aligned16(out))) &&
                   (!hi || (ldh % 4 == 0 && ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0));
#define LNR(NV4)                                                                                                          \
  if (vec && d == NV4 * 128) {                                                                                            \
    wxf_launch(ln_residual_kernel<NV4>, dim3(blocks), dim3(256), 0, st, x, ldx, res, ldr, out, ldo, hi, lo, ldh, g, b, M, d, eps); \
    WXF_CHECK_LAUNCH("layernorm_residual");                                                                               \
    return 0;                                                                                                             \
  }
  LNR(1) LNR(2) LNR(3) LNR(4) LNR(6) LNR(8) LNR(12)
#undef LNR
  wxf_launch(ln_residual_kernel<0>, dim3(blocks), dim3(256), 0, st, x, ldx, res, ldr, out, ldo, hi, lo, ldh, g, b, M, d, eps);
  WXF_CHECK_LAUNCH("layernorm_residual");
  return 0;
}

extern "C" int wxf_swin_window_attention(const float* qkv, int ldq, const float* bias, const float* logit_scale, void* out_hi,
                                         void* out_lo, float* out_f32, int ldh, int B, int H, int W, int d, int heads,
                                         int ws_h, int ws_w, int shift_h, int shift_w, int mask_shift_h, void* stream) {
  if (!qkv || !bias || !logit_scale || B <= 0 || H <= 0 || W <= 0 || d <= 0 || heads <= 0 || d % heads)
    WXF_FAIL(WXF_EINVAL, "swin_attention: bad arguments");
  if ((out_hi == nullptr) != (out_lo == nullptr) || (!out_hi && !out_f32) || (out_hi && out_f32))
    WXF_FAIL(WXF_EINVAL, "swin_attention: give either the plane pair or the fp32 output");
  const int dh = d / heads, L = ws_h * ws_w;
  if (ws_h <= 0 || ws_w <= 0 || H % ws_h || W % ws_w) WXF_FAIL(WXF_EINVAL, "swin_attention: grid %dx%d vs window %dx%d", H, W, ws_h, ws_w);
  if (L > 64) WXF_FAIL(WXF_EUNSUPPORTED, "swin_attention: %d tokens per window > 64", L);
  if (dh % 4 || ldq % 4 || ldh % 4 || ldq < 3 * d || ldh < d || !wxf_aligned16(qkv))
    WXF_FAIL(WXF_EALIGN, "swin_attention: head dim / strides must be multiples of 4, qkv 16-byte aligned");
  if (shift_h < 0 || shift_w < 0 || shift_h >= ws_h || shift_w >= ws_w) WXF_FAIL(WXF_EINVAL, "swin_attention: bad shift");
  if (mask_shift_h < 0) mask_shift_h = shift_h;
  if (mask_shift_h >= ws_h) WXF_FAIL(WXF_EINVAL, "swin_attention: bad mask shift");
  const size_t smem = (size_t)(3 * L * (dh + 4) + L * ((L + 4) & ~3)) * 4 + (size_t)((L + 1) & ~1) * 4 + (size_t)L * 8 + (size_t)2 * L * 4;
  if (smem > 227 * 1024) WXF_FAIL(WXF_EUNSUPPORTED, "swin_attention: window %d x head dim %d needs %zu bytes of shared memory", L, dh, smem);
  static WxfPerDevice<size_t> attr_pd;
  size_t& have = attr_pd.get();  // function attributes are per device
  if (smem > 48 * 1024 && smem > have) {
    cudaError_t e = cudaFuncSetAttribute(swin_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) WXF_FAIL((int)e, "swin_attention: cannot opt in to %zu bytes of shared memory", smem);
    have = smem;
  }
  const int nwy = H / ws_h, nwx = W / ws_w;
  const int64_t nwin = (int64_t)B * nwy * nwx;
  if (nwin > INT32_MAX || heads > 65535) WXF_FAIL(WXF_EINVAL, "swin_attention: grid too large");
  wxf_launch(swin_attention_kernel, dim3((unsigned)nwin, (unsigned)heads), dim3(SW_THREADS), smem, (cudaStream_t)stream, qkv, ldq,
             bias, logit_scale, (__half*)out_hi, (__half*)out_lo, out_f32, ldh, H, W, d, dh, ws_h, ws_w, shift_h, shift_w, nwx,
             nwy, mask_shift_h);
  WXF_CHECK_LAUNCH("swin_attention");
  return 0;
}

extern "C" int wxf_gather_rows_ex(const float* src, int ld_src, const int32_t* idx, float* dst, int ld_dst, void* hi, void* lo,
                                  int ldh, int h_off, int64_t n, int d, void* stream) {
  if (!src || !idx || n <= 0 || d <= 0 || (d & 3) || (ld_src & 3) || !wxf_aligned16(src))
    WXF_FAIL(WXF_EINVAL, "gather_rows_ex: bad arguments (d, strides multiples of 4; 16-byte aligned)");
  if (!dst && !hi) WXF_FAIL(WXF_EINVAL, "gather_rows_ex: no output");
  if ((hi == nullptr) != (lo == nullptr)) WXF_FAIL(WXF_EINVAL, "gather_rows_ex: hi/lo come together");
  if (dst && ((ld_dst & 3) || ld_dst < d || !wxf_aligned16(dst))) WXF_FAIL(WXF_EALIGN, "gather_rows_ex: dst stride/alignment");
  if (hi && ((ldh & 3) || (h_off & 3) || ldh < h_off + d || ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7)))
    WXF_FAIL(WXF_EALIGN, "gather_rows_ex: plane stride/alignment");
  const int64_t total = n * (d >> 2);
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  wxf_launch(gather_rows_ex_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, src, ld_src, idx, dst, ld_dst,
             hi ? (__half*)hi + h_off : nullptr, lo ? (__half*)lo + h_off : nullptr, ldh, n, d >> 2);
  WXF_CHECK_LAUNCH("gather_rows_ex");
  return 0;
}

extern "C" int wxf_unpatchify_unpad_resize_to_nchw(const float* y, float* out, int B, int C, int cp, int Lat, int Lon, int ph,
                                                   int pw, int top, int left, int Hc, int Wc, int Ho, int Wo, int o0,
                                                   int n_out, int lat0, void* stream) {
  if (!y || !out || B <= 0 || C <= 0 || cp < C || Lat <= 0 || Lon <= 0 || ph <= 0 || pw <= 0 || Hc <= 0 || Wc <= 0 || Ho <= 0 ||
      Wo <= 0 || top < 0 || left < 0 || left + Wc > Lon * pw)
    WXF_FAIL(WXF_EINVAL, "unpatchify_resize: bad dims");
  if (o0 < 0 || n_out < 0 || o0 + n_out > Ho) WXF_FAIL(WXF_EINVAL, "unpatchify_resize: rows [%d, %d) outside [0, %d)", o0, o0 + n_out, Ho);
  if (n_out == 0) return 0;
  const float sh = (float)Hc / (float)Ho, sw = (float)Wc / (float)Wo;
  {
    // the source rows the requested output rows touch must lie inside the buffer, which holds patch rows
    // [lat0, lat0 + Lat) (lat0 = 0 and Lat = the whole grid, or a latitude band with its halo rows)
    auto src_rows = [&](int dst, int& i0, int& i1) {
      float src = fmaf(sh, (float)dst + 0.5f, -0.5f);
      if (src < 0.f) src = 0.f;
      i0 = (int)src;
      if (i0 > Hc - 1) i0 = Hc - 1;
      i1 = i0 + ((i0 < Hc - 1) ? 1 : 0);
    };
    int a0, a1, b0, b1;
    src_rows(o0, a0, a1);
    src_rows(o0 + n_out - 1, b0, b1);
    if (a0 + top < lat0 * ph || b1 + top >= (lat0 + Lat) * ph)
      WXF_FAIL(WXF_EINVAL, "unpatchify_resize: output rows [%d, %d) read source rows [%d, %d] outside patch rows [%d, %d)", o0,
               o0 + n_out, a0 + top, b1 + top, lat0, lat0 + Lat);
  }
  const int cgroups = (C + 31) / 32;
  if (n_out > 65535 || (int64_t)B * cgroups > 65535) WXF_FAIL(WXF_EINVAL, "unpatchify_resize: grid too large");
  dim3 grid((Wo + 31) / 32, n_out, B * cgroups), block(32, 8);
  wxf_launch(unpatchify_resize_kernel, grid, block, 0, (cudaStream_t)stream, y, out, C, cp, Lat, Lon, ph, pw, top, left, Hc, Wc,
             Ho, Wo, sh, sw, cgroups, o0, lat0);
  WXF_CHECK_LAUNCH("unpatchify_resize");
  return 0;
}
