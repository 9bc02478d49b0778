// HBM-bound passes of the forecast step: boundary pad (+NCHW->pixel-major), channel LayerNorm,
// GroupNorm+SiLU, un-pad + bilinear resize (+pixel-major->NCHW), rollout channel copy.
// Each is one coalesced read and one coalesced write of its tensor; see DESIGN.md for the byte counts.
#include <stdlib.h>

#include "wxf_common.cuh"

thread_local char wxf_err_buf[512] = "";

bool wxf_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" int wxf_abi_version(void) { return WXF_ABI_VERSION; }
extern "C" const char* wxf_last_error(void) { return wxf_err_buf; }

// ------------------------------------------------------------------------------------------------
// boundary pad + transpose.  Index map of TensorPadding._earth_padding (boundary_padding.py:50-72):
// padded (r, c): j = (c - pl) mod W; r < pt -> x[pt-1-r, (j - W/2) mod W]; pt <= r < pt+H -> x[r-pt, j];
// else x[H-1-(r-pt-H), (j - W/2) mod W].  mirror (:98-117): circular lon, reflect lat (no edge repeat).

template <bool SPLIT>
__global__ void __launch_bounds__(256) pad_to_pixel_major_kernel(const float* __restrict__ x, float* __restrict__ xp,
                                                                  __half* __restrict__ xp_hi, __half* __restrict__ xp_lo,
                                                                  int CT, int H, int W, int pt, int pl, int mode,
                                                                  int ld, int Hp, int Wp, int cgroups, int row0) {
  __shared__ float tile[32][33];  // [channel][column]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int b = blockIdx.z / cgroups, cg = blockIdx.z % cgroups;
  const int r = row0 + blockIdx.y;
  const int c0 = blockIdx.x * 32;

  int sr;
  bool roll = false;
  if (mode == WXF_PAD_EARTH) {
    if (r < pt) {
      sr = pt - 1 - r;
      roll = true;
    } else if (r < pt + H) {
      sr = r - pt;
    } else {
      sr = H - 1 - (r - pt - H);
      roll = true;
    }
  } else {
    sr = r - pt;
    if (sr < 0) sr = -sr;
    if (sr >= H) sr = 2 * (H - 1) - sr;
  }

  const int col = c0 + tx;
  int sc = -1;
  if (col < Wp) {
    int j = (col - pl) % W;
    if (j < 0) j += W;
    if (roll) {
      j -= W / 2;
      if (j < 0) j += W;
    }
    sc = j;
  }
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int ch = cg * 32 + i;
    float v = 0.f;
    if (ch < CT && sc >= 0) v = __ldg(x + ((size_t)(b * CT + ch) * H + sr) * W + sc);
    tile[i][tx] = v;
  }
  __syncthreads();
  const int ch = cg * 32 + tx;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int cw = c0 + i;
    if (cw < Wp && ch < ld) {
      const size_t o = ((size_t)(b * Hp + r) * Wp + cw) * ld + ch;
      if constexpr (SPLIT) {
        __half hi, lo;
        wxf_split_f16x2(tile[tx][i], hi, lo);
        xp_hi[o] = hi;
        xp_lo[o] = lo;
      } else {
        xp[o] = tile[tx][i];
      }
    }
  }
}

// Same index map, one CTA = one padded row x 64 columns x a block of 64 channels: the column loop reads whole 128-byte
// runs of the NCHW rows, the transposed write stores 16 bytes per thread (8 fp16 of a plane / 4 fp32), i.e. whole
// 128-byte pixel rows.  Needs ld % 8 == 0 (otherwise the scalar kernel above runs).
// PRE: the per-step pre-blocks fused in: the source is not one tensor but a table of per-channel planes
// (ConcatToTensor, credit/preblock/concat.py:101-207: torch.cat of the sorted variables along the channel axis) and
// every value is z-scored on the way, (x - mean[c]) / max(std[c], 1e-12) (ERA5Normalizer, credit/preblock/norm.py:80-98).
struct PadPre {
  const float* const* chan;  // [B * C] device pointers to [T, H, W] planes (variable-major; nullptr: read x)
  const float* mean;         // [C] (0 for pass-through variables)
  const float* stdv;         // [C] (1 for pass-through variables)
  int T;
};

template <bool SPLIT, bool PRE = false>
__global__ void __launch_bounds__(256) pad_rows_vec_kernel(const float* __restrict__ x, float* __restrict__ xp,
                                                            __half* __restrict__ xp_hi, __half* __restrict__ xp_lo,
                                                            int CT, int H, int W, int pt, int pl, int mode, int ld,
                                                            int Hp, int Wp, int cblocks, int row0, PadPre pre) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  __shared__ float tile[64][65];  // [channel][column]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z / cblocks, cb = blockIdx.z % cblocks;
  const int r = row0 + blockIdx.y;
  const int c0 = blockIdx.x * 64;
  const int ch0 = cb * 64;
  const int nch = min(64, ld - ch0);

  int sr;
  bool roll = false;
  if (mode == WXF_PAD_EARTH) {
    if (r < pt) {
      sr = pt - 1 - r;
      roll = true;
    } else if (r < pt + H) {
      sr = r - pt;
    } else {
      sr = H - 1 - (r - pt - H);
      roll = true;
    }
  } else {
    sr = r - pt;
    if (sr < 0) sr = -sr;
    if (sr >= H) sr = 2 * (H - 1) - sr;
  }
  int sc[2];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int col = c0 + hf * 32 + lane;
    sc[hf] = -1;
    if (col < Wp) {
      int j = (col - pl) % W;
      if (j < 0) j += W;
      if (roll) {
        j -= W / 2;
        if (j < 0) j += W;
      }
      sc[hf] = j;
    }
  }
  for (int i = warp; i < nch; i += 8) {
    const int ch = ch0 + i;
    if constexpr (PRE) {
      const int c = ch / pre.T, t = ch - c * pre.T;
      const bool ok = ch < CT;
      const float* row = ok ? pre.chan[(size_t)b * (CT / pre.T) + c] + ((size_t)t * H + sr) * W : nullptr;
      const float m = ok ? __ldg(pre.mean + c) : 0.f, sd = ok ? fmaxf(__ldg(pre.stdv + c), 1e-12f) : 1.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) tile[i][hf * 32 + lane] = (ok && sc[hf] >= 0) ? (__ldg(row + sc[hf]) - m) / sd : 0.f;
    } else {
      const float* row = x + ((size_t)(b * CT + ch) * H + sr) * W;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) tile[i][hf * 32 + lane] = (ch < CT && sc[hf] >= 0) ? __ldg(row + sc[hf]) : 0.f;
    }
  }
  __syncthreads();
  const int ngrp = nch >> 3;  // groups of 8 channels
  for (int idx = tid; idx < 64 * ngrp; idx += 256) {
    const int cg = idx % ngrp, colx = idx / ngrp;
    const int cw = c0 + colx;
    if (cw >= Wp) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[cg * 8 + e][colx];
    const size_t o = ((size_t)(b * Hp + r) * Wp + cw) * ld + ch0 + cg * 8;
    if constexpr (SPLIT) {
      __align__(16) __half2 h8[4];
      __align__(16) __half2 l8[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) wxf_split2_f16x2(v[2 * e], v[2 * e + 1], h8[e], l8[e]);
      *reinterpret_cast<uint4*>(xp_hi + o) = *reinterpret_cast<const uint4*>(h8);
      *reinterpret_cast<uint4*>(xp_lo + o) = *reinterpret_cast<const uint4*>(l8);
    } else {
      *reinterpret_cast<float4*>(xp + o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(xp + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

static int pad_launch(const float* x, float* xp, void* xp_hi, void* xp_lo, int B, int C, int T, int H, int W, int pt,
                      int pb, int pl, int pr, int mode, int ld, int row0, int nrows, void* stream) {
  if (B <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0 || pt < 0 || pb < 0 || pl < 0 || pr < 0)
    WXF_FAIL(WXF_EINVAL, "pad: bad dims");
  if (ld < C * T) WXF_FAIL(WXF_EINVAL, "pad: ld %d < C*T %d", ld, C * T);
  if (mode != WXF_PAD_EARTH && mode != WXF_PAD_MIRROR) WXF_FAIL(WXF_EINVAL, "pad: bad mode %d", mode);
  if (mode == WXF_PAD_EARTH && (pt > H || pb > H)) WXF_FAIL(WXF_EINVAL, "pad: earth pad_lat larger than H");
  if (mode == WXF_PAD_MIRROR && (pt >= H || pb >= H)) WXF_FAIL(WXF_EINVAL, "pad: mirror pad_lat must be < H");
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  if (row0 < 0 || nrows < 0 || row0 + nrows > Hp) WXF_FAIL(WXF_EINVAL, "pad: rows [%d, %d) outside [0, %d)", row0, row0 + nrows, Hp);
  if (nrows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (ld % 8 == 0) && wxf_aligned16(xp_hi ? xp_hi : (void*)xp) && (!xp_lo || wxf_aligned16(xp_lo));
  if (vec) {
    const int cblocks = (ld + 63) / 64;
    if (nrows > 65535 || (int64_t)B * cblocks > 65535) WXF_FAIL(WXF_EINVAL, "pad: grid too large");
    dim3 grid((Wp + 63) / 64, nrows, B * cblocks);
    if (xp_hi)
      wxf_launch(pad_rows_vec_kernel<true>, dim3(grid), dim3(256), 0, st, x, nullptr, (__half*)xp_hi, (__half*)xp_lo, C * T, H, W, pt, pl, mode,
                                                      ld, Hp, Wp, cblocks, row0, PadPre{});
    else
      wxf_launch(pad_rows_vec_kernel<false>, dim3(grid), dim3(256), 0, st, x, xp, nullptr, nullptr, C * T, H, W, pt, pl, mode, ld, Hp, Wp,
                                                       cblocks, row0, PadPre{});
    WXF_CHECK_LAUNCH("pad_to_pixel_major");
    return 0;
  }
  const int cgroups = (ld + 31) / 32;
  if (nrows > 65535 || (int64_t)B * cgroups > 65535) WXF_FAIL(WXF_EINVAL, "pad: grid too large");
  dim3 grid((Wp + 31) / 32, nrows, B * cgroups), block(32, 8);
  if (xp_hi)
    pad_to_pixel_major_kernel<true><<<grid, block, 0, st>>>(x, nullptr, (__half*)xp_hi, (__half*)xp_lo, C * T, H, W, pt,
                                                            pl, mode, ld, Hp, Wp, cgroups, row0);
  else
    pad_to_pixel_major_kernel<false><<<grid, block, 0, st>>>(x, xp, nullptr, nullptr, C * T, H, W, pt, pl, mode, ld, Hp,
                                                             Wp, cgroups, row0);
  WXF_CHECK_LAUNCH("pad_to_pixel_major");
  return 0;
}

// Pre-blocks + padding in one pass: per-variable planes in, z-scored padded pixel-major planes out (f2).
extern "C" int wxf_preblock_pad_to_pixel_major(const void* const* chan_planes, const float* mean, const float* stdv, float* xp,
                                               void* xp_hi, void* xp_lo, int B, int C, int T, int H, int W, int pt, int pb,
                                               int pl, int pr, int mode, int ld, int row0, int nrows, void* stream) {
  if (!chan_planes || !mean || !stdv) WXF_FAIL(WXF_EINVAL, "preblock_pad: null table");
  if ((xp_hi == nullptr) != (xp_lo == nullptr) || (!xp && !xp_hi) || (xp && xp_hi))
    WXF_FAIL(WXF_EINVAL, "preblock_pad: give either the fp32 output or the plane pair");
  if (B <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0 || pt < 0 || pb < 0 || pl < 0 || pr < 0) WXF_FAIL(WXF_EINVAL, "preblock_pad: bad dims");
  if (ld < C * T || (ld & 7)) WXF_FAIL(WXF_EINVAL, "preblock_pad: ld %d must be a multiple of 8 and >= C*T %d", ld, C * T);
  if (mode != WXF_PAD_EARTH && mode != WXF_PAD_MIRROR) WXF_FAIL(WXF_EINVAL, "preblock_pad: bad mode %d", mode);
  if (mode == WXF_PAD_EARTH && (pt > H || pb > H)) WXF_FAIL(WXF_EINVAL, "preblock_pad: earth pad_lat larger than H");
  if (mode == WXF_PAD_MIRROR && (pt >= H || pb >= H)) WXF_FAIL(WXF_EINVAL, "preblock_pad: mirror pad_lat must be < H");
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  if (row0 < 0 || nrows < 0 || row0 + nrows > Hp) WXF_FAIL(WXF_EINVAL, "preblock_pad: rows outside the padded image");
  if (nrows == 0) return 0;
  if (!wxf_aligned16(xp_hi ? xp_hi : (void*)xp) || (xp_lo && !wxf_aligned16(xp_lo))) WXF_FAIL(WXF_EALIGN, "preblock_pad: output alignment");
  const int cblocks = (ld + 63) / 64;
  if (nrows > 65535 || (int64_t)B * cblocks > 65535) WXF_FAIL(WXF_EINVAL, "preblock_pad: grid too large");
  PadPre pre{reinterpret_cast<const float* const*>(chan_planes), mean, stdv, T};
  dim3 grid((Wp + 63) / 64, nrows, B * cblocks);
  cudaStream_t st = (cudaStream_t)stream;
  if (xp_hi)
    wxf_launch(pad_rows_vec_kernel<true, true>, dim3(grid), dim3(256), 0, st, (const float*)nullptr, (float*)nullptr, (__half*)xp_hi,
               (__half*)xp_lo, C * T, H, W, pt, pl, mode, ld, Hp, Wp, cblocks, row0, pre);
  else
    wxf_launch(pad_rows_vec_kernel<false, true>, dim3(grid), dim3(256), 0, st, (const float*)nullptr, xp, (__half*)nullptr,
               (__half*)nullptr, C * T, H, W, pt, pl, mode, ld, Hp, Wp, cblocks, row0, pre);
  WXF_CHECK_LAUNCH("preblock_pad_to_pixel_major");
  return 0;
}

extern "C" int wxf_pad_to_pixel_major(const float* x, float* xp, int B, int C, int T, int H, int W, int pt, int pb,
                                      int pl, int pr, int mode, int ld, int row0, int nrows, void* stream) {
  if (!xp) WXF_FAIL(WXF_EINVAL, "pad: null output");
  return pad_launch(x, xp, nullptr, nullptr, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows, stream);
}

extern "C" int wxf_pad_to_pixel_major_f16x2(const float* x, void* xp_hi, void* xp_lo, int B, int C, int T, int H, int W,
                                            int pt, int pb, int pl, int pr, int mode, int ld, int row0, int nrows,
                                            void* stream) {
  if (!xp_hi || !xp_lo) WXF_FAIL(WXF_EINVAL, "pad: null output planes");
  return pad_launch(x, nullptr, xp_hi, xp_lo, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows, stream);
}

// ------------------------------------------------------------------------------------------------
// channel LayerNorm (crossformer.py:182-192): one warp per pixel, row cached in registers.

template <int NV, bool SPLIT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                         int ldy, __half* __restrict__ y_hi, __half* __restrict__ y_lo,
                                                         const float* __restrict__ g, const float* __restrict__ bta,
                                                         int64_t M, int d, float eps) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * ldx;
  float v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    v[k] = (c < d) ? xr[c] : 0.f;
    s += v[k];
  }
  const float mean = wxf_warp_sum(s) / (float)d;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    const float t = (c < d) ? v[k] - mean : 0.f;
    ss += t * t;
  }
  const float var = wxf_warp_sum(ss) / (float)d;
  const float den = sqrtf(var + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    if (c < d) {
      const float o = (v[k] - mean) / den * __ldg(g + c) + __ldg(bta + c);
      if constexpr (SPLIT) {
        __half hi, lo;
        wxf_split_f16x2(o, hi, lo);
        y_hi[row * ldy + c] = hi;
        y_lo[row * ldy + c] = lo;
      } else {
        y[row * ldy + c] = o;
      }
    }
  }
}

// vectorised variant for d % 128 == 0: each lane owns float4 chunks, 16-byte loads, 8/16-byte stores.  A warp works on RPW
// rows at once (all their loads are issued before the first reduction: with one row per warp and d = 128 a lane has a single
// 16-byte load in flight, 32 KB per SM, below what hides the HBM latency), and the row's 1 / sqrt(var + eps) is computed once
// (one IEEE division per row instead of four per lane and chunk).
template <int NV4, bool SPLIT, int RPW>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                             int ldy, __half* __restrict__ y_hi, __half* __restrict__ y_lo,
                                                             const float* __restrict__ g, const float* __restrict__ bta,
                                                             int64_t M, float eps) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  constexpr int d = NV4 * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (row0 >= M) return;
  float4 v[RPW][NV4];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + r < M ? row0 + r : M - 1;  // a ragged last warp re-reads the last row and does not store it
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
#pragma unroll
    for (int k = 0; k < NV4; ++k) v[r][k] = xr[lane + 32 * k];
  }
  float mean[RPW], inv[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) s += (v[r][k].x + v[r][k].y) + (v[r][k].z + v[r][k].w);
    mean[r] = wxf_warp_sum(s) / (float)d;
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const float a = v[r][k].x - mean[r], b = v[r][k].y - mean[r], c = v[r][k].z - mean[r], e = v[r][k].w - mean[r];
      ss += (a * a + b * b) + (c * c + e * e);
    }
    const float var = wxf_warp_sum(ss) / (float)d;
    inv[r] = 1.0f / sqrtf(var + eps);
  }
#pragma unroll
  for (int k = 0; k < NV4; ++k) {
    const int c4 = lane + 32 * k;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + c4), bb = __ldg(reinterpret_cast<const float4*>(bta) + c4);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int64_t row = row0 + r;
      if (row >= M) break;
      float4 o;
      o.x = (v[r][k].x - mean[r]) * inv[r] * gg.x + bb.x;
      o.y = (v[r][k].y - mean[r]) * inv[r] * gg.y + bb.y;
      o.z = (v[r][k].z - mean[r]) * inv[r] * gg.z + bb.z;
      o.w = (v[r][k].w - mean[r]) * inv[r] * gg.w + bb.w;
      if constexpr (SPLIT) {
        __align__(8) __half2 h2[2];
        __align__(8) __half2 l2[2];
        wxf_split2_f16x2(o.x, o.y, h2[0], l2[0]);
        wxf_split2_f16x2(o.z, o.w, h2[1], l2[1]);
        *reinterpret_cast<uint2*>(y_hi + row * ldy + 4 * c4) = *reinterpret_cast<const uint2*>(h2);
        *reinterpret_cast<uint2*>(y_lo + row * ldy + 4 * c4) = *reinterpret_cast<const uint2*>(l2);
      } else {
        *reinterpret_cast<float4*>(y + row * ldy + 4 * c4) = o;
      }
    }
  }
}

template <bool SPLIT>
static int layernorm_launch(const float* x, int ldx, float* y, int ldy, void* y_hi, void* y_lo, const float* g,
                            const float* b, int64_t M, int d, float eps, void* stream) {
  if (M <= 0 || d <= 0 || ldx < d || ldy < d) WXF_FAIL(WXF_EINVAL, "layernorm: bad dims");
  if (d > 1024) WXF_FAIL(WXF_EUNSUPPORTED, "layernorm: d=%d > 1024", d);
  const int nv = (d + 31) / 32;
  const unsigned blocks = (unsigned)((M + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  // rows per warp of the vectorised kernel: 4 for d = 128, 2 up to d = 512 (register budget: RPW * d / 32 floats per lane)
  auto vec_blocks = [&](int rpw) { return dim3((unsigned)((M + 8 * rpw - 1) / (8 * rpw))); };
  const bool vec = (d % 128 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && wxf_aligned16(x) && wxf_aligned16(g) &&
                   wxf_aligned16(b) && (SPLIT ? ((reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) % 8 == 0)
                                              : wxf_aligned16(y));
  if (vec) {
#define LN_VEC(NV4, RPW)                                                                                              \
  if (d == NV4 * 128) {                                                                                               \
    wxf_launch(layernorm_vec_kernel<NV4, SPLIT, RPW>, vec_blocks(RPW), dim3(256), 0, st, x, ldx, y, ldy, (__half*)y_hi, (__half*)y_lo, g, b, M, eps); \
    WXF_CHECK_LAUNCH("layernorm");                                                                                    \
    return 0;                                                                                                         \
  }
    LN_VEC(1, 4) LN_VEC(2, 2) LN_VEC(3, 2) LN_VEC(4, 2) LN_VEC(6, 1) LN_VEC(8, 1)
#undef LN_VEC
  }
#define LN_CASE(NV)                                                                                         \
  if (nv <= NV) {                                                                                           \
    wxf_launch(layernorm_kernel<NV, SPLIT>, dim3(blocks), dim3(256), 0, st, x, ldx, y, ldy, (__half*)y_hi, (__half*)y_lo, g, b, M, d, eps); \
    WXF_CHECK_LAUNCH("layernorm");                                                                          \
    return 0;                                                                                               \
  }
  LN_CASE(1) LN_CASE(2) LN_CASE(4) LN_CASE(8) LN_CASE(16) LN_CASE(32)
#undef LN_CASE
  return WXF_EUNSUPPORTED;
}

extern "C" int wxf_layernorm(const float* x, int ldx, float* y, int ldy, const float* g, const float* b, int64_t M,
                             int d, float eps, void* stream) {
  return layernorm_launch<false>(x, ldx, y, ldy, nullptr, nullptr, g, b, M, d, eps, stream);
}

extern "C" int wxf_layernorm_f16x2(const float* x, int ldx, void* y_hi, void* y_lo, int ldh, const float* g,
                                   const float* b, int64_t M, int d, float eps, void* stream) {
  if (!y_hi || !y_lo) WXF_FAIL(WXF_EINVAL, "layernorm_f16x2: null planes");
  return layernorm_launch<true>(x, ldx, nullptr, ldh, y_hi, y_lo, g, b, M, d, eps, stream);
}

__global__ void __launch_bounds__(256) split_f16x2_kernel(const float* __restrict__ x, int ldx, __half* __restrict__ hi,
                                                           __half* __restrict__ lo, int ldh, int64_t M, int d) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int64_t total = M * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d;
    const int c = (int)(i - r * d);
    __half h, l;
    wxf_split_f16x2(x[r * ldx + c], h, l);
    hi[r * ldh + c] = h;
    lo[r * ldh + c] = l;
  }
}

extern "C" int wxf_split_f16x2(const float* x, int ldx, void* hi, void* lo, int ldh, int64_t M, int d, void* stream) {
  if (M <= 0 || d <= 0 || ldx < d || ldh < d || !hi || !lo) WXF_FAIL(WXF_EINVAL, "split_f16x2: bad args");
  int64_t blocks = (M * d + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wxf_launch(split_f16x2_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, ldx, (__half*)hi, (__half*)lo, ldh, M, d);
  WXF_CHECK_LAUNCH("split_f16x2");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm + SiLU (nn.GroupNorm(num_groups, C) + nn.SiLU, crossformer.py:96-100).
// Pass 1: per-block partial (sum, sumsq) per group -> scratch;  pass 2: fp64 combine -> (mean, rstd);
// pass 3: normalise + affine + SiLU (+ UpBlock shortcut).

// Pixels per statistics block: enough blocks to fill the GPU (>= ~4 per SM) but at least 32 pixels each.
static inline int gn_pix_per_block(int64_t HW) {
  int64_t ppb = HW / 592;
  if (ppb < 32) ppb = 32;
  if (ppb > 512) ppb = 512;
  return (int)ppb;
}

// partial (sum, sumsq) per group for a chunk of pixels; C % 4 == 0 and C/4 divides 256: float4 loads
__global__ void __launch_bounds__(256) gn_partial_vec_kernel(const float* __restrict__ x, int ldx, float2* __restrict__ part,
                                                              int64_t HW, int C, int G, int nchunk, int ppb) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  __shared__ float4 rs[256], rq[256];
  const int tid = threadIdx.x;
  const int TC = C >> 2;          // float4 columns
  const int lanes = 256 / TC;     // pixel lanes
  const int tc = tid % TC, pl = tid / TC;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t p0 = (int64_t)chunk * ppb;
  int64_t p1 = p0 + ppb;
  if (p1 > HW) p1 = HW;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  const float* xb = x + (int64_t)b * HW * ldx + 4 * tc;
#pragma unroll 4
  for (int64_t p = p0 + pl; p < p1; p += lanes) {  // unrolled: four independent 16-byte loads in flight per thread
    const float4 v = *reinterpret_cast<const float4*>(xb + p * ldx);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
  }
  rs[tid] = s;
  rq[tid] = q;
  __syncthreads();
  const int cpg = C / G;
  const float* fs = reinterpret_cast<const float*>(rs);
  const float* fq = reinterpret_cast<const float*>(rq);
  for (int gidx = tid; gidx < G; gidx += 256) {
    float ts = 0.f, tq = 0.f;
    for (int cc = 0; cc < cpg; ++cc) {
      const int c = gidx * cpg + cc;
      for (int l = 0; l < lanes; ++l) {
        ts += fs[(l * TC + (c >> 2)) * 4 + (c & 3)];
        tq += fq[(l * TC + (c >> 2)) * 4 + (c & 3)];
      }
    }
    part[((int64_t)b * nchunk + chunk) * G + gidx] = make_float2(ts, tq);
  }
}

__global__ void __launch_bounds__(256) gn_partial_kernel(const float* __restrict__ x, int ldx, float2* __restrict__ part,
                                                          int64_t HW, int C, int G, int nchunk, int ppb) {
  // channel slot of this thread: TC = min(C, 256) channels in flight, 256/TC pixel lanes
  __shared__ float rs[1024], rq[1024];
  const int tid = threadIdx.x;
  const int TC = C < 256 ? C : 256;
  const int lanes = 256 / TC;  // TC divides 256 (checked on host) or TC == 256
  const int KS = C / TC;       // channel slots per thread (<= 4)
  const int tc = tid % TC, pl = tid / TC;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t p0 = (int64_t)chunk * ppb;
  int64_t p1 = p0 + ppb;
  if (p1 > HW) p1 = HW;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    for (int64_t p = p0 + pl; p < p1; p += lanes) {
      const float* xr = x + ((int64_t)b * HW + p) * ldx;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < KS) {
          const float v = xr[tc + k * TC];
          s[k] += v;
          q[k] += v * v;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rs[tid + 256 * k] = s[k];
    rq[tid + 256 * k] = q[k];
  }
  __syncthreads();
  // per-channel totals across pixel lanes: channel c = tc + k*TC lives at rs[pl*TC + tc + 256*k]
  const int cpg = C / G;
  for (int gidx = tid; gidx < G; gidx += 256) {
    float ts = 0.f, tq = 0.f;
    for (int cc = 0; cc < cpg; ++cc) {
      const int c = gidx * cpg + cc;
      const int k = c / TC, t = c % TC;
      for (int l = 0; l < lanes; ++l) {
        ts += rs[l * TC + t + 256 * k];
        tq += rq[l * TC + t + 256 * k];
      }
    }
    part[((int64_t)b * nchunk + chunk) * G + gidx] = make_float2(ts, tq);
  }
}

// one warp per (image, group): lanes stride over the chunk partials, fp64 combine
__global__ void gn_finalize_kernel(const float2* __restrict__ part, float* __restrict__ stats, int B, int G, int nchunk,
                                   double count, float eps) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= B * G) return;
  const int b = i / G, g = i % G;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < nchunk; c += 32) {
    const float2 p = part[((int64_t)b * nchunk + c) * G + g];
    s += (double)p.x;
    q += (double)p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(256) gn_silu_kernel(const float* __restrict__ x, int ldx,
                                                       const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ res,
                                                       int ldr, float* __restrict__ y, __half* __restrict__ y_hi,
                                                       __half* __restrict__ y_lo, int ldy, int64_t HW, int C, int cpg,
                                                       int G, int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = idx / C;
    const int c = (int)(idx - pix * C);
    const int b = (int)(pix / HW);
    const int g = c / cpg;
    const float mean = __ldg(stats + 2 * (b * G + g)), rstd = __ldg(stats + 2 * (b * G + g) + 1);
    float v = (x[pix * ldx + c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    v = wxf_silu(v);
    if (res) v += res[pix * ldr + c];
    if constexpr (SPLIT) {
      __half hi, lo;
      wxf_split_f16x2(v, hi, lo);
      y_hi[pix * ldy + c] = hi;
      y_lo[pix * ldy + c] = lo;
    } else {
      y[pix * ldy + c] = v;
    }
  }
}

// vectorised apply: one image per blockIdx.y, per-channel scale/shift staged in shared memory, float4 in / out
template <bool SPLIT>
__global__ void __launch_bounds__(256) gn_silu_vec_kernel(const float* __restrict__ x, int ldx,
                                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ res,
                                                           int ldr, float* __restrict__ y, __half* __restrict__ y_hi,
                                                           __half* __restrict__ y_lo, int ldy, int64_t HW, int C, int cpg,
                                                           int G) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  __shared__ __align__(16) float sa[1024], sb[1024];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    const float mean = stats[2 * (b * G + g)], rstd = stats[2 * (b * G + g) + 1];
    sa[c] = rstd * gamma[c];
    sb[c] = beta[c];
  }
  __shared__ float smean[1024];
  for (int c = threadIdx.x; c < C; c += 256) smean[c] = stats[2 * (b * G + c / cpg)];
  __syncthreads();
  // a thread keeps its float4 channel column and walks pixels (C <= 1024: C/4 <= 256 columns, 256 / (C/4) pixels per CTA
  // pass), so there is no integer division per element; sigmoid = 1 / (1 + 2^(-x log2 e)) with the approximate ex2 / rcp
  // units (2-3 ulp, far inside the decoder's error budget) instead of expf + an IEEE division
  const int C4 = C >> 2;
  const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4;
  const int ppc = 256 / C4;                        // pixels per CTA pass; threads beyond ppc * C4 idle (C/4 not dividing 256)
  auto act = [](float t) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return t * r;
  };
  if (prow < ppc) {
    const int c = 4 * c4;
    const float4 a = *reinterpret_cast<const float4*>(sa + c), bb = *reinterpret_cast<const float4*>(sb + c);
    const float4 mm = *reinterpret_cast<const float4*>(smean + c);
#pragma unroll 2
    for (int64_t pl = (int64_t)blockIdx.x * ppc + prow; pl < HW; pl += (int64_t)gridDim.x * ppc) {
      const int64_t pix = (int64_t)b * HW + pl;
      const float4 v = *reinterpret_cast<const float4*>(x + pix * ldx + c);
      float4 o;
      o.x = act((v.x - mm.x) * a.x + bb.x);
      o.y = act((v.y - mm.y) * a.y + bb.y);
      o.z = act((v.z - mm.z) * a.z + bb.z);
      o.w = act((v.w - mm.w) * a.w + bb.w);
      if (res) {
        const float4 r = *reinterpret_cast<const float4*>(res + pix * ldr + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if constexpr (SPLIT) {
        __align__(8) __half2 h2[2];
        __align__(8) __half2 l2[2];
        wxf_split2_f16x2(o.x, o.y, h2[0], l2[0]);
        wxf_split2_f16x2(o.z, o.w, h2[1], l2[1]);
        *reinterpret_cast<uint2*>(y_hi + pix * ldy + c) = *reinterpret_cast<const uint2*>(h2);
        *reinterpret_cast<uint2*>(y_lo + pix * ldy + c) = *reinterpret_cast<const uint2*>(l2);
      } else {
        *reinterpret_cast<float4*>(y + pix * ldy + c) = o;
      }
    }
  }
}

extern "C" int64_t wxf_groupnorm_scratch_bytes(int B, int64_t HW, int C) {
  const int ppb = gn_pix_per_block(HW);
  const int64_t nchunk = (HW + ppb - 1) / ppb;
  return (int64_t)B * nchunk * C * (int64_t)sizeof(float2);  // G <= C
}

extern "C" int wxf_groupnorm_stats(const float* x, int ldx, float* stats, void* scratch, int B, int64_t HW, int C,
                                   int G, float eps, void* stream) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G) WXF_FAIL(WXF_EINVAL, "groupnorm: bad dims");
  if (C > 1024 || !((C <= 256 && 256 % C == 0) || (C % 256 == 0)))
    WXF_FAIL(WXF_EUNSUPPORTED, "groupnorm: C=%d must divide 256 or be a multiple of 256 (<=1024)", C);
  const int ppb = gn_pix_per_block(HW);
  const int nchunk = (int)((HW + ppb - 1) / ppb);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 4 == 0) && (256 % (C / 4) == 0) && (ldx % 4 == 0) && wxf_aligned16(x);
  if (vec)
    wxf_launch(gn_partial_vec_kernel, dim3(nchunk, B), dim3(256), 0, st, x, ldx, (float2*)scratch, HW, C, G, nchunk, ppb);
  else
    gn_partial_kernel<<<dim3(nchunk, B), 256, 0, st>>>(x, ldx, (float2*)scratch, HW, C, G, nchunk, ppb);
  WXF_CHECK_LAUNCH("gn_partial");
  wxf_launch(gn_finalize_kernel, dim3((B * G + 3) / 4), dim3(128), 0, st, (const float2*)scratch, stats, B, G, nchunk,
                                                          (double)HW * (double)(C / G), eps);
  WXF_CHECK_LAUNCH("gn_finalize");
  return 0;
}

__global__ void gn_sums_kernel(const float2* __restrict__ part, double* __restrict__ sums, int B, int G, int nchunk) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= B * G) return;
  const int b = i / G, g = i % G;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < nchunk; c += 32) {
    const float2 p = part[((int64_t)b * nchunk + c) * G + g];
    s += (double)p.x;
    q += (double)p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    sums[2 * i] = s;
    sums[2 * i + 1] = q;
  }
}

__global__ void gn_stats_from_sums_kernel(const double* __restrict__ sums, float* __restrict__ stats, int n, double count,
                                          float eps) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = sums[2 * i] / count;
  double var = sums[2 * i + 1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

extern "C" int wxf_groupnorm_sums(const float* x, int ldx, double* sums, void* scratch, int B, int64_t HW, int C, int G,
                                  void* stream) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G) WXF_FAIL(WXF_EINVAL, "groupnorm_sums: bad dims");
  if (C > 1024 || !((C <= 256 && 256 % C == 0) || (C % 256 == 0)))
    WXF_FAIL(WXF_EUNSUPPORTED, "groupnorm: C=%d must divide 256 or be a multiple of 256 (<=1024)", C);
  const int ppb = gn_pix_per_block(HW);
  const int nchunk = (int)((HW + ppb - 1) / ppb);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 4 == 0) && (256 % (C / 4) == 0) && (ldx % 4 == 0) && wxf_aligned16(x);
  if (vec)
    wxf_launch(gn_partial_vec_kernel, dim3(nchunk, B), dim3(256), 0, st, x, ldx, (float2*)scratch, HW, C, G, nchunk, ppb);
  else
    gn_partial_kernel<<<dim3(nchunk, B), 256, 0, st>>>(x, ldx, (float2*)scratch, HW, C, G, nchunk, ppb);
  WXF_CHECK_LAUNCH("gn_partial");
  wxf_launch(gn_sums_kernel, dim3((B * G + 3) / 4), dim3(128), 0, st, (const float2*)scratch, sums, B, G, nchunk);
  WXF_CHECK_LAUNCH("gn_sums");
  return 0;
}

extern "C" int wxf_groupnorm_stats_from_sums(const double* sums, float* stats, int B, int G, double count, float eps,
                                             void* stream) {
  if (B <= 0 || G <= 0 || count <= 0) WXF_FAIL(WXF_EINVAL, "groupnorm_stats_from_sums: bad dims");
  wxf_launch(gn_stats_from_sums_kernel, dim3((B * G + 127) / 128), dim3(128), 0, (cudaStream_t)stream, sums, stats, B * G, count, eps);
  WXF_CHECK_LAUNCH("gn_stats_from_sums");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// row gather (all-to-all pack / unpack of the domain decomposition)

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int ld_src,
                                                           const int32_t* __restrict__ idx, float* __restrict__ dst,
                                                           int ld_dst, int64_t n, int d4) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int64_t total = n * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d4;
    const int c = (int)(i - r * d4) * 4;
    *reinterpret_cast<float4*>(dst + r * ld_dst + c) =
        *reinterpret_cast<const float4*>(src + (int64_t)__ldg(idx + r) * ld_src + c);
  }
}

extern "C" int wxf_gather_rows(const float* src, int ld_src, const int32_t* idx, float* dst, int ld_dst, int64_t n, int d,
                               void* stream) {
  if (n < 0 || d <= 0 || (d & 3) || (ld_src & 3) || (ld_dst & 3) || ld_src < d || ld_dst < d || !wxf_aligned16(src) ||
      !wxf_aligned16(dst))
    WXF_FAIL(WXF_EINVAL, "gather_rows: bad arguments (d, strides multiples of 4; 16-byte aligned)");
  if (n == 0) return 0;
  int64_t blocks = (n * (d / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wxf_launch(gather_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, ld_src, idx, dst, ld_dst, n, d / 4);
  WXF_CHECK_LAUNCH("gather_rows");
  return 0;
}

static int gn_silu_launch(const float* x, int ldx, const float* stats, const float* gamma, const float* beta,
                          const float* res, int ldr, float* y, void* y_hi, void* y_lo, int ldy, int B, int64_t HW,
                          int C, int G, void* stream) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G) WXF_FAIL(WXF_EINVAL, "groupnorm_silu: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 4 == 0) && C <= 1024 && (ldx % 4 == 0) && (ldy % 4 == 0) && wxf_aligned16(x) &&
                   (!res || ((ldr % 4 == 0) && wxf_aligned16(res))) && B <= 65535 &&
                   (y_hi ? ((reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) % 8 == 0) : wxf_aligned16(y));
  if (vec) {
    int64_t blocks = (HW * (C / 4) + 255) / 256;
    const int64_t cap = (148 * 16 + B - 1) / B;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)B);
    if (y_hi)
      wxf_launch(gn_silu_vec_kernel<true>, dim3(grid), dim3(256), 0, st, x, ldx, stats, gamma, beta, res, ldr, nullptr, (__half*)y_hi,
                                                     (__half*)y_lo, ldy, HW, C, C / G, G);
    else
      wxf_launch(gn_silu_vec_kernel<false>, dim3(grid), dim3(256), 0, st, x, ldx, stats, gamma, beta, res, ldr, y, nullptr, nullptr, ldy, HW,
                                                      C, C / G, G);
    WXF_CHECK_LAUNCH("gn_silu");
    return 0;
  }
  const int64_t total = (int64_t)B * HW * C;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (y_hi)
    gn_silu_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, stats, gamma, beta, res, ldr, nullptr, (__half*)y_hi,
                                                           (__half*)y_lo, ldy, HW, C, C / G, G, total);
  else
    gn_silu_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, stats, gamma, beta, res, ldr, y, nullptr, nullptr,
                                                            ldy, HW, C, C / G, G, total);
  WXF_CHECK_LAUNCH("gn_silu");
  return 0;
}

extern "C" int wxf_groupnorm_silu(const float* x, int ldx, const float* stats, const float* gamma, const float* beta,
                                  const float* res, int ldr, float* y, int ldy, int B, int64_t HW, int C, int G,
                                  void* stream) {
  if (!y) WXF_FAIL(WXF_EINVAL, "groupnorm_silu: null output");
  return gn_silu_launch(x, ldx, stats, gamma, beta, res, ldr, y, nullptr, nullptr, ldy, B, HW, C, G, stream);
}

extern "C" int wxf_groupnorm_silu_f16x2(const float* x, int ldx, const float* stats, const float* gamma,
                                        const float* beta, const float* res, int ldr, void* y_hi, void* y_lo, int ldh,
                                        int h_off, int B, int64_t HW, int C, int G, void* stream) {
  if (!y_hi || !y_lo || h_off < 0 || ldh < h_off + C) WXF_FAIL(WXF_EINVAL, "groupnorm_silu_f16x2: bad planes");
  return gn_silu_launch(x, ldx, stats, gamma, beta, res, ldr, nullptr, (__half*)y_hi + h_off, (__half*)y_lo + h_off, ldh,
                        B, HW, C, G, stream);
}

// ------------------------------------------------------------------------------------------------
// un-pad + bilinear (align_corners=False) + pixel-major -> NCHW.
// src = (dst + 0.5) * in/out - 0.5 clamped at 0 (torch area_pixel_compute_source_index).

__device__ __forceinline__ void bilin_axis(int dst, float scale, int n_in, int& i0, int& i1, float& l0, float& l1) {
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

// POST: the per-step post-blocks fused into the epilogue: inverse scaling v * scale[c] + shift[c] (two
// roundings, like torch's y * std + mean, applications/rollout_to_netcdf.py:287) then the TracerFixer clamps
// (credit/postblock/conservation.py:88-115) max(v, lo[c]), min(v, hi[c]).
struct UnpadPost {
  const float* scale;
  const float* shift;
  const float* lo;
  const float* hi;
};
__device__ __forceinline__ float unpad_post(float v, const UnpadPost& p, int c) {
  v = __fadd_rn(__fmul_rn(v, __ldg(p.scale + c)), __ldg(p.shift + c));
  return fminf(fmaxf(v, __ldg(p.lo + c)), __ldg(p.hi + c));
}

template <bool POST>
__global__ void __launch_bounds__(256) unpad_resize_kernel(const float* __restrict__ y, int ld, float* __restrict__ out,
                                                            int C, int Hd, int Wd, int top, int left, int Hc, int Wc,
                                                            int Ho, int Wo, float sh, float sw, int cgroups, int o0,
                                                            UnpadPost post) {
  __shared__ float tile[32][33];  // [channel][column]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int b = blockIdx.z / cgroups, cg = blockIdx.z % cgroups;
  const int h = o0 + blockIdx.y, w0 = blockIdx.x * 32;
  int y0, y1;
  float ly0, ly1;
  bilin_axis(h, sh, Hc, y0, y1, ly0, ly1);
  const int ch = cg * 32 + tx;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int w = w0 + i;
    float v = 0.f;
    if (w < Wo && ch < C) {
      int x0, x1;
      float lx0, lx1;
      bilin_axis(w, sw, Wc, x0, x1, lx0, lx1);
      const float* r0 = y + ((size_t)(b * Hd + top + y0) * Wd + left) * ld + ch;
      const float* r1 = y + ((size_t)(b * Hd + top + y1) * Wd + left) * ld + ch;
      const float v00 = r0[(size_t)x0 * ld], v01 = r0[(size_t)x1 * ld];
      const float v10 = r1[(size_t)x0 * ld], v11 = r1[(size_t)x1 * ld];
      v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
      if constexpr (POST) v = unpad_post(v, post, ch);
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

// One CTA = one output row x 64 columns x a block of 64 channels: each thread interpolates 4 channels of a pixel from
// float4 loads (whole 128/256-byte pixel rows), the transposed write stores float4 runs along W.
template <bool POST>
__global__ void __launch_bounds__(256) unpad_resize_vec_kernel(const float* __restrict__ y, int ld, float* __restrict__ out,
                                                                int C, int Hd, int Wd, int top, int left, int Hc, int Wc,
                                                                int Ho, int Wo, float sh, float sw, int cblocks, int o0,
                                                                UnpadPost post) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  __shared__ float tile[64][65];  // [channel][column]
  const int tid = threadIdx.x;
  const int b = blockIdx.z / cblocks, cb = blockIdx.z % cblocks;
  const int h = o0 + blockIdx.y, w0 = blockIdx.x * 64;
  const int ch0 = cb * 64;
  const int nch = min(64, C - ch0);  // multiple of 4 (checked by the launcher)
  int y0, y1;
  float ly0, ly1;
  bilin_axis(h, sh, Hc, y0, y1, ly0, ly1);
  const int nq = nch >> 2;
  const float* base0 = y + ((size_t)(b * Hd + top + y0) * Wd + left) * ld + ch0;
  const float* base1 = y + ((size_t)(b * Hd + top + y1) * Wd + left) * ld + ch0;
  for (int idx = tid; idx < 64 * nq; idx += 256) {
    const int q = idx % nq, colx = idx / nq;
    const int w = w0 + colx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w < Wo) {
      int x0, x1;
      float lx0, lx1;
      bilin_axis(w, sw, Wc, x0, x1, lx0, lx1);
      const float4 v00 = *reinterpret_cast<const float4*>(base0 + (size_t)x0 * ld + 4 * q);
      const float4 v10 = *reinterpret_cast<const float4*>(base1 + (size_t)x0 * ld + 4 * q);
      float4 v01 = v00, v11 = v10;
      if (x1 != x0) {
        v01 = *reinterpret_cast<const float4*>(base0 + (size_t)x1 * ld + 4 * q);
        v11 = *reinterpret_cast<const float4*>(base1 + (size_t)x1 * ld + 4 * q);
      }
      v.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
      v.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
      v.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
      v.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
      if constexpr (POST) {
        const int c = ch0 + 4 * q;
        v.x = unpad_post(v.x, post, c);
        v.y = unpad_post(v.y, post, c + 1);
        v.z = unpad_post(v.z, post, c + 2);
        v.w = unpad_post(v.w, post, c + 3);
      }
    }
    tile[4 * q][colx] = v.x;
    tile[4 * q + 1][colx] = v.y;
    tile[4 * q + 2][colx] = v.z;
    tile[4 * q + 3][colx] = v.w;
  }
  __syncthreads();
  if ((Wo & 3) == 0) {
    for (int idx = tid; idx < nch * 16; idx += 256) {
      const int w4 = idx & 15, c = idx >> 4;
      const int w = w0 + 4 * w4;
      if (w < Wo)
        *reinterpret_cast<float4*>(out + ((size_t)(b * C + ch0 + c) * Ho + h) * Wo + w) =
            make_float4(tile[c][4 * w4], tile[c][4 * w4 + 1], tile[c][4 * w4 + 2], tile[c][4 * w4 + 3]);
    }
  } else {
    for (int idx = tid; idx < nch * 64; idx += 256) {
      const int wx = idx & 63, c = idx >> 6;
      if (w0 + wx < Wo) out[((size_t)(b * C + ch0 + c) * Ho + h) * Wo + w0 + wx] = tile[c][wx];
    }
  }
}

static int unpad_launch(const float* y, int ld, float* out, int B, int C, int Hd, int Wd, int top, int left, int Hc, int Wc,
                        int Ho, int Wo, int o0, int n_out, const UnpadPost* post, void* stream) {
  if (B <= 0 || C <= 0 || Hc <= 0 || Wc <= 0 || Ho <= 0 || Wo <= 0 || top < 0 || left < 0 || top + Hc > Hd ||
      left + Wc > Wd || ld < C)
    WXF_FAIL(WXF_EINVAL, "unpad_resize: bad dims");
  if (o0 < 0 || n_out < 0 || o0 + n_out > Ho) WXF_FAIL(WXF_EINVAL, "unpad_resize: rows [%d, %d) outside [0, %d)", o0, o0 + n_out, Ho);
  if (n_out == 0) return 0;
  const float sh = (float)Hc / (float)Ho, sw = (float)Wc / (float)Wo;
  const UnpadPost pp = post ? *post : UnpadPost{};
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 3) == 0 && (ld & 3) == 0 && wxf_aligned16(y) && wxf_aligned16(out)) {
    const int cblocks = (C + 63) / 64;
    if (n_out > 65535 || (int64_t)B * cblocks > 65535) WXF_FAIL(WXF_EINVAL, "unpad_resize: grid too large");
    dim3 grid((Wo + 63) / 64, n_out, B * cblocks);
    if (post)
      wxf_launch(unpad_resize_vec_kernel<true>, dim3(grid), dim3(256), 0, st, y, ld, out, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, sh, sw,
                 cblocks, o0, pp);
    else
      wxf_launch(unpad_resize_vec_kernel<false>, dim3(grid), dim3(256), 0, st, y, ld, out, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, sh, sw,
                 cblocks, o0, pp);
    WXF_CHECK_LAUNCH("unpad_resize");
    return 0;
  }
  const int cgroups = (C + 31) / 32;
  if (n_out > 65535 || (int64_t)B * cgroups > 65535) WXF_FAIL(WXF_EINVAL, "unpad_resize: grid too large");
  dim3 grid((Wo + 31) / 32, n_out, B * cgroups), block(32, 8);
  if (post)
    unpad_resize_kernel<true><<<grid, block, 0, st>>>(y, ld, out, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, sh, sw, cgroups, o0, pp);
  else
    unpad_resize_kernel<false><<<grid, block, 0, st>>>(y, ld, out, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, sh, sw, cgroups, o0, pp);
  WXF_CHECK_LAUNCH("unpad_resize");
  return 0;
}

extern "C" int wxf_unpad_resize_to_nchw(const float* y, int ld, float* out, int B, int C, int Hd, int Wd, int top,
                                        int left, int Hc, int Wc, int Ho, int Wo, int o0, int n_out, void* stream) {
  return unpad_launch(y, ld, out, B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, o0, n_out, nullptr, stream);
}

extern "C" int wxf_unpad_resize_post_to_nchw(const float* y, int ld, float* out, int B, int C, int Hd, int Wd, int top,
                                             int left, int Hc, int Wc, int Ho, int Wo, int o0, int n_out, const float* scale,
                                             const float* shift, const float* clamp_lo, const float* clamp_hi, void* stream) {
  if (!scale || !shift || !clamp_lo || !clamp_hi) WXF_FAIL(WXF_EINVAL, "unpad_resize_post: null per-channel table");
  const UnpadPost post{scale, shift, clamp_lo, clamp_hi};
  return unpad_launch(y, ld, out, B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, o0, n_out, &post, stream);
}

// ------------------------------------------------------------------------------------------------
// GlobalMassFixer (credit/postblock/conservation.py:118-176), hybrid-sigma grid, midpoint quantities:
//   sums[b] = ( sum_p area[p] * sum_l da[l] (1 - q[b,l,p]),  sum_p area[p] * sp[b,p] * sum_l db[l] (1 - q[b,l,p]) )
// One thread per pixel walks the column; block partials are combined in fp64 by the last block (deterministic order).
__global__ void __launch_bounds__(256) dry_mass_sums_kernel(const float* __restrict__ q, int64_t q_bstride, int64_t q_lstride,
                                                            const float* __restrict__ sp, int64_t sp_bstride,
                                                            const float* __restrict__ area, const float* __restrict__ da,
                                                            const float* __restrict__ db, int L, int64_t p0, int64_t np,
                                                            double* __restrict__ partial, double* __restrict__ sums,
                                                            unsigned int* __restrict__ counter) {
  const int b = blockIdx.y;
  double a = 0.0, c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = p0 + i;
    float pa = 0.f, pb = 0.f;
    for (int l = 0; l < L; ++l) {
      const float dry = 1.0f - q[b * q_bstride + l * q_lstride + p];
      pa = __fadd_rn(pa, __fmul_rn(__ldg(da + l), dry));   // ((da * (1 - q)).sum(1)): level order, fp32, no contraction
      pb = __fadd_rn(pb, __fmul_rn(__ldg(db + l), dry));
    }
    const float ar = __ldg(area + p);
    a += (double)__fmul_rn(pa, ar);
    c += (double)__fmul_rn(__fmul_rn(pb, sp[b * sp_bstride + p]), ar);
  }
  __shared__ double sa[256], sc[256];
  sa[threadIdx.x] = a;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sc[threadIdx.x] += sc[threadIdx.x + o];
    }
    __syncthreads();
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partial[((int64_t)b * gridDim.x + blockIdx.x) * 2] = sa[0];
    partial[((int64_t)b * gridDim.x + blockIdx.x) * 2 + 1] = sc[0];
    __threadfence();
    last = atomicAdd(counter + b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    double ta = 0.0, tc = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) {
      ta += partial[((int64_t)b * gridDim.x + k) * 2];
      tc += partial[((int64_t)b * gridDim.x + k) * 2 + 1];
    }
    sums[2 * b] = ta;
    sums[2 * b + 1] = tc;
    counter[b] = 0;
  }
}

extern "C" int64_t wxf_dry_mass_scratch_bytes(int B) { return (int64_t)B * 296 * 2 * 8 + (int64_t)B * 4 + 64; }

extern "C" int wxf_dry_mass_sums(const float* q, int64_t q_bstride, int64_t q_lstride, const float* sp, int64_t sp_bstride,
                                 const float* area, const float* da, const float* db, int B, int L, int64_t p0, int64_t np,
                                 double* sums, void* scratch, void* stream) {
  if (!q || !sp || !area || !da || !db || !sums || !scratch || B <= 0 || L <= 0 || p0 < 0 || np <= 0)
    WXF_FAIL(WXF_EINVAL, "dry_mass_sums: bad arguments");
  double* partial = reinterpret_cast<double*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + (int64_t)B * 296 * 2);
  int64_t blocks = (np + 255) / 256;
  if (blocks > 296) blocks = 296;
  dry_mass_sums_kernel<<<dim3((unsigned)blocks, B), 256, 0, (cudaStream_t)stream>>>(q, q_bstride, q_lstride, sp, sp_bstride, area, da, db, L,
                                                                                    p0, np, partial, sums, counter);
  WXF_CHECK_LAUNCH("dry_mass_sums");
  return 0;
}

__global__ void __launch_bounds__(256) scale_planes_kernel(float* __restrict__ x, int64_t bstride, int64_t n,
                                                           const float* __restrict__ ratio) {
  const int b = blockIdx.y;
  const float r = __ldg(ratio + b);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[b * bstride + i] *= r;
}

extern "C" int wxf_scale_planes(float* x, int64_t bstride, int64_t n, const float* ratio, int B, void* stream) {
  if (!x || !ratio || B <= 0 || n <= 0) WXF_FAIL(WXF_EINVAL, "scale_planes: bad arguments");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_planes_kernel<<<dim3((unsigned)blocks, B), 256, 0, (cudaStream_t)stream>>>(x, bstride, n, ratio);
  WXF_CHECK_LAUNCH("scale_planes");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GlobalWaterFixer / GlobalEnergyFixerUpDown (credit/postblock/conservation.py:179-236, 239-376), hybrid-sigma grid with
// midpoint quantities.  One thread per pixel walks the column of the input state (t0) and of the prediction (t1) and forms
// the per-pixel budget terms exactly as the reference does in fp32 (pressure = a + b * sp at the L + 1 interfaces, thickness =
// difference, integral = sum_l x_l * thickness_l); the area-weighted global sums are accumulated in fp64 per block and combined
// by the last block in a fixed order (deterministic; the reference sums in fp32: tolerance, not bit equality).

template <int NQ>
__device__ __forceinline__ void budget_reduce(const double (&v)[NQ], double* __restrict__ partial, double* __restrict__ sums,
                                              unsigned int* __restrict__ counter, int b) {
  __shared__ double red[NQ][256];
#pragma unroll
  for (int q = 0; q < NQ; ++q) red[q][threadIdx.x] = v[q];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
    }
    __syncthreads();
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) partial[((int64_t)b * gridDim.x + blockIdx.x) * NQ + q] = red[q][0];
    __threadfence();
    last = atomicAdd(counter + b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    double t[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) t[q] = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) t[q] += partial[((int64_t)b * gridDim.x + k) * NQ + q];
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) sums[NQ * b + q] = t[q];
    counter[b] = 0;
  }
}

// thickness of layer l at surface pressure sp: (a[l+1] + b[l+1] sp) - (a[l] + b[l] sp), each interface rounded like torch does
__device__ __forceinline__ float layer_dp(const float* __restrict__ ca, const float* __restrict__ cb, int l, float sp) {
  const float lo = __fadd_rn(__ldg(ca + l), __fmul_rn(__ldg(cb + l), sp));
  const float hi = __fadd_rn(__ldg(ca + l + 1), __fmul_rn(__ldg(cb + l + 1), sp));
  return __fsub_rn(hi, lo);
}

constexpr float WXF_GRAVITY = 9.80665f, WXF_RHO_WATER = 1000.0f, WXF_CP_DRY = 1004.64f, WXF_CP_VAPOR = 1810.0f,
                WXF_LH_WATER = 2.501e6f;

// sums[b] = ( sum area dTWC/dt, sum area evaporation flux, sum area precipitation flux )
__global__ void __launch_bounds__(256) water_budget_sums_kernel(const WxfWaterDesc d, double* __restrict__ partial,
                                                                double* __restrict__ sums, unsigned int* __restrict__ counter) {
  const int b = blockIdx.y;
  double acc[3] = {0.0, 0.0, 0.0};
  const float nsec = d.n_seconds;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.np; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = d.p0 + i;
    const float sp1 = d.sp_pred[b * d.sp_pred_bs + p], sp0 = d.sp_in[b * d.sp_in_bs + p];
    float w1 = 0.f, w0 = 0.f;
    for (int l = 0; l < d.L; ++l) {
      w1 = __fadd_rn(w1, __fmul_rn(d.q_pred[b * d.q_pred_bs + l * d.q_pred_ls + p], layer_dp(d.coef_a, d.coef_b, l, sp1)));
      w0 = __fadd_rn(w0, __fmul_rn(d.q_in[b * d.q_in_bs + l * d.q_in_ls + p], layer_dp(d.coef_a, d.coef_b, l, sp0)));
    }
    const float twc1 = __fdiv_rn(w1, WXF_GRAVITY), twc0 = __fdiv_rn(w0, WXF_GRAVITY);
    const float dtwc = __fdiv_rn(__fsub_rn(twc1, twc0), nsec);
    const float ef = __fdiv_rn(__fmul_rn(d.evapor[b * d.evapor_bs + p], WXF_RHO_WATER), nsec);
    const float pf = __fdiv_rn(__fmul_rn(d.precip[b * d.precip_bs + p], WXF_RHO_WATER), nsec);
    const float ar = __ldg(d.area + p);
    acc[0] += (double)__fmul_rn(dtwc, ar);
    acc[1] += (double)__fmul_rn(ef, ar);
    acc[2] += (double)__fmul_rn(pf, ar);
  }
  budget_reduce<3>(acc, partial, sums, counter, b);
}

__device__ __forceinline__ void energy_terms(float T, float q, float U, float V, float gph, float& cp, float& e_qgk) {
  cp = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, q), WXF_CP_DRY), __fmul_rn(q, WXF_CP_VAPOR));
  const float ken = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(U, U), __fmul_rn(V, V)));
  e_qgk = __fadd_rn(__fadd_rn(__fmul_rn(WXF_LH_WATER, q), gph), ken);
  (void)T;
}

// sums[b] = ( sum area R_T, sum area F_S, sum area TE(t0), sum area TE(t1) )
__global__ void __launch_bounds__(256) energy_budget_sums_kernel(const WxfEnergyDesc d, double* __restrict__ partial,
                                                                 double* __restrict__ sums, unsigned int* __restrict__ counter) {
  const int b = blockIdx.y;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const float nsec = d.n_seconds;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.np; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = d.p0 + i;
    const float sp1 = d.sp_pred[b * d.pred2_bs + p], sp0 = d.sp_in[b * d.sp_in_bs + p];
    const float gph = __ldg(d.gph_surf + p);
    float te1 = 0.f, te0 = 0.f;
    for (int l = 0; l < d.L; ++l) {
      const int64_t o1 = b * d.pred3_bs + l * d.pred3_ls + p, o0 = b * d.in3_bs + l * d.in3_ls + p;
      float cp, eq;
      energy_terms(0.f, d.q_pred[o1], d.u_pred[o1], d.v_pred[o1], gph, cp, eq);
      te1 = __fadd_rn(te1, __fmul_rn(__fadd_rn(__fmul_rn(cp, d.t_pred[o1]), eq), layer_dp(d.coef_a, d.coef_b, l, sp1)));
      energy_terms(0.f, d.q_in[o0], d.u_in[o0], d.v_in[o0], gph, cp, eq);
      te0 = __fadd_rn(te0, __fmul_rn(__fadd_rn(__fmul_rn(cp, d.t_in[o0]), eq), layer_dp(d.coef_a, d.coef_b, l, sp0)));
    }
    te1 = __fdiv_rn(te1, WXF_GRAVITY);
    te0 = __fdiv_rn(te0, WXF_GRAVITY);
    const int64_t o2 = b * d.pred2_bs + p;
    const float down = __fmul_rn(d.toa_down_in[b * d.toa_down_bs + p], nsec);
    const float rt = __fdiv_rn(__fsub_rn(__fsub_rn(down, __fmul_rn(d.toa_up_solar[o2], nsec)), __fmul_rn(d.toa_up_olr[o2], nsec)), nsec);
    float fs = __fsub_rn(d.surf_down_solar[o2], d.surf_up_solar[o2]);
    fs = __fadd_rn(fs, d.surf_down_lw[o2]);
    fs = __fsub_rn(fs, d.surf_up_lw[o2]);
    fs = __fadd_rn(fs, d.surf_sh[o2]);
    fs = __fadd_rn(fs, d.surf_lh[o2]);
    fs = __fdiv_rn(fs, nsec);
    const float ar = __ldg(d.area + p);
    acc[0] += (double)__fmul_rn(rt, ar);
    acc[1] += (double)__fmul_rn(fs, ar);
    acc[2] += (double)__fmul_rn(te0, ar);
    acc[3] += (double)__fmul_rn(te1, ar);
  }
  budget_reduce<4>(acc, partial, sums, counter, b);
}

// T_pred <- (E_level(t1) * ratio - E_qgk(t1)) / CP(t1), in place (conservation.py:368-372)
__global__ void __launch_bounds__(256) energy_fix_temperature_kernel(const WxfEnergyDesc d, const float* __restrict__ ratio) {
  const int b = blockIdx.y;
  const float r = __ldg(ratio + b);
  const int64_t n = (int64_t)d.L * d.np;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / d.np);
    const int64_t p = d.p0 + (i - (int64_t)l * d.np);
    const int64_t o1 = b * d.pred3_bs + l * d.pred3_ls + p;
    float cp, eq;
    energy_terms(0.f, d.q_pred[o1], d.u_pred[o1], d.v_pred[o1], __ldg(d.gph_surf + p), cp, eq);
    const float e = __fadd_rn(__fmul_rn(cp, d.t_pred[o1]), eq);
    d.t_pred[o1] = __fdiv_rn(__fsub_rn(__fmul_rn(e, r), eq), cp);
  }
}

extern "C" int64_t wxf_budget_scratch_bytes(int B) { return (int64_t)B * 296 * 4 * 8 + (int64_t)B * 4 + 64; }

extern "C" int wxf_water_budget_sums(const WxfWaterDesc* d, double* sums, void* scratch, void* stream) {
  if (!d || !sums || !scratch) WXF_FAIL(WXF_EINVAL, "water_budget_sums: null argument");
  if (!d->q_pred || !d->sp_pred || !d->q_in || !d->sp_in || !d->precip || !d->evapor || !d->area || !d->coef_a || !d->coef_b ||
      d->B <= 0 || d->L <= 0 || d->p0 < 0 || d->np <= 0 || !(d->n_seconds > 0.f))
    WXF_FAIL(WXF_EINVAL, "water_budget_sums: bad arguments");
  double* partial = reinterpret_cast<double*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + (int64_t)d->B * 296 * 4);
  int64_t blocks = (d->np + 255) / 256;
  if (blocks > 296) blocks = 296;
  water_budget_sums_kernel<<<dim3((unsigned)blocks, d->B), 256, 0, (cudaStream_t)stream>>>(*d, partial, sums, counter);
  WXF_CHECK_LAUNCH("water_budget_sums");
  return 0;
}

static int energy_desc_ok(const WxfEnergyDesc* d) {
  return d && d->t_pred && d->q_pred && d->u_pred && d->v_pred && d->sp_pred && d->t_in && d->q_in && d->u_in && d->v_in &&
         d->sp_in && d->gph_surf && d->toa_down_in && d->toa_up_solar && d->toa_up_olr && d->surf_down_solar &&
         d->surf_up_solar && d->surf_down_lw && d->surf_up_lw && d->surf_sh && d->surf_lh && d->area && d->coef_a && d->coef_b &&
         d->B > 0 && d->L > 0 && d->p0 >= 0 && d->np > 0 && d->n_seconds > 0.f;
}

extern "C" int wxf_energy_budget_sums(const WxfEnergyDesc* d, double* sums, void* scratch, void* stream) {
  if (!energy_desc_ok(d) || !sums || !scratch) WXF_FAIL(WXF_EINVAL, "energy_budget_sums: bad arguments");
  double* partial = reinterpret_cast<double*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + (int64_t)d->B * 296 * 4);
  int64_t blocks = (d->np + 255) / 256;
  if (blocks > 296) blocks = 296;
  energy_budget_sums_kernel<<<dim3((unsigned)blocks, d->B), 256, 0, (cudaStream_t)stream>>>(*d, partial, sums, counter);
  WXF_CHECK_LAUNCH("energy_budget_sums");
  return 0;
}

extern "C" int wxf_energy_fix_temperature(const WxfEnergyDesc* d, const float* ratio, void* stream) {
  if (!energy_desc_ok(d) || !ratio) WXF_FAIL(WXF_EINVAL, "energy_fix_temperature: bad arguments");
  int64_t blocks = ((int64_t)d->L * d->np + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  energy_fix_temperature_kernel<<<dim3((unsigned)blocks, d->B), 256, 0, (cudaStream_t)stream>>>(*d, ratio);
  WXF_CHECK_LAUNCH("energy_fix_temperature");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// rollout channel copy (update_x, datasets/gen_2/channel_utils.py:253-291)

struct CopyGroups {
  int32_t dst_c0[8], src_c0[8], len[8];
};

__global__ void __launch_bounds__(256) copy_channels_kernel(float* __restrict__ dst, int dst_C,
                                                             const float* __restrict__ src, int src_C, int64_t plane,
                                                             CopyGroups gr, int B) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int g = blockIdx.y / B, b = blockIdx.y % B;
  const int64_t n = (int64_t)gr.len[g] * plane;
  float* d = dst + ((int64_t)b * dst_C + gr.dst_c0[g]) * plane;
  const float* s = src + ((int64_t)b * src_C + gr.src_c0[g]) * plane;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)d | (uintptr_t)s) & 15u) == 0 && (n & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int64_t i = i0; i < (n >> 2); i += stride) d4[i] = s4[i];
  } else {
    for (int64_t i = i0; i < n; i += stride) d[i] = s[i];
  }
}

extern "C" int wxf_copy_channels(float* dst, int dst_C, const float* src, int src_C, int B, int64_t plane,
                                 const int32_t* dst_c0, const int32_t* src_c0, const int32_t* len, int n_groups,
                                 void* stream) {
  if (n_groups <= 0 || n_groups > 8 || B <= 0 || plane <= 0) WXF_FAIL(WXF_EINVAL, "copy_channels: bad dims");
  CopyGroups gr;
  for (int g = 0; g < n_groups; ++g) {
    if (dst_c0[g] < 0 || src_c0[g] < 0 || len[g] <= 0 || dst_c0[g] + len[g] > dst_C || src_c0[g] + len[g] > src_C)
      WXF_FAIL(WXF_EINVAL, "copy_channels: group %d out of range", g);
    gr.dst_c0[g] = dst_c0[g];
    gr.src_c0[g] = src_c0[g];
    gr.len[g] = len[g];
  }
  dim3 grid(148 * 4, n_groups * B);
  wxf_launch(copy_channels_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dst, dst_C, src, src_C, plane, gr, B);
  WXF_CHECK_LAUNCH("copy_channels");
  return 0;
}
