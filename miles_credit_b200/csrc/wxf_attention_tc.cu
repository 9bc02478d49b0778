// Cross-scale window attention on the tensor cores (tcgen05 + TMEM), operands gathered by TMA.
// Reference arithmetic: credit/models/crossformer.py:261-296, 301-314 (window / dilated token gather, q*scale,
// QK^T + dynamic position bias, softmax, PV, inverse gather).
//
// Two kernels share the tile layout below.  window_attention_tc2_kernel (further down; the default) pipelines tiles:
// one CTA per SM, a 3-stage TMA ring, S double buffered in TMEM, P kept in tensor memory, two softmax warpgroups.
// window_attention_tc_kernel (first, WXF_ATTN_V2=0) is the serial-per-tile version it replaced, kept for A/B.
//
// One CTA = one (window group, head) tile of up to 128 query rows: G = 128 / Lp windows of the same head are packed
// block-diagonally (Lp = L rounded up to 2 so every TMA box lands 128-byte aligned).  All operands are f16x2 planes:
//   TMA   : per window 6 boxes (Q, K, V) x (hi, lo) of the qkv planes [B, H, W, 3d]; short windows are a 4-D box
//           {32 ch, wsz, wsz, 1}; the dilated "long" groups are a 5-D box over the (l1, gh, l2, gw) view of the grid,
//           so the strided gather costs nothing.  64-byte rows, SWIZZLE_64B.
//   MMA 1 : S[128 x 128] = Q K^T, 3 passes x 2 K-steps (K-major A and B), accumulator in TMEM columns [0, 128)
//   bias  : the block-diagonal position-bias tile is identical for every tile of a launch: each CTA writes it once into
//           TMEM columns [128, 256) (tcgen05.st) and reads it back next to S - no global or shared traffic per tile
//   soft  : thread = row; two TMEM passes (max, then exp / sum); only the row's own window columns are kept, the
//           rest of the block-diagonal tile is written as exact zeros; P goes to shared memory as fp16 hi/lo planes in
//           the K-major SWIZZLE_128B layout the second MMA reads
//   MMA 2 : O[128 x 32] = P V, 3 passes x 8 K-steps; V is read as an MN-major B operand straight from the TMA tile;
//           O reuses the first 32 TMEM columns of S (dead once P is written), so two CTAs of 256 columns share an SM
//   epi   : O / rowsum -> fp16 hi/lo planes of the attention output at the token's pixel (the inverse gather)
#include <stdlib.h>

#include "wxf_fastdiv.h"
#include "wxf_tc_host.cuh"
#include "wxf_tc_ptx.cuh"

using namespace wxf_tc;

namespace {

constexpr int DH = 32;
constexpr int AT_THREADS = 192;
constexpr int ROWS = 128;
constexpr int QKV_PLANE = ROWS * 64;       // 8 KB: 128 rows x 64 B
constexpr int P_ATOM = ROWS * 128;         // 16 KB: 128 rows x 64 fp16
// K, V: hi plane then lo plane.  Q aliases the first 16 KB of the P area: Q is dead once S = Q K^T has completed,
// which is exactly when the softmax warps start writing P (2 CTAs of 97 KB fit one SM).
constexpr int OFF_K = 0, OFF_V = 2 * QKV_PLANE;
constexpr int OFF_P = 4 * QKV_PLANE;       // P_hi (2 atoms) then P_lo (2 atoms)
constexpr int OFF_Q = OFF_P;
constexpr int OFF_BAR = OFF_P + 4 * P_ATOM;
constexpr int AT_SMEM = OFF_BAR + 64 + 1024;
constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t IDESC_O = (1u << 4) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct AttnParams {
  const float* bias_tile;  // [128 columns][128 rows] fp32: bias*log2(e) inside the row's window, -1e30 elsewhere
  __half* out_hi;
  __half* out_lo;
  int ldh;
  int H, W, d, heads, wsz, kind;
  float scale;
  int L, Lp, G, nh, nw;
  int inter;          // long windows, one TMA box per plane: 1 = interleaved rows (row = token*G + window, odd L),
                      // 2 = window-major rows (row = window*L + token, even L: a warp's 32 rows see 32 + L columns, not L*G)
  int gpr;            // interleaved: tiles (groups of G consecutive gw) per window row
  float scale2;       // scale * log2(e): softmax runs in base 2
  int64_t nwin, ntiles;
  FastDiv fd_heads, fd_gpr, fd_img, fd_nw;   // divisors heads, gpr, nh*nw, nw (v2 kernel; nwin, ntiles < 2^31)
};

// Row r of a tile -> (window slot g, token i).  Window-major packing: rows [g*Lp, g*Lp + L); interleaved: r = i*G + g.
__device__ __forceinline__ void row_slot(const AttnParams& p, int r, int& g, int& i) {
  if (p.inter == 1) {
    i = r / p.G;
    g = r - i * p.G;
  } else {
    g = r / p.Lp;
    i = r - g * p.Lp;
  }
}

__global__ void __launch_bounds__(AT_THREADS, 2)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                           const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_load = base + OFF_BAR, bar_s = bar_load + 8, bar_p = bar_load + 16, bar_o = bar_load + 24,
                 bar_e = bar_load + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + OFF_BAR + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_e, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows the TMA boxes do not cover must be finite zeros (0 * garbage could be NaN in P V); the P columns the softmax
  // never writes (outside every window of the row's warp) must be exact zeros
  {
    uint4* z = reinterpret_cast<uint4*>(gen);
    for (int i = threadIdx.x; i < OFF_BAR / 16; i += AT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // position-bias tile (identical for every tile of the launch) -> TMEM columns [128, 256)
  if (warp >= 2) {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t bb[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) bb[j] = __float_as_uint(__ldg(p.bias_tile + (size_t)(c * 32 + j) * ROWS + r));
      tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + 128u + (uint32_t)(c * 32), bb);
    }
    tc_fence_before();
  }

  // per-row constants of the softmax warps (identical for every tile of the launch): the row's window slot / token, and
  // the warp-uniform range of 8-column groups that contain a window column of any of the warp's rows
  const int quarter = warp & 3;
  const int r = quarter * 32 + lane;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
  int g = 0, i = 0, ty = 0, tx = 0, g_lo = 0, g_hi = 0, c_lo = 0, c_hi = 0;
  bool row_ok = false;
  if (warp >= 2) {
    row_slot(p, r, g, i);
    row_ok = g < p.G && i < p.L;
    ty = i / p.wsz;
    tx = i - ty * p.wsz;
    const int col0 = p.inter == 1 ? 0 : g * p.Lp, col1 = p.inter == 1 ? p.L * p.G : col0 + p.L;
    g_lo = __reduce_min_sync(0xffffffffu, row_ok ? (col0 >> 3) : 16);
    g_hi = __reduce_max_sync(0xffffffffu, row_ok ? ((col1 + 7) >> 3) : 0);
    if (g_hi <= g_lo) g_lo = g_hi = 0;
    c_lo = g_lo >> 2;
    c_hi = (g_hi + 3) >> 2;
  }

  // persistent: each CTA (two per SM) walks tiles = (window group, head); every barrier completes once per tile, so
  // the wait parity is the tile iteration's low bit.  Stale rows of a previous tile are finite, masked data.
  for (int64_t tile = blockIdx.x, it = 0; tile < p.ntiles; tile += gridDim.x, ++it) {
  const uint32_t par = (uint32_t)it & 1u;
  const int head = (int)(tile % p.heads);
  const int64_t group = tile / p.heads;
  // first window of the tile and number of windows present
  int64_t w0;
  int nv;
  if (p.inter) {  // G consecutive gw of one (b, gh) row of groups
    const int64_t rowi = group / p.gpr;
    const int gw0 = (int)(group - rowi * p.gpr) * p.G;
    w0 = rowi * p.nw + gw0;
    nv = (p.nw - gw0) < p.G ? (p.nw - gw0) : p.G;
  } else {
    w0 = group * p.G;
    nv = (int)((p.nwin - w0) < p.G ? (p.nwin - w0) : p.G);
  }

  if (warp == 0) {
    if (lane == 0) {
      if (it > 0) mbar_wait(bar_o, par ^ 1u);  // previous tile's P V has finished reading this CTA's shared memory
      const int per_img = p.nh * p.nw;
      if (p.inter) {
        // one 5-D box {32 ch, G windows (gw), wsz (l2), 1, wsz (l1)} per plane: rows come out as token*G + window;
        // windows past the grid edge are TMA zero fill (the full box is always counted)
        mbar_expect_tx(bar_load, (uint32_t)(6 * p.L * p.G * 64));
        const int b = (int)(w0 / per_img);
        const int rem = (int)(w0 - (int64_t)b * per_img);
        const int gh = rem / p.nw, gw = rem - gh * p.nw;
#pragma unroll
        for (int which = 0; which < 3; ++which) {
          const int c0 = which * p.d + head * DH;
          const uint32_t dst = base + (uint32_t)(which == 0 ? OFF_Q : (which == 1 ? OFF_K : OFF_V));
          if (p.inter == 2) {  // window-major map: dims (c, l2, (b, l1), gw, gh)
            tma_load_5d(&tm_hi, bar_load, dst, c0, 0, b * p.wsz, gw, gh);
            tma_load_5d(&tm_lo, bar_load, dst + QKV_PLANE, c0, 0, b * p.wsz, gw, gh);
          } else {
            tma_load_5d(&tm_hi, bar_load, dst, c0, gw, 0, gh, b * p.wsz);
            tma_load_5d(&tm_lo, bar_load, dst + QKV_PLANE, c0, gw, 0, gh, b * p.wsz);
          }
        }
      } else {
      mbar_expect_tx(bar_load, (uint32_t)(nv * 6 * p.L * 64));
      for (int g = 0; g < nv; ++g) {
        const int64_t w = w0 + g;
        const int b = (int)(w / per_img);
        const int rem = (int)(w - (int64_t)b * per_img);
        const int gh = rem / p.nw, gw = rem - gh * p.nw;
        const uint32_t row_off = (uint32_t)(g * p.Lp * 64);
#pragma unroll
        for (int which = 0; which < 3; ++which) {  // q, k, v
          const int c0 = which * p.d + head * DH;
          const uint32_t dst = base + (uint32_t)(which == 0 ? OFF_Q : (which == 1 ? OFF_K : OFF_V)) + row_off;
          if (p.kind == WXF_ATTN_SHORT) {
            tma_load_4d(&tm_hi, bar_load, dst, c0, gw * p.wsz, gh * p.wsz, b);
            tma_load_4d(&tm_lo, bar_load, dst + QKV_PLANE, c0, gw * p.wsz, gh * p.wsz, b);
          } else {
            tma_load_5d(&tm_hi, bar_load, dst, c0, gw, 0, gh, b * p.wsz);
            tma_load_5d(&tm_lo, bar_load, dst + QKV_PLANE, c0, gw, 0, gh, b * p.wsz);
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(bar_load, par);
      if (it > 0) mbar_wait(bar_e, par ^ 1u);  // previous tile's O (aliases S columns 0-31) has been read
      tc_fence_after();
      const uint32_t q_hi = base + OFF_Q, q_lo = q_hi + QKV_PLANE, k_hi = base + OFF_K, k_lo = k_hi + QKV_PLANE;
#pragma unroll
      for (int k = 0; k < 2; ++k) {  // dh = 32 = two K=16 steps (32 bytes each inside the 64-byte row)
        tc_mma_f16(tmem_base, umma_desc_sw64_kmajor(q_hi + k * 32), umma_desc_sw64_kmajor(k_lo + k * 32), IDESC_S, k ? 1u : 0u);
        tc_mma_f16(tmem_base, umma_desc_sw64_kmajor(q_lo + k * 32), umma_desc_sw64_kmajor(k_hi + k * 32), IDESC_S, 1u);
        tc_mma_f16(tmem_base, umma_desc_sw64_kmajor(q_hi + k * 32), umma_desc_sw64_kmajor(k_hi + k * 32), IDESC_S, 1u);
      }
      tc_commit(bar_s);
      mbar_wait(bar_p, par);  // P planes written by the softmax warps
      tc_fence_after();
      const uint32_t p_hi = base + OFF_P, p_lo = p_hi + 2 * P_ATOM, v_hi = base + OFF_V, v_lo = v_hi + QKV_PLANE;
      const uint32_t d_o = tmem_base;  // O aliases S[:, 0:32]
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // 128 key rows = eight K=16 steps
        const uint32_t pa = (uint32_t)((ks >> 2) * P_ATOM + (ks & 3) * 32);
        const uint32_t va = (uint32_t)(ks * 16 * 64);
        tc_mma_f16(d_o, umma_desc_sw128(p_hi + pa), umma_desc_sw64_mnmajor(v_lo + va), IDESC_O, ks ? 1u : 0u);
        tc_mma_f16(d_o, umma_desc_sw128(p_lo + pa), umma_desc_sw64_mnmajor(v_hi + va), IDESC_O, 1u);
        tc_mma_f16(d_o, umma_desc_sw128(p_hi + pa), umma_desc_sw64_mnmajor(v_hi + va), IDESC_O, 1u);
      }
      tc_commit(bar_o);
    }
  } else {
    // ---- softmax + epilogue: thread = query row ----
    mbar_wait(bar_s, par);
    tc_fence_after();
    // s2 = S * (scale*log2 e) + bias*log2 e (masked columns: -1e30); p = 2^(s2 - max).  Only the 8-column groups
    // [g_lo, g_hi) that hold a window column of one of this warp's rows are touched; everything else of P stays zero.
    const float2 sc2 = make_float2(p.scale2, p.scale2);
    float mx;
    {
      float m0 = -3.0e38f, m1 = -3.0e38f, m2 = -3.0e38f, m3 = -3.0e38f;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t rr[32], bb[32];
        tmem_ld32_nowait(lane_base + (uint32_t)(c * 32), rr);
        tmem_ld32_nowait(lane_base + 128u + (uint32_t)(c * 32), bb);
        tmem_ld_wait();
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          const int gq = c * 4 + q8;
          if (gq >= g_lo && gq < g_hi) {  // warp-uniform
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
              const int j = q8 * 8 + e;
              const float2 t0 = __ffma2_rn(make_float2(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), sc2,
                                           make_float2(__uint_as_float(bb[j]), __uint_as_float(bb[j + 1])));
              const float2 t1 = __ffma2_rn(make_float2(__uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3])), sc2,
                                           make_float2(__uint_as_float(bb[j + 2]), __uint_as_float(bb[j + 3])));
              m0 = fmaxf(m0, t0.x); m1 = fmaxf(m1, t0.y); m2 = fmaxf(m2, t1.x); m3 = fmaxf(m3, t1.y);
            }
          }
        }
      }
      mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    }
    const float2 nmx2 = make_float2(-mx, -mx);
    float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
    uint8_t* prow_hi = gen + OFF_P + r * 128;
    uint8_t* prow_lo = prow_hi + 2 * P_ATOM;
    // groups outside the warp's range that alias Q (P_hi columns 0..63) were overwritten by this tile's Q load
#pragma unroll 1
    for (int gq = 0; gq < 8; ++gq)
      if (gq < g_lo || gq >= g_hi) *reinterpret_cast<uint4*>(prow_hi + ((gq ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t rr[32], bb[32];
      tmem_ld32_nowait(lane_base + (uint32_t)(c * 32), rr);
      tmem_ld32_nowait(lane_base + 128u + (uint32_t)(c * 32), bb);
      tmem_ld_wait();
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) {  // 8 columns = one 16-byte chunk of the swizzled row
        const int gq = c * 4 + q8;
        if (gq >= g_lo && gq < g_hi) {  // warp-uniform
          __align__(16) uint32_t h2[4];
          __align__(16) uint32_t l2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q8 * 8 + 2 * e;
            float2 t = __ffma2_rn(make_float2(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), sc2,
                                  make_float2(__uint_as_float(bb[j]), __uint_as_float(bb[j + 1])));
            t = __fadd2_rn(t, nmx2);
            float2 pe;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe.x) : "f"(t.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe.y) : "f"(t.y));
            if (e & 1) sum_b = __fadd2_rn(sum_b, pe); else sum_a = __fadd2_rn(sum_a, pe);
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2[e]) : "f"(pe.y), "f"(pe.x));
            const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h2[e]));
            const float2 dlt = __ffma2_rn(back, make_float2(-1.0f, -1.0f), pe);
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l2[e]) : "f"(dlt.y), "f"(dlt.x));
          }
          const int cc = (c & 1) * 4 + q8;                 // 16-byte chunk index inside the 64-column atom
          const int off = (c >> 1) * P_ATOM + ((cc ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(prow_hi + off) = *reinterpret_cast<const uint4*>(h2);
          *reinterpret_cast<uint4*>(prow_lo + off) = *reinterpret_cast<const uint4*>(l2);
        }
      }
    }
    const float lsum = (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
    fence_proxy_async();   // generic-proxy writes of P -> visible to the tensor core (async proxy)
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, par);
    tc_fence_after();
    uint32_t oo[32];
    tmem_ld32(lane_base, oo);
    tc_fence_before();
    mbar_arrive(bar_e);  // O consumed: the next tile's QK^T may overwrite these TMEM columns
    if (row_ok && g < nv) {
      const int64_t w = w0 + g;
      const int per_img = p.nh * p.nw;
      const int b = (int)(w / per_img);
      const int rem = (int)(w - (int64_t)b * per_img);
      const int gh = rem / p.nw, gw = rem - gh * p.nw;
      int y, x;
      if (p.kind == WXF_ATTN_SHORT) {
        y = gh * p.wsz + ty;
        x = gw * p.wsz + tx;
      } else {
        y = ty * p.nh + gh;
        x = tx * p.nw + gw;
      }
      const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
      const float inv = 1.0f / lsum;
      uint4* hp = reinterpret_cast<uint4*>(p.out_hi + pix * p.ldh + head * DH);
      uint4* lp = reinterpret_cast<uint4*>(p.out_lo + pix * p.ldh + head * DH);
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        __align__(16) __half2 h8[4];
        __align__(16) __half2 l8[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          wxf_split2_f16x2(__uint_as_float(oo[8 * c + 2 * e]) * inv, __uint_as_float(oo[8 * c + 2 * e + 1]) * inv, h8[e],
                           l8[e]);
        hp[c] = *reinterpret_cast<const uint4*>(h8);
        lp[c] = *reinterpret_cast<const uint4*>(l8);
      }
    }
    tc_fence_before();  // this row's TMEM reads are done before the next tile's MMAs overwrite S / O
  }
  }  // tile loop

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ---- v2: one CTA per SM, decoupled load pipeline, P kept in tensor memory ---------------------------------------
// The v1 kernel above runs load -> QK^T -> softmax -> PV -> store serially per CTA (two CTAs per SM hide part of it);
// its profile is the softmax warps waiting for the TMA load and the MMA round trips.  Here:
//   * Q/K/V tiles go through an A2_NST-deep shared-memory ring filled by the TMA warp, so loads run tiles ahead;
//   * S is double buffered in TMEM and two softmax warpgroups alternate tiles (QK^T of tile i+1 is issued before PV of i);
//   * P never touches shared memory: each softmax thread writes its row of P (fp16 hi / lo, two per 32-bit column) over
//     its own S row with tcgen05.st, and PV reads the A operand from tensor memory;
//   * the output row goes through a per-warp swizzled staging slab so every store instruction writes whole 64-byte runs,
//     one tile late (four O buffers in TMEM), so a softmax group never waits for its own PV round trip.
//   warp 0: TMA producer     warp 1: tcgen05.mma issuer     warps 2-5: softmax group 0     warps 6-9: softmax group 1
constexpr int A2_THREADS = 320;
constexpr int A2_NST = 3;
constexpr int A2_QKV = 6 * QKV_PLANE;            // 48 KB per stage: K hi/lo, V hi/lo, Q hi/lo
constexpr int A2_OFF_K = 0, A2_OFF_V = 2 * QKV_PLANE, A2_OFF_Q = 4 * QKV_PLANE;
constexpr int A2_OFF_STG = A2_NST * A2_QKV;      // 8 softmax warps x 4 KB output staging
constexpr int A2_OFF_BAR = A2_OFF_STG + 8 * 4096;
constexpr int A2_SMEM = A2_OFF_BAR + 256 + 1024;   // 14 mbarriers + the TMEM address slot
constexpr uint32_t A2_COL_BIAS = 256, A2_COL_O = 384;

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// 16 loaded score / bias columns [16h, 16h+16) of the row -> 8 packed fp16x2 words of P_hi and of P_lo (zeros for the
// 8-column groups outside [g_lo, g_hi))
__device__ __forceinline__ void softmax_half_math(const uint32_t (&rr)[16], const uint32_t (&bb)[16], int h, int g_lo, int g_hi,
                                                  float2 sc2, float2 nmx2, float2& sum_a, float2& sum_b, uint32_t (&hi)[8],
                                                  uint32_t (&lo)[8]) {
#pragma unroll
  for (int q8 = 0; q8 < 2; ++q8) {
    const int gq = h * 2 + q8;
    if (gq >= g_lo && gq < g_hi) {  // warp-uniform
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = q8 * 8 + 2 * e;
        float2 t = __ffma2_rn(make_float2(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), sc2,
                              make_float2(__uint_as_float(bb[j]), __uint_as_float(bb[j + 1])));
        t = __fadd2_rn(t, nmx2);
        float2 pe;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe.x) : "f"(t.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe.y) : "f"(t.y));
        if (e & 1) sum_b = __fadd2_rn(sum_b, pe); else sum_a = __fadd2_rn(sum_a, pe);
        uint32_t hw, lw;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hw) : "f"(pe.y), "f"(pe.x));  // low half = even key
        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hw));
        const float2 dlt = __ffma2_rn(back, make_float2(-1.0f, -1.0f), pe);
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lw) : "f"(dlt.y), "f"(dlt.x));
        hi[q8 * 4 + e] = hw;
        lo[q8 * 4 + e] = lw;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) hi[q8 * 4 + e] = lo[q8 * 4 + e] = 0u;
    }
  }
}

// a half is active when one of its two 8-column groups holds a window column of one of the warp's rows (warp-uniform)
__device__ __forceinline__ bool half_active(int h, int g_lo, int g_hi) { return h * 2 < g_hi && h * 2 + 2 > g_lo; }

// load + arithmetic of one half; inactive halves come out as zeros without touching tensor memory
__device__ __forceinline__ void softmax_half(uint32_t s_addr, uint32_t b_addr, int h, int g_lo, int g_hi, float2 sc2,
                                             float2 nmx2, float2& sum_a, float2& sum_b, uint32_t (&hi)[8],
                                             uint32_t (&lo)[8]) {
  if (!half_active(h, g_lo, g_hi)) {
#pragma unroll
    for (int e = 0; e < 8; ++e) hi[e] = lo[e] = 0u;
    return;
  }
  uint32_t rr[16], bb[16];
  tmem_ld16_nowait(s_addr + (uint32_t)(h * 16), rr);
  tmem_ld16_nowait(b_addr + (uint32_t)(h * 16), bb);
  tmem_ld_wait();
  softmax_half_math(rr, bb, h, g_lo, g_hi, sc2, nmx2, sum_a, sum_b, hi, lo);
}

__global__ void __launch_bounds__(A2_THREADS, 1)
window_attention_tc2_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                            const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + A2_OFF_BAR;
  auto bar_qkv_full = [&](int s) { return bars + 8u * s; };
  auto bar_qkv_empty = [&](int s) { return bars + 8u * (A2_NST + s); };
  auto bar_s_full = [&](int b) { return bars + 8u * (2 * A2_NST) + 8u * b; };
  auto bar_p_full = [&](int b) { return bars + 8u * (2 * A2_NST + 2) + 8u * b; };
  auto bar_o_full = [&](int q) { return bars + 8u * (2 * A2_NST + 4) + 8u * q; };    // q = tile & 3: four O buffers
  auto bar_o_empty = [&](int q) { return bars + 8u * (2 * A2_NST + 8) + 8u * q; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + A2_OFF_BAR + 8 * (2 * A2_NST + 12));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  wxf_pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < A2_NST; ++s) {
      mbar_init(bar_qkv_full(s), 1);
      mbar_init(bar_qkv_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s_full(b), 1);
      mbar_init(bar_p_full(b), 128);
    }
    for (int q = 0; q < 4; ++q) {
      mbar_init(bar_o_full(q), 1);
      mbar_init(bar_o_empty(q), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // operand rows the TMA boxes never cover must be finite zeros (0 * garbage could be NaN in P V)
  {
    uint4* z = reinterpret_cast<uint4*>(gen);
    for (int i = threadIdx.x; i < A2_OFF_STG / 16; i += A2_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  wxf_pdl_wait();  // barrier init, TMEM allocation and the shared-memory zero fill overlap the previous kernel's tail

  // position-bias tile (identical for every tile of the launch) -> TMEM columns [256, 384), written by group 0
  if (warp >= 2 && warp < 6) {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t bb[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) bb[j] = __float_as_uint(__ldg(p.bias_tile + (size_t)(c * 32 + j) * ROWS + r));
      tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + A2_COL_BIAS + (uint32_t)(c * 32), bb);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // tiles of this CTA: tile(it) = blockIdx.x + it * gridDim.x; ring stage it % A2_NST; S/P/O buffer it & 1
  const int my_tiles = p.ntiles > (int64_t)blockIdx.x ? (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  // tile -> (head, first window, windows present); 32-bit, divisions by multiply-high (the host checks ntiles, nwin < 2^31)
  auto tile_windows = [&](uint32_t tile, int& head, uint32_t& w0, int& nv) {
    const uint32_t group = fdiv(tile, p.fd_heads);
    head = (int)(tile - group * (uint32_t)p.heads);
    if (p.inter) {  // G consecutive gw of one (b, gh) row of groups
      const uint32_t rowi = fdiv(group, p.fd_gpr);
      const int gw0 = (int)(group - rowi * (uint32_t)p.gpr) * p.G;
      w0 = rowi * (uint32_t)p.nw + (uint32_t)gw0;
      nv = (p.nw - gw0) < p.G ? (p.nw - gw0) : p.G;
    } else {
      w0 = group * (uint32_t)p.G;
      const uint32_t left = (uint32_t)p.nwin - w0;
      nv = left < (uint32_t)p.G ? (int)left : p.G;
    }
  };
  // window -> (image, group row, group column)
  auto window_pos = [&](uint32_t w, int& bi, int& gh, int& gw) {
    const uint32_t b = fdiv(w, p.fd_img);
    const uint32_t rem = w - b * p.fd_img.d;
    const uint32_t h = fdiv(rem, p.fd_nw);
    bi = (int)b; gh = (int)h; gw = (int)(rem - h * p.fd_nw.d);
  };

  if (warp == 0) {
    // ---- TMA producer: the warp runs converged (tile decode on the uniform datapath), an elected lane issues ----
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % A2_NST;
      const uint32_t k = (uint32_t)(it / A2_NST);
      int head, nv;
      uint32_t w0;
      tile_windows(blockIdx.x + (uint32_t)it * gridDim.x, head, w0, nv);
      if (k > 0) mbar_wait(bar_qkv_empty(st), (k - 1u) & 1u);  // PV of the tile A2_NST back has read this stage
      const uint32_t buf = base + (uint32_t)(st * A2_QKV);
      const uint32_t full = bar_qkv_full(st);
      if (p.inter) {
        int bi, gh, gw;
        window_pos(w0, bi, gh, gw);
        if (elect_one()) {
          mbar_expect_tx(full, (uint32_t)(6 * p.L * p.G * 64));
#pragma unroll
          for (int which = 0; which < 3; ++which) {
            const int c0 = which * p.d + head * DH;
            const uint32_t dst = buf + (uint32_t)(which == 0 ? A2_OFF_Q : (which == 1 ? A2_OFF_K : A2_OFF_V));
            if (p.inter == 2) {  // window-major map: dims (c, l2, (b, l1), gw, gh)
              tma_load_5d(&tm_hi, full, dst, c0, 0, bi * p.wsz, gw, gh);
              tma_load_5d(&tm_lo, full, dst + QKV_PLANE, c0, 0, bi * p.wsz, gw, gh);
            } else {
              tma_load_5d(&tm_hi, full, dst, c0, gw, 0, gh, bi * p.wsz);
              tma_load_5d(&tm_lo, full, dst + QKV_PLANE, c0, gw, 0, gh, bi * p.wsz);
            }
          }
        }
        __syncwarp();
      } else {
        if (elect_one()) mbar_expect_tx(full, (uint32_t)(nv * 6 * p.L * 64));
        __syncwarp();
        // consecutive windows of a tile: (bi, gh, gw) advance incrementally, one division per tile
        int bi, gh, gw;
        window_pos(w0, bi, gh, gw);
        for (int g = 0; g < nv; ++g) {
          const uint32_t row_off = (uint32_t)(g * p.Lp * 64);
          if (elect_one()) {
#pragma unroll
            for (int which = 0; which < 3; ++which) {  // q, k, v
              const int c0 = which * p.d + head * DH;
              const uint32_t dst = buf + (uint32_t)(which == 0 ? A2_OFF_Q : (which == 1 ? A2_OFF_K : A2_OFF_V)) + row_off;
              if (p.kind == WXF_ATTN_SHORT) {
                tma_load_4d(&tm_hi, full, dst, c0, gw * p.wsz, gh * p.wsz, bi);
                tma_load_4d(&tm_lo, full, dst + QKV_PLANE, c0, gw * p.wsz, gh * p.wsz, bi);
              } else {
                tma_load_5d(&tm_hi, full, dst, c0, gw, 0, gh, bi * p.wsz);
                tma_load_5d(&tm_lo, full, dst + QKV_PLANE, c0, gw, 0, gh, bi * p.wsz);
              }
            }
          }
          __syncwarp();
          if (++gw == p.nw) {
            gw = 0;
            if (++gh == p.nh) {
              gh = 0;
              ++bi;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: converged warp, an elected lane issues; operand descriptors = constant high word + an integer add
    //      on the low word (the 36 tcgen05.mma of a tile are small: N = 32 for P V, so the issue cost is what counts) ----
    constexpr uint32_t HI64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);  // SBO 512 B, version 1, SWIZZLE_64B
    // S_b = Q K^T of tile `it`.  S_b / P_b is free: tcgen05.mma instructions of one thread execute in issue order, and
    // PV(it-2), which read P_b, was issued before this.
    auto issue_qk = [&](int it) {
      const int st = it % A2_NST;
      const uint32_t k = (uint32_t)(it / A2_NST);
      mbar_wait(bar_qkv_full(st), k & 1u);
      tc_fence_after();
      const uint32_t buf = base + (uint32_t)(st * A2_QKV);
      const uint32_t q_hi = umma_desc_lo(buf + A2_OFF_Q), q_lo = q_hi + (QKV_PLANE >> 4);  // K-major: LBO field = 1
      const uint32_t k_hi = umma_desc_lo(buf + A2_OFF_K), k_lo = k_hi + (QKV_PLANE >> 4);
      const uint32_t d_s = tmem_base + (uint32_t)((it & 1) * 128);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {  // dh = 32 = two K=16 steps (32 bytes each inside the 64-byte row)
          tc_mma_f16(d_s, umma_desc_make(q_hi + 2 * ks, HI64), umma_desc_make(k_lo + 2 * ks, HI64), IDESC_S, ks ? 1u : 0u);
          tc_mma_f16(d_s, umma_desc_make(q_lo + 2 * ks, HI64), umma_desc_make(k_hi + 2 * ks, HI64), IDESC_S, 1u);
          tc_mma_f16(d_s, umma_desc_make(q_hi + 2 * ks, HI64), umma_desc_make(k_hi + 2 * ks, HI64), IDESC_S, 1u);
        }
        tc_commit(bar_s_full((int)(it & 1)));
      }
      __syncwarp();
    };
    // O_b = P_b V: A = P from tensor memory (K-step ks = 16 keys: hi words in S_b columns [16 ks, 16 ks + 8), lo words in the next 8)
    auto issue_pv = [&](int it) {
      const int b = it & 1, q = it & 3;
      const uint32_t k2 = (uint32_t)(it >> 1), k4 = (uint32_t)(it >> 2);
      const int st = it % A2_NST;
      mbar_wait(bar_p_full(b), k2 & 1u);                        // P_b written by softmax group b
      if (k4 > 0) mbar_wait(bar_o_empty(q), (k4 - 1u) & 1u);    // O_q of the tile four back has been read
      tc_fence_after();
      const uint32_t buf = base + (uint32_t)(st * A2_QKV);
      // MN-major V: LBO field = 512 B >> 4; one K = 16 step = 16 rows of 64 B = 1024 B = 64 descriptor units
      const uint32_t v_hi = (((buf + A2_OFF_V) & 0x3FFFFu) >> 4) | ((uint32_t)(512 >> 4) << 16), v_lo = v_hi + (QKV_PLANE >> 4);
      const uint32_t p_hi = tmem_base + (uint32_t)(b * 128), p_lo = p_hi + 8u;
      const uint32_t d_o = tmem_base + A2_COL_O + (uint32_t)(q * 32);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // 128 key rows = eight K=16 steps
          tc_mma_f16_ts(d_o, p_hi + ks * 16, umma_desc_make(v_lo + 64 * ks, HI64), IDESC_O, ks ? 1u : 0u);
          tc_mma_f16_ts(d_o, p_lo + ks * 16, umma_desc_make(v_hi + 64 * ks, HI64), IDESC_O, 1u);
          tc_mma_f16_ts(d_o, p_hi + ks * 16, umma_desc_make(v_hi + 64 * ks, HI64), IDESC_O, 1u);
        }
        tc_commit(bar_o_full(q));
        if (it + A2_NST < my_tiles) tc_commit(bar_qkv_empty(st));  // stage may be refilled (nobody waits after the last use)
      }
      __syncwarp();
    };
    if (my_tiles > 0) issue_qk(0);
    for (int it = 0; it < my_tiles; ++it) {
      if (it + 1 < my_tiles) issue_qk(it + 1);
      issue_pv(it);
    }
  } else {
    // ---- softmax + epilogue: thread = query row; group wg handles tiles it = wg, wg + 2, ... ----
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    int g, i;
    row_slot(p, r, g, i);
    const bool row_ok = g < p.G && i < p.L;
    const int ty = i / p.wsz, tx = i - ty * p.wsz;
    // warp-uniform range of 8-column groups that hold a window column of any of the warp's rows
    int g_lo, g_hi;
    {
      const int col0 = p.inter == 1 ? 0 : g * p.Lp, col1 = p.inter == 1 ? p.L * p.G : col0 + p.L;
      g_lo = __reduce_min_sync(0xffffffffu, row_ok ? (col0 >> 3) : 16);
      g_hi = __reduce_max_sync(0xffffffffu, row_ok ? ((col1 + 7) >> 3) : 0);
      if (g_hi <= g_lo) g_lo = g_hi = 0;
    }
    const int c_lo = g_lo >> 2, c_hi = (g_hi + 3) >> 2;
    const float2 sc2 = make_float2(p.scale2, p.scale2);
    const uint32_t s_addr = tmem_base + lane_off + (uint32_t)(wg * 128);
    const uint32_t b_addr = tmem_base + lane_off + A2_COL_BIAS;
    uint8_t* stg = gen + A2_OFF_STG + (warp - 2) * 4096;  // this warp's 32 rows x (64 B hi | 64 B lo)
    const int swz = ((lane >> 1) & 3) ^ ((lane & 1) << 2);

    // Output of tile `it` (O buffer it & 3): O / rowsum -> fp16 hi/lo -> swizzled staging slab -> 64-byte runs.  It runs one
    // iteration late (after the softmax of this group's NEXT tile), so the PV round trip is never waited for.
    auto epilogue = [&](int it, int64_t pix, float inv, int head) {
      const int q = it & 3;
      mbar_wait(bar_o_full(q), (uint32_t)(it >> 2) & 1u);
      tc_fence_after();
      uint32_t oo[32];
      tmem_ld32(tmem_base + lane_off + A2_COL_O + (uint32_t)(q * 32), oo);
      tc_fence_before();
      mbar_arrive(bar_o_empty(q));  // O_q consumed: PV of the tile four ahead may overwrite it
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        __align__(16) __half2 h8[4];
        __align__(16) __half2 l8[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          wxf_split2_f16x2(__uint_as_float(oo[8 * c + 2 * e]) * inv, __uint_as_float(oo[8 * c + 2 * e + 1]) * inv, h8[e],
                           l8[e]);
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ swz) << 4)) = *reinterpret_cast<const uint4*>(h8);
        *reinterpret_cast<uint4*>(stg + lane * 128 + (((c + 4) ^ swz) << 4)) = *reinterpret_cast<const uint4*>(l8);
      }
      __syncwarp();
      // 4 lanes write one row's 64 bytes: 8 rows per store instruction, hi plane then lo plane
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int row = ps * 8 + (lane >> 2), c = lane & 3;
        const int64_t rp = __shfl_sync(0xffffffffu, pix, row);
        const int rs = ((row >> 1) & 3) ^ ((row & 1) << 2);
        const uint4 vh = *reinterpret_cast<const uint4*>(stg + row * 128 + ((c ^ rs) << 4));
        const uint4 vl = *reinterpret_cast<const uint4*>(stg + row * 128 + (((c + 4) ^ rs) << 4));
        if (rp >= 0) {
          *reinterpret_cast<uint4*>(p.out_hi + rp * p.ldh + head * DH + c * 8) = vh;
          *reinterpret_cast<uint4*>(p.out_lo + rp * p.ldh + head * DH + c * 8) = vl;
        }
      }
      __syncwarp();  // the slab is rewritten by the next epilogue of this warp
    };

    // the epilogue of a tile runs one iteration late for every tile shape (measured at 0.25 deg: L = 25: 109 -> 94 us, L = 4:
    // 72 -> 64 us; since the decodes left the critical path also L = 100: 162 -> 154 us)
    int pend_it = -1;
    int64_t pend_pix = -1;
    float pend_inv = 0.f;
    int pend_head = 0;
    for (int it = wg; it < my_tiles; it += 2) {
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      int head, nv;
      uint32_t w0;
      tile_windows(blockIdx.x + (uint32_t)it * gridDim.x, head, w0, nv);

      mbar_wait(bar_s_full(wg), par);
      tc_fence_after();
      // s2 = S * (scale*log2 e) + bias*log2 e (masked columns: -1e30); p = 2^(s2 - max)
      float mx;
      {
        float m0 = -3.0e38f, m1 = -3.0e38f, m2 = -3.0e38f, m3 = -3.0e38f;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t rr[32], bb[32];
          tmem_ld32_nowait(s_addr + (uint32_t)(c * 32), rr);
          tmem_ld32_nowait(b_addr + (uint32_t)(c * 32), bb);
          tmem_ld_wait();
          // every loaded column counts: a column of another window carries bias -1e30 and cannot win (measured: the
          // per-group range tests cost more in branches than the few columns they skip)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float2 t0 = __ffma2_rn(make_float2(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), sc2,
                                         make_float2(__uint_as_float(bb[j]), __uint_as_float(bb[j + 1])));
            const float2 t1 = __ffma2_rn(make_float2(__uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3])), sc2,
                                         make_float2(__uint_as_float(bb[j + 2]), __uint_as_float(bb[j + 3])));
            m0 = fmaxf(m0, t0.x); m1 = fmaxf(m1, t0.y); m2 = fmaxf(m2, t1.x); m3 = fmaxf(m3, t1.y);
          }
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      // P over S, in place: the 16 score columns of half h become its 8 hi words (columns [16h, 16h+8)) and 8 lo words
      // ([16h+8, 16h+16)) - a half only overwrites what it has just read, so no word waits in registers for another half
      const float2 nmx2 = make_float2(-mx, -mx);
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        uint32_t hi[8], lo[8];
        softmax_half(s_addr, b_addr, h, g_lo, g_hi, sc2, nmx2, sum_a, sum_b, hi, lo);
        tmem_st8_nowait(s_addr + (uint32_t)(16 * h), hi);
        tmem_st8_nowait(s_addr + (uint32_t)(16 * h + 8), lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      const float lsum = (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
      tc_fence_before();
      mbar_arrive(bar_p_full(wg));

      if (pend_it >= 0) epilogue(pend_it, pend_pix, pend_inv, pend_head);  // the previous tile of this group

      // this row's output pixel
      int64_t pix = -1;
      if (row_ok && g < nv) {
        int bi, gh, gw;
        window_pos(w0 + (uint32_t)g, bi, gh, gw);
        int y, x;
        if (p.kind == WXF_ATTN_SHORT) {
          y = gh * p.wsz + ty;
          x = gw * p.wsz + tx;
        } else {
          y = ty * p.nh + gh;
          x = tx * p.nw + gw;
        }
        pix = ((int64_t)bi * p.H + y) * p.W + x;
      }
      const float inv = __fdividef(1.0f, lsum);  // lsum >= 1 (the row maximum contributes 2^0): no division slow path needed
      pend_it = it; pend_pix = pix; pend_inv = inv; pend_head = head;
    }
    if (pend_it >= 0) epilogue(pend_it, pend_pix, pend_inv, pend_head);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// tile[c*128 + r] = bias[i(r)][j(c)] * log2(e) if row r and column c belong to the same window slot, else -1e30
__global__ void attention_bias_tile_kernel(const float* __restrict__ biasT, float* __restrict__ tile, AttnParams p) {
  const int c = blockIdx.x, r = threadIdx.x;
  int g, i, gc, jc;
  row_slot(p, r, g, i);
  row_slot(p, c, gc, jc);
  const bool ok = g < p.G && i < p.L && gc == g && jc < p.L;
  tile[c * ROWS + r] = ok ? biasT[(size_t)jc * p.L + i] * 1.4426950408889634f : -1.0e30f;
}

// how windows are packed into a 128-row tile (shared by the tile builder and the attention launch)
void attn_packing(AttnParams& p, int wsz, int kind, int nw) {
  const int L = wsz * wsz;
  p.wsz = wsz; p.kind = kind; p.nw = nw;
  p.L = L; p.Lp = (L + 1) & ~1; p.G = ROWS / p.Lp;
  p.inter = (kind == WXF_ATTN_LONG && ROWS / L >= 2) ? ((L & 1) ? 1 : 2) : 0;
  p.gpr = 1;
  if (p.inter) {
    p.G = ROWS / L < nw ? ROWS / L : nw;
    p.gpr = (nw + p.G - 1) / p.G;
  }
}

}  // namespace

extern "C" int wxf_attention_bias_tile(const float* biasT, float* tile, int W, int wsz, int kind, void* stream) {
  if (!biasT || !tile || wsz <= 0 || W <= 0 || W % wsz) WXF_FAIL(WXF_EINVAL, "attention_bias_tile: bad arguments");
  if (wsz * wsz > ROWS) WXF_FAIL(WXF_EUNSUPPORTED, "attention_bias_tile: window %d too large", wsz);
  if (kind != WXF_ATTN_SHORT && kind != WXF_ATTN_LONG) WXF_FAIL(WXF_EINVAL, "attention_bias_tile: bad kind");
  AttnParams p{};
  attn_packing(p, wsz, kind, W / wsz);
  attention_bias_tile_kernel<<<ROWS, ROWS, 0, (cudaStream_t)stream>>>(biasT, tile, p);
  WXF_CHECK_LAUNCH("attention_bias_tile");
  return 0;
}

extern "C" int wxf_window_attention_tc(const void* qkv_hi, const void* qkv_lo, int ldq, const float* bias_tile, void* out_hi,
                                       void* out_lo, int ldh, int B, int H, int W, int d, int dh, int wsz, int kind,
                                       float scale, void* stream) {
  if (!qkv_hi || !qkv_lo || !bias_tile || !out_hi || !out_lo) WXF_FAIL(WXF_EINVAL, "attention_tc: null pointer");
  if (B <= 0 || H <= 0 || W <= 0 || d <= 0 || wsz <= 0) WXF_FAIL(WXF_EINVAL, "attention_tc: bad dims");
  if (dh != DH) WXF_FAIL(WXF_EUNSUPPORTED, "attention_tc: dim_head must be 32, got %d", dh);
  if (d % dh) WXF_FAIL(WXF_EINVAL, "attention_tc: d %% dh != 0");
  if (H % wsz || W % wsz) WXF_FAIL(WXF_EINVAL, "attention_tc: grid %dx%d not divisible by window %d", H, W, wsz);
  if (kind != WXF_ATTN_SHORT && kind != WXF_ATTN_LONG) WXF_FAIL(WXF_EINVAL, "attention_tc: bad kind");
  const int L = wsz * wsz;
  if (L > ROWS) WXF_FAIL(WXF_EUNSUPPORTED, "attention_tc: window %d (L=%d) > %d tokens", wsz, L, ROWS);
  if (ldq < 3 * d || (ldq & 7) || ldh < d || (ldh & 7) || !wxf_aligned16(qkv_hi) || !wxf_aligned16(qkv_lo) ||
      !wxf_aligned16(out_hi) || !wxf_aligned16(out_lo))
    WXF_FAIL(WXF_EALIGN, "attention_tc: strides must be multiples of 8 and planes 16-byte aligned");
  const int nh = H / wsz, nw = W / wsz;
  const uint64_t ldb = (uint64_t)ldq * 2;
  AttnParams p{};
  attn_packing(p, wsz, kind, nw);
  CUtensorMap tm_hi, tm_lo;
  int rc;
  if (kind == WXF_ATTN_SHORT) {
    const uint64_t dims[4] = {(uint64_t)3 * d, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {ldb, (uint64_t)W * ldb, (uint64_t)H * W * ldb};
    const uint32_t box[4] = {(uint32_t)DH, (uint32_t)wsz, (uint32_t)wsz, 1}, es[4] = {1, 1, 1, 1};
    if ((rc = make_map(&tm_hi, qkv_hi, 4, dims, strides, box, es, 64))) return rc;
    if ((rc = make_map(&tm_lo, qkv_lo, 4, dims, strides, box, es, 64))) return rc;
  } else if (p.inter == 2) {
    // window-major rows (even L): pixel (y, x) = (l1*nh + gh, l2*nw + gw) viewed as dims (c, l2, (b, l1), gw, gh); the box
    // {32 ch, wsz, wsz, G windows, 1} lands as row = window*L + l1*wsz + l2
    const uint64_t dims[5] = {(uint64_t)3 * d, (uint64_t)wsz, (uint64_t)B * wsz, (uint64_t)nw, (uint64_t)nh};
    const uint64_t strides[4] = {(uint64_t)nw * ldb, (uint64_t)nh * W * ldb, ldb, (uint64_t)W * ldb};
    const uint32_t box[5] = {(uint32_t)DH, (uint32_t)wsz, (uint32_t)wsz, (uint32_t)p.G, 1}, es[5] = {1, 1, 1, 1, 1};
    if ((rc = make_map(&tm_hi, qkv_hi, 5, dims, strides, box, es, 64))) return rc;
    if ((rc = make_map(&tm_lo, qkv_lo, 5, dims, strides, box, es, 64))) return rc;
  } else {
    // pixel (y, x) = (l1*nh + gh, l2*nw + gw): dims (c, gw, l2, gh, (b, l1)); a group's tokens are one box
    const uint64_t dims[5] = {(uint64_t)3 * d, (uint64_t)nw, (uint64_t)wsz, (uint64_t)nh, (uint64_t)B * wsz};
    const uint64_t strides[4] = {ldb, (uint64_t)nw * ldb, (uint64_t)W * ldb, (uint64_t)nh * W * ldb};
    const uint32_t box[5] = {(uint32_t)DH, (uint32_t)(p.inter ? p.G : 1), (uint32_t)wsz, 1, (uint32_t)wsz}, es[5] = {1, 1, 1, 1, 1};
    if ((rc = make_map(&tm_hi, qkv_hi, 5, dims, strides, box, es, 64))) return rc;
    if ((rc = make_map(&tm_lo, qkv_lo, 5, dims, strides, box, es, 64))) return rc;
  }
  p.bias_tile = bias_tile;
  p.out_hi = reinterpret_cast<__half*>(out_hi);
  p.out_lo = reinterpret_cast<__half*>(out_lo);
  p.ldh = ldh;
  p.H = H; p.W = W; p.d = d; p.heads = d / dh; p.scale = scale;
  p.nh = nh;
  p.scale2 = scale * 1.4426950408889634f;
  p.nwin = (int64_t)B * nh * nw;
  const int64_t groups = p.inter ? (int64_t)B * nh * p.gpr : (p.nwin + p.G - 1) / p.G;
  p.ntiles = groups * p.heads;
  if (p.ntiles >= (int64_t)1 << 31 || p.nwin >= (int64_t)1 << 31) WXF_FAIL(WXF_EUNSUPPORTED, "attention_tc: more than 2^31 windows");
  p.fd_heads = make_fastdiv((uint32_t)p.heads);
  p.fd_gpr = make_fastdiv((uint32_t)p.gpr);
  p.fd_img = make_fastdiv((uint32_t)(nh * nw));
  p.fd_nw = make_fastdiv((uint32_t)nw);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static int use_v2 = -1;
  if (use_v2 < 0) {
    const char* e = getenv("WXF_ATTN_V2");
    use_v2 = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_v2) {
    static WxfPerDevice<bool> attr2_set_pd;
    bool& attr2_set = attr2_set_pd.get();  // function attributes are per device
    if (!attr2_set) {
      cudaError_t e = cudaFuncSetAttribute(window_attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
      if (e != cudaSuccess) WXF_FAIL((int)e, "attention_tc: cannot opt in to %d bytes of shared memory", A2_SMEM);
      attr2_set = true;
    }
    const int64_t blocks2 = p.ntiles < (int64_t)sms ? p.ntiles : (int64_t)sms;
    wxf_launch(window_attention_tc2_kernel, dim3((unsigned)blocks2), dim3(A2_THREADS), A2_SMEM, (cudaStream_t)stream, tm_hi, tm_lo, p);
    WXF_CHECK_LAUNCH("window_attention_tc2");
    return 0;
  }
  const int64_t blocks = p.ntiles < 2 * (int64_t)sms ? p.ntiles : 2 * (int64_t)sms;
  static WxfPerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.get();  // function attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) WXF_FAIL((int)e, "attention_tc: cannot opt in to %d bytes of shared memory", AT_SMEM);
    attr_set = true;
  }
  window_attention_tc_kernel<<<(unsigned)blocks, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tm_hi, tm_lo, p);
  WXF_CHECK_LAUNCH("window_attention_tc");
  return 0;
}
