// FeedForward of a d = 128 stage as ONE kernel: fc1 -> +bias -> exact-erf GELU -> fc2 -> +bias -> +residual, the
// 4d-wide hidden activation never leaves the SM.  Reference arithmetic: FeedForward.forward + residual,
// credit/models/crossformer.py:195-207, 361-363 (the LayerNorm in front is the caller's layernorm_f16x2 launch).
//
// ROUND-2 CANDIDATE: selected only with WXF_FF_FUSED=1 (miles_credit_b200/model.py); written at the end of round 1 when
// no GPU time was left, so it has NOT run on hardware yet.  The default path is the two tc_persistent_kernel launches.
//
// Why: at stage 0 (d = 128, 320 000 pixels) fc1 writes and fc2 re-reads 655 MB of hidden planes per FeedForward, and the
// fc1 epilogue (GELU + plane split + TMA stores) is issue-bound (profiles/README.md): 0.31 + 0.20 ms per FeedForward for
// 84 GFLOP.  Here a CTA owns a 128-pixel tile and walks the hidden dimension in 4 chunks of 128:
//     H   = A W1[chunk]^T            tcgen05 SS MMAs (f16x2, 3 passes), accumulators in TMEM columns [0, 256)
//     P   = split(gelu(H s1 + b1))   8 warps; written back IN PLACE as fp16 hi/lo words (TMEM columns [0, 128))
//     Y  += P W2[:, chunk]^T         tcgen05 MMAs with the A operand read from tensor memory, accumulators in [256, 512)
// and finally out = Y s2 + b2 + residual through the same swizzled-staging TMA-store epilogue as the GEMM kernel.
// Shared memory: A tile 64 KB + W1 chunk 64 KB + W2 chunk 64 KB (each single-buffered; the next chunk's weights load
// while the other MMA / the GELU run) + 32 KB staging.
#include <stdlib.h>

#include "wxf_tc_host.cuh"
#include "wxf_tc_ptx.cuh"

using namespace wxf_tc;

namespace {

constexpr int FF_D = 128;                         // model width this kernel is built for (stage 0 of WXFormer-6h)
constexpr int FF_CHUNKS = 4;                      // hidden = 4 d, walked in chunks of 128
constexpr int FF_EW = 8;                          // GELU / epilogue warps
constexpr int FF_THREADS = 64 + 32 * FF_EW;
constexpr int FF_PLANE = 128 * 64 * 2;            // 16 KB: 128 rows x 64 fp16 (one K-step of one plane)
constexpr int FF_KSTEP = 2 * FF_PLANE;            // hi | lo of one K-step: 32 KB
constexpr int FF_OFF_A = 0, FF_OFF_W1 = 2 * FF_KSTEP, FF_OFF_W2 = 4 * FF_KSTEP, FF_OFF_STG = 6 * FF_KSTEP;
constexpr int FF_OFF_BAR = FF_OFF_STG + FF_EW * 4096;
constexpr int FF_SMEM = FF_OFF_BAR + 128 + 1024;
constexpr uint32_t FF_COL_H = 0, FF_COL_Y = 256;
constexpr uint32_t FF_IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t FF_IDESC2 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct FfParams {
  const float* b1;
  const float* b2;
  const float* res;
  int64_t M;
  int ldr;
  float s1, s2;       // 2^-k: undo the power-of-two pre-scale of the weight planes
  int has_out, has_planes;
};

__device__ __forceinline__ void ff_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(FF_THREADS, 1)
ff_fused_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
                const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO_hi,
                const __grid_constant__ CUtensorMap tmO_lo, const __grid_constant__ FfParams p, const int m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + FF_OFF_BAR;
  const uint32_t a_full = bars, a_empty = bars + 8, w1_full = bars + 16, w1_empty = bars + 24, w2_full = bars + 32,
                 w2_empty = bars + 40, h_full = bars + 48, p_full = bars + 56, y_full = bars + 64, y_empty = bars + 72;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + FF_OFF_BAR + 80);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  wxf_pdl_trigger();
  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(w1_full, 1);
    mbar_init(w1_empty, 1);
    mbar_init(w2_full, 1);
    mbar_init(w2_empty, 1);
    mbar_init(h_full, 1);
    mbar_init(p_full, 32 * FF_EW);
    mbar_init(y_full, 1);
    mbar_init(y_empty, FF_EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  wxf_pdl_wait();

  // tile i of this CTA = m tile blockIdx.x + i * gridDim.x; chunk counter g = 4 i + hc (barrier parities)
  if (warp == 0) {
    if (lane == 0) {
      uint32_t i = 0;
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++i) {
        const int m0 = t * 128;
        if (i > 0) mbar_wait(a_empty, (i - 1u) & 1u);  // fc1 MMAs of the previous tile have read the A tile
        mbar_expect_tx(a_full, 2u * FF_KSTEP);
        for (int ks = 0; ks < 2; ++ks) {
          tma_load_2d(&tmA_hi, a_full, base + FF_OFF_A + ks * FF_KSTEP, ks * 64, m0);
          tma_load_2d(&tmA_lo, a_full, base + FF_OFF_A + ks * FF_KSTEP + FF_PLANE, ks * 64, m0);
        }
        for (int hc = 0; hc < FF_CHUNKS; ++hc) {
          const uint32_t g = i * FF_CHUNKS + hc;
          if (g > 0) mbar_wait(w1_empty, (g - 1u) & 1u);
          mbar_expect_tx(w1_full, 2u * FF_KSTEP);
          for (int ks = 0; ks < 2; ++ks) {  // W1 rows [128 hc, 128 hc + 128), K-step ks
            tma_load_2d(&tmW1_hi, w1_full, base + FF_OFF_W1 + ks * FF_KSTEP, ks * 64, hc * 128);
            tma_load_2d(&tmW1_lo, w1_full, base + FF_OFF_W1 + ks * FF_KSTEP + FF_PLANE, ks * 64, hc * 128);
          }
          if (g > 0) mbar_wait(w2_empty, (g - 1u) & 1u);
          mbar_expect_tx(w2_full, 2u * FF_KSTEP);
          for (int ks = 0; ks < 2; ++ks) {  // W2 rows [0, 128) (output channels), hidden columns [128 hc + 64 ks, + 64)
            tma_load_2d(&tmW2_hi, w2_full, base + FF_OFF_W2 + ks * FF_KSTEP, hc * 128 + ks * 64, 0);
            tma_load_2d(&tmW2_lo, w2_full, base + FF_OFF_W2 + ks * FF_KSTEP + FF_PLANE, hc * 128 + ks * 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t d_h = tmem_base + FF_COL_H, d_y = tmem_base + FF_COL_Y;
      uint32_t i = 0;
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++i) {
        mbar_wait(a_full, i & 1u);
        for (int hc = 0; hc < FF_CHUNKS; ++hc) {
          const uint32_t g = i * FF_CHUNKS + hc;
          // ---- H = A W1[chunk]^T.  The H / P columns are free: P(g-1) was complete before its fc2 MMAs were issued, and
          //      tcgen05.mma instructions of one thread execute in issue order.
          mbar_wait(w1_full, g & 1u);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t sa = base + FF_OFF_A + ks * FF_KSTEP + k * 32, sw = base + FF_OFF_W1 + ks * FF_KSTEP + k * 32;
              const uint32_t acc = (ks | k) ? 1u : 0u;
              tc_mma_f16(d_h, umma_desc_sw128(sa), umma_desc_sw128(sw), FF_IDESC2, acc);          // A_hi [W_hi | W_lo]
              tc_mma_f16(d_h + 128u, umma_desc_sw128(sa + FF_PLANE), umma_desc_sw128(sw), FF_IDESC, 1u);  // + A_lo W_hi
            }
          }
          tc_commit(w1_empty);
          if (hc == FF_CHUNKS - 1) tc_commit(a_empty);
          tc_commit(h_full);
          // ---- Y += P W2[:, chunk]^T with A = P from tensor memory (hi words columns [0, 64), lo words [64, 128))
          mbar_wait(p_full, g & 1u);
          if (hc == 0 && i > 0) mbar_wait(y_empty, (i - 1u) & 1u);  // the previous tile's Y has been read
          mbar_wait(w2_full, g & 1u);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t kk = (uint32_t)(ks * 4 + k);
              const uint32_t sw = base + FF_OFF_W2 + ks * FF_KSTEP + k * 32;
              const uint32_t acc = (hc | ks | k) ? 1u : 0u;
              ff_mma_ts(d_y, d_h + kk * 8u, umma_desc_sw128(sw), FF_IDESC2, acc);          // P_hi [W_hi | W_lo]
              ff_mma_ts(d_y + 128u, d_h + 64u + kk * 8u, umma_desc_sw128(sw), FF_IDESC, 1u);  // + P_lo W_hi
            }
          }
          tc_commit(w2_empty);
        }
        tc_commit(y_full);
      }
    }
  } else {
    // ---- 8 GELU / epilogue warps: quarter = TMEM lane quarter, half = which 64 of the 128 columns ----
    const int ew = warp - 2;
    const int quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint8_t* stg_b = gen + FF_OFF_STG + ew * 4096;
    const uint32_t stg_a = base + FF_OFF_STG + ew * 4096;
    uint32_t i = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++i) {
      const int64_t m0 = (int64_t)t * 128;
      const int64_t m = m0 + row;
      for (int hc = 0; hc < FF_CHUNKS; ++hc) {
        const uint32_t g = i * FF_CHUNKS + hc;
        mbar_wait(h_full, g & 1u);
        tc_fence_after();
        uint32_t hw[32], lw[32];  // this thread's 64 hidden values as packed fp16 hi / lo words
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r0[32], r1[32];
          const uint32_t col = (uint32_t)(half * 64 + c * 32);
          tmem_ld32_nowait(tmem_base + lane_off + FF_COL_H + col, r0);         // main
          tmem_ld32_nowait(tmem_base + lane_off + FF_COL_H + 128u + col, r1);  // cross terms
          tmem_ld_wait();
          const float2 sc = make_float2(p.s1, p.s1);
          const float* bp = p.b1 + hc * 128 + half * 64 + c * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + 4 * q));
            float2 a0 = __fadd2_rn(make_float2(__uint_as_float(r0[4 * q]), __uint_as_float(r0[4 * q + 1])),
                                   make_float2(__uint_as_float(r1[4 * q]), __uint_as_float(r1[4 * q + 1])));
            float2 a1 = __fadd2_rn(make_float2(__uint_as_float(r0[4 * q + 2]), __uint_as_float(r0[4 * q + 3])),
                                   make_float2(__uint_as_float(r1[4 * q + 2]), __uint_as_float(r1[4 * q + 3])));
            a0 = wxf_gelu_erf2(__ffma2_rn(a0, sc, make_float2(b4.x, b4.y)));
            a1 = wxf_gelu_erf2(__ffma2_rn(a1, sc, make_float2(b4.z, b4.w)));
            __half2 h0, l0, h1, l1;
            wxf_split2_f16x2(a0.x, a0.y, h0, l0);  // low half = even hidden index
            wxf_split2_f16x2(a1.x, a1.y, h1, l1);
            hw[c * 16 + 2 * q] = *reinterpret_cast<uint32_t*>(&h0);
            hw[c * 16 + 2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
            lw[c * 16 + 2 * q] = *reinterpret_cast<uint32_t*>(&l0);
            lw[c * 16 + 2 * q + 1] = *reinterpret_cast<uint32_t*>(&l1);
          }
        }
        // P overwrites H columns other warps of this lane quarter read: both warps of the quarter must be done reading
        tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        tc_fence_after();
        tmem_st32(tmem_base + lane_off + FF_COL_H + (uint32_t)(half * 32), hw);        // P_hi words [32 half, +32)
        tmem_st32(tmem_base + lane_off + FF_COL_H + 64u + (uint32_t)(half * 32), lw);  // P_lo words
        tc_fence_before();
        mbar_arrive(p_full);
      }

      // ---- tile epilogue: out = Y s2 + b2 + residual (fp32 tile and / or fp16 planes through TMA stores) ----
      float4 rres[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        rres[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.res && m < p.M) rres[q] = *reinterpret_cast<const float4*>(p.res + m * p.ldr + half * 64 + 4 * q);
      }
      mbar_wait(y_full, i & 1u);
      tc_fence_after();
      float v[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r0[32], r1[32];
        const uint32_t col = (uint32_t)(half * 64 + c * 32);
        tmem_ld32_nowait(tmem_base + lane_off + FF_COL_Y + col, r0);
        tmem_ld32_nowait(tmem_base + lane_off + FF_COL_Y + 128u + col, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_empty);  // Y may be overwritten by the next tile's first fc2 MMA
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        const int nb = half * 64 + cb * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + nb + 4 * q));
          const float4 rr = rres[cb * 8 + q];
          float* vv = v + cb * 32 + 4 * q;
          vv[0] = fmaf(vv[0], p.s2, b4.x) + rr.x;
          vv[1] = fmaf(vv[1], p.s2, b4.y) + rr.y;
          vv[2] = fmaf(vv[2], p.s2, b4.z) + rr.z;
          vv[3] = fmaf(vv[3], p.s2, b4.w) + rr.w;
        }
        if (p.has_out) {
          if (lane == 0) bulk_wait_read0();  // the previous TMA store has finished reading the staging buffer
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)  // fp32 row of 128 B, SWIZZLE_128B: 16-byte chunk ^= row % 8
            *reinterpret_cast<float4*>(stg_b + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_float4(v[cb * 32 + 4 * q], v[cb * 32 + 4 * q + 1], v[cb * 32 + 4 * q + 2], v[cb * 32 + 4 * q + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO, stg_a, nb, (int)(m0 + quarter * 32));
            bulk_commit();
          }
        }
        if (p.has_planes) {
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // fp16 rows of 64 B, SWIZZLE_64B: 16-byte chunk ^= (row / 2) % 4
            __align__(16) __half2 h8[4];
            __align__(16) __half2 l8[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              wxf_split2_f16x2(v[cb * 32 + 8 * q + 2 * e], v[cb * 32 + 8 * q + 2 * e + 1], h8[e], l8[e]);
            const int off = lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(stg_b + off) = *reinterpret_cast<const uint4*>(h8);
            *reinterpret_cast<uint4*>(stg_b + 2048 + off) = *reinterpret_cast<const uint4*>(l8);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO_hi, stg_a, nb, (int)(m0 + quarter * 32));
            tma_store_2d(&tmO_lo, stg_a + 2048, nb, (int)(m0 + quarter * 32));
            bulk_commit();
          }
        }
      }
    }
    if (lane == 0) bulk_wait0();  // all stores of this warp have landed before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int wxf_ff_fused_f16x2_tc(const WxfFfDesc* d, void* stream) {
  if (!d || !d->a_hi || !d->a_lo || !d->w1_hi || !d->w1_lo || !d->w2_hi || !d->w2_lo || !d->b1 || !d->b2)
    WXF_FAIL(WXF_EINVAL, "ff_fused: null operand");
  if (d->d != FF_D) WXF_FAIL(WXF_EUNSUPPORTED, "ff_fused: built for d = %d, got %d", FF_D, d->d);
  if (d->M <= 0 || d->lda < FF_D || (d->lda & 7)) WXF_FAIL(WXF_EINVAL, "ff_fused: bad dims");
  if (!d->out && !d->out_hi) WXF_FAIL(WXF_EINVAL, "ff_fused: no output");
  if ((d->out_hi == nullptr) != (d->out_lo == nullptr)) WXF_FAIL(WXF_EINVAL, "ff_fused: out_hi/out_lo must come together");
  if (d->out && (d->ldc < FF_D || (d->ldc & 3) || !wxf_aligned16(d->out))) WXF_FAIL(WXF_EALIGN, "ff_fused: out stride/alignment");
  if (d->res && (d->ldr < FF_D || (d->ldr & 3) || !wxf_aligned16(d->res))) WXF_FAIL(WXF_EALIGN, "ff_fused: res stride/alignment");
  if (d->out_hi && (d->ldh < FF_D || (d->ldh & 7) || !wxf_aligned16(d->out_hi) || !wxf_aligned16(d->out_lo)))
    WXF_FAIL(WXF_EALIGN, "ff_fused: plane stride/alignment");
  if (!wxf_aligned16(d->a_hi) || !wxf_aligned16(d->a_lo) || !wxf_aligned16(d->w1_hi) || !wxf_aligned16(d->w1_lo) ||
      !wxf_aligned16(d->w2_hi) || !wxf_aligned16(d->w2_lo) || !wxf_aligned16(d->b1) || !wxf_aligned16(d->b2))
    WXF_FAIL(WXF_EALIGN, "ff_fused: operands must be 16-byte aligned");
  if (d->M > (int64_t)INT32_MAX - 128) WXF_FAIL(WXF_EINVAL, "ff_fused: M too large");
  int rc;
  CUtensorMap ta_hi, ta_lo, t1_hi, t1_lo, t2_hi, t2_lo;
  if ((rc = make_map_2d(&ta_hi, d->a_hi, (uint64_t)d->M, FF_D, (uint64_t)d->lda, 128))) return rc;
  if ((rc = make_map_2d(&ta_lo, d->a_lo, (uint64_t)d->M, FF_D, (uint64_t)d->lda, 128))) return rc;
  if ((rc = make_map_2d(&t1_hi, d->w1_hi, 4 * FF_D, FF_D, FF_D, 128))) return rc;
  if ((rc = make_map_2d(&t1_lo, d->w1_lo, 4 * FF_D, FF_D, FF_D, 128))) return rc;
  if ((rc = make_map_2d(&t2_hi, d->w2_hi, FF_D, 4 * FF_D, 4 * FF_D, 128))) return rc;
  if ((rc = make_map_2d(&t2_lo, d->w2_lo, FF_D, 4 * FF_D, 4 * FF_D, 128))) return rc;
  CUtensorMap to = ta_hi, to_hi = ta_hi, to_lo = ta_hi;  // placeholders when an output is absent (never dereferenced)
  const uint32_t box[2] = {32, 32}, es[2] = {1, 1};
  if (d->out) {
    const uint64_t dims[2] = {(uint64_t)FF_D, (uint64_t)d->M}, strides[1] = {(uint64_t)d->ldc * 4};
    if ((rc = make_map(&to, d->out, 2, dims, strides, box, es, 128, true))) return rc;
  }
  if (d->out_hi) {
    const uint64_t dims[2] = {(uint64_t)FF_D, (uint64_t)d->M}, strides[1] = {(uint64_t)d->ldh * 2};
    if ((rc = make_map(&to_hi, d->out_hi, 2, dims, strides, box, es, 64))) return rc;
    if ((rc = make_map(&to_lo, d->out_lo, 2, dims, strides, box, es, 64))) return rc;
  }
  FfParams p{};
  p.b1 = d->b1;
  p.b2 = d->b2;
  p.res = d->res;
  p.M = d->M;
  p.ldr = d->ldr;
  p.s1 = ldexpf(1.0f, -d->w1_scale_log2);
  p.s2 = ldexpf(1.0f, -d->w2_scale_log2);
  p.has_out = d->out ? 1 : 0;
  p.has_planes = d->out_hi ? 1 : 0;
  static WxfPerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.get();  // function attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ff_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM);
    if (e != cudaSuccess) WXF_FAIL((int)e, "ff_fused: cannot opt in to %d bytes of shared memory: %s", FF_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int m_tiles = (int)((d->M + 127) / 128);
  const int grid = m_tiles < sms ? m_tiles : sms;
  wxf_launch(ff_fused_kernel, dim3(grid), dim3(FF_THREADS), FF_SMEM, (cudaStream_t)stream, ta_hi, ta_lo, t1_hi, t1_lo, t2_hi,
             t2_lo, to, to_hi, to_lo, p, m_tiles);
  WXF_CHECK_LAUNCH("ff_fused");
  return 0;
}
