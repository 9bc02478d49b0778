// Contractions on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
//   GEMM mode : C[M,N] = epi(A[M,K] * W[N,K]^T)                       (1x1 convs: to_qkv, to_out, FeedForward)
//   CONV mode : implicit GEMM over (tap, channel-block) K-steps; the A tile of a K-step is ONE 4-D TMA box
//               {64 channels, bw pixels (element stride s), bh rows (element stride s), 1 image} of the
//               pixel-major input, so there is no im2col buffer and zero padding is TMA's out-of-bounds fill.
//               Transposed convolutions run as 4 output-parity phases (gridDim.z).
//   TOEP mode : the stage-0 cross-embed (k in {4,8,16,32}, stride 2, only 16..64 output channels).  A direct
//               implicit GEMM would have N = 16: the tensor core idles on operand traffic.  Instead the kernel
//               column kx = 2j + r is split; the tap index j is folded into the GEMM's N dimension:
//                   P[oy, m, (j, c)] = sum_{ky, r, ci} X[2 oy + ky - p, 2 m + r - p, ci] * W[c, ci, ky, 2j + r]
//                   out[oy, ox, c]   = sum_j P[oy, ox + j, (j, c)]
//               so N = (k/2) * channels (128..256), K-steps = 2k taps (ky, r) of 64 channels, and the A tile of a
//               K-step is again one TMA box (128 pixels at element stride 2).  The epilogue stages P in shared
//               memory (reusing the operand ring) and does the diagonal sum over j.
//
// Precision scheme "f16x2": every fp32 operand is carried as two fp16 planes (hi = fp16(x), lo = fp16(x - hi),
// 22 significant bits); a product is three tcgen05.mma passes  A_hi W_lo + A_lo W_hi + A_hi W_hi  with fp32
// accumulation in TMEM.  The tensor-core accumulator truncates on every accumulate (measured: 1.8e-5 rel at
// K=4096 with one accumulator), so the small cross terms go to their own accumulator and the main term is
// spread over several accumulators along K; they are summed in fp32 (round-to-nearest) in the epilogue.
//
// CTA = 128 x BN output tile, 192 threads:
//   warp 0    : TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier ring)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction)
//   warps 2-5 : epilogue: tcgen05.ld -> registers -> padded smem transpose -> bias/GELU/residual ->
//               coalesced 128-byte row segments (fp32 and/or fp16 hi/lo planes)
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "wxf_tc_host.cuh"
#include "wxf_tc_ptx.cuh"

using namespace wxf_tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                        // 64 fp16 = one 128-byte swizzle row
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB: one 128-row operand plane of one stage
constexpr int NUM_THREADS = 192;
constexpr int STG_LD = 36;                         // floats per staged row (32 + 4 pad: conflict-free float4)
constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;     // 4 epilogue warps
constexpr int MAX_TAPS = 64;
constexpr int MODE_GEMM = 0, MODE_CONV = 1, MODE_TOEP = 2;

struct TcParams {
  const float* bias;
  const float* res;
  float* out;
  __half* out_hi;
  __half* out_lo;
  int64_t M;          // GEMM mode: rows
  int N, num_ksteps;  // K-steps of 64
  int ldc, ldr, ldh;
  int act;
  float w_scale;      // 2^-k: undoes the power-of-two pre-scale of the weight planes
  // conv mode
  int cblocks;        // channel blocks of 64 per tap
  int cin_pad;        // weight K stride per tap
  int bw, bh, tiles_x, tiles_y;
  int Ho, Wo, stride, out_scale;
  uint32_t a_bytes;   // bytes of one A plane box
  int bias_zs;        // bias index = phase * bias_zs + n
  int step, J, ch;    // Toeplitz mode: tile step along x (128 - (J-1)), taps folded into N, channels of the branch
  uint32_t w_bytes;   // persistent kernel: bytes of one W box (fewer than 128 rows when N < 128: up_block4 has 64 channels)
  uint32_t park_ns;   // specialised epilogues: suspend-time hint of the far-away mbarrier waits (0 = plain spin)
  int16_t taps[4][MAX_TAPS][2];  // [phase][tap] = (dy, dx); only [0] used when phases == 1
};

// ---- kernel --------------------------------------------------------------------------------------

template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_contract_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                   const __grid_constant__ TcParams p) {
  constexpr int W_BYTES = BN * BLOCK_K * 2;
  constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * W_BYTES;
  constexpr int NACC = 512 / BN;        // TMEM accumulators: [0] cross terms, [1..] main term split along K
  constexpr int NMAIN = NACC - 1;
  constexpr bool CONV = MODE != MODE_GEMM;   // operand staging of TOEP is the conv staging (taps = (ky, r))
  // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle-128B atoms need 1024-byte alignment
  uint8_t* gen = smem_raw + (base - raw);
  float* staging = reinterpret_cast<float*>(gen + STAGES * STAGE_BYTES);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES + STG_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + STG_BYTES + 8 * (2 * STAGES + 1));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int z = blockIdx.z;
  const int num_k = p.num_ksteps;

  // tile origin
  int64_t m0 = 0;
  int tb = 0, oy0 = 0, ox0 = 0;
  if constexpr (CONV) {
    const int per_img = p.tiles_x * p.tiles_y;
    tb = blockIdx.y / per_img;
    const int rem = blockIdx.y - tb * per_img;
    oy0 = (rem / p.tiles_x) * p.bh;
    ox0 = (rem % p.tiles_x) * (MODE == MODE_TOEP ? p.step : p.bw);
  } else {
    m0 = (int64_t)blockIdx.y * BLOCK_M;
  }

  wxf_pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // all 512 TMEM columns: NACC accumulators of BN fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  wxf_pdl_wait();  // everything above overlaps the previous kernel's tail (programmatic dependent launch)

  if (warp == 0) {
    // ---- TMA producer: converged warp, an elected lane issues (see elect_one in wxf_tc_ptx.cuh) ----
    const uint32_t stage_tx = CONV ? (2u * p.a_bytes + 2u * W_BYTES) : (uint32_t)STAGE_BYTES;
    int s = 0;
    uint32_t ph = 0;
    int t = 0, cb = 0;  // K-step = (tap, channel block), advanced without a division per step
    for (int ks = 0; ks < num_k; ++ks) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t st = base + s * STAGE_BYTES;
      int wk, c0 = 0, c1 = 0, c2 = 0;
      if constexpr (CONV) {
        c2 = oy0 * p.stride + p.taps[z][t][0];
        c1 = ox0 * p.stride + p.taps[z][t][1];
        c0 = cb * BLOCK_K;
        wk = t * p.cin_pad + cb * BLOCK_K;
      } else {
        wk = ks * BLOCK_K;
      }
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), stage_tx);
        if constexpr (CONV) {
          tma_load_4d(&tmA_hi, full_bar(s), st, c0, c1, c2, tb);
          tma_load_4d(&tmA_lo, full_bar(s), st + TILE_BYTES, c0, c1, c2, tb);
        } else {
          tma_load_2d(&tmA_hi, full_bar(s), st, wk, (int)m0);
          tma_load_2d(&tmA_lo, full_bar(s), st + TILE_BYTES, wk, (int)m0);
        }
        tma_load_2d(&tmW_hi, full_bar(s), st + 2 * TILE_BYTES, wk, z * p.N + n0);
        tma_load_2d(&tmW_lo, full_bar(s), st + 2 * TILE_BYTES + W_BYTES, wk, z * p.N + n0);
      }
      __syncwarp();
      if constexpr (CONV) {
        if (++cb == p.cblocks) {
          cb = 0;
          ++t;
        }
      }
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: converged warp, descriptors advance by integer adds on the low word, an elected lane issues ----
    uint32_t started = 0;  // bit a: accumulator a already holds data
    int s = 0;
    uint32_t ph = 0;
    int am = 1, am_rem = 0;  // main accumulator of this K-step = 1 + ks * NMAIN / num_k, advanced without a division
    for (int ks = 0; ks < num_k; ++ks) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t lo = umma_desc_lo(base + s * STAGE_BYTES);
      const uint32_t d_cross = tmem_base, d_main = tmem_base + (uint32_t)(am * BN);
      const uint32_t acc_cross = started & 1u, acc_main = (started >> am) & 1u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t a_hi = umma_desc_make(lo + 2 * k, UMMA_SW128_HI);
          const uint64_t a_lo = umma_desc_make(lo + (TILE_BYTES >> 4) + 2 * k, UMMA_SW128_HI);
          const uint64_t w_hi = umma_desc_make(lo + (2 * TILE_BYTES >> 4) + 2 * k, UMMA_SW128_HI);
          const uint64_t w_lo = umma_desc_make(lo + ((2 * TILE_BYTES + W_BYTES) >> 4) + 2 * k, UMMA_SW128_HI);
          tc_mma_f16(d_cross, a_hi, w_lo, IDESC, k ? 1u : acc_cross);
          tc_mma_f16(d_cross, a_lo, w_hi, IDESC, 1u);
          tc_mma_f16(d_main, a_hi, w_hi, IDESC, k ? 1u : acc_main);
        }
        tc_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
        if (ks == num_k - 1) tc_commit(tmem_full_bar);  // accumulators complete
      }
      __syncwarp();
      started |= 1u | (1u << am);
      am_rem += NMAIN;
      while (am_rem >= num_k) {
        am_rem -= num_k;
        ++am;
      }
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ----
    if constexpr (MODE == MODE_TOEP) {
      const int quarter = warp & 3;
      const int used_main = num_k < NMAIN ? num_k : NMAIN;
      constexpr int SLD = BN + 4;
      float* S = reinterpret_cast<float*>(gen);  // the operand ring is idle once the accumulators are complete
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int trow = quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (c * 32 >= p.N) break;
        float v[32];
        uint32_t r[32];
        tmem_ld32(lane_base + (uint32_t)(c * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll 1
        for (int a = 1; a <= used_main; ++a) {
          tmem_ld32(lane_base + (uint32_t)(a * BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(S + trow * SLD + c * 32 + j) =
              make_float4(v[j] * p.w_scale, v[j + 1] * p.w_scale, v[j + 2] * p.w_scale, v[j + 3] * p.w_scale);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
      // diagonal sum: out[ox0 + i, c] = sum_j P[i + j, j*ch + c]
      const int i = trow;
      const int ox = ox0 + i;
      if (i < p.step && ox < p.Wo) {
        const int64_t pix = ((int64_t)tb * p.Ho + oy0) * p.Wo + ox;
        float* orow = p.out + pix * p.ldc;
        for (int c4 = 0; c4 < p.ch; c4 += 4) {
          float4 acc = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j = 0; j < p.J; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(S + (i + j) * SLD + j * p.ch + c4);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          *reinterpret_cast<float4*>(orow + c4) = acc;
        }
      }
    } else {
    const int quarter = warp & 3;
    float* stg = staging + quarter * (32 * STG_LD);
    const int r4 = lane >> 3, c4 = lane & 7;
    const int used_main = num_k < NMAIN ? num_k : NMAIN;

    // output pixel of the 8 rows this lane stores (row = it*4 + r4 of this warp's 32-row slab)
    int64_t opix[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = quarter * 32 + it * 4 + r4;
      if constexpr (CONV) {
        const int i = row / p.bw, j = row - i * p.bw;
        const int oy = oy0 + i, ox = ox0 + j;
        if (i < p.bh && oy < p.Ho && ox < p.Wo) {
          const int Hout = p.Ho * p.out_scale, Wout = p.Wo * p.out_scale;
          opix[it] = ((int64_t)tb * Hout + oy * p.out_scale + (z >> 1)) * Wout + ox * p.out_scale + (z & 1);
        } else {
          opix[it] = -1;
        }
      } else {
        const int64_t m = m0 + row;
        opix[it] = m < p.M ? m : -1;
      }
    }

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;  // warp-uniform
      float v[32];
      {
        uint32_t r[32];
        tmem_ld32(lane_base + (uint32_t)(c * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll 1
        for (int a = 1; a <= used_main; ++a) {
          tmem_ld32(lane_base + (uint32_t)(a * BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
        }
      }
      // lane = row: write the 32 columns of this row into the padded staging slab
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(stg + lane * STG_LD + j) =
            make_float4(v[j] * p.w_scale, v[j + 1] * p.w_scale, v[j + 2] * p.w_scale, v[j + 3] * p.w_scale);
      __syncwarp();
      // transposed pass: 8 lanes cover one row's 32 columns (128 B), 4 rows per instruction
      const int n = nb + c4 * 4;
      if (n < p.N) {
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + z * p.bias_zs + n));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (opix[it] < 0) continue;
          float4 t = *reinterpret_cast<const float4*>(stg + (it * 4 + r4) * STG_LD + c4 * 4);
          t.x += bias4.x; t.y += bias4.y; t.z += bias4.z; t.w += bias4.w;
          if (p.act == WXF_ACT_GELU_ERF) {
            t.x = wxf_gelu_erf(t.x); t.y = wxf_gelu_erf(t.y); t.z = wxf_gelu_erf(t.z); t.w = wxf_gelu_erf(t.w);
          }
          if (p.res) {
            const float4 q = *reinterpret_cast<const float4*>(p.res + opix[it] * p.ldr + n);
            t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
          }
          if (p.out) *reinterpret_cast<float4*>(p.out + opix[it] * p.ldc + n) = t;
          if (p.out_hi) {
            __align__(8) __half2 h4[2];
            __align__(8) __half2 l4[2];
            wxf_split2_f16x2(t.x, t.y, h4[0], l4[0]);
            wxf_split2_f16x2(t.z, t.w, h4[1], l4[1]);
            *reinterpret_cast<uint2*>(p.out_hi + opix[it] * p.ldh + n) = *reinterpret_cast<const uint2*>(h4);
            *reinterpret_cast<uint2*>(p.out_lo + opix[it] * p.ldh + n) = *reinterpret_cast<const uint2*>(l4);
          }
        }
      }
      __syncwarp();
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---- persistent kernel: one CTA per SM loops over output tiles; TMEM holds two accumulator slots so the epilogue of
//      tile i overlaps the TMA loads and MMAs of tile i+1; 8 epilogue warps (two per TMEM lane quarter) -------------

constexpr int P_BN = 128;
constexpr int P_STAGE_BYTES = 2 * TILE_BYTES + 2 * P_BN * BLOCK_K * 2;  // 64 KB
// Two shapes: <8 epilogue warps, 3 stages> for K >= 512 (main loop bound) and <16 epilogue warps, 2 stages> for the
// small-K launches of stages 0-1, whose cost is the epilogue (issue-bound with 2 warps per scheduler).
constexpr int P_STG_LD = 20;                                            // 16 columns + 4 pad per staged row
template <int EW, int STAGES>
constexpr int p_smem() {
  return STAGES * P_STAGE_BYTES + EW * 4096 + 8 * (2 * STAGES + 4) + 16 + 1024;
}

// Specialised GEMM epilogues (EPI != 0; EPI = 0 is the generic one with run-time switches).  The ncu source page of the
// stage-0 fc1 launch showed 950 issued instructions per warp and tile where ~500 are needed: predicated-off residual code,
// spilled residual registers, eight predicated bias loads with their address arithmetic, unpacked accumulator adds, a
// two-division tile decode.  With the switches as template parameters none of that is compiled in.
#ifndef WXF_ABLATE
#define WXF_ABLATE 0  // diagnosis builds only (tools/build_ablate.sh): 1 = no TMA stores, 2 = no GELU, 4 = drain only; results are wrong
#endif
constexpr int EPI_GELU = 1;    // erf GELU
constexpr int EPI_RED = 2;     // in-place residual (out == res): the fp32 tile leaves as a TMA REDUCE-ADD store, the L2 does x += f(x)
constexpr int EPI_F32 = 4;     // fp32 output tile
constexpr int EPI_PLANES = 8;  // fp16 hi / lo operand planes

// Toeplitz epilogue, staging of one pass: the thread's 64 accumulator columns col = half*64 + e hold tap j = col / CH, channel
// c = col % CH; pass Q takes the channels [Q CH/4, (Q+1) CH/4) and writes them to st[row][j * CH/4 + (c - Q CH/4)].  With CH
// and Q compile-time every register index and shared-memory offset is static: 16 stores per pass.
template <int CH, int Q>
__device__ __forceinline__ void toep_stage_pass(const float (&v)[64], float* st_row, int half) {
  constexpr int JL = 64 / CH, C4 = CH / 4;  // taps per 64-column half, channels per pass
#pragma unroll
  for (int jl = 0; jl < JL; ++jl) {
    float* dst = st_row + (half * JL + jl) * C4;
#pragma unroll
    for (int cq = 0; cq < C4; ++cq) dst[cq] = v[jl * CH + Q * C4 + cq];
  }
}
template <int CH>
__device__ __forceinline__ void toep_stage(const float (&v)[64], float* st_row, int half, int q) {
  switch (q) {
    case 0: toep_stage_pass<CH, 0>(v, st_row, half); break;
    case 1: toep_stage_pass<CH, 1>(v, st_row, half); break;
    case 2: toep_stage_pass<CH, 2>(v, st_row, half); break;
    default: toep_stage_pass<CH, 3>(v, st_row, half); break;
  }
}

// ---- specialised GEMM epilogue of the persistent kernels (EPI != 0): lane = output row, 32-column chunks, every switch a
//      template parameter.  Shared by tc_persistent_kernel and tc_resident_w_kernel (same TMEM slot / barrier protocol).
template <int EW, int EPI>
__device__ __forceinline__ void gemm_epilogue_fast(const TcParams& p, const CUtensorMap& tmO, const CUtensorMap& tmO_hi,
                                                   const CUtensorMap& tmO_lo, const uint32_t tmem_base, const uint32_t tfull0,
                                                   const uint32_t tempty0, uint8_t* staging_bytes, const uint32_t staging_addr,
                                                   const int n_tiles, const int total_tiles, const int warp, const int lane) {
  constexpr int BN = P_BN;
  constexpr int CW = 128 / (EW / 4);
  constexpr int NCH = CW / 32;
  const int ew = warp - 2;
  const int quarter = warp & 3, half = ew >> 2;
  auto tfull_bar = [&](int s) { return tfull0 + 8u * s; };
  auto tempty_bar = [&](int s) { return tempty0 + 8u * s; };
  // ---- specialised GEMM epilogue: lane = output row, 32-column chunks, everything decided at compile time ----
  constexpr bool E_GELU = (EPI & EPI_GELU) != 0, E_RED = (EPI & EPI_RED) != 0, E_F32 = (EPI & EPI_F32) != 0,
                 E_PLANES = (EPI & EPI_PLANES) != 0;
  static_assert(E_F32 || E_PLANES, "an epilogue needs an output");
  static_assert(!E_RED || E_F32, "the reduce-add store carries the fp32 tile");
  uint8_t* stg_b = staging_bytes + ew * 4096;
  const uint32_t stg_a = staging_addr + ew * 4096;
  const bool has_bias = p.bias != nullptr;  // (the host takes this path only when N % 32 == 0: no column predicates)
  // tile coordinates advance incrementally (n fastest): no division per tile
  int nt = (int)(blockIdx.x % (unsigned)n_tiles), mt = (int)(blockIdx.x / (unsigned)n_tiles);
  const int dn = (int)(gridDim.x % (unsigned)n_tiles), dm = (int)(gridDim.x / (unsigned)n_tiles);
  const float2 sc = make_float2(p.w_scale, p.w_scale);
  int i = 0;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
    const int slot = i & 1;
    const int nb0 = nt * BN + half * CW;
    const int row0 = mt * BLOCK_M + quarter * 32;
    nt += dn;
    mt += dm;
    if (nt >= n_tiles) {
      nt -= n_tiles;
      ++mt;
    }
    mbar_wait_parked(tfull_bar(slot), ((uint32_t)i >> 1) & 1u, p.park_ns);
    tc_fence_after();
    float2 acc[NCH][16];
    {
      const uint32_t tb_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 2 * BN + half * CW);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t ra[32], rb[32];
        tmem_ld32_nowait(tb_addr + (uint32_t)(c * 32), ra);
        tmem_ld32_nowait(tb_addr + (uint32_t)(BN + c * 32), rb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          acc[c][j] = __fadd2_rn(make_float2(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1])),
                                 make_float2(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1])));
      }
    }
    tc_fence_before();
    __syncwarp();
    if (elect_one()) mbar_arrive(tempty_bar(slot));  // TMEM slot free for the MMA of tile i+2

#if WXF_ABLATE & 4
    continue;  // diagnosis build: accumulator drained, nothing else
#endif
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int nb = nb0 + c * 32;
      if (nb >= p.N) break;  // warp-uniform (columns beyond N inside a chunk are clipped by the TMA store)
      float2* a = acc[c];
      if (has_bias) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + nb);  // the same address in every lane: one transaction
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(bp + q);
          a[2 * q] = __ffma2_rn(a[2 * q], sc, make_float2(b4.x, b4.y));
          a[2 * q + 1] = __ffma2_rn(a[2 * q + 1], sc, make_float2(b4.z, b4.w));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = __fmul2_rn(a[j], sc);
      }
#if !(WXF_ABLATE & 2)
      if constexpr (E_GELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = wxf_gelu_erf2_relu(a[j]);
      }
#endif
      if constexpr (E_F32) {
        if (elect_one()) bulk_wait_read0();  // the previous TMA store has finished reading the staging buffer
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)  // fp32 row of 128 B, SWIZZLE_128B: 16-byte chunk ^= row % 8
          *reinterpret_cast<float4*>(stg_b + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              make_float4(a[2 * q].x, a[2 * q].y, a[2 * q + 1].x, a[2 * q + 1].y);
        fence_proxy_async();
        __syncwarp();
#if !(WXF_ABLATE & 1)
        if (elect_one()) {
          if constexpr (E_RED) tma_reduce_add_2d(&tmO, stg_a, nb, row0);
          else tma_store_2d(&tmO, stg_a, nb, row0);
          bulk_commit();
        }
#endif
      }
      if constexpr (E_PLANES) {
        if (elect_one()) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // fp16 rows of 64 B, SWIZZLE_64B: 16-byte chunk ^= (row / 2) % 4
          __align__(16) __half2 h8[4];
          __align__(16) __half2 l8[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) wxf_split2_f16x2(a[4 * q + e].x, a[4 * q + e].y, h8[e], l8[e]);
          const int off = lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(stg_b + off) = *reinterpret_cast<const uint4*>(h8);
          *reinterpret_cast<uint4*>(stg_b + 2048 + off) = *reinterpret_cast<const uint4*>(l8);
        }
        fence_proxy_async();
        __syncwarp();
#if !(WXF_ABLATE & 1)
        if (elect_one()) {
          tma_store_2d(&tmO_hi, stg_a, nb, row0);
          tma_store_2d(&tmO_lo, stg_a + 2048, nb, row0);
          bulk_commit();
        }
#endif
      }
    }
  }
  if (elect_one()) bulk_wait0();  // all stores of this warp have landed before the CTA exits
}

template <int MODE, int EW, int STAGES, int EPI = 0>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                     const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO_hi,
                     const __grid_constant__ CUtensorMap tmO_lo, const __grid_constant__ TcParams p, const int n_tiles,
                     const int m_tiles, const int total_tiles) {
  constexpr bool TOEP = MODE == MODE_TOEP;                  // Toeplitz-lifted stage-0 cross-embed: conv staging, diagonal-sum epilogue
  constexpr bool CONV = MODE == MODE_CONV || TOEP;
  constexpr int BN = P_BN, STAGE_BYTES = P_STAGE_BYTES, W_BYTES = P_BN * BLOCK_K * 2;
  constexpr int P_STG_BYTES = EW * 4096;   // one 4 KB (1024-byte aligned) staging buffer per epilogue warp
  constexpr int CW = 128 / (EW / 4);       // columns per epilogue warp: 64 (8 warps) or 32 (16 warps)
  constexpr int NCH = CW / 32;             // 32-column chunks per warp
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  float* staging = reinterpret_cast<float*>(gen + STAGES * STAGE_BYTES);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES + P_STG_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + P_STG_BYTES + 8 * (2 * STAGES + 4));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.num_ksteps;

  wxf_pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EW);  // one arrival per epilogue warp
    }
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
  wxf_pdl_wait();  // everything above overlaps the previous kernel's tail (programmatic dependent launch)

  // tile t -> (n tile, phase, m tile); n fastest so CTAs running concurrently share the A tile in L2, then the output-parity
  // phases of a transposed / sub-pixel convolution: the 4 phases of an M tile read the same input pixels, and with the
  // phase as the slowest index every phase swept the whole input through DRAM again (ncu: up_block4 moved 1.65 GB for
  // 0.65 GB of tensors)
  const int n_phases = total_tiles / (n_tiles * m_tiles);
  auto decode = [&](int t, int& n0, int& z, int64_t& m0, int& tb, int& oy0, int& ox0) {
    const int nt = t % n_tiles;
    const int rest = t / n_tiles;
    z = rest % n_phases;
    const int mt = rest / n_phases;
    n0 = nt * BN;
    m0 = 0; tb = 0; oy0 = 0; ox0 = 0;
    if constexpr (CONV) {
      const int per_img = p.tiles_x * p.tiles_y;
      tb = mt / per_img;
      const int rem = mt - tb * per_img;
      oy0 = (rem / p.tiles_x) * p.bh;
      ox0 = (rem % p.tiles_x) * (TOEP ? p.step : p.bw);
    } else {
      m0 = (int64_t)mt * BLOCK_M;
    }
  };

  if (warp == 0) {
    // ---- TMA producer: the whole warp runs the loop converged, one elected lane issues (see elect_one) ----
    const uint32_t stage_tx = (CONV ? 2u * p.a_bytes : 2u * (uint32_t)TILE_BYTES) + 2u * p.w_bytes;
    int s = 0;
    uint32_t ph = 0;  // ring position: stage and its phase parity
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n0, z, tb, oy0, ox0;
      int64_t m0;
      decode(t, n0, z, m0, tb, oy0, ox0);
      int tp = 0, cb = 0;  // K-step = (tap, channel block), advanced without a division per step
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait_parked(empty_bar(s), ph ^ 1u, EPI != 0 ? p.park_ns : 0u);
        const uint32_t st = base + s * STAGE_BYTES;
        int wk, c0 = 0, c1 = 0, c2 = 0;
        if constexpr (CONV) {
          c2 = oy0 * p.stride + p.taps[z][tp][0];
          c1 = ox0 * p.stride + p.taps[z][tp][1];
          c0 = cb * BLOCK_K;
          wk = tp * p.cin_pad + cb * BLOCK_K;
        } else {
          wk = ks * BLOCK_K;
        }
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), stage_tx);
          if constexpr (CONV) {
            tma_load_4d(&tmA_hi, full_bar(s), st, c0, c1, c2, tb);
            tma_load_4d(&tmA_lo, full_bar(s), st + TILE_BYTES, c0, c1, c2, tb);
          } else {
            tma_load_2d(&tmA_hi, full_bar(s), st, wk, (int)m0);
            tma_load_2d(&tmA_lo, full_bar(s), st + TILE_BYTES, wk, (int)m0);
          }
          tma_load_2d(&tmW_hi, full_bar(s), st + 2 * TILE_BYTES, wk, z * p.N + n0);
          tma_load_2d(&tmW_lo, full_bar(s), st + 2 * TILE_BYTES + W_BYTES, wk, z * p.N + n0);
        }
        __syncwarp();
        if constexpr (CONV) {
          if (++cb == p.cblocks) {
            cb = 0;
            ++tp;
          }
        }
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: converged warp, descriptors advance by integer adds on their low word, one elected lane issues ----
    int s = 0;
    uint32_t ph = 0;
    int i = 0;  // local tile counter
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int slot = i & 1;
      mbar_wait_parked(tempty_bar(slot), (((uint32_t)i >> 1) & 1u) ^ 1u, EPI != 0 ? p.park_ns : 0u);  // epilogue drained the slot
      tc_fence_after();
      const uint32_t d_cross = tmem_base + (uint32_t)(slot * 2 * BN), d_main = d_cross + BN;
      // a ragged last N tile (N % 128 != 0, e.g. the 64 output channels of up_block4) issues MMAs of only the columns that
      // exist, rounded up to 16 (the weight rows beyond N are TMA zero fill; the epilogue never stores those columns)
      const int n_left = p.N - (t % n_tiles) * BN;
      const int n_mma = n_left >= BN ? BN : ((n_left + 15) & ~15);
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t lo = umma_desc_lo(base + s * STAGE_BYTES);
        if (elect_one()) {
          if (n_mma == BN) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t a_hi = umma_desc_make(lo + 2 * k, UMMA_SW128_HI);
              const uint64_t a_lo = umma_desc_make(lo + (TILE_BYTES >> 4) + 2 * k, UMMA_SW128_HI);
              const uint64_t w_hi = umma_desc_make(lo + (2 * TILE_BYTES >> 4) + 2 * k, UMMA_SW128_HI);
              const uint32_t acc = (ks | k) ? 1u : 0u;
              // W_hi and W_lo tiles are adjacent in the stage: ONE N=256 instruction computes A_hi [W_hi | W_lo] into the
              // slot's two accumulators (A_hi is read from shared memory once instead of twice), then A_lo W_hi is added
              // to the second one.  The epilogue sums both accumulators, so which one holds "main" does not matter.
              tc_mma_f16(d_cross, a_hi, w_hi, IDESC2, acc);
              tc_mma_f16(d_main, a_lo, w_hi, IDESC, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t a_hi = umma_desc_make(lo + 2 * k, UMMA_SW128_HI);
              const uint64_t a_lo = umma_desc_make(lo + (TILE_BYTES >> 4) + 2 * k, UMMA_SW128_HI);
              const uint64_t w_hi = umma_desc_make(lo + (2 * TILE_BYTES >> 4) + 2 * k, UMMA_SW128_HI);
              const uint64_t w_lo = umma_desc_make(lo + ((2 * TILE_BYTES + W_BYTES) >> 4) + 2 * k, UMMA_SW128_HI);
              const uint32_t acc = (ks | k) ? 1u : 0u;
              tc_mma_f16(d_cross, a_hi, w_hi, idesc_n, acc);
              tc_mma_f16(d_main, a_hi, w_lo, idesc_n, acc);
              tc_mma_f16(d_main, a_lo, w_hi, idesc_n, 1u);
            }
          }
          tc_commit(empty_bar(s));
          if (ks == num_k - 1) tc_commit(tfull_bar(slot));
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ---- epilogue warps: quarter = TMEM lane quarter, half = which CW-column slice of the 128 columns ----
    const int ew = warp - 2;
    const int quarter = warp & 3, half = ew >> 2;
    if constexpr (!CONV && EPI != 0) {
      gemm_epilogue_fast<EW, EPI>(p, tmO, tmO_hi, tmO_lo, tmem_base, tfull_bar(0), tempty_bar(0),
                                  reinterpret_cast<uint8_t*>(staging), base + STAGES * STAGE_BYTES, n_tiles, total_tiles, warp,
                                  lane);
    } else if constexpr (TOEP) {
      // ---- Toeplitz epilogue: out[ox0 + i, c] = bias[c] + sum_j P[i + j, j*ch + c] (P = the 128 x 128 accumulator tile).
      // The diagonal sum crosses rows, i.e. TMEM lanes and warps: P goes through shared memory in four passes of ch/4
      // channels (32 of the 128 columns each, [128][33] floats in the staging area), 256 threads sum and store.
      static_assert(EW == 8, "written for 8 epilogue warps");
      float* st = staging;
      constexpr int SLD = 33;
      const int row = quarter * 32 + lane;
      const int tid = ew * 32 + lane;        // 0..255
      const int di = tid & 127, dh = tid >> 7;
      const int ch = p.ch, ch4 = p.ch >> 2, ch8 = p.ch >> 3;
      float* bias_s = st + 128 * SLD + 32;     // the branch's bias (<= 64 floats) behind the staging tile, loaded once
      if (tid < 64) bias_s[tid] = (p.bias && tid < ch) ? __ldg(p.bias + tid) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int i = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
        const int slot = i & 1;
        int n0, z, tb, oy0, ox0;
        int64_t m0;
        decode(t, n0, z, m0, tb, oy0, ox0);
        mbar_wait(tfull_bar(slot), ((uint32_t)i >> 1) & 1u);
        tc_fence_after();
        float v[64];
        {
          const uint32_t tb_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 2 * BN + half * 64);
          uint32_t ra[32], rb[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld32_nowait(tb_addr + (uint32_t)(c * 32), ra);
            tmem_ld32_nowait(tb_addr + (uint32_t)(BN + c * 32), rb);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] = (__uint_as_float(ra[j]) + __uint_as_float(rb[j])) * p.w_scale;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(tempty_bar(slot));  // TMEM slot free for the MMA of tile i+2
        const int64_t pix = ((int64_t)tb * p.Ho + oy0) * p.Wo + ox0 + di;
        const bool live = di < p.step && ox0 + di < p.Wo;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          // ch is 64, 32 or 16 (N = J ch = 128 with J = 2, 4, 8): static register indices and offsets per case
          if (ch == 64) toep_stage<64>(v, st + row * SLD, half, q);
          else if (ch == 32) toep_stage<32>(v, st + row * SLD, half, q);
          else toep_stage<16>(v, st + row * SLD, half, q);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (live) {
            const int c0 = dh * ch8;  // this thread's channels of the pass: [c0, c0 + ch/8)
            for (int cc = 0; cc < ch8; cc += 2) {
              const int c = q * ch4 + c0 + cc;
              float2 acc = *reinterpret_cast<const float2*>(bias_s + c);
              for (int j = 0; j < p.J; ++j) {
                const float* sp = st + (di + j) * SLD + j * ch4 + c0 + cc;
                acc.x += sp[0];
                acc.y += sp[1];
              }
              *reinterpret_cast<float2*>(p.out + pix * p.ldc + c) = acc;
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
    } else if constexpr (!CONV) {
      // GEMM mode: lane = output row.  Accumulator row -> registers -> scale/bias/GELU/residual -> 32-column chunks
      // staged in swizzled shared memory -> TMA store (fp32 tile and/or fp16 hi/lo plane tiles).
      uint8_t* stg_b = reinterpret_cast<uint8_t*>(staging) + ew * 4096;
      const uint32_t stg_a = base + STAGES * STAGE_BYTES + ew * 4096;
      const int row = quarter * 32 + lane;
      int i = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
        const int slot = i & 1;
        int n0, z, tb, oy0, ox0;
        int64_t m0;
        decode(t, n0, z, m0, tb, oy0, ox0);
        const int nb0 = n0 + half * CW;
        const int64_t m = m0 + row;
        // residual prefetch (the row's 64 columns), issued before waiting for the accumulator
        float4 rres[CW / 4];
        if (p.res) {
#pragma unroll
          for (int q = 0; q < CW / 4; ++q) {
            rres[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < p.M && nb0 + 4 * q < p.N) rres[q] = *reinterpret_cast<const float4*>(p.res + m * p.ldr + nb0 + 4 * q);
          }
        }
        mbar_wait(tfull_bar(slot), ((uint32_t)i >> 1) & 1u);
        tc_fence_after();
        float v[CW];
        {
          const uint32_t tb_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 2 * BN + half * CW);
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            tmem_ld32(tb_addr + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r[j]);
            tmem_ld32(tb_addr + (uint32_t)(BN + c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(slot));  // TMEM slot free for the MMA of tile i+2

#pragma unroll
        for (int cb = 0; cb < NCH; ++cb) {
          const int nb = nb0 + cb * 32;
          if (nb >= p.N) break;  // warp-uniform
          {
            const float2 sc = make_float2(p.w_scale, p.w_scale);
            float2* v2 = reinterpret_cast<float2*>(v + cb * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {  // packed fp32 math (FFMA2): scale + bias, GELU, residual
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias && nb + 4 * q < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + 4 * q));
              float2 a0 = __ffma2_rn(v2[2 * q], sc, make_float2(b4.x, b4.y));
              float2 a1 = __ffma2_rn(v2[2 * q + 1], sc, make_float2(b4.z, b4.w));
              if (p.act == WXF_ACT_GELU_ERF) {
                a0 = wxf_gelu_erf2(a0);
                a1 = wxf_gelu_erf2(a1);
              }
              if (p.res) {
                const float4 rr = rres[cb * 8 + q];
                a0 = __fadd2_rn(a0, make_float2(rr.x, rr.y));
                a1 = __fadd2_rn(a1, make_float2(rr.z, rr.w));
              }
              v2[2 * q] = a0;
              v2[2 * q + 1] = a1;
            }
          }
          if (p.out) {
            if (lane == 0) bulk_wait_read0();  // previous TMA store has finished reading the staging buffer
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
          if (p.out_hi) {
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
    } else {
    static_assert(!CONV || EW == 8, "conv epilogue is written for 8 epilogue warps");
    float* stg = staging + ew * (32 * P_STG_LD);
    const int r8 = lane >> 2, c4 = lane & 3;  // transposed pass: 8 rows x 4 float4 (16 columns) per instruction
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int slot = i & 1;
      int n0, z, tb, oy0, ox0;
      int64_t m0;
      decode(t, n0, z, m0, tb, oy0, ox0);
      const int nb0 = n0 + half * 64;

      int64_t opix[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = quarter * 32 + it * 8 + r8;
        if constexpr (CONV) {
          const int ii = row / p.bw, jj = row - ii * p.bw;
          const int oy = oy0 + ii, ox = ox0 + jj;
          if (ii < p.bh && oy < p.Ho && ox < p.Wo) {
            const int Hout = p.Ho * p.out_scale, Wout = p.Wo * p.out_scale;
            opix[it] = ((int64_t)tb * Hout + oy * p.out_scale + (z >> 1)) * Wout + ox * p.out_scale + (z & 1);
          } else {
            opix[it] = -1;
          }
        } else {
          const int64_t m = m0 + row;
          opix[it] = m < p.M ? m : -1;
        }
      }
      // residual prefetch: issued before waiting for the accumulator so it overlaps the main loop
      float4 rres[4][4];
      if (p.res) {
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const int n = nb0 + sub * 16 + c4 * 4;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            rres[sub][it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < p.N && opix[it] >= 0) rres[sub][it] = *reinterpret_cast<const float4*>(p.res + opix[it] * p.ldr + n);
          }
        }
      }

      mbar_wait(tfull_bar(slot), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      float v[64];
      {
        const uint32_t tb_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 2 * BN + half * 64);
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld32(tb_addr + (uint32_t)(c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r[j]);
          tmem_ld32(tb_addr + (uint32_t)(BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[c * 32 + j] += __uint_as_float(r[j]);
        }
      }
      // this warp no longer needs the TMEM slot: release it before the (slow) global stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(slot)) : "memory");

#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        const int nb = nb0 + sub * 16;
        if (nb >= p.N) break;  // warp-uniform
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(stg + lane * P_STG_LD + j) =
              make_float4(v[sub * 16 + j] * p.w_scale, v[sub * 16 + j + 1] * p.w_scale, v[sub * 16 + j + 2] * p.w_scale,
                          v[sub * 16 + j + 3] * p.w_scale);
        __syncwarp();
        const int n = nb + c4 * 4;
        if (n < p.N) {
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + z * p.bias_zs + n));
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (opix[it] < 0) continue;
            float4 tt = *reinterpret_cast<const float4*>(stg + (it * 8 + r8) * P_STG_LD + c4 * 4);
            tt.x += bias4.x; tt.y += bias4.y; tt.z += bias4.z; tt.w += bias4.w;
            if (p.act == WXF_ACT_GELU_ERF) {
              tt.x = wxf_gelu_erf(tt.x); tt.y = wxf_gelu_erf(tt.y); tt.z = wxf_gelu_erf(tt.z); tt.w = wxf_gelu_erf(tt.w);
            }
            if (p.res) {
              tt.x += rres[sub][it].x; tt.y += rres[sub][it].y; tt.z += rres[sub][it].z; tt.w += rres[sub][it].w;
            }
            if (p.out) *reinterpret_cast<float4*>(p.out + opix[it] * p.ldc + n) = tt;
            if (p.out_hi) {
              __align__(8) __half2 h4[2];
              __align__(8) __half2 l4[2];
              wxf_split2_f16x2(tt.x, tt.y, h4[0], l4[0]);
              wxf_split2_f16x2(tt.z, tt.w, h4[1], l4[1]);
              *reinterpret_cast<uint2*>(p.out_hi + opix[it] * p.ldh + n) = *reinterpret_cast<const uint2*>(h4);
              *reinterpret_cast<uint2*>(p.out_lo + opix[it] * p.ldh + n) = *reinterpret_cast<const uint2*>(l4);
            }
          }
        }
        __syncwarp();
      }
    }
    }  // CONV epilogue (instantiated with 8 epilogue warps only)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- small-K GEMM (K <= 128: to_qkv / to_out / FF fc1 of stage 0) with the weight tile RESIDENT in shared memory ------
// One TMA tile load takes ~2500 cycles (profiles/r1_tma_microbench.txt) and a K = 128 tile is only 2 K-steps = 1536 cycles
// of MMA: with the 2-stage ring of the generic kernel every tile exposes the full load latency (ff1 at stage 0: 8000
// cycles per tile).  Here the grid is a multiple of the number of N tiles, so a CTA keeps ONE N tile for its whole life:
// W (hi|lo per K-step, 64 KB) is loaded once, and the ring streams only A (32 KB per K-step, 3 stages = 1.5 tiles ahead).
// Same MMA issue, TMEM double buffering and 16-warp TMA-store epilogue as tc_persistent_kernel<GEMM, 16, 2>.
constexpr int RW_STA = 3;                                  // A stages (hi + lo planes of one K-step: 32 KB)
constexpr int RW_EW = 16;
constexpr int RW_A_BYTES = 2 * TILE_BYTES;
constexpr int RW_W_BYTES = 2 * P_BN * BLOCK_K * 2;         // W_hi | W_lo of one K-step: 32 KB
constexpr int RW_SMEM = 2 * RW_W_BYTES + RW_STA * RW_A_BYTES + RW_EW * 4096 + 8 * (2 * RW_STA + 5) + 16 + 1024;

template <int EPI>
__global__ void __launch_bounds__(64 + 32 * RW_EW, 1)
tc_resident_w_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                     const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO_hi,
                     const __grid_constant__ CUtensorMap tmO_lo, const __grid_constant__ TcParams p, const int n_tiles,
                     const int m_tiles, const int total_tiles) {
  constexpr int EW = RW_EW, STAGES = RW_STA, BN = P_BN, W_BYTES = P_BN * BLOCK_K * 2;
  constexpr int CW = 128 / (EW / 4);
  constexpr int NCH = CW / 32;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  constexpr int OFF_A = 2 * RW_W_BYTES;                    // [W k-step 0 | W k-step 1 | A ring | staging | barriers]
  constexpr int OFF_STG = OFF_A + STAGES * RW_A_BYTES;
  constexpr int OFF_BARS = OFF_STG + EW * 4096;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  float* staging = reinterpret_cast<float*>(gen + OFF_STG);
  const uint32_t bar_base = base + OFF_BARS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + OFF_BARS + 8 * (2 * STAGES + 5));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.num_ksteps;  // 1 or 2

  wxf_pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EW);
    }
    mbar_init(w_bar, 1);
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
  wxf_pdl_wait();  // everything above overlaps the previous kernel's tail (programmatic dependent launch)

  // gridDim.x is a multiple of n_tiles: t % n_tiles == blockIdx.x % n_tiles for every tile of this CTA
  auto decode = [&](int t, int& n0, int& z, int64_t& m0, int& tb, int& oy0, int& ox0) {
    n0 = (t % n_tiles) * BN;
    m0 = (int64_t)(t / n_tiles) * BLOCK_M;
    z = 0; tb = 0; oy0 = 0; ox0 = 0;
  };
  (void)m_tiles;

  if (warp == 0) {
    if (lane == 0) {
      const int n0 = (int)(blockIdx.x % n_tiles) * BN;
      mbar_expect_tx(w_bar, (uint32_t)(num_k * RW_W_BYTES));
      for (int ks = 0; ks < num_k; ++ks) {
        tma_load_2d(&tmW_hi, w_bar, base + ks * RW_W_BYTES, ks * BLOCK_K, n0);
        tma_load_2d(&tmW_lo, w_bar, base + ks * RW_W_BYTES + W_BYTES, ks * BLOCK_K, n0);
      }
      uint32_t g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BLOCK_M;
        for (int ks = 0; ks < num_k; ++ks, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = base + OFF_A + s * RW_A_BYTES;
          mbar_expect_tx(full_bar(s), (uint32_t)RW_A_BYTES);
          tma_load_2d(&tmA_hi, full_bar(s), st, ks * BLOCK_K, m0);
          tma_load_2d(&tmA_lo, full_bar(s), st + TILE_BYTES, ks * BLOCK_K, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(w_bar, 0);
      uint32_t g = 0;
      int i = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
        const int slot = i & 1;
        mbar_wait(tempty_bar(slot), (((uint32_t)i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(slot * 2 * BN), d1 = d0 + BN;
        for (int ks = 0; ks < num_k; ++ks, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = base + OFF_A + s * RW_A_BYTES, sw = base + ks * RW_W_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t a_hi = umma_desc_sw128(sa + k * 32);
            const uint64_t a_lo = umma_desc_sw128(sa + TILE_BYTES + k * 32);
            const uint64_t w_hi = umma_desc_sw128(sw + k * 32);
            const uint32_t acc = (ks | k) ? 1u : 0u;
            tc_mma_f16(d0, a_hi, w_hi, IDESC2, acc);  // A_hi [W_hi | W_lo] -> both accumulators of the slot
            tc_mma_f16(d1, a_lo, w_hi, IDESC, 1u);    // + A_lo W_hi
          }
          tc_commit(empty_bar(s));
        }
        tc_commit(tfull_bar(slot));
      }
    }
  } else if constexpr (EPI != 0) {
    gemm_epilogue_fast<EW, EPI>(p, tmO, tmO_hi, tmO_lo, tmem_base, tfull_bar(0), tempty_bar(0),
                                reinterpret_cast<uint8_t*>(staging), base + OFF_STG, n_tiles, total_tiles, warp, lane);
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3, half = ew >> 2;
      // GEMM mode: lane = output row.  Accumulator row -> registers -> scale/bias/GELU/residual -> 32-column chunks
      // staged in swizzled shared memory -> TMA store (fp32 tile and/or fp16 hi/lo plane tiles).
      uint8_t* stg_b = reinterpret_cast<uint8_t*>(staging) + ew * 4096;
      const uint32_t stg_a = base + OFF_STG + ew * 4096;
      const int row = quarter * 32 + lane;
      int i = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
        const int slot = i & 1;
        int n0, z, tb, oy0, ox0;
        int64_t m0;
        decode(t, n0, z, m0, tb, oy0, ox0);
        const int nb0 = n0 + half * CW;
        const int64_t m = m0 + row;
        // residual prefetch (the row's 64 columns), issued before waiting for the accumulator
        float4 rres[CW / 4];
        if (p.res) {
#pragma unroll
          for (int q = 0; q < CW / 4; ++q) {
            rres[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < p.M && nb0 + 4 * q < p.N) rres[q] = *reinterpret_cast<const float4*>(p.res + m * p.ldr + nb0 + 4 * q);
          }
        }
        mbar_wait(tfull_bar(slot), ((uint32_t)i >> 1) & 1u);
        tc_fence_after();
        float v[CW];
        {
          const uint32_t tb_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 2 * BN + half * CW);
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            tmem_ld32(tb_addr + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r[j]);
            tmem_ld32(tb_addr + (uint32_t)(BN + c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(slot));  // TMEM slot free for the MMA of tile i+2

#pragma unroll
        for (int cb = 0; cb < NCH; ++cb) {
          const int nb = nb0 + cb * 32;
          if (nb >= p.N) break;  // warp-uniform
          {
            const float2 sc = make_float2(p.w_scale, p.w_scale);
            float2* v2 = reinterpret_cast<float2*>(v + cb * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {  // packed fp32 math (FFMA2): scale + bias, GELU, residual
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias && nb + 4 * q < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + 4 * q));
              float2 a0 = __ffma2_rn(v2[2 * q], sc, make_float2(b4.x, b4.y));
              float2 a1 = __ffma2_rn(v2[2 * q + 1], sc, make_float2(b4.z, b4.w));
              if (p.act == WXF_ACT_GELU_ERF) {
                a0 = wxf_gelu_erf2(a0);
                a1 = wxf_gelu_erf2(a1);
              }
              if (p.res) {
                const float4 rr = rres[cb * 8 + q];
                a0 = __fadd2_rn(a0, make_float2(rr.x, rr.y));
                a1 = __fadd2_rn(a1, make_float2(rr.z, rr.w));
              }
              v2[2 * q] = a0;
              v2[2 * q + 1] = a1;
            }
          }
          if (p.out) {
            if (lane == 0) bulk_wait_read0();  // previous TMA store has finished reading the staging buffer
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
          if (p.out_hi) {
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

template <int BN, int STAGES>
constexpr int smem_bytes() {
  return STAGES * (2 * TILE_BYTES + 2 * BN * BLOCK_K * 2) + STG_BYTES + 8 * (2 * STAGES + 1) + 16 + 1024;
}

template <int BN, int STAGES, int MODE>
int launch(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tw_hi, const CUtensorMap& tw_lo,
           const TcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int SMEM = smem_bytes<BN, STAGES>();
  static WxfPerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.get();  // function attributes are per device
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(tc_contract_kernel<BN, STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) WXF_FAIL((int)e, "tc: cannot opt in to %d bytes of shared memory: %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  wxf_launch(tc_contract_kernel<BN, STAGES, MODE>, grid, dim3(NUM_THREADS), SMEM, st, ta_hi, ta_lo, tw_hi, tw_lo, p);
  WXF_CHECK_LAUNCH("tc_contract");
  return 0;
}

int num_sms() {
  static WxfPerDevice<int> n_pd;
  int& n = n_pd.get();
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

bool resident_w_enabled() {  // WXF_GEMM_RESIDENT_W=1: K <= 128 GEMMs keep their weight tile in shared memory (round-2 candidate)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_GEMM_RESIDENT_W");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

bool fast_epilogue_enabled() {  // WXF_GEMM_EPI=0: every GEMM takes the generic run-time-switched epilogue (A/B, debugging)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_GEMM_EPI");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

uint32_t park_ns() {  // WXF_MBAR_PARK_NS: suspend-time hint (ns) of the far-away mbarrier waits; 0 = plain try_wait spin
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_MBAR_PARK_NS");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  return (uint32_t)v;
}

int ew16_max_k() {  // WXF_GEMM_EW16_MAXK: largest K that takes the <16 epilogue warps, 2 stages> shape (default 128; K = 256 measured
                    // 4 % faster on <8, 3> once the specialised epilogues had cut the issue load)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_GEMM_EW16_MAXK");
    v = e ? atoi(e) : 128;
  }
  return v;
}

int toep_persistent_max_k() {  // WXF_TOEP_PERSISTENT_MAXK: largest K of a Toeplitz branch on the persistent kernel (0 = never)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_TOEP_PERSISTENT_MAXK");
    v = e ? atoi(e) : 2048;
  }
  return v;
}

bool persistent_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXF_TC_PERSISTENT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int MODE, int EW, int STAGES, int EPI = 0>
int launch_persistent(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tw_hi, const CUtensorMap& tw_lo,
                      const CUtensorMap& to, const CUtensorMap& to_hi, const CUtensorMap& to_lo, const TcParams& p,
                      int n_tiles, int m_tiles, int phases, cudaStream_t st) {
  static WxfPerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.get();  // function attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_persistent_kernel<MODE, EW, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         p_smem<EW, STAGES>());
    if (e != cudaSuccess)
      WXF_FAIL((int)e, "tc: cannot opt in to %d bytes of shared memory: %s", p_smem<EW, STAGES>(), cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t total = (int64_t)n_tiles * m_tiles * phases;
  if (total > INT32_MAX) WXF_FAIL(WXF_EINVAL, "tc: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  wxf_launch(tc_persistent_kernel<MODE, EW, STAGES, EPI>, dim3(grid), dim3(64 + 32 * EW), p_smem<EW, STAGES>(), st, ta_hi, ta_lo,
             tw_hi, tw_lo, to, to_hi, to_lo, p, n_tiles, m_tiles, (int)total);
  WXF_CHECK_LAUNCH("tc_persistent");
  return 0;
}

int check_epilogue(const char* who, int N, const float* bias, const float* res, const float* out, const void* out_hi,
                   const void* out_lo, int ldc, int c_off, int ldr, int r_off, int ldh, int h_off) {
  if (!out && !out_hi) WXF_FAIL(WXF_EINVAL, "%s: no output", who);
  if ((out_hi == nullptr) != (out_lo == nullptr)) WXF_FAIL(WXF_EINVAL, "%s: out_hi/out_lo must come together", who);
  if (N & 3) WXF_FAIL(WXF_EUNSUPPORTED, "%s: N=%d must be a multiple of 4", who, N);
  if (out && (ldc < c_off + N || (ldc & 3) || (c_off & 3) || !wxf_aligned16(out)))
    WXF_FAIL(WXF_EALIGN, "%s: out stride/alignment", who);
  if (res && (ldr < r_off + N || (ldr & 3) || (r_off & 3) || !wxf_aligned16(res)))
    WXF_FAIL(WXF_EALIGN, "%s: res stride/alignment", who);
  if (out_hi && (ldh < h_off + N || (ldh & 3) || (h_off & 3) || (reinterpret_cast<uintptr_t>(out_hi) & 7) ||
                 (reinterpret_cast<uintptr_t>(out_lo) & 7)))
    WXF_FAIL(WXF_EALIGN, "%s: plane stride/alignment", who);
  if (bias && !wxf_aligned16(bias)) WXF_FAIL(WXF_EALIGN, "%s: bias alignment", who);
  return 0;
}

}  // namespace

extern "C" int wxf_gemm_f16x2_tc(const WxfGemmDesc* d, void* stream) {
  if (!d || !d->a_hi || !d->a_lo || !d->w_hi || !d->w_lo) WXF_FAIL(WXF_EINVAL, "gemm_tc: null operand");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0 || d->lda < d->K) WXF_FAIL(WXF_EINVAL, "gemm_tc: bad dims");
  if ((d->lda & 7) || (d->K & 7)) WXF_FAIL(WXF_EALIGN, "gemm_tc: K and lda must be multiples of 8 (16-byte TMA rows)");
  if (!wxf_aligned16(d->a_hi) || !wxf_aligned16(d->a_lo) || !wxf_aligned16(d->w_hi) || !wxf_aligned16(d->w_lo))
    WXF_FAIL(WXF_EALIGN, "gemm_tc: operand planes must be 16-byte aligned");
  int rc = check_epilogue("gemm_tc", d->N, d->bias, d->res, d->out, d->out_hi, d->out_lo, d->ldc, d->c_off, d->ldr,
                          d->r_off, d->ldh, 0);
  if (rc) return rc;
  if (d->M > (int64_t)65535 * BLOCK_M) WXF_FAIL(WXF_EINVAL, "gemm_tc: M too large for one launch");
  if (d->out_hi && ((d->ldh & 7) || !wxf_aligned16(d->out_hi) || !wxf_aligned16(d->out_lo)))
    WXF_FAIL(WXF_EALIGN, "gemm_tc: output planes need ldh %% 8 == 0 and 16-byte alignment (TMA store)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool persistent = persistent_enabled();
  const int BN = (!persistent && d->N > 128) ? 256 : 128;
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  if ((rc = make_map_2d(&ta_hi, d->a_hi, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, BLOCK_M))) return rc;
  if ((rc = make_map_2d(&ta_lo, d->a_lo, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, BLOCK_M))) return rc;
  // a single ragged N tile loads only the weight rows that exist (rounded up to the MMA's N granularity of 16)
  const bool plain_persistent = persistent && !resident_w_enabled();
  const int w_rows = (plain_persistent && d->N < BN) ? ((d->N + 15) & ~15) : BN;
  if ((rc = make_map_2d(&tw_hi, d->w_hi, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->K, (uint32_t)w_rows))) return rc;
  if ((rc = make_map_2d(&tw_lo, d->w_lo, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->K, (uint32_t)w_rows))) return rc;
  TcParams p{};
  p.w_bytes = (uint32_t)(w_rows * BLOCK_K * 2);
  p.bias = d->bias;
  p.res = d->res ? d->res + d->r_off : nullptr;
  p.out = d->out ? d->out + d->c_off : nullptr;
  p.out_hi = reinterpret_cast<__half*>(d->out_hi);
  p.out_lo = reinterpret_cast<__half*>(d->out_lo);
  p.M = d->M; p.N = d->N; p.num_ksteps = (d->K + BLOCK_K - 1) / BLOCK_K;
  p.ldc = d->ldc; p.ldr = d->ldr; p.ldh = d->ldh; p.act = d->act;
  p.w_scale = ldexpf(1.0f, -d->w_scale_log2);
  if (persistent) {
    // output tiles leave through TMA stores: 32-row x 32-column boxes of the fp32 matrix and of the fp16 planes
    CUtensorMap to = ta_hi, to_hi = ta_hi, to_lo = ta_hi;  // placeholders when an output is absent (never dereferenced)
    const uint32_t box[2] = {32, 32}, es[2] = {1, 1};
    if (d->out) {
      const uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M}, strides[1] = {(uint64_t)d->ldc * 4};
      if ((rc = make_map(&to, d->out + d->c_off, 2, dims, strides, box, es, 128, true))) return rc;
    }
    if (d->out_hi) {
      const uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M}, strides[1] = {(uint64_t)d->ldh * 2};
      if ((rc = make_map(&to_hi, d->out_hi, 2, dims, strides, box, es, 64))) return rc;
      if ((rc = make_map(&to_lo, d->out_lo, 2, dims, strides, box, es, 64))) return rc;
    }
    const int nt = (d->N + BN - 1) / BN, mt = (int)((d->M + BLOCK_M - 1) / BLOCK_M);
    // specialised epilogues for the combinations the forecast plans use; anything else takes the generic one
    int epi = 0;
    p.park_ns = park_ns();
    if (fast_epilogue_enabled()) {
      const bool gelu = d->act == WXF_ACT_GELU_ERF, planes = d->out_hi != nullptr, f32 = d->out != nullptr;
      const bool inplace = d->res && f32 && d->res + d->r_off == d->out + d->c_off && d->ldr == d->ldc;
      if ((d->act == WXF_ACT_NONE || gelu) && d->N % 32 == 0) {
        if (!d->res && planes && !f32) epi = EPI_PLANES | (gelu ? EPI_GELU : 0);
        else if (!d->res && f32 && !planes && !gelu) epi = EPI_F32;
        else if (inplace && !planes && !gelu) epi = EPI_F32 | EPI_RED;
      }
    }
    if (d->K <= 128 && nt <= num_sms() && resident_w_enabled()) {
      const int64_t total = (int64_t)nt * mt;
      int64_t grid = (num_sms() / nt) * nt;   // a multiple of the N-tile count: every CTA keeps one N tile
      if (grid > total) grid = total;         // total is a multiple of nt as well
#define WXF_RW_LAUNCH(EPI_)                                                                                              \
  {                                                                                                                      \
    static WxfPerDevice<bool> attr_set_pd;                                                                               \
    bool& attr_set = attr_set_pd.get(); /* function attributes are per device */                                         \
    if (!attr_set) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(tc_resident_w_kernel<EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, RW_SMEM); \
      if (e != cudaSuccess) WXF_FAIL((int)e, "tc: cannot opt in to %d bytes of shared memory: %s", RW_SMEM, cudaGetErrorString(e)); \
      attr_set = true;                                                                                                   \
    }                                                                                                                    \
    wxf_launch(tc_resident_w_kernel<EPI_>, dim3((unsigned)grid), dim3(64 + 32 * RW_EW), RW_SMEM, st, ta_hi, ta_lo, tw_hi, \
               tw_lo, to, to_hi, to_lo, p, nt, mt, (int)total);                                                          \
    WXF_CHECK_LAUNCH("tc_resident_w");                                                                                   \
    return 0;                                                                                                            \
  }
      switch (epi) {
        case EPI_PLANES: WXF_RW_LAUNCH(EPI_PLANES)
        case EPI_PLANES | EPI_GELU: WXF_RW_LAUNCH(EPI_PLANES | EPI_GELU)
        case EPI_F32: WXF_RW_LAUNCH(EPI_F32)
        case EPI_F32 | EPI_RED: WXF_RW_LAUNCH(EPI_F32 | EPI_RED)
        default: WXF_RW_LAUNCH(0)
      }
#undef WXF_RW_LAUNCH
    }
#define WXF_P_LAUNCH(EW_, ST_, EPI_) \
  return launch_persistent<MODE_GEMM, EW_, ST_, EPI_>(ta_hi, ta_lo, tw_hi, tw_lo, to, to_hi, to_lo, p, nt, mt, 1, st)
    if (d->K <= ew16_max_k()) {
      switch (epi) {
        case EPI_PLANES: WXF_P_LAUNCH(16, 2, EPI_PLANES);
        case EPI_PLANES | EPI_GELU: WXF_P_LAUNCH(16, 2, EPI_PLANES | EPI_GELU);
        case EPI_F32: WXF_P_LAUNCH(16, 2, EPI_F32);
        case EPI_F32 | EPI_RED: WXF_P_LAUNCH(16, 2, EPI_F32 | EPI_RED);
        default: WXF_P_LAUNCH(16, 2, 0);
      }
    }
    switch (epi) {
      case EPI_PLANES: WXF_P_LAUNCH(8, 3, EPI_PLANES);
      case EPI_PLANES | EPI_GELU: WXF_P_LAUNCH(8, 3, EPI_PLANES | EPI_GELU);
      case EPI_F32: WXF_P_LAUNCH(8, 3, EPI_F32);
      case EPI_F32 | EPI_RED: WXF_P_LAUNCH(8, 3, EPI_F32 | EPI_RED);
      default: WXF_P_LAUNCH(8, 3, 0);
    }
#undef WXF_P_LAUNCH
  }
  dim3 grid((unsigned)((d->N + BN - 1) / BN), (unsigned)((d->M + BLOCK_M - 1) / BLOCK_M), 1);
  if (BN == 256) return launch<256, 2, MODE_GEMM>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  return launch<128, 3, MODE_GEMM>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
}

extern "C" int wxf_conv_f16x2_tc(const WxfConvTcDesc* d, void* stream) {
  if (!d || !d->in_hi || !d->in_lo || !d->w_hi || !d->w_lo || !d->taps) WXF_FAIL(WXF_EINVAL, "conv_tc: null operand");
  if (d->B <= 0 || d->Hi <= 0 || d->Wi <= 0 || d->Cin <= 0 || d->N <= 0 || d->T <= 0 || d->Ho <= 0 || d->Wo <= 0 ||
      d->lda < d->Cin)
    WXF_FAIL(WXF_EINVAL, "conv_tc: bad dims");
  if (d->T > MAX_TAPS) WXF_FAIL(WXF_EUNSUPPORTED, "conv_tc: %d taps > %d", d->T, MAX_TAPS);
  if (d->stride != 1 && d->stride != 2 && d->stride != 4) WXF_FAIL(WXF_EUNSUPPORTED, "conv_tc: stride must be 1, 2 or 4");
  if (d->phases != 1 && d->phases != 4) WXF_FAIL(WXF_EINVAL, "conv_tc: phases must be 1 or 4");
  if ((d->phases == 4) != (d->out_scale == 2) || (d->phases == 1 && d->out_scale != 1))
    WXF_FAIL(WXF_EINVAL, "conv_tc: phases/out_scale mismatch");
  if ((d->lda & 7) || (d->cin_pad & 63) || d->cin_pad < d->Cin) WXF_FAIL(WXF_EALIGN, "conv_tc: lda %% 8, cin_pad %% 64");
  if (d->bias_phase_stride & 3) WXF_FAIL(WXF_EALIGN, "conv_tc: bias_phase_stride %% 4");
  if (!wxf_aligned16(d->in_hi) || !wxf_aligned16(d->in_lo) || !wxf_aligned16(d->w_hi) || !wxf_aligned16(d->w_lo))
    WXF_FAIL(WXF_EALIGN, "conv_tc: operand planes must be 16-byte aligned");
  int rc = check_epilogue("conv_tc", d->N, d->bias, d->res, d->out, d->out_hi, d->out_lo, d->ldc, d->c_off, d->ldr,
                          d->r_off, d->ldh, d->h_off);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;

  // M tile = bh x bw output pixels (<= 128) chosen to minimise the number of tiles
  const int s = d->stride;
  int best_bw = 1, best_bh = 1;
  int64_t best_tiles = INT64_MAX;
  for (int bw = (d->Wo < 128 ? d->Wo : 128); bw >= 1; --bw) {
    int bh = 128 / bw;
    if (bh > d->Ho) bh = d->Ho;
    if (bh * s > 256 || bw * s > 256) continue;
    const int64_t tiles = (int64_t)((d->Wo + bw - 1) / bw) * ((d->Ho + bh - 1) / bh);
    if (tiles < best_tiles) {
      best_tiles = tiles;
      best_bw = bw;
      best_bh = bh;
    }
  }
  const int bw = best_bw, bh = best_bh;
  const bool persistent = persistent_enabled();
  const int BN = (!persistent && d->N > 128) ? 256 : 128;
  const int K = d->T * d->cin_pad;

  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  {
    const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Wi, (uint64_t)d->Hi, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->lda * 2, (uint64_t)d->Wi * d->lda * 2,
                                 (uint64_t)d->Hi * d->Wi * d->lda * 2};
    // with an element stride s TMA loads ceil(boxDim/s) elements: boxDim = n*s picks n (driver API documentation)
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)(bw * s), (uint32_t)(bh * s), 1};
    const uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
    if ((rc = make_map(&ta_hi, d->in_hi, 4, dims, strides, box, es))) return rc;
    if ((rc = make_map(&ta_lo, d->in_lo, 4, dims, strides, box, es))) return rc;
  }
  // a single ragged N tile loads only the weight rows that exist (up_block4: 64 of 128), rounded up to the MMA's N step
  const int w_rows = (persistent && d->N < BN) ? ((d->N + 15) & ~15) : BN;
  if ((rc = make_map_2d(&tw_hi, d->w_hi, (uint64_t)d->phases * d->N, (uint64_t)K, (uint64_t)K, (uint32_t)w_rows))) return rc;
  if ((rc = make_map_2d(&tw_lo, d->w_lo, (uint64_t)d->phases * d->N, (uint64_t)K, (uint64_t)K, (uint32_t)w_rows))) return rc;

  TcParams p{};
  p.w_bytes = (uint32_t)(w_rows * BLOCK_K * 2);
  p.bias = d->bias;
  p.res = d->res ? d->res + d->r_off : nullptr;
  p.out = d->out ? d->out + d->c_off : nullptr;
  p.out_hi = d->out_hi ? reinterpret_cast<__half*>(d->out_hi) + d->h_off : nullptr;
  p.out_lo = d->out_lo ? reinterpret_cast<__half*>(d->out_lo) + d->h_off : nullptr;
  p.M = 0; p.N = d->N;
  p.cblocks = d->cin_pad / BLOCK_K;
  p.cin_pad = d->cin_pad;
  p.num_ksteps = d->T * p.cblocks;
  p.ldc = d->ldc; p.ldr = d->ldr; p.ldh = d->ldh; p.act = d->act;
  p.w_scale = ldexpf(1.0f, -d->w_scale_log2);
  p.bw = bw; p.bh = bh;
  p.tiles_x = (d->Wo + bw - 1) / bw;
  p.tiles_y = (d->Ho + bh - 1) / bh;
  p.Ho = d->Ho; p.Wo = d->Wo; p.stride = s; p.out_scale = d->out_scale;
  p.bias_zs = d->bias_phase_stride;
  p.a_bytes = (uint32_t)(bw * bh * BLOCK_K * 2);
  for (int z = 0; z < d->phases; ++z)
    for (int t = 0; t < d->T; ++t) {
      p.taps[z][t][0] = (int16_t)d->taps[(z * d->T + t) * 2];
      p.taps[z][t][1] = (int16_t)d->taps[(z * d->T + t) * 2 + 1];
    }
  const int64_t ntiles = (int64_t)d->B * p.tiles_x * p.tiles_y;
  if (persistent)
    return launch_persistent<MODE_CONV, 8, 3>(ta_hi, ta_lo, tw_hi, tw_lo, ta_hi, ta_hi, ta_hi, p, (d->N + BN - 1) / BN,
                                              (int)ntiles, d->phases, st);
  if (ntiles > 65535) WXF_FAIL(WXF_EINVAL, "conv_tc: too many tiles for one launch");
  dim3 grid((unsigned)((d->N + BN - 1) / BN), (unsigned)ntiles, (unsigned)d->phases);
  if (BN == 256) return launch<256, 2, MODE_CONV>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  return launch<128, 3, MODE_CONV>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
}

extern "C" int wxf_cross_embed_toeplitz_tc(const WxfToeplitzDesc* d, void* stream) {
  if (!d || !d->in_hi || !d->in_lo || !d->w_hi || !d->w_lo || !d->out) WXF_FAIL(WXF_EINVAL, "toeplitz: null pointer");
  if (d->B <= 0 || d->Hi <= 0 || d->Wi <= 0 || d->Cin <= 0 || d->ch <= 0 || d->kernel <= 0 || d->Ho <= 0 || d->Wo <= 0 ||
      d->lda < d->Cin || d->pad < 0)
    WXF_FAIL(WXF_EINVAL, "toeplitz: bad dims");
  if (d->kernel & 1) WXF_FAIL(WXF_EUNSUPPORTED, "toeplitz: kernel size must be even (stride 2)");
  const int J = d->kernel / 2, T = 2 * d->kernel, N = J * d->ch;
  if (T > MAX_TAPS) WXF_FAIL(WXF_EUNSUPPORTED, "toeplitz: kernel %d too large", d->kernel);
  if (N > 256 || (d->ch & 3)) WXF_FAIL(WXF_EUNSUPPORTED, "toeplitz: (k/2)*channels = %d must be <= 256, channels %% 4 == 0", N);
  if (d->cin_pad != 64 || d->Cin > 64) WXF_FAIL(WXF_EUNSUPPORTED, "toeplitz: at most 64 input channels (cin_pad = 64)");
  if ((d->lda & 7)) WXF_FAIL(WXF_EALIGN, "toeplitz: lda %% 8");
  if (!wxf_aligned16(d->in_hi) || !wxf_aligned16(d->in_lo) || !wxf_aligned16(d->w_hi) || !wxf_aligned16(d->w_lo))
    WXF_FAIL(WXF_EALIGN, "toeplitz: operand planes must be 16-byte aligned");
  if (d->ldc < d->c_off + d->ch || (d->ldc & 3) || (d->c_off & 3) || !wxf_aligned16(d->out) ||
      (d->bias && !wxf_aligned16(d->bias)))
    WXF_FAIL(WXF_EALIGN, "toeplitz: output stride/alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int BN = N > 128 ? 256 : 128;
  const int K = T * 64;
  int rc;
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  {
    const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Wi, (uint64_t)d->Hi, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->lda * 2, (uint64_t)d->Wi * d->lda * 2,
                                 (uint64_t)d->Hi * d->Wi * d->lda * 2};
    const uint32_t box[4] = {(uint32_t)BLOCK_K, 256, 1, 1};  // 128 pixels at element stride 2
    const uint32_t es[4] = {1, 2, 1, 1};
    if ((rc = make_map(&ta_hi, d->in_hi, 4, dims, strides, box, es))) return rc;
    if ((rc = make_map(&ta_lo, d->in_lo, 4, dims, strides, box, es))) return rc;
  }
  if ((rc = make_map_2d(&tw_hi, d->w_hi, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN))) return rc;
  if ((rc = make_map_2d(&tw_lo, d->w_lo, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN))) return rc;
  TcParams p{};
  p.bias = d->bias;
  p.out = d->out + d->c_off;
  p.N = N;
  p.cblocks = 1;
  p.cin_pad = 64;
  p.num_ksteps = T;
  p.ldc = d->ldc;
  p.w_scale = ldexpf(1.0f, -d->w_scale_log2);
  p.bw = 128; p.bh = 1;
  p.step = 128 - (J - 1);
  p.J = J; p.ch = d->ch;
  p.tiles_x = (d->Wo + p.step - 1) / p.step;
  p.tiles_y = d->Ho;
  p.Ho = d->Ho; p.Wo = d->Wo; p.stride = 2; p.out_scale = 1;
  p.a_bytes = (uint32_t)TILE_BYTES;
  for (int ky = 0; ky < d->kernel; ++ky)
    for (int r = 0; r < 2; ++r) {
      p.taps[0][ky * 2 + r][0] = (int16_t)(ky - d->pad + 2 * d->oy_off);
      p.taps[0][ky * 2 + r][1] = (int16_t)(r - d->pad);
    }
  const int64_t ntiles = (int64_t)d->B * p.tiles_x * p.tiles_y;
  if (ntiles > 65535) WXF_FAIL(WXF_EINVAL, "toeplitz: too many tiles for one launch");
  dim3 grid(1, (unsigned)ntiles, 1);
  if (BN == 256) return launch<256, 2, MODE_TOEP>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  // the N = 128 branches (k = 4, 8, 16: 8 / 16 / 32 K-steps per tile) spend much of a one-tile CTA's life outside the main
  // loop (launch, TMEM allocation, pipeline fill, the diagonal-sum epilogue): they take the persistent kernel (tile loop,
  // TMEM double buffering, epilogue of tile i under the MMAs of tile i+1).  One main accumulator there: at K = 2048 (k = 16)
  // the branch's error vs fp64 goes 6.7e-7 -> 1.9e-6 (accumulator truncation), the full-grid forward stays at 4.15e-6.
  if (N == 128 && K <= toep_persistent_max_k() && persistent_enabled() && (d->ch == 64 || d->ch == 32 || d->ch == 16)) {
    p.w_bytes = (uint32_t)(BN * BLOCK_K * 2);
    return launch_persistent<MODE_TOEP, 8, 3>(ta_hi, ta_lo, tw_hi, tw_lo, ta_hi, ta_hi, ta_hi, p, 1, (int)ntiles, 1, st);
  }
  return launch<128, 3, MODE_TOEP>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
}
