// Pointwise (1x1-conv) GEMM on the 5th-generation tensor cores: C[M,N] = epi(A[M,K] * W[N,K]^T).
//
// Precision scheme "f16x2": every fp32 operand is carried as two fp16 planes (hi = fp16(x), lo = fp16(x - hi),
// 22 significant bits) and the product is three tcgen05.mma passes into one fp32 TMEM accumulator:
//     A W^T ~= A_hi W_lo^T + A_lo W_hi^T + A_hi W_hi^T          (the lo*lo term is below fp32 resolution)
// Measured end to end this stays within 4e-6 rel-max of the fp32 reference forward (budget 1e-4), where a
// single tf32/fp16 pass is at 1e-3.  Operand planes cost the same HBM bytes as fp32.
//
// Kernel shape: one CTA per 128 x BN output tile (BN = 128 or 256), 192 threads:
//   warp 0    : TMA producer  (cp.async.bulk.tensor 2D, 128B swizzle, 3/2-stage mbarrier ring)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction)
//   warps 2-5 : epilogue, tcgen05.ld 32 lanes x 32 columns -> registers -> scale/bias/GELU/residual -> global
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>
#include <unordered_map>

#include "wxf_common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // 64 fp16 = one 128-byte swizzle row
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB: one 128-row operand plane of one stage
constexpr int NUM_THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 22;          // mbarrier waits trap instead of hanging the GPU

struct GemmTcParams {
  const float* bias;
  const float* res;
  float* out;
  __half* out_hi;
  __half* out_lo;
  int64_t M;
  int N, K;
  int ldc, ldr, ldh;
  int act;
  float w_scale;  // 2^-k: undoes the power-of-two pre-scale of the weight planes
};

// ---- PTX wrappers ------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) {
      printf("wxf_gemm_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 128 B, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}


// ---- kernel --------------------------------------------------------------------------------------

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_f16x2_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                     const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                     const GemmTcParams p) {
  constexpr int W_BYTES = BN * BLOCK_K * 2;
  constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * W_BYTES;
  // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle-128B atoms need 1024-byte alignment
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int64_t m0 = (int64_t)blockIdx.y * BLOCK_M;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: BN fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + s * STAGE_BYTES;
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        tma_load_2d(&tmA_hi, full_bar(s), st, kb * BLOCK_K, (int)m0);
        tma_load_2d(&tmA_lo, full_bar(s), st + TILE_BYTES, kb * BLOCK_K, (int)m0);
        tma_load_2d(&tmW_hi, full_bar(s), st + 2 * TILE_BYTES, kb * BLOCK_K, n0);
        tma_load_2d(&tmW_lo, full_bar(s), st + 2 * TILE_BYTES + W_BYTES, kb * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t a_hi = umma_desc_sw128(st + k * 32);
          const uint64_t a_lo = umma_desc_sw128(st + TILE_BYTES + k * 32);
          const uint64_t w_hi = umma_desc_sw128(st + 2 * TILE_BYTES + k * 32);
          const uint64_t w_lo = umma_desc_sw128(st + 2 * TILE_BYTES + W_BYTES + k * 32);
          tc_mma_f16(tmem_base, a_hi, w_lo, IDESC, (kb | k) ? 1u : 0u);
          tc_mma_f16(tmem_base, a_lo, w_hi, IDESC, 1u);
          tc_mma_f16(tmem_base, a_hi, w_hi, IDESC, 1u);
        }
        tc_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
      }
      tc_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int64_t m = m0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool vec4 = ((p.ldc & 3) == 0) && (!p.res || (p.ldr & 3) == 0) && ((p.N & 3) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
      if (m < p.M) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + c * 32 + j;
          float t = __uint_as_float(r[j]) * p.w_scale;
          if (n < p.N) {
            if (p.bias) t += __ldg(p.bias + n);
            if (p.act == WXF_ACT_GELU_ERF) t = wxf_gelu_erf(t);
          }
          v[j] = t;
        }
        const int nb = n0 + c * 32;
        if (p.res) {
          const float* rr = p.res + m * p.ldr + nb;
          if (vec4 && nb + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = *reinterpret_cast<const float4*>(rr + j);
              v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) v[j] += rr[j];
          }
        }
        if (p.out) {
          float* oo = p.out + m * p.ldc + nb;
          if (vec4 && nb + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(oo + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) oo[j] = v[j];
          }
        }
        if (p.out_hi) {
          __half* hh = p.out_hi + m * p.ldh + nb;
          __half* ll = p.out_lo + m * p.ldh + nb;
          if ((p.ldh & 7) == 0 && nb + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __align__(16) __half h8[8];
              __align__(16) __half l8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) wxf_split_f16x2(v[j + e], h8[e], l8[e]);
              *reinterpret_cast<uint4*>(hh + j) = *reinterpret_cast<const uint4*>(h8);
              *reinterpret_cast<uint4*>(ll + j) = *reinterpret_cast<const uint4*>(l8);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) wxf_split_f16x2(v[j], hh[j], ll[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

// ---- host side: TMA descriptors -----------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<uint64_t>()(k.rows * 1315423911ull + k.cols * 2654435761ull + k.ld * 97ull + k.box_rows);
    return h;
  }
};

// fp16 [rows, cols] row-major with row stride ld elements; box = box_rows x 64 columns, 128B swizzle, zero OOB fill
int make_map(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) WXF_FAIL(WXF_EUNSUPPORTED, "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) WXF_FAIL(WXF_EINVAL, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

template <int BN, int STAGES>
int launch(const WxfGemmDesc* d, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * BN * BLOCK_K * 2;
  constexpr int SMEM = STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16x2_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) WXF_FAIL((int)e, "gemm_tc: cannot opt in to %d bytes of shared memory: %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  int rc;
  if ((rc = make_map(&ta_hi, d->a_hi, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, BLOCK_M))) return rc;
  if ((rc = make_map(&ta_lo, d->a_lo, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, BLOCK_M))) return rc;
  if ((rc = make_map(&tw_hi, d->w_hi, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->K, BN))) return rc;
  if ((rc = make_map(&tw_lo, d->w_lo, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->K, BN))) return rc;
  GemmTcParams p;
  p.bias = d->bias;
  p.res = d->res ? d->res + d->r_off : nullptr;
  p.out = d->out ? d->out + d->c_off : nullptr;
  p.out_hi = reinterpret_cast<__half*>(d->out_hi);
  p.out_lo = reinterpret_cast<__half*>(d->out_lo);
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.ldc = d->ldc; p.ldr = d->ldr; p.ldh = d->ldh; p.act = d->act;
  p.w_scale = ldexpf(1.0f, -d->w_scale_log2);
  dim3 grid((unsigned)((d->N + BN - 1) / BN), (unsigned)((d->M + BLOCK_M - 1) / BLOCK_M));
  gemm_f16x2_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, SMEM, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
  WXF_CHECK_LAUNCH("gemm_f16x2_tc");
  return 0;
}

}  // namespace

extern "C" int wxf_gemm_f16x2_tc(const WxfGemmDesc* d, void* stream) {
  if (!d || !d->a_hi || !d->a_lo || !d->w_hi || !d->w_lo) WXF_FAIL(WXF_EINVAL, "gemm_tc: null operand");
  if (!d->out && !d->out_hi) WXF_FAIL(WXF_EINVAL, "gemm_tc: no output");
  if ((d->out_hi == nullptr) != (d->out_lo == nullptr)) WXF_FAIL(WXF_EINVAL, "gemm_tc: out_hi/out_lo must come together");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0 || d->lda < d->K) WXF_FAIL(WXF_EINVAL, "gemm_tc: bad dims");
  if ((d->lda & 7) || (d->K & 7)) WXF_FAIL(WXF_EALIGN, "gemm_tc: K and lda must be multiples of 8 (16-byte TMA rows)");
  if (!wxf_aligned16(d->a_hi) || !wxf_aligned16(d->a_lo) || !wxf_aligned16(d->w_hi) || !wxf_aligned16(d->w_lo))
    WXF_FAIL(WXF_EALIGN, "gemm_tc: operand planes must be 16-byte aligned");
  if (d->out && (d->ldc < d->c_off + d->N)) WXF_FAIL(WXF_EINVAL, "gemm_tc: ldc");
  if (d->res && (d->ldr < d->r_off + d->N)) WXF_FAIL(WXF_EINVAL, "gemm_tc: ldr");
  if (d->out_hi && d->ldh < d->N) WXF_FAIL(WXF_EINVAL, "gemm_tc: ldh");
  if (d->out_hi && (!wxf_aligned16(d->out_hi) || !wxf_aligned16(d->out_lo))) WXF_FAIL(WXF_EALIGN, "gemm_tc: out planes alignment");
  if (d->out && ((d->c_off & 3) || !wxf_aligned16(d->out))) WXF_FAIL(WXF_EALIGN, "gemm_tc: out alignment");
  if (d->res && ((d->r_off & 3) || !wxf_aligned16(d->res))) WXF_FAIL(WXF_EALIGN, "gemm_tc: res alignment");
  if (d->M > (int64_t)65535 * BLOCK_M) WXF_FAIL(WXF_EINVAL, "gemm_tc: M too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->N > 128) return launch<256, 2>(d, st);
  return launch<128, 3>(d, st);
}
