// Exact-fp32 implicit-GEMM convolution on CUDA cores (FFMA), the arithmetic reference of the library.
//
// It serves (a) every contraction whose shape does not suit the tensor-core path (few output channels,
// odd channel counts) and (b) as the on-device ground truth the tcgen05 kernels are tested against.
// Tiles BM x BN x 16, 256 threads, TM x TN register micro-tiles, register-prefetch double buffering.
#include "wxf_common.cuh"

namespace {

struct ConvP {
  const float* in;
  const float* w;
  const int32_t* taps;
  const float* bias;
  const float* res;
  float* out;
  int B, Hi, Wi, lda, Cin;
  int N, T, stride;
  int Ho, Wo;
  int out_scale;
  int ldc, c_off, ldr, r_off, act, bias_zs;
  int K;
  int64_t M;
};

constexpr int BK = 16;
constexpr int NT = 256;

template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(NT) conv_igemm_kernel(const ConvP p) {
  static_assert((BM / TM) * (BN / TN) == NT, "thread tiling");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile");
  constexpr int TXN = BN / TN;         // threads along N
  constexpr int RG = TM / 4;           // row groups of 4 per thread, spaced BM/RG apart
  constexpr int CG = TN / 4;           // col groups of 4 per thread, spaced BN/CG apart
  constexpr int A_ITEMS = VEC ? (BM * BK / 4 + NT - 1) / NT : (BM * BK + NT - 1) / NT;
  constexpr int B_ITEMS = VEC ? (BN * BK / 4 + NT - 1) / NT : (BN * BK + NT - 1) / NT;

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int z = blockIdx.z;
  const float* __restrict__ wz = p.w + (size_t)z * p.N * p.K;
  const int32_t* __restrict__ taps = p.taps + (size_t)z * p.T * 2;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HoWo = p.Ho * p.Wo;

  // fixed (row -> input origin) mapping of the A items this thread loads
  int a_b[A_ITEMS], a_y[A_ITEMS], a_x[A_ITEMS];
#pragma unroll
  for (int i = 0; i < A_ITEMS; ++i) {
    const int idx = tid + i * NT;
    const int row = VEC ? idx / (BK / 4) : idx / BK;
    const int64_t m = m0 + row;
    if (row < BM && m < p.M) {
      const int b = (int)(m / HoWo);
      const int rem = (int)(m - (int64_t)b * HoWo);
      const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
      a_b[i] = b;
      a_y[i] = oy * p.stride;
      a_x[i] = ox * p.stride;
    } else {
      a_b[i] = -1;
      a_y[i] = a_x[i] = 0;
    }
  }

  float4 a_reg[VEC ? A_ITEMS : 1];
  float4 b_reg[VEC ? B_ITEMS : 1];
  float a_sc[VEC ? 1 : A_ITEMS];
  float b_sc[VEC ? 1 : B_ITEMS];

  auto load_tiles = [&](int k0) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < A_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int kq = idx % (BK / 4);
        const int k = k0 + kq * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_b[i] >= 0 && k < p.K) {
          const int t = k / p.Cin, c = k - t * p.Cin;
          const int iy = a_y[i] + __ldg(taps + 2 * t), ix = a_x[i] + __ldg(taps + 2 * t + 1);
          if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi)
            v = __ldg(reinterpret_cast<const float4*>(p.in + ((size_t)(a_b[i] * p.Hi + iy) * p.Wi + ix) * p.lda + c));
        }
        a_reg[i] = v;
      }
#pragma unroll
      for (int i = 0; i < B_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / (BK / 4), kq = idx % (BK / 4);
        const int k = k0 + kq * 4, n = n0 + row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < BN && n < p.N && k < p.K) v = __ldg(reinterpret_cast<const float4*>(wz + (size_t)n * p.K + k));
        b_reg[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int kk = idx % BK;
        const int k = k0 + kk;
        float v = 0.f;
        if (a_b[i] >= 0 && k < p.K) {
          const int t = k / p.Cin, c = k - t * p.Cin;
          const int iy = a_y[i] + __ldg(taps + 2 * t), ix = a_x[i] + __ldg(taps + 2 * t + 1);
          if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi)
            v = __ldg(p.in + ((size_t)(a_b[i] * p.Hi + iy) * p.Wi + ix) * p.lda + c);
        }
        a_sc[i] = v;
      }
#pragma unroll
      for (int i = 0; i < B_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / BK, kk = idx % BK;
        const int k = k0 + kk, n = n0 + row;
        float v = 0.f;
        if (row < BN && n < p.N && k < p.K) v = __ldg(wz + (size_t)n * p.K + k);
        b_sc[i] = v;
      }
    }
  };

  auto store_tiles = [&](int buf) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < A_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / (BK / 4), kq = idx % (BK / 4);
        if (row < BM) {
          As[buf][kq * 4 + 0][row] = a_reg[i].x;
          As[buf][kq * 4 + 1][row] = a_reg[i].y;
          As[buf][kq * 4 + 2][row] = a_reg[i].z;
          As[buf][kq * 4 + 3][row] = a_reg[i].w;
        }
      }
#pragma unroll
      for (int i = 0; i < B_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / (BK / 4), kq = idx % (BK / 4);
        if (row < BN) {
          Bs[buf][kq * 4 + 0][row] = b_reg[i].x;
          Bs[buf][kq * 4 + 1][row] = b_reg[i].y;
          Bs[buf][kq * 4 + 2][row] = b_reg[i].z;
          Bs[buf][kq * 4 + 3][row] = b_reg[i].w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / BK, kk = idx % BK;
        if (row < BM) As[buf][kk][row] = a_sc[i];
      }
#pragma unroll
      for (int i = 0; i < B_ITEMS; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / BK, kk = idx % BK;
        if (row < BN) Bs[buf][kk][row] = b_sc[i];
      }
    }
  };

  const int tyy = tid / TXN, txx = tid % TXN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ktiles = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < RG; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][g * (BM / RG) + tyy * 4]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * (BN / CG) + txx * 4]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias -> activation -> residual -> store (scatter for transposed-conv phases)
  const int Hout = p.Ho * p.out_scale, Wout = p.Wo * p.out_scale;
  const int pz_y = z >> 1, pz_x = z & 1;
  const bool vec_out = ((p.ldc | p.c_off) & 3) == 0 && wxf_aligned16(p.out) &&
                       (!p.res || (((p.ldr | p.r_off) & 3) == 0 && wxf_aligned16(p.res)));
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = (i / 4) * (BM / RG) + tyy * 4 + (i % 4);
    const int64_t m = m0 + row;
    if (m >= p.M) continue;
    int64_t opix = m;
    if (p.out_scale != 1 || pz_y || pz_x) {
      const int b = (int)(m / HoWo);
      const int rem = (int)(m - (int64_t)b * HoWo);
      const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
      opix = ((int64_t)b * Hout + oy * p.out_scale + pz_y) * Wout + ox * p.out_scale + pz_x;
    }
    float* orow = p.out + opix * p.ldc + p.c_off;
    const float* rrow = p.res ? p.res + opix * p.ldr + p.r_off : nullptr;
#pragma unroll
    for (int g = 0; g < CG; ++g) {
      const int n = n0 + g * (BN / CG) + txx * 4;
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[i][g * 4 + j];
        if (n + j < p.N) {
          if (p.bias) t += __ldg(p.bias + z * p.bias_zs + n + j);
          if (p.act == WXF_ACT_GELU_ERF) t = wxf_gelu_erf(t);
        }
        v[j] = t;
      }
      if (vec_out && n + 3 < p.N) {
        if (rrow) {
          const float4 r = *reinterpret_cast<const float4*>(rrow + n);
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        *reinterpret_cast<float4*>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) orow[n + j] = v[j] + (rrow ? rrow[n + j] : 0.f);
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
int launch_cfg(const ConvP& p, int phases, bool vec, cudaStream_t st) {
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN), (unsigned)phases);
  if (vec)
    conv_igemm_kernel<BM, BN, TM, TN, true><<<grid, NT, 0, st>>>(p);
  else
    conv_igemm_kernel<BM, BN, TM, TN, false><<<grid, NT, 0, st>>>(p);
  WXF_CHECK_LAUNCH("conv_igemm_f32");
  return 0;
}

}  // namespace

extern "C" int wxf_conv_igemm_f32(const WxfConvDesc* d, void* stream) {
  if (!d || !d->in || !d->w || !d->taps || !d->out) WXF_FAIL(WXF_EINVAL, "conv: null pointer");
  if (d->B <= 0 || d->Hi <= 0 || d->Wi <= 0 || d->Cin <= 0 || d->N <= 0 || d->T <= 0 || d->stride <= 0 || d->Ho <= 0 ||
      d->Wo <= 0 || d->lda < d->Cin || d->ldc < d->c_off + d->N)
    WXF_FAIL(WXF_EINVAL, "conv: bad dims");
  if (d->phases != 1 && d->phases != 4) WXF_FAIL(WXF_EINVAL, "conv: phases must be 1 or 4");
  if (d->phases == 4 && d->out_scale != 2) WXF_FAIL(WXF_EINVAL, "conv: 4 phases need out_scale 2");
  if (d->phases == 1 && d->out_scale != 1) WXF_FAIL(WXF_EINVAL, "conv: 1 phase needs out_scale 1");
  if (d->res && d->ldr < d->r_off + d->N) WXF_FAIL(WXF_EINVAL, "conv: residual stride");
  ConvP p;
  p.in = d->in; p.w = d->w; p.taps = d->taps; p.bias = d->bias; p.res = d->res; p.out = d->out;
  p.B = d->B; p.Hi = d->Hi; p.Wi = d->Wi; p.lda = d->lda; p.Cin = d->Cin;
  p.N = d->N; p.T = d->T; p.stride = d->stride; p.Ho = d->Ho; p.Wo = d->Wo; p.out_scale = d->out_scale;
  p.ldc = d->ldc; p.c_off = d->c_off; p.ldr = d->ldr; p.r_off = d->r_off; p.act = d->act;
  p.bias_zs = d->bias_phase_stride;
  p.K = d->T * d->Cin;
  p.M = (int64_t)d->B * d->Ho * d->Wo;
  const bool vec = (d->Cin % 4 == 0) && (d->lda % 4 == 0) && wxf_aligned16(d->in) && wxf_aligned16(d->w);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->N > 64) return launch_cfg<128, 128, 8, 8>(p, d->phases, vec, st);
  if (d->N > 32) return launch_cfg<128, 64, 8, 4>(p, d->phases, vec, st);
  if (d->N > 16) return launch_cfg<128, 32, 4, 4>(p, d->phases, vec, st);
  return launch_cfg<256, 16, 4, 4>(p, d->phases, vec, st);
}
