// Micro-benchmark: per-SM throughput of TMA tile loads / stores as a function of the box row length (swizzle span) and
// row count.  Motivation: the kernels whose TMA boxes have 64-byte rows (attention Q/K/V gathers, fp16 plane stores of the
// GEMM epilogue) run at ~2 TB/s while the ones with 128-byte rows reach 4-5 TB/s.  Build: see tools/gpu_tma_bench.sh.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../miles_credit_b200/csrc/wxf_tc_host.cuh"
#include "../miles_credit_b200/csrc/wxf_tc_ptx.cuh"

thread_local char wxf_err_buf[512] = "";
using namespace wxf_tc;

constexpr int NBUF = 3;
constexpr int BUF_BYTES = 32768;

struct LoadParams {
  int rank;            // 2 or 4
  int iters;           // boxes per CTA
  uint32_t box_bytes;
  int tiles0, tiles1, tiles2;  // tile grid (2D: inner tiles, row tiles; 4D: c tiles, x tiles, y tiles)
  int step0, step1, step2;     // coordinate step per tile index
};

__global__ void __launch_bounds__(128, 2) tma_load_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ LoadParams p,
                                                          long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + NBUF * BUF_BYTES;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long total_tiles = (long long)p.tiles0 * p.tiles1 * p.tiles2;
    auto issue = [&](int i) {
      long long t = ((long long)blockIdx.x + (long long)i * gridDim.x) % total_tiles;
      const int t0 = (int)(t % p.tiles0);
      t /= p.tiles0;
      const int t1 = (int)(t % p.tiles1);
      const int t2 = (int)(t / p.tiles1);
      const uint32_t bar = bars + 8 * (i % NBUF), dst = base + (i % NBUF) * BUF_BYTES;
      mbar_expect_tx(bar, p.box_bytes);
      if (p.rank == 2)
        tma_load_2d(&tm, bar, dst, t0 * p.step0, t1 * p.step1);
      else
        tma_load_4d(&tm, bar, dst, t0 * p.step0, t1 * p.step1, t2 * p.step2, 0);
    };
    const long long c0 = clock64();
    for (int i = 0; i < NBUF && i < p.iters; ++i) issue(i);
    for (int i = 0; i < p.iters; ++i) {
      mbar_wait(bars + 8 * (i % NBUF), (uint32_t)(i / NBUF) & 1u);
      if (i + NBUF < p.iters) issue(i + NBUF);
    }
    cycles[blockIdx.x] = clock64() - c0;
  }
}

__global__ void __launch_bounds__(128, 2) tma_store_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ LoadParams p,
                                                           long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < NBUF * BUF_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = i;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long total_tiles = (long long)p.tiles0 * p.tiles1;
    const long long c0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      long long t = ((long long)blockIdx.x + (long long)i * gridDim.x) % total_tiles;
      const int t0 = (int)(t % p.tiles0), t1 = (int)(t / p.tiles0);
      tma_store_2d(&tm, base + (i % NBUF) * BUF_BYTES, t0 * p.step0, t1 * p.step1);
      bulk_commit();
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
    }
    bulk_wait0();
    cycles[blockIdx.x] = clock64() - c0;
  }
}

static double run(bool store, const CUtensorMap& tm, const LoadParams& p, int grid, long long* d_cyc, float* ms_out) {
  const int smem = NBUF * BUF_BYTES + 64 + 1024;
  cudaFuncSetAttribute(tma_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (store)
      tma_store_kernel<<<grid, 128, smem>>>(tm, p, d_cyc);
    else
      tma_load_kernel<<<grid, 128, smem>>>(tm, p, d_cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      exit(1);
    }
  }
  cudaEventElapsedTime(ms_out, e0, e1);
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (auto v : h) sum += (double)v;
  return sum / grid;
}

int main() {
  const size_t big = (size_t)640 << 20;  // > L2
  void* buf;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  long long* d_cyc;
  cudaMalloc(&d_cyc, 1024 * sizeof(long long));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  printf("%-44s %5s %9s %9s %10s %10s %9s\n", "case", "grid", "rows/box", "B/row", "cyc/box/SM", "cyc/row/SM", "GB/s");
  struct Case { const char* name; int inner; int rows; int swz; size_t span; };
  // 2-D loads from a [rows][4096] fp16 matrix; span = bytes of the matrix actually walked (L2-resident vs HBM)
  const Case cases[] = {
      {"load2d 128B x128 rows (GEMM tile) HBM", 64, 128, 128, big},   {"load2d 128B x128 rows L2", 64, 128, 128, (size_t)48 << 20},
      {"load2d  64B x128 rows HBM", 32, 128, 64, big},               {"load2d  64B x128 rows L2", 32, 128, 64, (size_t)48 << 20},
      {"load2d 128B x 32 rows HBM", 64, 32, 128, big},               {"load2d  64B x 32 rows HBM", 32, 32, 64, big},
      {"load2d 128B x256 rows HBM", 64, 256, 128, big},
  };
  const int K = 4096;
  for (const Case& c : cases) {
    for (int mult = 1; mult <= 2; ++mult) {
      const uint64_t M = c.span / (K * 2);
      const uint64_t dims[2] = {(uint64_t)K, M}, strides[1] = {(uint64_t)K * 2};
      const uint32_t box[2] = {(uint32_t)c.inner, (uint32_t)c.rows}, es[2] = {1, 1};
      CUtensorMap tm;
      if (make_map(&tm, buf, 2, dims, strides, box, es, c.swz)) { printf("map failed: %s\n", wxf_err_buf); return 1; }
      LoadParams p{};
      p.rank = 2;
      p.box_bytes = (uint32_t)(c.inner * 2 * c.rows);
      p.tiles0 = K / c.inner; p.tiles1 = (int)(M / c.rows); p.tiles2 = 1;
      p.step0 = c.inner; p.step1 = c.rows;
      p.iters = (int)((size_t)256 * 1024 * 1024 / p.box_bytes / (sms * mult));  // ~256 MB per launch
      if (p.iters > 4000) p.iters = 4000;
      float ms;
      const double cyc = run(false, tm, p, sms * mult, d_cyc, &ms);
      const double bytes = (double)p.box_bytes * p.iters * sms * mult;
      printf("%-44s %5d %9d %9d %10.1f %10.2f %9.0f\n", c.name, sms * mult, c.rows, c.inner * 2, cyc / p.iters / mult,
             cyc / p.iters / c.rows / mult, bytes / ms / 1e6);
    }
  }
  // 4-D window gathers from a [1, 400, 800, 384] fp16 qkv plane (attention stage 0): {32|64 ch, 10, 10, 1}
  for (int inner : {32, 64}) {
    for (int mult = 1; mult <= 2; ++mult) {
      const int H = 400, W = 800, C = 384;
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, 1};
      const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
      const uint32_t box[4] = {(uint32_t)inner, 10, 10, 1}, es[4] = {1, 1, 1, 1};
      CUtensorMap tm;
      if (make_map(&tm, buf, 4, dims, strides, box, es, inner * 2)) { printf("map failed: %s\n", wxf_err_buf); return 1; }
      LoadParams p{};
      p.rank = 4;
      p.box_bytes = (uint32_t)(inner * 2 * 100);
      p.tiles0 = C / inner; p.tiles1 = W / 10; p.tiles2 = H / 10;
      p.step0 = inner; p.step1 = 10; p.step2 = 10;
      p.iters = (int)((size_t)245760000 / p.box_bytes / (sms * mult));
      float ms;
      const double cyc = run(false, tm, p, sms * mult, d_cyc, &ms);
      const double bytes = (double)p.box_bytes * p.iters * sms * mult;
      char name[64];
      snprintf(name, sizeof(name), "load4d window 10x10 x %dB rows", inner * 2);
      printf("%-44s %5d %9d %9d %10.1f %10.2f %9.0f\n", name, sms * mult, 100, inner * 2, cyc / p.iters / mult,
             cyc / p.iters / 100 / mult, bytes / ms / 1e6);
    }
  }
  // 2-D stores into a [M][512] fp16 plane: the GEMM epilogue's 32-row boxes with 64-byte / 128-byte rows
  for (int inner : {32, 64}) {
    for (int mult = 1; mult <= 2; ++mult) {
      const int N = 512;
      const uint64_t M = 320000;
      const uint64_t dims[2] = {(uint64_t)N, M}, strides[1] = {(uint64_t)N * 2};
      const uint32_t box[2] = {(uint32_t)inner, 32}, es[2] = {1, 1};
      CUtensorMap tm;
      if (make_map(&tm, buf, 2, dims, strides, box, es, inner * 2)) { printf("map failed: %s\n", wxf_err_buf); return 1; }
      LoadParams p{};
      p.rank = 2;
      p.box_bytes = (uint32_t)(inner * 2 * 32);
      p.tiles0 = N / inner; p.tiles1 = (int)(M / 32); p.tiles2 = 1;
      p.step0 = inner; p.step1 = 32;
      p.iters = (int)((size_t)M * N * 2 / p.box_bytes / (sms * mult));
      float ms;
      const double cyc = run(true, tm, p, sms * mult, d_cyc, &ms);
      const double bytes = (double)p.box_bytes * p.iters * sms * mult;
      char name[64];
      snprintf(name, sizeof(name), "store2d 32 rows x %dB (plane tile)", inner * 2);
      printf("%-44s %5d %9d %9d %10.1f %10.2f %9.0f\n", name, sms * mult, 32, inner * 2, cyc / p.iters / mult,
             cyc / p.iters / 32 / mult, bytes / ms / 1e6);
    }
  }
  return 0;
}
