// Exchanges of the domain decomposition over NVLink peer memory (no NCCL on the data path).
//
// Every rank allocates one arena with cudaMalloc, exports it with cudaIpc and opens its peers' arenas: inside one box every
// GPU reaches every peer through NVSwitch, so a halo row, a re-layout of the residual stream or 2 x 128 GroupNorm sums are
// plain stores into the consumer's memory, issued by the producer's own stream as soon as the rows exist.  Arrival is
// signalled by one counter per (exchange site, sender) in the consumer's arena: the sender adds 1 with system scope after
// its data (release), the consumer's stream runs a one-warp kernel that polls until the counter reaches the consumer's
// step number (acquire).  Nothing else synchronises; all of it captures into a CUDA graph as ordinary kernels.
//
//   reference: credit/domain_parallel/halo_exchange.py:56-67 (batch_isend_irecv of 4 P2POps per convolution),
//              credit/domain_parallel/layers.py:507-518 (two all-reduces per GroupNorm).
#include "wxf_common.cuh"

namespace {

constexpr int MAX_SEG = 8;
constexpr int MAX_SIG = 16;

struct PeerPut {
  const void* src[MAX_SEG];
  void* dst[MAX_SEG];      // peer (or local) addresses
  int64_t bytes[MAX_SEG];  // multiples of 16
  uint32_t* sig[MAX_SIG];  // counters in the consumers' arenas, +1 each after all segments are written
  int nseg, nsig;
};

__device__ __forceinline__ void signal_add(uint32_t* p) {
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ uint32_t signal_load(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// done: device counter of this rank (self-resetting) that elects the last CTA to publish the signals
__global__ void __launch_bounds__(256) peer_put_kernel(PeerPut p, unsigned int* done) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  for (int s = 0; s < p.nseg; ++s) {
    const int64_t n16 = p.bytes[s] >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(p.src[s]);
    uint4* dst = reinterpret_cast<uint4*>(p.dst[s]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int i = 0; i < p.nsig; ++i) signal_add(p.sig[i]);
    }
  }
}

struct PeerScatter {
  float* base[8];     // destination buffer of every rank (peer addresses; own buffer for the own rank)
  uint32_t* sig[8];   // counter "from me" in every rank's arena (nullptr for ranks that receive nothing)
  int world;
};

// dst_rank[i] / dst_idx[i]: where row src_idx[i] of the local tensor goes.  Rows are d floats, float4 granularity.
__global__ void __launch_bounds__(256) peer_scatter_rows_kernel(const float* __restrict__ src, int ld_src,
                                                                const int32_t* __restrict__ src_idx,
                                                                const int32_t* __restrict__ dst_rank,
                                                                const int32_t* __restrict__ dst_idx, PeerScatter p, int ld_dst,
                                                                int64_t n, int d4, unsigned int* done) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const int64_t total = n * d4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / d4;
    const int c = (int)(e - i * d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)__ldg(src_idx + i) * ld_src + c);
    *reinterpret_cast<float4*>(p.base[__ldg(dst_rank + i)] + (int64_t)__ldg(dst_idx + i) * ld_dst + c) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int r = 0; r < p.world; ++r)
        if (p.sig[r]) signal_add(p.sig[r]);
    }
  }
}

struct PeerWait {
  const uint32_t* sig[MAX_SIG];
  int nsig;
};

// One warp: lane i polls counter i until it reaches *epoch.  ~2^26 polls (seconds) then trap: a protocol bug becomes a CUDA
// error instead of a hung box.
__global__ void __launch_bounds__(32) peer_wait_kernel(PeerWait w, const uint32_t* __restrict__ epoch) {
  wxf_pdl_trigger();
  wxf_pdl_wait();
  const uint32_t want = *epoch;
  if ((int)threadIdx.x < w.nsig) {
    const uint32_t* s = w.sig[threadIdx.x];
    uint32_t polls = 0;
    while ((int32_t)(signal_load(s) - want) < 0) {
      if (++polls > (1u << 26)) {
        printf("wxf_peer_wait: timeout on counter %d (have %u, want %u)\n", (int)threadIdx.x, signal_load(s), want);
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  __threadfence_system();
}

__global__ void peer_epoch_kernel(uint32_t* epoch) {
  if (threadIdx.x == 0) *epoch += 1;
}

// sums[i] = sum_r slots[r * n + i]  (fixed order: bit-identical on every rank)
__global__ void __launch_bounds__(256) sum_rank_slots_kernel(const double* __restrict__ slots, double* __restrict__ sums, int world,
                                                             int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int r = 0; r < world; ++r) s += slots[(int64_t)r * n + i];
  sums[i] = s;
}

}  // namespace

extern "C" int wxf_peer_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) WXF_FAIL(WXF_EINVAL, "peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) WXF_FAIL((int)e, "peer_alloc: cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) WXF_FAIL((int)e, "peer_alloc: cudaMemset: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int wxf_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) WXF_FAIL((int)e, "peer_free: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int wxf_peer_export(const void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) WXF_FAIL((int)e, "peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int wxf_peer_open(const void* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) WXF_FAIL((int)e, "peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int wxf_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) WXF_FAIL((int)e, "peer_close: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int wxf_peer_put(const void* const* src, void* const* dst, const int64_t* bytes, int nseg, void* const* signals,
                            int nsig, void* done_counter, void* stream) {
  if (nseg < 0 || nseg > MAX_SEG || nsig < 0 || nsig > MAX_SIG || !done_counter) WXF_FAIL(WXF_EINVAL, "peer_put: bad counts");
  PeerPut p{};
  int64_t most = 0;
  for (int i = 0; i < nseg; ++i) {
    if (!src[i] || !dst[i] || bytes[i] <= 0 || (bytes[i] & 15) || !wxf_aligned16(src[i]) || !wxf_aligned16(dst[i]))
      WXF_FAIL(WXF_EALIGN, "peer_put: segment %d must be a non-empty multiple of 16 bytes, 16-byte aligned", i);
    p.src[i] = src[i];
    p.dst[i] = dst[i];
    p.bytes[i] = bytes[i];
    if (bytes[i] > most) most = bytes[i];
  }
  for (int i = 0; i < nsig; ++i) p.sig[i] = reinterpret_cast<uint32_t*>(signals[i]);
  p.nseg = nseg;
  p.nsig = nsig;
  int64_t blocks = (most / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  wxf_launch(peer_put_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, p, reinterpret_cast<unsigned int*>(done_counter));
  WXF_CHECK_LAUNCH("peer_put");
  return 0;
}

extern "C" int wxf_peer_scatter_rows(const float* src, int ld_src, const int32_t* src_idx, const int32_t* dst_rank,
                                     const int32_t* dst_idx, void* const* dst_base, void* const* signals, int world, int ld_dst,
                                     int64_t n, int d, void* done_counter, void* stream) {
  if (!src || !src_idx || !dst_rank || !dst_idx || !dst_base || !signals || world <= 0 || world > 8 || n < 0 || d <= 0 ||
      (d & 3) || (ld_src & 3) || (ld_dst & 3) || ld_src < d || ld_dst < d || !wxf_aligned16(src) || !done_counter)
    WXF_FAIL(WXF_EINVAL, "peer_scatter_rows: bad arguments");
  PeerScatter p{};
  p.world = world;
  for (int r = 0; r < world; ++r) {
    p.base[r] = reinterpret_cast<float*>(dst_base[r]);
    p.sig[r] = reinterpret_cast<uint32_t*>(signals[r]);
    if (p.base[r] && !wxf_aligned16(p.base[r])) WXF_FAIL(WXF_EALIGN, "peer_scatter_rows: destination %d not 16-byte aligned", r);
  }
  int64_t blocks = (n * (d / 4) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  wxf_launch(peer_scatter_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, ld_src, src_idx, dst_rank,
             dst_idx, p, ld_dst, n, d / 4, reinterpret_cast<unsigned int*>(done_counter));
  WXF_CHECK_LAUNCH("peer_scatter_rows");
  return 0;
}

extern "C" int wxf_peer_wait(void* const* signals, int nsig, const void* epoch, void* stream) {
  if (nsig < 0 || nsig > MAX_SIG || !epoch) WXF_FAIL(WXF_EINVAL, "peer_wait: bad arguments");
  if (nsig == 0) return 0;
  PeerWait w{};
  for (int i = 0; i < nsig; ++i) w.sig[i] = reinterpret_cast<const uint32_t*>(signals[i]);
  w.nsig = nsig;
  wxf_launch(peer_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, w, reinterpret_cast<const uint32_t*>(epoch));
  WXF_CHECK_LAUNCH("peer_wait");
  return 0;
}

extern "C" int wxf_peer_epoch_advance(void* epoch, void* stream) {
  if (!epoch) WXF_FAIL(WXF_EINVAL, "peer_epoch_advance: null");
  peer_epoch_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t*>(epoch));
  WXF_CHECK_LAUNCH("peer_epoch_advance");
  return 0;
}

extern "C" int wxf_sum_rank_slots(const double* slots, double* sums, int world, int n, void* stream) {
  if (!slots || !sums || world <= 0 || n <= 0) WXF_FAIL(WXF_EINVAL, "sum_rank_slots: bad arguments");
  sum_rank_slots_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(slots, sums, world, n);
  WXF_CHECK_LAUNCH("sum_rank_slots");
  return 0;
}
