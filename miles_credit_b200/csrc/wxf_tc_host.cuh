// Host-side TMA tensor-map construction shared by the tcgen05 kernels (driver entry point looked up at run time,
// so the library does not link against libcuda).
#pragma once
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "wxf_common.cuh"

namespace wxf_tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
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
  uint64_t d[5], s[4];
  uint32_t box[5], es[5], rank, swz;
  bool operator==(const MapKey& o) const {
    if (ptr != o.ptr || rank != o.rank || swz != o.swz) return false;
    for (int i = 0; i < 5; ++i)
      if (d[i] != o.d[i] || box[i] != o.box[i] || es[i] != o.es[i]) return false;
    for (int i = 0; i < 4; ++i)
      if (s[i] != o.s[i]) return false;
    return true;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull + k.rank * 131 + k.swz;
    for (int i = 0; i < 5; ++i) h = (h ^ (k.d[i] * 1315423911ull + k.box[i] * 2654435761ull + k.es[i])) * 0x100000001B3ull;
    for (int i = 0; i < 4; ++i) h = (h ^ k.s[i]) * 0x100000001B3ull;
    return (size_t)h;
  }
};

// fp16 tensor map (rank <= 5), 128B or 64B swizzle, zero OOB fill; dims innermost-first, strides in bytes for dims 1..
inline int make_map(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box, const uint32_t* estr, uint32_t swizzle_bytes = 128, bool f32 = false) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{};
  key.ptr = ptr;
  key.rank = rank;
  key.swz = swizzle_bytes + (f32 ? 1u : 0u);
  for (uint32_t i = 0; i < rank; ++i) {
    key.d[i] = dims[i];
    key.box[i] = box[i];
    key.es[i] = estr[i];
    if (i + 1 < rank) key.s[i] = strides[i];
  }
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) WXF_FAIL(WXF_EUNSUPPORTED, "tc: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t b[5], e[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    b[i] = box[i];
    e[i] = estr[i];
    if (i + 1 < rank) gstr[i] = strides[i];
  }
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) WXF_FAIL(WXF_EINVAL, "tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

inline int make_map_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows}, strides[1] = {ld * 2};
  const uint32_t box[2] = {64u, box_rows}, es[2] = {1, 1};
  return make_map(out, ptr, 2, dims, strides, box, es);
}


}  // namespace wxf_tc
