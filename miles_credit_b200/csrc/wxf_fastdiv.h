// Division of a 32-bit unsigned integer by a launch-constant divisor without a division sequence (Granlund-Montgomery
// round-up form: exact for EVERY 32-bit n and every divisor 1 <= d < 2^32).  The tile -> (head, window) and window ->
// (image, row, column) decodes of the window-attention kernel run once per tile on the softmax warps' critical path; the
// compiler's division (I2F, MUFU.RCP, F2I, two IMAD.HI, fix-ups: ~25 dependent instructions) cost 25 % of their time.
// Plain C++ on the host side so that tests/test_fastdiv_cpu.py can compile it with g++ and check it against `/`.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WXF_FD_HD __host__ __device__ __forceinline__
#else
#define WXF_FD_HD inline
#endif

struct FastDiv {
  uint32_t m, sh1, sh2, d;
};

inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;  // ceil(log2 d)
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l > 0 ? l - 1 : 0;
  f.d = d;
  return f;
}

// high 32 bits of a 32 x 32 -> 64 bit product
WXF_FD_HD uint32_t wxf_umulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

WXF_FD_HD uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = wxf_umulhi(n, f.m);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
