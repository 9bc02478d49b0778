"""The multiply-high division the window-attention kernel decodes tiles with (csrc/wxf_fastdiv.h) against `/` on the host:
the header is plain C++ outside nvcc, so g++ compiles the very functions the kernel calls."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <cstdint>
#include <initializer_list>
#include "wxf_fastdiv.h"
static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint32_t next32() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 16); }
static long check(uint32_t d) {
  const FastDiv f = make_fastdiv(d);
  long bad = 0;
  const uint32_t edge[] = {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffffu, 0x80000000u, 0xffffffffu, 0xffffffffu - d};
  for (uint32_t n : edge) bad += fdiv(n, f) != n / d;
  for (int i = 0; i < 2000; ++i) { const uint32_t n = next32(); bad += fdiv(n, f) != n / d; }
  return bad;
}
int main() {
  long bad = 0;
  for (uint32_t d = 1; d <= 5000; ++d) bad += check(d);                 // every divisor a forecast grid produces
  for (uint32_t n = 0; n < 200000; ++n) for (uint32_t d : {1u, 2u, 3u, 4u, 7u, 32u, 100u, 3200u, 12800u}) bad += fdiv(n, make_fastdiv(d)) != n / d;
  for (int i = 0; i < 20000; ++i) { uint32_t d = next32(); if (!d) d = 1; bad += check(d); }  // arbitrary 32-bit divisors
  for (uint32_t d : {0x7fffffffu, 0x80000000u, 0x80000001u, 0xfffffffeu, 0xffffffffu}) bad += check(d);
  std::printf("%ld\n", bad);
  return bad != 0;
}
"""


def test_fastdiv_matches_integer_division(tmp_path):
    src = tmp_path / "fd.cpp"
    src.write_text(SRC)
    exe = tmp_path / "fd"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "miles_credit_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout + out.stderr
