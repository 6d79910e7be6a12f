// Host side of the compact output transfer: widen the fp16 rows that came over PCIe back into the reference's
// (N, C) float32 layout (np.save / .lab payload, predict_folds.py:240, kw_utils.py:4-12).
//
//   dst[r][c] = float(src16[r][c]) + row_ref[r]
//
// The D2H copy of the (N, 1909) float32 matrix is the end-to-end bottleneck of the path (7,636 B per frame).
// In the 16-bit modes the head kernel therefore emits fp16 offsets from each row's maximum plus the maximum itself
// (csrc/head.cu, nnam_head_f16): half the bytes over PCIe.  This function is plain C++ (no CUDA), multi-threaded, F16C +
// AVX2 when the CPU has them, with non-temporal stores so that the 2 x larger output does not cost a read-for-ownership.
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/nnam_b200.h"

namespace {

inline float half_to_float(uint16_t h) {
  const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {  // subnormal: renormalise
      int e = -1;
      do {
        ++e;
        man <<= 1;
      } while ((man & 0x400u) == 0);
      bits = sign | static_cast<uint32_t>(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | man << 13;
  } else {
    bits = sign | (exp + 112u) << 23 | man << 13;
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

void widen_rows_scalar(const uint16_t* src, long long ld16, const float* ref, float* dst, long long ld_dst, long long r0,
                       long long r1, int cols) {
  for (long long r = r0; r < r1; ++r) {
    const uint16_t* s = src + r * ld16;
    float* d = dst + r * ld_dst;
    const float a = ref[r];
    for (int c = 0; c < cols; ++c) d[c] = half_to_float(s[c]) + a;
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2,f16c"))) void widen_rows_f16c(const uint16_t* src, long long ld16, const float* ref, float* dst,
                                                         long long ld_dst, long long r0, long long r1, int cols) {
  for (long long r = r0; r < r1; ++r) {
    const uint16_t* s = src + r * ld16;
    float* d = dst + r * ld_dst;
    const float a = ref[r];
    const __m256 va = _mm256_set1_ps(a);
    int c = 0;
    // scalar head until the destination is 32-byte aligned (rows of 1909 floats start at any 4-byte boundary)
    while (c < cols && (reinterpret_cast<uintptr_t>(d + c) & 31u) != 0) {
      d[c] = _cvtsh_ss(s[c]) + a;
      ++c;
    }
    for (; c + 8 <= cols; c += 8) {
      const __m128i h = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + c));
      _mm256_stream_ps(d + c, _mm256_add_ps(_mm256_cvtph_ps(h), va));
    }
    for (; c < cols; ++c) d[c] = _cvtsh_ss(s[c]) + a;
  }
  _mm_sfence();
}
#endif

}  // namespace

extern "C" int nnam_widen_f16_host(const void* src16_host, long long ld16, const float* row_ref_host, float* dst_host,
                                   long long ld_dst, long long rows, int cols, int threads) {
  if (rows < 0 || cols <= 0 || ld16 < cols || ld_dst < cols || !src16_host || !row_ref_host || !dst_host) return NNAM_ERR_ARG;
  if (rows == 0) return NNAM_OK;
  const uint16_t* src = static_cast<const uint16_t*>(src16_host);
  auto run = [&](long long r0, long long r1) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c")) {
      widen_rows_f16c(src, ld16, row_ref_host, dst_host, ld_dst, r0, r1, cols);
      return;
    }
#endif
    widen_rows_scalar(src, ld16, row_ref_host, dst_host, ld_dst, r0, r1, cols);
  };
  long long t = threads < 1 ? 1 : threads;
  if (t > rows / 256 + 1) t = rows / 256 + 1;  // not worth a thread for a few rows
  if (t == 1) {
    run(0, rows);
    return NNAM_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(static_cast<size_t>(t));
  for (long long i = 0; i < t; ++i) pool.emplace_back(run, rows * i / t, rows * (i + 1) / t);
  for (auto& th : pool) th.join();
  return NNAM_OK;
}
