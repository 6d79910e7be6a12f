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

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
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

void widen_rows_scalar(const uint16_t* src, long long ld16, const float* ref, float* dst, long long ld_dst,
                       const long long* dst_rows, long long r0, long long r1, int cols) {
  for (long long r = r0; r < r1; ++r) {
    const uint16_t* s = src + r * ld16;
    float* d = dst + (dst_rows ? dst_rows[r] : r) * ld_dst;
    const float a = ref[r];
    for (int c = 0; c < cols; ++c) d[c] = half_to_float(s[c]) + a;
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2,f16c"))) void widen_rows_f16c(const uint16_t* src, long long ld16, const float* ref, float* dst,
                                                         long long ld_dst, const long long* dst_rows, long long r0,
                                                         long long r1, int cols) {
  for (long long r = r0; r < r1; ++r) {
    const uint16_t* s = src + r * ld16;
    float* d = dst + (dst_rows ? dst_rows[r] : r) * ld_dst;
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

// Persistent worker pool: a widening call covers ~30 MB (a piece small enough to still sit in the last-level cache when
// it is read back, see engine._ChunkWriter), i.e. ~0.5 ms of work on 16 threads -- creating the threads per call would
// cost as much again.  One job at a time (calls from several Python threads are serialised by the job mutex).
class Pool {
 public:
  static Pool& get() {
    static Pool p;
    return p;
  }
  // run fn(part) for part in [0, parts) on up to `threads` threads (the caller is one of them)
  void run(int parts, int threads, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> job_lock(job_mu_);
    grow(threads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      parts_ = parts;
      next_.store(0);
      pending_ = parts;
      helpers_ = threads - 1;
      ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  Pool() = default;
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void grow(int n) {
    while (static_cast<int>(workers_.size()) < n) {
      const int id = static_cast<int>(workers_.size());
      workers_.emplace_back([this, id] {
        unsigned long long seen = 0;
        for (;;) {
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && id < helpers_); });
            if (stop_) return;
            seen = epoch_;
          }
          work();
        }
      });
    }
  }
  void work() {
    for (;;) {
      const int part = next_.fetch_add(1);
      if (part >= parts_) break;
      (*fn_)(part);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  std::mutex job_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int parts_ = 0, pending_ = 0, helpers_ = 0;
  unsigned long long epoch_ = 0;
  bool stop_ = false;
};

}  // namespace

extern "C" int nnam_widen_f16_host(const void* src16_host, long long ld16, const float* row_ref_host, float* dst_host,
                                   long long ld_dst, const long long* dst_rows_host, long long rows, int cols,
                                   int threads) {
  if (rows < 0 || cols <= 0 || ld16 < cols || ld_dst < cols || !src16_host || !row_ref_host || !dst_host) return NNAM_ERR_ARG;
  if (rows == 0) return NNAM_OK;
  const uint16_t* src = static_cast<const uint16_t*>(src16_host);
  auto run = [&](long long r0, long long r1) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c")) {
      widen_rows_f16c(src, ld16, row_ref_host, dst_host, ld_dst, dst_rows_host, r0, r1, cols);
      return;
    }
#endif
    widen_rows_scalar(src, ld16, row_ref_host, dst_host, ld_dst, dst_rows_host, r0, r1, cols);
  };
  long long t = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  if (t > rows / 64 + 1) t = rows / 64 + 1;  // not worth a thread for a few rows
  if (t == 1) {
    run(0, rows);
    return NNAM_OK;
  }
  // twice as many parts as threads: a thread that was descheduled does not hold the call up for a whole share
  const int parts = static_cast<int>(2 * t);
  Pool::get().run(parts, static_cast<int>(t), [&](int part) { run(rows * part / parts, rows * (part + 1) / parts); });
  return NNAM_OK;
}
