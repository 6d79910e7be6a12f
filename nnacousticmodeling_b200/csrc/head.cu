// K4 -- fused output head: (ensemble mean ->) (RPL4 ->) minus log-prior -> log-softmax.
//
// Reference semantics:
//   y - logsum(y, axis=1)                       predict_folds.py:57,88 with kw_utils.py:38-43
//   y = y - ap ; y = y - logsum(y, axis=1)      evaluateModelForTest.py:75-77,110-112 (ap = ap_coef * log_ap)
//   NNWithRPL.__call__ logit averaging          evaluate.py:35-51
//   RPL4.__call__                               RPL.py:68-74
//   dev-mode mean of fold log-softmax outputs   predict_folds.py:199-219
//
// HBM-bound: one warp owns one row, the whole row lives in registers (C <= 2048), reductions are warp
// shuffles, so every logit is read once and every output written once: (n_inputs + 1) * C * 4 B per frame.
#include <math_constants.h>

#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

constexpr int HEAD_MAX_INPUTS = 16;
constexpr int HEAD_THREADS = 256;

struct HeadParams {
  const float* in[HEAD_MAX_INPUTS];
  float w[HEAD_MAX_INPUTS];
  int n_inputs;
  long long ld_in;
  int pre_normalize;
  const float* rpl_w;
  const float* rpl_b;
  const float* rpl_lb;
  const float* prior;
  float prior_scale;
  int final_normalize;
  float* out;
  long long ld_out;
  long long rows;
  int n_classes;
  const int* out_row_map;  // optional scatter: logits row r -> out row map[r] (negative: drop)
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// log(sum(exp(v))) over the row held as v[i] <-> column lane + 32*i (invalid columns are skipped)
template <int NV>
__device__ __forceinline__ float row_logsumexp(const float (&v)[NV], int lane, int n_classes) {
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < n_classes) mx = fmaxf(mx, v[i]);
  mx = warp_max(mx);
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < n_classes) s += expf(v[i] - mx);
  s = warp_sum(s);
  return mx + logf(s);
}

template <int NV, bool MULTI>
__global__ void __launch_bounds__(HEAD_THREADS, (MULTI || NV < 64) ? 1 : 2) head_kernel(const HeadParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * HEAD_THREADS + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * HEAD_THREADS) >> 5;
  const int C = p.n_classes;
  for (long long row = warp_global; row < p.rows; row += n_warps) {
    float h[NV];
    if (!MULTI) {  // one input, no per-input normalisation: the common predict()/evaluate path
      const float* src = p.in[0] + row * p.ld_in;
      const float w0 = p.w[0];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        h[i] = c < C ? __ldg(src + c) * w0 : 0.0f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) h[i] = 0.0f;
      for (int k = 0; k < p.n_inputs; ++k) {
        const float* src = p.in[k] + row * p.ld_in;
        float t[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + 32 * i;
          t[i] = c < C ? __ldg(src + c) : 0.0f;
        }
        const float lse = p.pre_normalize ? row_logsumexp<NV>(t, lane, C) : 0.0f;
        const float wk = p.w[k];
#pragma unroll
        for (int i = 0; i < NV; ++i) h[i] += wk * (t[i] - lse);
      }
    }
    if (p.rpl_w != nullptr) {
      const float lse = row_logsumexp<NV>(h, lane, C);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
          const float x = h[i] - lse;
          const float g = x + x * __ldg(p.rpl_w + c) + __ldg(p.rpl_b + c);
          const float lb = __ldg(p.rpl_lb + c);
          const float mx = fmaxf(g, lb), mn = fminf(g, lb);
          h[i] = mx + logf(1.0f + expf(mn - mx));
        }
      }
    }
    if (p.prior != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) h[i] -= p.prior_scale * __ldg(p.prior + c);
      }
    }
    const float lse = p.final_normalize ? row_logsumexp<NV>(h, lane, C) : 0.0f;
    long long out_row = row;
    if (p.out_row_map != nullptr) {
      out_row = __ldg(p.out_row_map + row);
      if (out_row < 0) continue;  // warp-uniform: one row per warp
    }
    float* dst = p.out + out_row * p.ld_out;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) dst[c] = h[i] - lse;
    }
  }
}

int head(const float* const* logits_host, const float* weights_host, int n_inputs, long long ld_in, int pre_normalize,
         const float* rpl_w, const float* rpl_b, const float* rpl_lb, const float* prior, float prior_scale,
         int final_normalize, float* out, long long ld_out, long long rows, int n_classes, const int* out_row_map,
         cudaStream_t stream) {
  if (n_inputs < 1 || n_inputs > HEAD_MAX_INPUTS)
    return set_error(NNAM_ERR_ARG, "head: n_inputs must be in [1, %d]", HEAD_MAX_INPUTS);
  if (n_classes < 1 || n_classes > 2048) return set_error(NNAM_ERR_UNSUPPORTED, "head: n_classes must be in [1, 2048]");
  if (ld_in < n_classes || ld_out < n_classes) return set_error(NNAM_ERR_ARG, "head: leading dimension < n_classes");
  if (rows < 0) return set_error(NNAM_ERR_ARG, "head: negative row count");
  if (rows == 0) return NNAM_OK;
  if ((rpl_w != nullptr) != (rpl_b != nullptr) || (rpl_w != nullptr) != (rpl_lb != nullptr))
    return set_error(NNAM_ERR_ARG, "head: RPL4 needs W, b and lb together");
  HeadParams p;
  for (int k = 0; k < n_inputs; ++k) {
    if (logits_host[k] == nullptr) return set_error(NNAM_ERR_ARG, "head: logits[%d] is NULL", k);
    p.in[k] = logits_host[k];
    p.w[k] = weights_host ? weights_host[k] : 1.0f / n_inputs;
  }
  p.n_inputs = n_inputs;
  p.ld_in = ld_in;
  p.pre_normalize = pre_normalize;
  p.rpl_w = rpl_w;
  p.rpl_b = rpl_b;
  p.rpl_lb = rpl_lb;
  p.prior = prior;
  p.prior_scale = prior_scale;
  p.final_normalize = final_normalize;
  p.out = out;
  p.ld_out = ld_out;
  p.rows = rows;
  p.n_classes = n_classes;
  p.out_row_map = out_row_map;
  const int warps_per_block = HEAD_THREADS / 32;
  long long blocks = (rows + warps_per_block - 1) / warps_per_block;
  const long long cap = static_cast<long long>(sm_count()) * 8;  // grid-stride beyond 8 resident blocks per SM
  if (blocks > cap) blocks = cap;
  const bool multi = n_inputs > 1 || pre_normalize;
  const unsigned g = static_cast<unsigned>(blocks);
#define NNAM_HEAD_LAUNCH(NV)                                              \
  do {                                                                    \
    if (multi)                                                            \
      head_kernel<NV, true><<<g, HEAD_THREADS, 0, stream>>>(p);           \
    else                                                                  \
      head_kernel<NV, false><<<g, HEAD_THREADS, 0, stream>>>(p);          \
  } while (0)
  if (n_classes <= 128)
    NNAM_HEAD_LAUNCH(4);
  else if (n_classes <= 1024)
    NNAM_HEAD_LAUNCH(32);
  else
    NNAM_HEAD_LAUNCH(64);
#undef NNAM_HEAD_LAUNCH
  return check_launch("head_kernel");
}

}  // namespace nnam
