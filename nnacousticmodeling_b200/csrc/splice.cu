// K1 -- context splice + Kaldi feature transform (+ i-vector append), and fp32 -> bf16(/split) staging.
//
// Reference semantics (bit-exact in fp32): prepareBatch, scripts/util/kw_nn_utils.py:19-43 (== splicing,
// scripts/util/kw_utils.py:24-36); applyKaldiFeatureTransform, kw_nn_utils.py:13-17 ((x + addShift) * rescale
// as two separately rounded fp32 ops); i-vector concatenate AFTER the transform, evaluate.py:169-171.
//
// HBM-bound gather.  A block owns TILE_F consecutive frames.  Because the splice window of frame f is the
// CONTIGUOUS raw span rows [f-S, f+S] (dim floats each), the block stages rows [f0-S, f0+TILE_F+S) in shared
// memory with coalesced 128-bit loads (clamped at the ends of the whole array), and output element (f, c) is
// simply smem[(f - f0) * dim + c]: every output row is produced with 128-bit shared loads and coalesced
// 128-bit global stores.  Algorithmic traffic: read dim*4 (+ivec) and write (winlen*dim + ivec)*elem bytes
// per frame.
#include <stdlib.h>

#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

// the "hi" plane of a 16-bit output: fp16 for NNAM_OUT_F16, bf16 for NNAM_OUT_BF16 / NNAM_OUT_BF16_SPLIT
template <int OUT_KIND>
__device__ __forceinline__ uint32_t pack_hi(float a, float b) {
  return OUT_KIND == NNAM_OUT_F16 ? pack_f16x2(a, b) : pack_bf16x2(a, b);
}

constexpr int SPLICE_TILE_F = 64;
constexpr int SPLICE_THREADS = 256;

struct SpliceParams {
  const float* x;
  long long x_row0, n_total;
  int dim, splice, winlen;
  const float* add_shift;
  const float* rescale;
  const float* ivec;
  int ivec_dim;
  long long f0, f1;
  void* out_hi;
  void* out_lo;
  long long ldo;
  int spl_cols;  // winlen * dim
  int tile_f;    // frames per CTA (<= SPLICE_TILE_F), chosen per launch so that the grid fills whole waves
};

// value of output column c for tile-local frame r (raw rows staged at s_x, transform at s_add / s_mul)
__device__ __forceinline__ float splice_elem(const SpliceParams& p, const float* s_x, const float* s_add,
                                             const float* s_mul, int r, long long f, int c) {
  if (c < p.spl_cols) {
    float v = s_x[r * p.dim + c];
    if (s_add != nullptr) v = __fmul_rn(__fadd_rn(v, s_add[c]), s_mul[c]);
    return v;
  }
  c -= p.spl_cols;
  if (c < p.ivec_dim) return __ldg(p.ivec + (f - p.f0) * p.ivec_dim + c);
  return 0.0f;
}

template <int OUT_KIND, bool VEC>
__global__ void __launch_bounds__(SPLICE_THREADS) splice_transform_kernel(const SpliceParams p) {
  extern __shared__ __align__(16) float s_mem[];
  const int halo_rows = p.tile_f + 2 * p.splice;
  float* s_x = s_mem;                        // halo_rows * dim
  float* s_add = s_x + halo_rows * p.dim;    // spl_cols (only when a transform is given)
  float* s_mul = s_add + p.spl_cols;
  float* s_iv = s_mul + p.spl_cols;          // VEC only: the tile's i-vector rows, tile_f x ivec_dim
  const bool has_ft = p.add_shift != nullptr;
  const long long tile_f0 = p.f0 + static_cast<long long>(blockIdx.x) * p.tile_f;
  const int tile_n = static_cast<int>(min(static_cast<long long>(p.tile_f), p.f1 - tile_f0));
  const int need_rows = tile_n + 2 * p.splice;

  // ---- stage raw rows (clamped to the whole array), the tile's i-vector rows and the transform vectors
  if (VEC) {
    // Every thread first ISSUES all the loads of its share (feature rows, i-vector rows, transform) and only then
    // stores them: written as plain copy loops, each load -> store pair waited for its own DRAM round trip (13 of them
    // back to back per thread), and that latency chain, not bandwidth, set the kernel's duration.
    constexpr int UX = 4, UI = 8, UT = 2;
    const int tid = threadIdx.x;
    const int vpr = p.dim >> 2;
    const int nx4 = need_rows * vpr;
    const int n4 = p.ivec_dim > 0 ? (tile_n * p.ivec_dim) >> 2 : 0;
    // i-vector rows of the tile are contiguous in global memory
    const float4* iv4 = reinterpret_cast<const float4*>(p.ivec + (tile_f0 - p.f0) * p.ivec_dim);
    auto x_src = [&](int i) {
      const int row = i / vpr, v = i - row * vpr;
      long long g = tile_f0 - p.splice + row;
      g = g < 0 ? 0 : (g >= p.n_total ? p.n_total - 1 : g);
      return reinterpret_cast<const float4*>(p.x + (g - p.x_row0) * p.dim) + v;
    };
    float4 xr[UX], ir[UI];
    float ar[UT], mr[UT];
#pragma unroll
    for (int u = 0; u < UX; ++u) {
      const int i = tid + u * SPLICE_THREADS;
      if (i < nx4) xr[u] = __ldg(x_src(i));
    }
#pragma unroll
    for (int u = 0; u < UI; ++u) {
      const int i = tid + u * SPLICE_THREADS;
      if (i < n4) ir[u] = __ldg(iv4 + i);
    }
    if (has_ft) {
#pragma unroll
      for (int u = 0; u < UT; ++u) {
        const int i = tid + u * SPLICE_THREADS;
        if (i < p.spl_cols) {
          ar[u] = __ldg(p.add_shift + i);
          mr[u] = __ldg(p.rescale + i);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UX; ++u) {
      const int i = tid + u * SPLICE_THREADS;
      if (i < nx4) reinterpret_cast<float4*>(s_x)[i] = xr[u];
    }
#pragma unroll
    for (int u = 0; u < UI; ++u) {
      const int i = tid + u * SPLICE_THREADS;
      if (i < n4) reinterpret_cast<float4*>(s_iv)[i] = ir[u];
    }
    if (has_ft) {
#pragma unroll
      for (int u = 0; u < UT; ++u) {
        const int i = tid + u * SPLICE_THREADS;
        if (i < p.spl_cols) {
          s_add[i] = ar[u];
          s_mul[i] = mr[u];
        }
      }
    }
    // whatever does not fit the first round (wide windows, long i-vectors)
    for (int i = tid + UX * SPLICE_THREADS; i < nx4; i += SPLICE_THREADS) reinterpret_cast<float4*>(s_x)[i] = __ldg(x_src(i));
    for (int i = tid + UI * SPLICE_THREADS; i < n4; i += SPLICE_THREADS) reinterpret_cast<float4*>(s_iv)[i] = __ldg(iv4 + i);
    if (has_ft) {
      for (int i = tid + UT * SPLICE_THREADS; i < p.spl_cols; i += SPLICE_THREADS) {
        s_add[i] = __ldg(p.add_shift + i);
        s_mul[i] = __ldg(p.rescale + i);
      }
    }
  } else {
    for (int i = threadIdx.x; i < need_rows * p.dim; i += SPLICE_THREADS) {
      const int row = i / p.dim, d = i - row * p.dim;
      long long g = tile_f0 - p.splice + row;
      g = g < 0 ? 0 : (g >= p.n_total ? p.n_total - 1 : g);
      s_x[i] = __ldg(p.x + (g - p.x_row0) * p.dim + d);
    }
    if (has_ft) {
      for (int i = threadIdx.x; i < p.spl_cols; i += SPLICE_THREADS) {
        s_add[i] = __ldg(p.add_shift + i);
        s_mul[i] = __ldg(p.rescale + i);
      }
    }
  }
  __syncthreads();
  const float* t_add = has_ft ? s_add : nullptr;

  // ---- emit
  if (OUT_KIND == NNAM_OUT_F32) {
    float* out = static_cast<float*>(p.out_hi) + (tile_f0 - p.f0) * p.ldo;
    if (VEC) {
      const int vpr = static_cast<int>(p.ldo >> 2);
      for (int i = threadIdx.x; i < tile_n * vpr; i += SPLICE_THREADS) {
        const int r = i / vpr, c = (i - r * vpr) << 2;
        float4 o;
        if (c + 4 <= p.spl_cols) {
          const float4 v = *reinterpret_cast<const float4*>(s_x + r * p.dim + c);
          if (has_ft) {
            const float4 a = *reinterpret_cast<const float4*>(s_add + c);
            const float4 m = *reinterpret_cast<const float4*>(s_mul + c);
            o.x = __fmul_rn(__fadd_rn(v.x, a.x), m.x);
            o.y = __fmul_rn(__fadd_rn(v.y, a.y), m.y);
            o.z = __fmul_rn(__fadd_rn(v.z, a.z), m.z);
            o.w = __fmul_rn(__fadd_rn(v.w, a.w), m.w);
          } else {
            o = v;
          }
        } else if (c >= p.spl_cols && c + 4 <= p.spl_cols + p.ivec_dim) {
          o = *reinterpret_cast<const float4*>(s_iv + r * p.ivec_dim + (c - p.spl_cols));
        } else {
          o.x = splice_elem(p, s_x, t_add, s_mul, r, tile_f0 + r, c);
          o.y = splice_elem(p, s_x, t_add, s_mul, r, tile_f0 + r, c + 1);
          o.z = splice_elem(p, s_x, t_add, s_mul, r, tile_f0 + r, c + 2);
          o.w = splice_elem(p, s_x, t_add, s_mul, r, tile_f0 + r, c + 3);
        }
        reinterpret_cast<float4*>(out + static_cast<long long>(r) * p.ldo)[c >> 2] = o;
      }
    } else {
      const int ld = static_cast<int>(p.ldo);
      for (int i = threadIdx.x; i < tile_n * ld; i += SPLICE_THREADS) {
        const int r = i / ld, c = i - r * ld;
        out[static_cast<long long>(r) * p.ldo + c] = splice_elem(p, s_x, t_add, s_mul, r, tile_f0 + r, c);
      }
    }
  } else {
    // bf16 (hi) or bf16 hi/lo split: 8 elements (16 B) per thread per store; ldo % 8 == 0 is required.
    __nv_bfloat16* out_hi = static_cast<__nv_bfloat16*>(p.out_hi) + (tile_f0 - p.f0) * p.ldo;
    __nv_bfloat16* out_lo =
        OUT_KIND == NNAM_OUT_BF16_SPLIT ? static_cast<__nv_bfloat16*>(p.out_lo) + (tile_f0 - p.f0) * p.ldo : nullptr;
    // A thread owns ONE 16-byte output chunk column (8 elements) and walks the tile's rows: the transform of its
    // columns stays in registers, so a chunk costs two 128-bit shared loads and one 128-bit store (the first version
    // re-read the transform from shared memory for every chunk and ran at 94 % L1/shared-pipe utilisation).
    const int vpr = static_cast<int>(p.ldo >> 3);            // chunks per output row
    const bool narrow = vpr <= SPLICE_THREADS;               // the usual case: one chunk column per thread
    const int lanes_r = narrow ? SPLICE_THREADS / vpr : 1;   // rows processed per sweep of the block
    const int tid = static_cast<int>(threadIdx.x);
    const int r_first = narrow ? tid / vpr : 0;
    const bool active = !narrow || tid < vpr * lanes_r;
    for (int cc = narrow ? tid % vpr : tid; active && cc < vpr; cc += narrow ? vpr : SPLICE_THREADS) {
      const int c = cc << 3;
      const int kind = (VEC && c + 8 <= p.spl_cols) ? 0
                       : (VEC && c >= p.spl_cols && c + 8 <= p.spl_cols + p.ivec_dim && (p.spl_cols & 7) == 0) ? 1 : 2;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, m0 = make_float4(1.f, 1.f, 1.f, 1.f), m1 = m0;
      if (kind == 0 && has_ft) {
        a0 = *reinterpret_cast<const float4*>(s_add + c);
        a1 = *reinterpret_cast<const float4*>(s_add + c + 4);
        m0 = *reinterpret_cast<const float4*>(s_mul + c);
        m1 = *reinterpret_cast<const float4*>(s_mul + c + 4);
      }
      // row loops specialised per column kind, with pointer increments instead of per-row address arithmetic (the
      // generic form cost ~125 instructions per 16-byte chunk and made the kernel issue-bound at 42 % of the slots)
      const long long row_step = static_cast<long long>(lanes_r) * p.ldo;
      __nv_bfloat16* dh = out_hi + static_cast<long long>(r_first) * p.ldo + c;
      __nv_bfloat16* dl = OUT_KIND == NNAM_OUT_BF16_SPLIT ? out_lo + static_cast<long long>(r_first) * p.ldo + c : nullptr;
      auto emit = [&](const float4& v0, const float4& v1) {
        *reinterpret_cast<uint4*>(dh) = make_uint4(pack_hi<OUT_KIND>(v0.x, v0.y), pack_hi<OUT_KIND>(v0.z, v0.w),
                                                   pack_hi<OUT_KIND>(v1.x, v1.y), pack_hi<OUT_KIND>(v1.z, v1.w));
        dh += row_step;
        if (OUT_KIND == NNAM_OUT_BF16_SPLIT) {
          *reinterpret_cast<uint4*>(dl) = make_uint4(
              pack_bf16x2(v0.x - bf16_round(v0.x), v0.y - bf16_round(v0.y)),
              pack_bf16x2(v0.z - bf16_round(v0.z), v0.w - bf16_round(v0.w)),
              pack_bf16x2(v1.x - bf16_round(v1.x), v1.y - bf16_round(v1.y)),
              pack_bf16x2(v1.z - bf16_round(v1.z), v1.w - bf16_round(v1.w)));
          dl += row_step;
        }
      };
      if (kind == 0) {
        const float* src = s_x + r_first * p.dim + c;
        const int src_step = lanes_r * p.dim;
        if (has_ft) {
          for (int r = r_first; r < tile_n; r += lanes_r, src += src_step) {
            float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
            v0.x = __fmul_rn(__fadd_rn(v0.x, a0.x), m0.x);
            v0.y = __fmul_rn(__fadd_rn(v0.y, a0.y), m0.y);
            v0.z = __fmul_rn(__fadd_rn(v0.z, a0.z), m0.z);
            v0.w = __fmul_rn(__fadd_rn(v0.w, a0.w), m0.w);
            v1.x = __fmul_rn(__fadd_rn(v1.x, a1.x), m1.x);
            v1.y = __fmul_rn(__fadd_rn(v1.y, a1.y), m1.y);
            v1.z = __fmul_rn(__fadd_rn(v1.z, a1.z), m1.z);
            v1.w = __fmul_rn(__fadd_rn(v1.w, a1.w), m1.w);
            emit(v0, v1);
          }
        } else {
          for (int r = r_first; r < tile_n; r += lanes_r, src += src_step)
            emit(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4));
        }
      } else if (kind == 1) {  // i-vector columns: from the staged tile, 128-bit
        const float* src = s_iv + r_first * p.ivec_dim + (c - p.spl_cols);
        const int src_step = lanes_r * p.ivec_dim;
        for (int r = r_first; r < tile_n; r += lanes_r, src += src_step)
          emit(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4));
      } else {  // ragged columns (transform tail, padding): element by element
        for (int r = r_first; r < tile_n; r += lanes_r) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = splice_elem(p, s_x, t_add, s_mul, r, tile_f0 + r, c + j);
          emit(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Streaming variant for the product path (bf16 / bf16 hi+lo output, dim and ivec_dim multiples of 4, spliced width a
// multiple of 8): persistent CTAs walk tiles of `tile_f` frames; the raw rows of a tile (+ halo) and its i-vector rows
// are two contiguous spans of global memory, so ONE thread fetches them with two bulk copies (cp.async.bulk) into a
// double buffer while the block emits the previous tile.  Everything that depends only on a thread's output column --
// its transform values, which kind of column it is, its pointers' strides -- is set up once per CTA instead of once
// per tile (ncu on the tile-per-CTA kernel: 75 % of its 16 M warp instructions were such set-up and staging copies).
// Tiles that touch the ends of the whole array (clamped rows, quirk Q1) are staged by the threads themselves.
constexpr int STREAM_THREADS = 256;

template <int OUT_KIND>
__global__ void __launch_bounds__(STREAM_THREADS) splice_stream_kernel(const SpliceParams p, const int n_tiles) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int halo_rows = p.tile_f + 2 * p.splice;
  const int x_floats = (halo_rows * p.dim + 31) & ~31;          // keep every region 128-byte aligned
  const int iv_floats = (p.tile_f * p.ivec_dim + 31) & ~31;
  const int t_floats = (p.spl_cols + 31) & ~31;
  float* s_add = reinterpret_cast<float*>(s_raw);
  float* s_mul = s_add + t_floats;
  float* s_x0 = s_mul + t_floats;            // two buffers of x_floats
  float* s_iv0 = s_x0 + 2 * x_floats;        // two buffers of iv_floats
  uint64_t* full = reinterpret_cast<uint64_t*>(s_iv0 + 2 * iv_floats);  // [2]
  const int tid = static_cast<int>(threadIdx.x);
  const bool has_ft = p.add_shift != nullptr;

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  if (has_ft) {
    for (int i = tid; i < p.spl_cols; i += STREAM_THREADS) {
      s_add[i] = __ldg(p.add_shift + i);
      s_mul[i] = __ldg(p.rescale + i);
    }
  }
  __syncthreads();

  // ---- per-thread column set-up (once per CTA)
  const int vpr = static_cast<int>(p.ldo >> 3);   // 16-byte chunks per output row (<= STREAM_THREADS, checked by the host)
  const int lanes_r = STREAM_THREADS / vpr;       // rows per sweep of the block
  const bool active = tid < vpr * lanes_r;
  const int cc = tid % vpr;
  const int r_first = tid / vpr;
  const int c = cc << 3;
  // 0: spliced + transformed columns, 1: i-vector columns, 2: i-vector tail and zero padding (element-wise)
  const int kind = c + 8 <= p.spl_cols ? 0 : (c + 8 <= p.spl_cols + p.ivec_dim ? 1 : 2);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, m0 = make_float4(1.f, 1.f, 1.f, 1.f), m1 = m0;
  if (kind == 0 && has_ft) {
    a0 = *reinterpret_cast<const float4*>(s_add + c);
    a1 = *reinterpret_cast<const float4*>(s_add + c + 4);
    m0 = *reinterpret_cast<const float4*>(s_mul + c);
    m1 = *reinterpret_cast<const float4*>(s_mul + c + 4);
  }
  const long long row_step = static_cast<long long>(lanes_r) * p.ldo;
  const int src_off = kind == 0 ? r_first * p.dim + c : r_first * p.ivec_dim + (c - p.spl_cols);
  const int src_step = lanes_r * (kind == 0 ? p.dim : p.ivec_dim);

  auto tile_is_edge = [&](long long tf0, int tn) {
    return tf0 - p.splice < 0 || tf0 + tn + p.splice > p.n_total;
  };
  // one thread: both spans of tile `t` into buffer `b`
  auto issue = [&](int t, int b) {
    const long long tf0 = p.f0 + static_cast<long long>(t) * p.tile_f;
    const int tn = static_cast<int>(min(static_cast<long long>(p.tile_f), p.f1 - tf0));
    if (tile_is_edge(tf0, tn)) return;  // staged synchronously by the whole block when its turn comes
    const uint32_t bx = static_cast<uint32_t>((tn + 2 * p.splice) * p.dim) * 4u;
    const uint32_t bi = static_cast<uint32_t>(tn * p.ivec_dim) * 4u;
    fence_proxy_async_smem();  // the buffer's previous contents were read (or, for an edge tile, written) by threads
    mbar_expect_tx(&full[b], bx + bi);
    bulk_load_1d(s_x0 + b * x_floats, p.x + (tf0 - p.splice - p.x_row0) * p.dim, bx, &full[b]);
    if (bi) bulk_load_1d(s_iv0 + b * iv_floats, p.ivec + (tf0 - p.f0) * p.ivec_dim, bi, &full[b]);
  };

  uint32_t phase = 0u;  // bit b: parity of buffer b's next fill
  int b = 0;
  if (tid == 0 && static_cast<int>(blockIdx.x) < n_tiles) issue(static_cast<int>(blockIdx.x), 0);
  for (int t = static_cast<int>(blockIdx.x); t < n_tiles; t += static_cast<int>(gridDim.x), b ^= 1) {
    const long long tf0 = p.f0 + static_cast<long long>(t) * p.tile_f;
    const int tn = static_cast<int>(min(static_cast<long long>(p.tile_f), p.f1 - tf0));
    const int t_next = t + static_cast<int>(gridDim.x);
    if (tid == 0 && t_next < n_tiles) issue(t_next, b ^ 1);  // buffer b ^ 1 was released by the barrier below
    float* xb = s_x0 + b * x_floats;
    float* ivb = s_iv0 + b * iv_floats;
    if (tile_is_edge(tf0, tn)) {
      const int vx = p.dim >> 2;
      for (int i = tid; i < (tn + 2 * p.splice) * vx; i += STREAM_THREADS) {
        const int row = i / vx, v = i - row * vx;
        long long g = tf0 - p.splice + row;
        g = g < 0 ? 0 : (g >= p.n_total ? p.n_total - 1 : g);
        reinterpret_cast<float4*>(xb)[i] = __ldg(reinterpret_cast<const float4*>(p.x + (g - p.x_row0) * p.dim) + v);
      }
      const float4* iv4 = reinterpret_cast<const float4*>(p.ivec + (tf0 - p.f0) * p.ivec_dim);
      for (int i = tid; i < (tn * p.ivec_dim) >> 2; i += STREAM_THREADS) reinterpret_cast<float4*>(ivb)[i] = __ldg(iv4 + i);
      __syncthreads();
    } else {
      mbar_wait(&full[b], (phase >> b) & 1u);
      phase ^= 1u << b;
    }

    if (active) {
      __nv_bfloat16* dh = static_cast<__nv_bfloat16*>(p.out_hi) + (tf0 - p.f0 + r_first) * p.ldo + c;
      __nv_bfloat16* dl = OUT_KIND == NNAM_OUT_BF16_SPLIT
                              ? static_cast<__nv_bfloat16*>(p.out_lo) + (tf0 - p.f0 + r_first) * p.ldo + c
                              : nullptr;
      auto emit = [&](const float4& v0, const float4& v1) {
        *reinterpret_cast<uint4*>(dh) = make_uint4(pack_hi<OUT_KIND>(v0.x, v0.y), pack_hi<OUT_KIND>(v0.z, v0.w),
                                                   pack_hi<OUT_KIND>(v1.x, v1.y), pack_hi<OUT_KIND>(v1.z, v1.w));
        dh += row_step;
        if (OUT_KIND == NNAM_OUT_BF16_SPLIT) {
          *reinterpret_cast<uint4*>(dl) = make_uint4(
              pack_bf16x2(v0.x - bf16_round(v0.x), v0.y - bf16_round(v0.y)),
              pack_bf16x2(v0.z - bf16_round(v0.z), v0.w - bf16_round(v0.w)),
              pack_bf16x2(v1.x - bf16_round(v1.x), v1.y - bf16_round(v1.y)),
              pack_bf16x2(v1.z - bf16_round(v1.z), v1.w - bf16_round(v1.w)));
          dl += row_step;
        }
      };
      if (kind == 0) {
        const float* src = xb + src_off;
        if (has_ft) {
#pragma unroll 2
          for (int r = r_first; r < tn; r += lanes_r, src += src_step) {
            float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
            v0.x = __fmul_rn(__fadd_rn(v0.x, a0.x), m0.x);
            v0.y = __fmul_rn(__fadd_rn(v0.y, a0.y), m0.y);
            v0.z = __fmul_rn(__fadd_rn(v0.z, a0.z), m0.z);
            v0.w = __fmul_rn(__fadd_rn(v0.w, a0.w), m0.w);
            v1.x = __fmul_rn(__fadd_rn(v1.x, a1.x), m1.x);
            v1.y = __fmul_rn(__fadd_rn(v1.y, a1.y), m1.y);
            v1.z = __fmul_rn(__fadd_rn(v1.z, a1.z), m1.z);
            v1.w = __fmul_rn(__fadd_rn(v1.w, a1.w), m1.w);
            emit(v0, v1);
          }
        } else {
#pragma unroll 2
          for (int r = r_first; r < tn; r += lanes_r, src += src_step)
            emit(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4));
        }
      } else if (kind == 1) {
        const float* src = ivb + src_off;
#pragma unroll 2
        for (int r = r_first; r < tn; r += lanes_r, src += src_step)
          emit(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4));
      } else {
        const int n_iv = p.spl_cols + p.ivec_dim - c;  // i-vector elements left in this chunk (may be <= 0: all padding)
        const float* src = ivb + src_off;
        if (n_iv == 4 || n_iv <= 0) {  // ivec_dim % 4 == 0: the tail is one aligned float4 or nothing
          const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int r = r_first; r < tn; r += lanes_r, src += src_step)
            emit(n_iv == 4 ? *reinterpret_cast<const float4*>(src) : zero, zero);
        } else
        for (int r = r_first; r < tn; r += lanes_r, src += src_step) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = j < n_iv ? src[j] : 0.0f;
          emit(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
        }
      }
    }
    __syncthreads();  // buffer b may be refilled (by the issue at the top of the next-but-one iteration)
  }
}

static size_t splice_stream_smem(const SpliceParams& p, int tile_f) {
  const size_t x_floats = (static_cast<size_t>(tile_f + 2 * p.splice) * p.dim + 31) & ~static_cast<size_t>(31);
  const size_t iv_floats = (static_cast<size_t>(tile_f) * p.ivec_dim + 31) & ~static_cast<size_t>(31);
  const size_t t_floats = (static_cast<size_t>(p.spl_cols) + 31) & ~static_cast<size_t>(31);
  return (2 * t_floats + 2 * x_floats + 2 * iv_floats) * sizeof(float) + 16;
}

// true if the streaming kernel applies (the caller has checked the 16-byte alignment of x / ivec / out)
static bool splice_stream_applies(const SpliceParams& p) {
  if (getenv("NNAM_SPLICE_STREAM") && getenv("NNAM_SPLICE_STREAM")[0] == '0') return false;
  if (p.dim % 4 || p.ivec_dim % 4 || p.spl_cols % 8 || p.ldo % 8) return false;
  const long long vpr = p.ldo >> 3;
  return vpr >= 1 && vpr <= STREAM_THREADS;
}

template <int OUT_KIND>
static int launch_splice_stream(const SpliceParams& p_in, cudaStream_t stream) {
  SpliceParams p = p_in;
  const long long frames = p.f1 - p.f0;
  // eight sweeps of the block over the tile's rows (cfg2: 68 chunk columns -> 3 rows per sweep -> 24 frames; measured
  // 23.6 us per 65,536 frames against 25.5 / 26.7 / 27.6 us for 32 / 48 / 64 frames)
  const int lanes_r = STREAM_THREADS / static_cast<int>(p.ldo >> 3);
  p.tile_f = lanes_r * 8 > 64 ? 64 : (lanes_r * 8 < 8 ? 8 : lanes_r * 8);
  if (const char* v = getenv("NNAM_SPLICE_TILE")) {  // tuning aid
    const int t = atoi(v);
    if (t >= 8 && t <= 128) p.tile_f = t;
  }
  const size_t smem = splice_stream_smem(p, p.tile_f);
  if (smem > 100 * 1024) return -1;  // caller falls back to the tile-per-CTA kernel
  // per-device attribute: set on every launch (cheap) rather than behind a process-wide flag, which left every
  // device but the first without the opt-in when predict(gpu=[0, 1, ...]) runs one host thread per GPU
  if (cudaFuncSetAttribute(splice_stream_kernel<OUT_KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) !=
      cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, splice_stream_kernel<OUT_KIND>, STREAM_THREADS, smem) !=
          cudaSuccess || occ <= 0) {
    cudaGetLastError();
    return -1;
  }
  const long long n_tiles = (frames + p.tile_f - 1) / p.tile_f;
  if (n_tiles > 0x7fffffffLL) return -1;
  const long long slots = static_cast<long long>(occ) * sm_count();
  const unsigned grid = static_cast<unsigned>(n_tiles < slots ? n_tiles : slots);
  splice_stream_kernel<OUT_KIND><<<grid, STREAM_THREADS, smem, stream>>>(p, static_cast<int>(n_tiles));
  return check_launch("splice_stream_kernel");
}

static size_t splice_smem_bytes(const SpliceParams& p, bool vec, int tile_f) {
  return (static_cast<size_t>(tile_f + 2 * p.splice) * p.dim + 2 * static_cast<size_t>(p.spl_cols) +
          (vec ? static_cast<size_t>(tile_f) * p.ivec_dim : 0)) *
         sizeof(float);
}

template <int OUT_KIND>
static int launch_splice(const SpliceParams& p_in, bool vec, cudaStream_t stream) {
  SpliceParams p = p_in;
  const long long frames = p.f1 - p.f0;
  // Frames per CTA: the kernel is one pass over HBM, so a partly filled last wave costs its full duration (64-frame
  // tiles on a 65,536-frame chunk: 1024 CTAs on 740 resident slots = 1.38 waves, 31 % of the second one idle).  Pick
  // the tile that fills whole waves best, preferring larger tiles (less halo and transform staging per frame).
  int best_tile = SPLICE_TILE_F;
  double best_score = -1.0;
  for (int t = SPLICE_TILE_F; t >= 16; t -= 4) {
    const size_t sm = splice_smem_bytes(p, vec, t);
    if (sm > 200 * 1024) continue;
    int occ = 0;
    cudaError_t eo = vec ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, splice_transform_kernel<OUT_KIND, true>,
                                                                         SPLICE_THREADS, sm)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, splice_transform_kernel<OUT_KIND, false>,
                                                                         SPLICE_THREADS, sm);
    if (eo != cudaSuccess || occ <= 0) {
      cudaGetLastError();
      continue;
    }
    const long long slots = static_cast<long long>(occ) * sm_count();
    const long long tiles = (frames + t - 1) / t;
    const long long waves = (tiles + slots - 1) / slots;
    const double eff = static_cast<double>(tiles) / static_cast<double>(waves * slots);
    const double score = eff * t / (t + 4.0);
    if (score > best_score) {
      best_score = score;
      best_tile = t;
    }
  }
  p.tile_f = best_tile;
  const long long blocks = (frames + p.tile_f - 1) / p.tile_f;
  if (blocks > 0x7fffffffLL) return set_error(NNAM_ERR_ARG, "splice: too many frames for one launch");
  const size_t smem = splice_smem_bytes(p, vec, p.tile_f);
  if (smem > 200 * 1024) return set_error(NNAM_ERR_UNSUPPORTED, "splice: window too large for shared memory");
  if (smem > 48 * 1024) {
    cudaError_t e;
    if (vec)
      e = cudaFuncSetAttribute(splice_transform_kernel<OUT_KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    else
      e = cudaFuncSetAttribute(splice_transform_kernel<OUT_KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "splice: cudaFuncSetAttribute");
  }
  if (vec)
    splice_transform_kernel<OUT_KIND, true><<<static_cast<unsigned>(blocks), SPLICE_THREADS, smem, stream>>>(p);
  else
    splice_transform_kernel<OUT_KIND, false><<<static_cast<unsigned>(blocks), SPLICE_THREADS, smem, stream>>>(p);
  return check_launch("splice_transform_kernel");
}

int splice_transform(const float* x, long long x_row0, long long x_rows, long long n_total, int dim, int splice,
                     const float* add_shift, const float* rescale, const float* ivec, int ivec_dim, long long f0,
                     long long f1, void* out_hi, void* out_lo, long long ldo, int out_kind, cudaStream_t stream) {
  if (f1 < f0 || f0 < 0 || f1 > n_total) return set_error(NNAM_ERR_ARG, "splice: bad frame range [%lld,%lld)", f0, f1);
  if (f1 == f0) return NNAM_OK;  // empty input: nothing to do
  if (dim <= 0 || splice < 0 || ivec_dim < 0) return set_error(NNAM_ERR_ARG, "splice: bad dims");
  if ((add_shift == nullptr) != (rescale == nullptr))
    return set_error(NNAM_ERR_ARG, "splice: add_shift and rescale must be given together");
  if (ivec_dim > 0 && ivec == nullptr) return set_error(NNAM_ERR_ARG, "splice: ivec_dim > 0 but ivec is NULL");
  const long long lo = f0 - splice < 0 ? 0 : f0 - splice;
  const long long hi = f1 + splice > n_total ? n_total : f1 + splice;
  if (x_row0 > lo || x_row0 + x_rows < hi)
    return set_error(NNAM_ERR_ARG, "splice: x rows [%lld,%lld) do not cover the halo [%lld,%lld)", x_row0,
                     x_row0 + x_rows, lo, hi);
  SpliceParams p;
  p.x = x;
  p.x_row0 = x_row0;
  p.n_total = n_total;
  p.dim = dim;
  p.splice = splice;
  p.winlen = 2 * splice + 1;
  p.add_shift = add_shift;
  p.rescale = rescale;
  p.ivec = ivec;
  p.ivec_dim = ivec_dim;
  p.f0 = f0;
  p.f1 = f1;
  p.out_hi = out_hi;
  p.out_lo = out_lo;
  p.ldo = ldo;
  p.spl_cols = p.winlen * dim;
  const int cols = p.spl_cols + ivec_dim;
  if (ldo < cols) return set_error(NNAM_ERR_ARG, "splice: ldo %lld < %d output columns", ldo, cols);
  const bool aligned_in = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                          (ivec_dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(ivec) & 15) == 0);
  switch (out_kind) {
    case NNAM_OUT_F32: {
      const bool vec = aligned_in && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_hi) & 15) == 0);
      return launch_splice<NNAM_OUT_F32>(p, vec, stream);
    }
    case NNAM_OUT_BF16:
    case NNAM_OUT_F16:
    case NNAM_OUT_BF16_SPLIT: {
      if (ldo % 8 || (reinterpret_cast<uintptr_t>(out_hi) & 15))
        return set_error(NNAM_ERR_ARG, "splice: 16-bit output needs ldo %% 8 == 0 and a 16-byte aligned buffer");
      if (out_kind == NNAM_OUT_BF16_SPLIT) {
        if (!out_lo || (reinterpret_cast<uintptr_t>(out_lo) & 15))
          return set_error(NNAM_ERR_ARG, "splice: split output needs an aligned out_lo");
        if (aligned_in && splice_stream_applies(p)) {
          const int rc = launch_splice_stream<NNAM_OUT_BF16_SPLIT>(p, stream);
          if (rc >= 0) return rc;
        }
        return launch_splice<NNAM_OUT_BF16_SPLIT>(p, aligned_in, stream);
      }
      if (out_kind == NNAM_OUT_F16) {
        if (aligned_in && splice_stream_applies(p)) {
          const int rc = launch_splice_stream<NNAM_OUT_F16>(p, stream);
          if (rc >= 0) return rc;
        }
        return launch_splice<NNAM_OUT_F16>(p, aligned_in, stream);
      }
      if (aligned_in && splice_stream_applies(p)) {
        const int rc = launch_splice_stream<NNAM_OUT_BF16>(p, stream);
        if (rc >= 0) return rc;
      }
      return launch_splice<NNAM_OUT_BF16>(p, aligned_in, stream);
    }
    default:
      return set_error(NNAM_ERR_ARG, "splice: unknown out_kind %d", out_kind);
  }
}

// ---------------------------------------------------------------------------------------------------
// Row gather + transform (+ i-vector) for the recurrent path: out[r] = (x[map[r]] + add) * mul ++ ivec[map[r]].
// 8 output elements per thread, 16-byte stores (bf16) / two float4 stores (fp32).
template <int OUT_KIND>
__global__ void gather_transform_kernel(const float* __restrict__ x, int dim, const float* __restrict__ add,
                                        const float* __restrict__ mul, const float* __restrict__ ivec, int ivec_dim,
                                        const int* __restrict__ row_map, long long n_rows, void* out_hi_v,
                                        void* out_lo_v, long long ldo, long long n_src) {
  const long long vpr = ldo >> 3;
  const long long total = n_rows * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vpr;
    const int c = static_cast<int>(i - r * vpr) << 3;
    const long long src = __ldg(row_map + r);
    const bool in_range = src >= 0 && src < n_src;  // a map entry outside the source gives a zero row, never a stray read
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = c + j;
      if (!in_range) {
        v[j] = 0.0f;
      } else if (cc < dim) {
        float t = __ldg(x + src * dim + cc);
        if (add != nullptr) t = __fmul_rn(__fadd_rn(t, __ldg(add + cc)), __ldg(mul + cc));
        v[j] = t;
      } else if (cc < dim + ivec_dim) {
        v[j] = __ldg(ivec + src * ivec_dim + (cc - dim));
      } else {
        v[j] = 0.0f;
      }
    }
    if (OUT_KIND == NNAM_OUT_F32) {
      float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out_hi_v) + r * ldo + c);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint32_t h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = pack_hi<OUT_KIND>(v[2 * j], v[2 * j + 1]);
      reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out_hi_v) + r * ldo)[c >> 3] =
          make_uint4(h[0], h[1], h[2], h[3]);
      if (OUT_KIND == NNAM_OUT_BF16_SPLIT) {
        uint32_t l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          l[j] = pack_bf16x2(v[2 * j] - bf16_round(v[2 * j]), v[2 * j + 1] - bf16_round(v[2 * j + 1]));
        reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out_lo_v) + r * ldo)[c >> 3] =
            make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

int gather_transform(const float* x, long long n_src, int dim, const float* add_shift, const float* rescale,
                     const float* ivec, int ivec_dim, const int* row_map, long long n_rows, void* out_hi, void* out_lo,
                     long long ldo, int out_kind, cudaStream_t stream) {
  if (n_rows < 0 || n_src <= 0 || dim <= 0 || ivec_dim < 0) return set_error(NNAM_ERR_ARG, "gather: bad shape");
  if (n_rows == 0) return NNAM_OK;
  if (!row_map) return set_error(NNAM_ERR_ARG, "gather: row_map is NULL");
  if ((add_shift == nullptr) != (rescale == nullptr)) return set_error(NNAM_ERR_ARG, "gather: add_shift/rescale");
  if (ivec_dim > 0 && !ivec) return set_error(NNAM_ERR_ARG, "gather: ivec is NULL");
  if (ldo < dim + ivec_dim || ldo % 8) return set_error(NNAM_ERR_ARG, "gather: ldo must be >= columns and %% 8 == 0");
  if (reinterpret_cast<uintptr_t>(out_hi) & 15) return set_error(NNAM_ERR_ARG, "gather: out_hi alignment");
  const long long total = n_rows * (ldo >> 3);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  const unsigned g = static_cast<unsigned>(blocks);
  switch (out_kind) {
    case NNAM_OUT_F32:
      gather_transform_kernel<NNAM_OUT_F32><<<g, 256, 0, stream>>>(x, dim, add_shift, rescale, ivec, ivec_dim, row_map,
                                                                  n_rows, out_hi, out_lo, ldo, n_src);
      break;
    case NNAM_OUT_BF16:
      gather_transform_kernel<NNAM_OUT_BF16><<<g, 256, 0, stream>>>(x, dim, add_shift, rescale, ivec, ivec_dim,
                                                                   row_map, n_rows, out_hi, out_lo, ldo, n_src);
      break;
    case NNAM_OUT_F16:
      gather_transform_kernel<NNAM_OUT_F16><<<g, 256, 0, stream>>>(x, dim, add_shift, rescale, ivec, ivec_dim,
                                                                  row_map, n_rows, out_hi, out_lo, ldo, n_src);
      break;
    case NNAM_OUT_BF16_SPLIT:
      if (!out_lo || (reinterpret_cast<uintptr_t>(out_lo) & 15)) return set_error(NNAM_ERR_ARG, "gather: out_lo");
      gather_transform_kernel<NNAM_OUT_BF16_SPLIT><<<g, 256, 0, stream>>>(x, dim, add_shift, rescale, ivec, ivec_dim,
                                                                         row_map, n_rows, out_hi, out_lo, ldo, n_src);
      break;
    default:
      return set_error(NNAM_ERR_ARG, "gather: unknown out_kind %d", out_kind);
  }
  return check_launch("gather_transform_kernel");
}

// ---------------------------------------------------------------------------------------------------
// fp32 -> bf16 / bf16 split staging of weights and pre-spliced inputs (8 elements per thread).
template <int OUT_KIND>
__global__ void convert_f32_kernel(const float* __restrict__ src, long long rows, int cols, long long lds,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldd) {
  const long long vpr = ldd >> 3;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vpr;
    const int c = static_cast<int>(i - r * vpr) << 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? __ldg(src + r * lds + c + j) : 0.0f;
    uint32_t h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = pack_hi<OUT_KIND>(v[2 * j], v[2 * j + 1]);
    reinterpret_cast<uint4*>(hi + r * ldd)[c >> 3] = make_uint4(h[0], h[1], h[2], h[3]);
    if (OUT_KIND == NNAM_OUT_BF16_SPLIT) {
      uint32_t l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        l[j] = pack_bf16x2(v[2 * j] - bf16_round(v[2 * j]), v[2 * j + 1] - bf16_round(v[2 * j + 1]));
      reinterpret_cast<uint4*>(lo + r * ldd)[c >> 3] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

int convert_f32(const float* src, long long rows, int cols, long long lds, void* dst_hi, void* dst_lo, long long ldd,
                int out_kind, cudaStream_t stream) {
  if (rows < 0 || cols <= 0 || lds < cols || ldd < cols) return set_error(NNAM_ERR_ARG, "convert: bad shape");
  if (rows == 0) return NNAM_OK;
  if (ldd % 8 || (reinterpret_cast<uintptr_t>(dst_hi) & 15)) return set_error(NNAM_ERR_ARG, "convert: ldd %% 8 / alignment");
  const long long total = rows * (ldd >> 3);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (out_kind == NNAM_OUT_BF16_SPLIT) {
    if (!dst_lo || (reinterpret_cast<uintptr_t>(dst_lo) & 15)) return set_error(NNAM_ERR_ARG, "convert: dst_lo");
    convert_f32_kernel<NNAM_OUT_BF16_SPLIT><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        src, rows, cols, lds, static_cast<__nv_bfloat16*>(dst_hi), static_cast<__nv_bfloat16*>(dst_lo), ldd);
  } else if (out_kind == NNAM_OUT_BF16) {
    convert_f32_kernel<NNAM_OUT_BF16><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        src, rows, cols, lds, static_cast<__nv_bfloat16*>(dst_hi), nullptr, ldd);
  } else if (out_kind == NNAM_OUT_F16) {
    convert_f32_kernel<NNAM_OUT_F16><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        src, rows, cols, lds, static_cast<__nv_bfloat16*>(dst_hi), nullptr, ldd);
  } else {
    return set_error(NNAM_ERR_ARG, "convert: out_kind must be BF16, F16 or BF16_SPLIT");
  }
  return check_launch("convert_f32_kernel");
}

}  // namespace nnam
