// K3 -- persistent recurrence kernel (LSTM family; GRU family below) for a whole layer of a whole shard.
//
// Reference semantics: one time step of L.LSTM / F.lstm per Python loop iteration
// (scripts/common/chainer_networks.py:44-62 driven by predict_folds.py:49-61 and
// evaluateModelForTest.py:67-80): gates = upward(x) + lateral(h); the 4H axis is interleaved
// unit-major / gate-minor [a, i, f, o]; c = tanh(a) s(i) + s(f) c; h = s(o) tanh(c); s(x) = tanh(x/2)/2 + 1/2.
// Here upward(x)+b for ALL frames is one K2 GEMM (gx), and this kernel runs the sequential part.
//
// Decomposition
//   * utterances are sorted by length and cut into batches of NB; a batch is stored time-major ("packed"):
//     row(t, u) = row0 + base[t] + u, active utterances at time t are the prefix u < base[t+1]-base[t];
//   * a GROUP of G = 4H / M_ROWS co-resident CTAs owns one (batch, direction) work item at a time; CTA r keeps the
//     M_ROWS lateral-weight rows [r*M_ROWS, (r+1)*M_ROWS) (= M_ROWS/4 whole units, thanks to Chainer's interleaved
//     layout) resident in shared memory as the A operand of tcgen05.mma (K-major, SWIZZLE_128B, loaded by TMA);
//   * per step: h_{t-1} (n_t x H, bf16 [hi, lo]) is read from the layer's own output buffer in L2 into a swizzled
//     smem tile (the B operand), D[M_ROWS x NB] = W_slice . h^T accumulates in TMEM (3 MMAs passes in bf16x3 mode),
//     each thread owns one gate row (TMEM lane), adds gx, applies the nonlinearity, the 4 gates of a unit are
//     exchanged inside a lane quad with a 4x4 shuffle transpose, the cell state lives in registers for the whole
//     utterance, and the new h slice is written (bf16 hi/lo) straight into the layer output rows -- which is what
//     the other CTAs of the group read next step.  A release/acquire counter per group orders the exchange.
//   * a bidirectional layer is ONE launch: forward and backward items are spread over the groups and run
//     concurrently; the backward direction walks t = len-1-s and addresses rows through base[] the same way.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

constexpr int RNN_THREADS = 256;

struct RnnTmaps {
  CUtensorMap w_hi[2];
  CUtensorMap w_lo[2];
};

struct RnnParams {
  int hidden;     // H
  int n_dirs;
  int n_groups;   // groups that have work
  int group_ctas; // G
  long long gx_ld, h_ld;
  const float* gx[2];     // per direction: (rows, gx_ld) fp32, gate-interleaved columns
  const float* u_bias[2]; // GRU family only
  __nv_bfloat16* h_hi;    // (rows, h_ld); direction d owns columns [d*H, (d+1)*H)
  __nv_bfloat16* h_lo;
  __nv_bfloat16* aux_hi;  // GRU reset-gate variants: r*h exchange buffer, same shape as h
  __nv_bfloat16* aux_lo;
  const int* item_batch;
  const int* item_dir;
  const int* group_item_start;  // n_groups + 1
  const int* batch_row0;
  const int* batch_steps;
  const int* batch_nutt;
  const int* batch_base_off;
  const int* base;     // concatenated per-batch prefix sums (steps + 1 entries each), relative to batch_row0
  const int* utt_len;  // steps per utterance, sorted order, batch b owns [b*NB, b*NB + nutt)
  const __nv_bfloat16* h0_hi;  // optional initial state (n_utts_sorted, H * n_dirs)
  const __nv_bfloat16* h0_lo;
  const float* c0;             // optional (n_utts_sorted, H * n_dirs)
  float* c_out;                // optional final cell state, same shape
  unsigned int* counters;      // one per group, zero on entry
  int gru_flags;
  long long* prof;             // optional: 8 cycle accumulators per CTA (thread 0), phases of a step
};

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FAST>
__device__ __forceinline__ float tanh_sel(float x) {
  return FAST ? tanh_fast(x) : tanhf(x);
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int RNN_BASE_SMEM = 2048;  // steps of a batch whose prefix-sum table is mirrored in shared memory

// 4x4 transpose inside a lane quad: on entry thread k of the quad holds x[i] = (gate k, utterance i); on exit it
// holds x[g] = (gate g, utterance k).
__device__ __forceinline__ void quad_transpose(float (&x)[4], int k) {
#pragma unroll
  for (int m = 1; m <= 2; m <<= 1) {
    const bool up = (k & m) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i & m) continue;
      const float send = up ? x[i] : x[i | m];
      const float recv = __shfl_xor_sync(0xffffffffu, send, m);
      if (up)
        x[i] = recv;
      else
        x[i | m] = recv;
    }
  }
}

// byte offset of 16-byte chunk `c16` (0..7) of row `r` inside one [rows x 128 B] SWIZZLE_128B K-major block
__device__ __forceinline__ uint32_t sw128_offset(int r, int c16) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

// KBT: compile-time number of 64-element k-blocks (H / 64) so that the MMA issue loop fully unrolls with
// immediate descriptor offsets; KBT = 0 selects the generic runtime-H variant.
template <int M_ROWS, int NB, int NSPLIT, bool FAST_TANH, int KBT>
__global__ void __launch_bounds__(RNN_THREADS, 1)
    lstm_seq_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int group = blockIdx.x / p.group_ctas;
  const int rank = blockIdx.x % p.group_ctas;
  if (group >= p.n_groups) return;

  const int H = KBT > 0 ? KBT * 64 : p.hidden;
  const int KB = KBT > 0 ? KBT : (H >> 6);  // 64-element k-blocks
  constexpr int W_BLOCK = M_ROWS * 128;
  constexpr int H_BLOCK = NB * 128;
  constexpr int NBH = NB / 2;  // utterance slots per thread: warps 0-3 own slots [0, NBH), warps 4-7 [NBH, NB)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* w_hi_s = smem;
  uint8_t* w_lo_s = w_hi_s + (NSPLIT == 3 ? KB * W_BLOCK : 0);
  uint8_t* h_hi_s = w_lo_s + KB * W_BLOCK;
  uint8_t* h_lo_s = h_hi_s + (NSPLIT == 3 ? KB * H_BLOCK : 0);
  uint8_t* tail = h_lo_s + KB * H_BLOCK;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  int* s_len = reinterpret_cast<int*>(tmem_slot + 2);  // NB ints
  int* s_base = s_len + NB;                            // RNN_BASE_SMEM + 1 ints

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int quarter = warp & 3;  // TMEM lane quarter this warp may read
  const int half = warp >> 2;    // which half of the utterance slots (TMEM columns) this warp owns
  const int u_lo = half * NBH;
  constexpr int TMEM_COLS = NB < 32 ? 32 : NB;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // TMEM lane <-> gate row of this CTA's slice.  M_ROWS = 128: lane index = row.  M_ROWS = 64: rows 16q..16q+15
  // sit in lanes 32q..32q+15 (the upper half of every subpartition is unused).
  const bool row_valid = (M_ROWS == 128) || (lane < 16);
  const int my_row = (M_ROWS == 128) ? (quarter * 32 + lane) : (quarter * 16 + (lane & 15));
  const int gate = my_row & 3;                           // a, i, f, o
  const int unit = rank * (M_ROWS / 4) + (my_row >> 2);  // hidden unit index in [0, H)
  const int gate_col = rank * M_ROWS + my_row;           // column of gx / row of W_lat
  const uint32_t tmem_lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + u_lo;
  const uint32_t idesc = make_idesc_bf16_f32(M_ROWS, NB);

  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#define PROF_START() do { if (p.prof != nullptr && tid == 0) prof_t = clock64(); } while (0)
#define PROF_MARK(i) do { if (p.prof != nullptr && tid == 0) { const long long now = clock64(); prof_acc[i] += now - prof_t; prof_t = now; } } while (0)
  unsigned int steps_done = 0;
  uint32_t w_phase = 0, mma_phase = 0;
  int cur_dir = -1;
  unsigned int* counter = p.counters + group;

  for (int it = p.group_item_start[group]; it < p.group_item_start[group + 1]; ++it) {
    const int b = p.item_batch[it];
    const int d = p.item_dir[it];
    const bool bwd = d == 1;
    if (d != cur_dir) {  // (re)load this CTA's slice of the lateral weights
      __syncthreads();
      if (tid == 0) {
        mbar_expect_tx(bar_w, static_cast<uint32_t>(KB * W_BLOCK * (NSPLIT == 3 ? 2 : 1)));
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_2d(w_hi_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * M_ROWS);
          if (NSPLIT == 3) tma_load_2d(w_lo_s + kb * W_BLOCK, &tmaps.w_lo[d], bar_w, kb * 64, rank * M_ROWS);
        }
      }
      mbar_wait(bar_w, w_phase);
      w_phase ^= 1;
      cur_dir = d;
    }
    const long long row0 = p.batch_row0[b];
    const int T = p.batch_steps[b];
    const int nutt = p.batch_nutt[b];
    const int* base = p.base + p.batch_base_off[b];
    const int* len = p.utt_len + b * NB;
    const float* gx = p.gx[d] + gate_col;
    const int h_col0 = d * H;
    __syncthreads();  // previous item's readers of s_len / s_base are done
    if (tid < NB) s_len[tid] = tid < nutt ? len[tid] : 0;
    const bool base_in_smem = T <= RNN_BASE_SMEM;
    if (base_in_smem)
      for (int i = tid; i <= T; i += RNN_THREADS) s_base[i] = __ldg(base + i);
    __syncthreads();
    const int* bp = base_in_smem ? s_base : base;  // prefix sums: shared-memory mirror when it fits

    // cell state of (utterance u_lo + 4m + gate, unit) lives in this thread for the whole item
    float c_reg[NBH / 4];
#pragma unroll
    for (int m = 0; m < NBH / 4; ++m) {
      const int u = u_lo + 4 * m + gate;
      c_reg[m] = (p.c0 != nullptr && row_valid && u < nutt)
                     ? p.c0[(static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0 + unit]
                     : 0.0f;
    }
    const bool has_h0 = p.h0_hi != nullptr;

    for (int s = 0; s < T; ++s) {
      PROF_START();
      const int base_s = bp[s];
      const int n_s = bp[s + 1] - base_s;  // active utterances (prefix of the batch), >= 1

      // ---- prefetch the input projection of my gate row for my utterance slots: independent loads
      // (slots beyond n_s re-read the last active row, so no load is predicated or dependent on another)
      float gxr[NBH];
#pragma unroll
      for (int j = 0; j < NBH; ++j) {
        const int u = u_lo + j;
        const int uu = u < n_s ? u : n_s - 1;
        const long long row = row0 + (bwd ? bp[s_len[uu] - 1 - s] : base_s) + uu;
        gxr[j] = __ldg(gx + row * p.gx_ld);
      }
      PROF_MARK(0);  // gx prefetch issue
      const bool do_mma = (s > 0) || has_h0;
      float acc[NBH];
      if (do_mma) {
        if (s > 0) {
          if (tid == 0) {
            const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
            while (ld_acquire_gpu(counter) < target) {
            }
          }
          __syncthreads();
        }
        PROF_MARK(1);  // wait for the group
        // ---- h_{s-1} rows of the active utterances -> swizzled smem (B operand): one warp per row, 16-byte
        // cp.async chunks, all in flight at once
        const int chunks_per_row = H >> 3;
        const uint32_t h_hi_sa = smem_u32(h_hi_s), h_lo_sa = smem_u32(h_lo_s);
        for (int u = warp; u < n_s; u += RNN_THREADS / 32) {
          long long off;
          const __nv_bfloat16 *src_hi, *src_lo;
          if (s == 0) {
            off = (static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0;
            src_hi = p.h0_hi;
            src_lo = p.h0_lo;
          } else {
            off = (row0 + bp[bwd ? (s_len[u] - s) : (s - 1)] + u) * p.h_ld + h_col0;
            src_hi = p.h_hi;
            src_lo = p.h_lo;
          }
          const uint32_t row_so = static_cast<uint32_t>((u >> 3) * 1024 + (u & 7) * 128);
          for (int c = lane; c < chunks_per_row; c += 32) {
            const uint32_t so = static_cast<uint32_t>((c >> 3) * H_BLOCK) + row_so + (((c & 7) ^ (u & 7)) << 4);
            cp_async_16(h_hi_sa + so, src_hi + off + c * 8);
            if (NSPLIT == 3) cp_async_16(h_lo_sa + so, src_lo + off + c * 8);
          }
        }
        cp_async_wait_all();
        fence_proxy_async_smem();
        __syncthreads();
        PROF_MARK(2);  // h load
        if (tid == 0) {
          tc_fence_after();
          const uint64_t wd_hi = make_sw128_kmajor_desc(smem_u32(w_hi_s));
          const uint64_t wd_lo = make_sw128_kmajor_desc(smem_u32(w_lo_s));
          const uint64_t hd_hi = make_sw128_kmajor_desc(smem_u32(h_hi_s));
          const uint64_t hd_lo = make_sw128_kmajor_desc(smem_u32(h_lo_s));
          uint32_t accum = 0;
#pragma unroll
          for (int pass = 0; pass < NSPLIT; ++pass) {
            const uint64_t wa = pass == 2 ? wd_lo : wd_hi;
            const uint64_t ha = pass == 1 ? hd_lo : hd_hi;
            if (KBT > 0) {
#pragma unroll
              for (int kb = 0; kb < KBT; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem_base, wa + ((kb * W_BLOCK + k * 32) >> 4), ha + ((kb * H_BLOCK + k * 32) >> 4), idesc,
                            accum);
                  accum = 1;
                }
              }
            } else {
              for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem_base, wa + ((kb * W_BLOCK + k * 32) >> 4), ha + ((kb * H_BLOCK + k * 32) >> 4), idesc,
                            accum);
                  accum = 1;
                }
              }
            }
          }
          umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        PROF_MARK(3);  // MMA issue + completion
#pragma unroll
        for (int c0 = 0; c0 < NBH; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(r[j]);
        }
        tc_fence_before();
        PROF_MARK(4);  // TMEM -> registers
      } else {
#pragma unroll
        for (int j = 0; j < NBH; ++j) acc[j] = 0.0f;
      }

      // ---- gates, quad transpose, cell update, write h
#pragma unroll
      for (int m = 0; m < NBH / 4; ++m) {
        if (u_lo + 4 * m >= n_s) break;  // warp-uniform
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = acc[4 * m + i] + gxr[4 * m + i];
          const float t = tanh_sel<FAST_TANH>(gate == 0 ? v : 0.5f * v);
          x[i] = gate == 0 ? t : fmaf(t, 0.5f, 0.5f);
        }
        quad_transpose(x, gate);  // x = {a, i, f, o} of utterance u = u_lo + 4m + gate
        const int u = u_lo + 4 * m + gate;
        const float c_new = fmaf(x[0], x[1], x[2] * c_reg[m]);
        const float h_new = x[3] * tanh_sel<FAST_TANH>(c_new);
        if (row_valid && u < n_s) {
          c_reg[m] = c_new;
          const int t_idx = bwd ? (s_len[u] - 1 - s) : s;
          const long long off = (row0 + bp[t_idx] + u) * p.h_ld + h_col0 + unit;
          const __nv_bfloat16 hb = __float2bfloat16_rn(h_new);
          p.h_hi[off] = hb;
          if (NSPLIT == 3) p.h_lo[off] = __float2bfloat16_rn(h_new - __bfloat162float(hb));
          if (p.c_out != nullptr && s == s_len[u] - 1)
            p.c_out[(static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0 + unit] = c_new;
        }
      }
      PROF_MARK(5);  // gates, cell update, h stores
      __threadfence();
      __syncthreads();
      if (tid == 0) red_release_gpu_add(counter, 1u);
      PROF_MARK(6);  // fence + publish
      ++steps_done;
    }
  }

  if (p.prof != nullptr && tid == 0) {
    prof_acc[7] = steps_done;
    for (int i = 0; i < 8; ++i) p.prof[blockIdx.x * 8 + i] = prof_acc[i];
  }
#undef PROF_START
#undef PROF_MARK
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================
// LSTM over a thread-block CLUSTER (bf16 mode, 128 gate rows per CTA): the group IS the cluster, and the hidden state is
// exchanged through distributed shared memory instead of L2.
//
// The global-memory variant above spends about half of a step (profiles/r01_k3_phase_cycles.md: ~4.1 k of 8.2 k
// cycles) on __threadfence + counter publish, the acquire spin and pulling the h rows back from L2.  Here every CTA
// PUSHES its freshly computed h slice (NB utterances x 32 units, bf16) straight into the B-operand tile of all G CTAs
// of the cluster with st.shared::cluster (already in the SWIZZLE_128B layout tcgen05.mma reads), and then arrives
// (release.cluster) on an mbarrier in every peer; a step starts as soon as the local mbarrier has collected G
// arrivals.  Tiles and mbarriers are double-buffered by the parity of the global step counter g: step g reads tile
// g & 1 and writes tile (g + 1) & 1.  That is WAR-safe without a second barrier because a peer can only be writing
// tile (g + 1) & 1 after it has collected all arrivals of step g - 1, and every arrival of step g - 1 was sent after
// its sender finished the MMA that read that tile.  The layer output rows (needed by the next layer's GEMM) are
// written from the same staging tile with 16-byte coalesced stores.
template <int NB, bool FAST_TANH, int KBT>
__global__ void __launch_bounds__(RNN_THREADS, 1)
    lstm_seq_cluster_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int M_ROWS = 128;
  constexpr int UNITS = M_ROWS / 4;  // hidden units per CTA
  const int G = p.group_ctas;        // == cluster size
  const int group = blockIdx.x / G;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool active = group < p.n_groups;  // uniform over the cluster

  const int H = KBT > 0 ? KBT * 64 : p.hidden;
  const int KB = KBT > 0 ? KBT : (H >> 6);
  constexpr int W_BLOCK = M_ROWS * 128;
  constexpr int H_BLOCK = NB * 128;
  constexpr int NBH = NB / 2;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* w_s = smem;                               // KB * W_BLOCK
  uint8_t* h_s = w_s + KB * W_BLOCK;                 // 2 tiles of KB * H_BLOCK
  uint8_t* stage = h_s + 2 * KB * H_BLOCK;           // NB rows x 64 B: this CTA's new h slice
  uint8_t* tail = stage + NB * UNITS * 2;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;
  uint64_t* bar_h = bar_mma + 1;                     // [2]: "h tiles of global step g are complete"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_h + 2);
  int* s_len = reinterpret_cast<int*>(tmem_slot + 2);
  int* s_base = s_len + NB;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int quarter = warp & 3;
  const int half = warp >> 2;
  const int u_lo = half * NBH;
  constexpr int TMEM_COLS = NB < 32 ? 32 : NB;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_init(&bar_h[0], static_cast<uint32_t>(G));
    mbar_init(&bar_h[1], static_cast<uint32_t>(G));
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // every CTA's mbarriers are initialised before anyone sends a remote arrive
  cluster_arrive_release();
  cluster_wait_acquire();

  const int my_row = quarter * 32 + lane;
  const int gate = my_row & 3;
  const int unit_local = my_row >> 2;
  const int gate_col = rank * M_ROWS + my_row;
  const uint32_t tmem_lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + u_lo;
  const uint32_t idesc = make_idesc_bf16_f32(M_ROWS, NB);
  const uint32_t h_sa = smem_u32(h_s);
  const uint32_t stage_sa = smem_u32(stage);
  const uint32_t tile_bytes = static_cast<uint32_t>(KB * H_BLOCK);
  // push role: thread -> (utterance slot, 16-byte chunk of the 64-byte slice), peers tid>>7, +2, +4, ...
  const int push_u = (tid & 127) >> 2;
  const int push_j = tid & 3;
  const uint32_t push_off = static_cast<uint32_t>((rank >> 1) * H_BLOCK) +
                            sw128_offset(push_u, (rank & 1) * 4 + push_j);

  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#define PROF_START() do { if (p.prof != nullptr && tid == 0) prof_t = clock64(); } while (0)
#define PROF_MARK(i) do { if (p.prof != nullptr && tid == 0) { const long long now = clock64(); prof_acc[i] += now - prof_t; prof_t = now; } } while (0)
  unsigned int g = 0;  // global step counter of this cluster (runs across work items)
  uint32_t w_phase = 0, mma_phase = 0;
  int cur_dir = -1;

  if (active) {
    for (int it = p.group_item_start[group]; it < p.group_item_start[group + 1]; ++it) {
      const int b = p.item_batch[it];
      const int d = p.item_dir[it];
      const bool bwd = d == 1;
      if (d != cur_dir) {
        __syncthreads();
        if (tid == 0) {
          mbar_expect_tx(bar_w, static_cast<uint32_t>(KB * W_BLOCK));
          for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * M_ROWS);
        }
        mbar_wait(bar_w, w_phase);
        w_phase ^= 1;
        cur_dir = d;
      }
      const long long row0 = p.batch_row0[b];
      const int T = p.batch_steps[b];
      const int nutt = p.batch_nutt[b];
      const int* base = p.base + p.batch_base_off[b];
      const int* len = p.utt_len + b * NB;
      const float* gx = p.gx[d] + gate_col;
      const int h_col0 = d * H;
      __syncthreads();
      if (tid < NB) s_len[tid] = tid < nutt ? len[tid] : 0;
      const bool base_in_smem = T <= RNN_BASE_SMEM;
      if (base_in_smem)
        for (int i = tid; i <= T; i += RNN_THREADS) s_base[i] = __ldg(base + i);
      __syncthreads();
      const int* bp = base_in_smem ? s_base : base;

      float c_reg[NBH / 4];
#pragma unroll
      for (int m = 0; m < NBH / 4; ++m) c_reg[m] = 0.0f;

      for (int s = 0; s < T; ++s, ++g) {
        PROF_START();
        const int base_s = bp[s];
        const int n_s = bp[s + 1] - base_s;
        float gxr[NBH];
#pragma unroll
        for (int j = 0; j < NBH; ++j) {
          const int u = u_lo + j;
          const int uu = u < n_s ? u : n_s - 1;
          const long long row = row0 + (bwd ? bp[s_len[uu] - 1 - s] : base_s) + uu;
          gxr[j] = __ldg(gx + row * p.gx_ld);
        }
        PROF_MARK(0);
        // all h slices of global step g-1 have landed in tile g&1 (also orders tile / staging reuse, see header)
        if (g > 0 && tid == 0) mbar_wait_cluster_acquire(&bar_h[(g - 1) & 1], ((g - 1) >> 1) & 1);
        __syncthreads();
        PROF_MARK(1);
        float acc[NBH];
        if (s > 0) {
          if (tid == 0) {
            tc_fence_after();
            const uint64_t wd = make_sw128_kmajor_desc(smem_u32(w_s));
            const uint64_t hd = make_sw128_kmajor_desc(h_sa + (g & 1) * tile_bytes);
            uint32_t accum = 0;
            if (KBT > 0) {
#pragma unroll
              for (int kb = 0; kb < KBT; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem_base, wd + ((kb * W_BLOCK + k * 32) >> 4), hd + ((kb * H_BLOCK + k * 32) >> 4), idesc,
                            accum);
                  accum = 1;
                }
            } else {
              for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem_base, wd + ((kb * W_BLOCK + k * 32) >> 4), hd + ((kb * H_BLOCK + k * 32) >> 4), idesc,
                            accum);
                  accum = 1;
                }
            }
            umma_commit(bar_mma);
          }
          mbar_wait(bar_mma, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          PROF_MARK(3);
#pragma unroll
          for (int c0 = 0; c0 < NBH; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_lane_addr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(r[j]);
          }
          tc_fence_before();
          PROF_MARK(4);
        } else {
#pragma unroll
          for (int j = 0; j < NBH; ++j) acc[j] = 0.0f;
        }

        // ---- gates, quad transpose, cell update; the new h goes to the staging tile (utterance-major, 64 B rows)
#pragma unroll
        for (int m = 0; m < NBH / 4; ++m) {
          if (u_lo + 4 * m >= n_s) break;
          float x[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v = acc[4 * m + i] + gxr[4 * m + i];
            const float t = tanh_sel<FAST_TANH>(gate == 0 ? v : 0.5f * v);
            x[i] = gate == 0 ? t : fmaf(t, 0.5f, 0.5f);
          }
          quad_transpose(x, gate);
          const int u = u_lo + 4 * m + gate;
          const float c_new = fmaf(x[0], x[1], x[2] * c_reg[m]);
          const float h_new = x[3] * tanh_sel<FAST_TANH>(c_new);
          if (u < n_s) {
            c_reg[m] = c_new;
            reinterpret_cast<__nv_bfloat16*>(stage)[u * UNITS + unit_local] = __float2bfloat16_rn(h_new);
          }
        }
        __syncthreads();
        PROF_MARK(5);
        // ---- push my slice into tile (g+1)&1 of every CTA of the cluster, and into the layer output rows
        if (push_u < n_s) {
          const uint4 v = *reinterpret_cast<const uint4*>(stage + push_u * (UNITS * 2) + push_j * 16);
          const uint32_t dst = h_sa + ((g + 1) & 1) * tile_bytes + push_off;
          for (int r = tid >> 7; r < G; r += 2) st_cluster_v4(mapa_shared(dst, static_cast<uint32_t>(r)), v);
          if (tid < 128) {
            const int t_idx = bwd ? (s_len[push_u] - 1 - s) : s;
            const long long off = (row0 + bp[t_idx] + push_u) * p.h_ld + h_col0 + rank * UNITS + push_j * 8;
            *reinterpret_cast<uint4*>(p.h_hi + off) = v;
          }
        }
        fence_proxy_async_all();  // my generic-proxy tile writes -> visible to the peers' tcgen05.mma (async proxy)
        __syncthreads();
        if (tid < G) mbar_arrive_cluster_release(mapa_shared(smem_u32(&bar_h[g & 1]), static_cast<uint32_t>(tid)));
        PROF_MARK(6);
      }
    }
  }
  if (p.prof != nullptr && tid == 0) {
    prof_acc[7] = g;
    for (int i = 0; i < 8; ++i) p.prof[blockIdx.x * 8 + i] = prof_acc[i];
  }
#undef PROF_START
#undef PROF_MARK
  // nobody may exit while peers can still write into its shared memory / arrive on its mbarriers
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================
// GRU family (scripts/common/MGRU.py:67-85; Chainer's L.GRU is the reset-gate + tanh instance).
//
// Same decomposition as the LSTM kernel.  The three hidden-side matrices are interleaved 4 rows per unit,
// row 4j+0 = U_z[j], 4j+1 = U_r[j] (zero without reset gate), 4j+2 = U[j], 4j+3 = zero padding, so a CTA slice again
// holds whole units and the quad transpose applies unchanged; gx holds [W_z x + b, W_r x + b, W x + b, 0] the same
// way and u_bias the hidden-side biases, which -- like every U term -- only exist from the second step on
// (MGRU.py:70-83: h is None on the first step).
//   no reset gate : one exchange per step:  z, hbar from [U_z; U] h
//   reset gate    : r = s(W_r x + U_r h) must be applied BEFORE the candidate matmul (MGRU.py:73-74), so a step has
//                   two phases: (1) [U_z; U_r] h -> r, publish r*h;  (2) U (r*h) -> hbar, publish h.
// h itself stays in fp32 registers for the interpolation h' = z*hbar + (1-z)*h.
template <int M_ROWS, int NB, int NSPLIT, bool FAST_TANH, int KBT>
__global__ void __launch_bounds__(RNN_THREADS, 1)
    gru_seq_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int group = blockIdx.x / p.group_ctas;
  const int rank = blockIdx.x % p.group_ctas;
  if (group >= p.n_groups) return;

  const int H = KBT > 0 ? KBT * 64 : p.hidden;
  const int KB = KBT > 0 ? KBT : (H >> 6);
  constexpr int W_BLOCK = M_ROWS * 128;
  constexpr int H_BLOCK = NB * 128;
  constexpr int NBH = NB / 2;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* w_hi_s = smem;
  uint8_t* w_lo_s = w_hi_s + (NSPLIT == 3 ? KB * W_BLOCK : 0);
  uint8_t* h_hi_s = w_lo_s + KB * W_BLOCK;
  uint8_t* h_lo_s = h_hi_s + (NSPLIT == 3 ? KB * H_BLOCK : 0);
  uint8_t* tail = h_lo_s + KB * H_BLOCK;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  int* s_len = reinterpret_cast<int*>(tmem_slot + 2);
  int* s_base = s_len + NB;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int quarter = warp & 3;
  const int half = warp >> 2;
  const int u_lo = half * NBH;
  constexpr int TMEM_COLS = NB < 32 ? 32 : NB;
  const bool reset = (p.gru_flags & 1) != 0;
  const int act = (p.gru_flags >> 1) & 3;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool row_valid = (M_ROWS == 128) || (lane < 16);
  const int my_row = (M_ROWS == 128) ? (quarter * 32 + lane) : (quarter * 16 + (lane & 15));
  const int gate = my_row & 3;  // z, r, candidate, pad
  const int unit = rank * (M_ROWS / 4) + (my_row >> 2);
  const int gate_col = rank * M_ROWS + my_row;
  const uint32_t tmem_lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + u_lo;
  const uint32_t idesc = make_idesc_bf16_f32(M_ROWS, NB);

  unsigned int steps_done = 0;  // arrivals on the group counter so far
  uint32_t w_phase = 0, mma_phase = 0;
  int cur_dir = -1;
  unsigned int* counter = p.counters + group;
  const uint32_t h_hi_sa = smem_u32(h_hi_s), h_lo_sa = smem_u32(h_lo_s);

  auto group_wait = [&]() {
    if (tid == 0) {
      const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
      while (ld_acquire_gpu(counter) < target) {
      }
    }
    __syncthreads();
  };
  auto group_publish = [&]() {
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(counter, 1u);
    ++steps_done;
  };
  // D[M_ROWS x NB] = W_slice . tile^T, result of my (row, slots) in acc[]
  auto mma_tile = [&](float (&acc)[NBH]) {
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t wd_hi = make_sw128_kmajor_desc(smem_u32(w_hi_s));
      const uint64_t wd_lo = make_sw128_kmajor_desc(smem_u32(w_lo_s));
      const uint64_t hd_hi = make_sw128_kmajor_desc(h_hi_sa);
      const uint64_t hd_lo = make_sw128_kmajor_desc(h_lo_sa);
      uint32_t accum = 0;
#pragma unroll
      for (int pass = 0; pass < NSPLIT; ++pass) {
        const uint64_t wa = pass == 2 ? wd_lo : wd_hi;
        const uint64_t ha = pass == 1 ? hd_lo : hd_hi;
        if (KBT > 0) {
#pragma unroll
          for (int kb = 0; kb < KBT; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tmem_base, wa + ((kb * W_BLOCK + k * 32) >> 4), ha + ((kb * H_BLOCK + k * 32) >> 4), idesc, accum);
              accum = 1;
            }
        } else {
          for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tmem_base, wa + ((kb * W_BLOCK + k * 32) >> 4), ha + ((kb * H_BLOCK + k * 32) >> 4), idesc, accum);
              accum = 1;
            }
        }
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < NBH; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_lane_addr + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  };

  for (int it = p.group_item_start[group]; it < p.group_item_start[group + 1]; ++it) {
    const int b = p.item_batch[it];
    const int d = p.item_dir[it];
    const bool bwd = d == 1;
    if (d != cur_dir) {
      __syncthreads();
      if (tid == 0) {
        mbar_expect_tx(bar_w, static_cast<uint32_t>(KB * W_BLOCK * (NSPLIT == 3 ? 2 : 1)));
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_2d(w_hi_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * M_ROWS);
          if (NSPLIT == 3) tma_load_2d(w_lo_s + kb * W_BLOCK, &tmaps.w_lo[d], bar_w, kb * 64, rank * M_ROWS);
        }
      }
      mbar_wait(bar_w, w_phase);
      w_phase ^= 1;
      cur_dir = d;
    }
    const long long row0 = p.batch_row0[b];
    const int T = p.batch_steps[b];
    const int nutt = p.batch_nutt[b];
    const int* base = p.base + p.batch_base_off[b];
    const int* len = p.utt_len + b * NB;
    const float* gx = p.gx[d] + gate_col;
    const float ub = row_valid ? __ldg(p.u_bias[d] + gate_col) : 0.0f;
    const int h_col0 = d * H;
    __syncthreads();
    if (tid < NB) s_len[tid] = tid < nutt ? len[tid] : 0;
    const bool base_in_smem = T <= RNN_BASE_SMEM;
    if (base_in_smem)
      for (int i = tid; i <= T; i += RNN_THREADS) s_base[i] = __ldg(base + i);
    __syncthreads();
    const int* bp = base_in_smem ? s_base : base;
    const bool has_h0 = p.h0_hi != nullptr;

    // fp32 hidden state of (utterance u_lo + 4m + gate, unit)
    float h_reg[NBH / 4];
#pragma unroll
    for (int m = 0; m < NBH / 4; ++m) {
      const int u = u_lo + 4 * m + gate;
      h_reg[m] = 0.0f;
      if (has_h0 && row_valid && u < nutt) {
        const long long o = (static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0 + unit;
        h_reg[m] = __bfloat162float(p.h0_hi[o]) + (NSPLIT == 3 ? __bfloat162float(p.h0_lo[o]) : 0.0f);
      }
    }
    const int chunks_per_row = H >> 3;
    // stage rows of `src` (hi/lo) for the active utterances into the swizzled B-operand tile
    auto load_tile = [&](const __nv_bfloat16* src_hi, const __nv_bfloat16* src_lo, int n_act, int s, int mode) {
      // mode 0: initial state rows; 1: rows of the previous step; 2: rows of the current step
      for (int u = warp; u < n_act; u += RNN_THREADS / 32) {
        long long off;
        if (mode == 0) {
          off = (static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0;
        } else {
          const int t_idx = mode == 1 ? (bwd ? (s_len[u] - s) : (s - 1)) : (bwd ? (s_len[u] - 1 - s) : s);
          off = (row0 + bp[t_idx] + u) * p.h_ld + h_col0;
        }
        const uint32_t row_so = static_cast<uint32_t>((u >> 3) * 1024 + (u & 7) * 128);
        for (int c = lane; c < chunks_per_row; c += 32) {
          const uint32_t so = static_cast<uint32_t>((c >> 3) * H_BLOCK) + row_so + (((c & 7) ^ (u & 7)) << 4);
          cp_async_16(h_hi_sa + so, src_hi + off + c * 8);
          if (NSPLIT == 3) cp_async_16(h_lo_sa + so, src_lo + off + c * 8);
        }
      }
    };

    for (int s = 0; s < T; ++s) {
      const int base_s = bp[s];
      const int n_s = bp[s + 1] - base_s;
      float gxr[NBH];
#pragma unroll
      for (int j = 0; j < NBH; ++j) {
        const int u = u_lo + j;
        const int uu = u < n_s ? u : n_s - 1;
        const long long row = row0 + (bwd ? bp[s_len[uu] - 1 - s] : base_s) + uu;
        gxr[j] = __ldg(gx + row * p.gx_ld);
      }
      const bool have_h = (s > 0) || has_h0;
      float acc[NBH];
      if (have_h) {
        if (s > 0) group_wait();
        if (s == 0)
          load_tile(p.h0_hi, p.h0_lo, n_s, s, 0);
        else
          load_tile(p.h_hi, p.h_lo, n_s, s, 1);
        mma_tile(acc);
      } else {
#pragma unroll
        for (int j = 0; j < NBH; ++j) acc[j] = 0.0f;
      }
      const float ubs = have_h ? ub : 0.0f;  // U biases exist only when h does (MGRU.py:70-83)
      const bool two_phase = reset && have_h;

      float zp[NBH / 4], hp[NBH / 4];
#pragma unroll
      for (int m = 0; m < NBH / 4; ++m) {
        zp[m] = hp[m] = 0.0f;
        if (u_lo + 4 * m >= n_s) break;
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          x[i] = gxr[4 * m + i] + ubs + ((two_phase && gate == 2) ? 0.0f : acc[4 * m + i]);
        quad_transpose(x, gate);  // x = {z_pre, r_pre, cand_pre, pad} of utterance u_lo + 4m + gate
        zp[m] = x[0];
        hp[m] = x[2];
        if (two_phase) {
          const int u = u_lo + 4 * m + gate;
          const float r = fmaf(tanh_sel<FAST_TANH>(0.5f * x[1]), 0.5f, 0.5f);
          const float rh = r * h_reg[m];
          if (row_valid && u < n_s) {
            const int t_idx = bwd ? (s_len[u] - 1 - s) : s;
            const long long off = (row0 + bp[t_idx] + u) * p.h_ld + h_col0 + unit;
            const __nv_bfloat16 hb = __float2bfloat16_rn(rh);
            p.aux_hi[off] = hb;
            if (NSPLIT == 3) p.aux_lo[off] = __float2bfloat16_rn(rh - __bfloat162float(hb));
          }
        }
      }
      if (two_phase) {
        group_publish();  // r*h slices are out
        group_wait();
        load_tile(p.aux_hi, p.aux_lo, n_s, s, 2);
        mma_tile(acc);    // row 4j+2 now holds U (r*h)
      }
#pragma unroll
      for (int m = 0; m < NBH / 4; ++m) {
        if (u_lo + 4 * m >= n_s) break;
        float cand = hp[m];
        if (two_phase) {
          float y[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = gate == 2 ? acc[4 * m + i] : 0.0f;
          quad_transpose(y, gate);
          cand += y[2];
        }
        const int u = u_lo + 4 * m + gate;
        const float z = fmaf(tanh_sel<FAST_TANH>(0.5f * zp[m]), 0.5f, 0.5f);
        const float hb = act == NNAM_ACT_RELU ? fmaxf(cand, 0.0f)
                                              : (act == NNAM_ACT_SIGMOID
                                                     ? fmaf(tanh_sel<FAST_TANH>(0.5f * cand), 0.5f, 0.5f)
                                                     : (act == NNAM_ACT_TANH ? tanh_sel<FAST_TANH>(cand) : cand));
        const float h_new = have_h ? fmaf(z, hb, (1.0f - z) * h_reg[m]) : z * hb;
        if (row_valid && u < n_s) {
          h_reg[m] = h_new;
          const int t_idx = bwd ? (s_len[u] - 1 - s) : s;
          const long long off = (row0 + bp[t_idx] + u) * p.h_ld + h_col0 + unit;
          const __nv_bfloat16 hbf = __float2bfloat16_rn(h_new);
          p.h_hi[off] = hbf;
          if (NSPLIT == 3) p.h_lo[off] = __float2bfloat16_rn(h_new - __bfloat162float(hbf));
        }
      }
      group_publish();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
template <int M_ROWS, int NB, int NSPLIT, bool FAST, int KBT>
static int launch_rnn(int cell, const RnnTmaps& tm, const RnnParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = cell == NNAM_CELL_GRU ? gru_seq_kernel<M_ROWS, NB, NSPLIT, FAST, KBT>
                                    : lstm_seq_kernel<M_ROWS, NB, NSPLIT, FAST, KBT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute");
  void* args[] = {const_cast<RnnTmaps*>(&tm), const_cast<RnnParams*>(&p)};
  // cooperative launch: every CTA of a group must be co-resident (they wait on one another every step)
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(RNN_THREADS), args, smem, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaLaunchCooperativeKernel");
  return NNAM_OK;
}

// ---- cluster (DSMEM) variant: LSTM, bf16, 128 rows per CTA, no carried state
size_t rnn_cluster_smem_bytes(int nb, int hidden) {
  const size_t kb = hidden / 64;
  return kb * (128 * 128 + 2 * static_cast<size_t>(nb) * 128) + static_cast<size_t>(nb) * 64 + 64 + nb * 4 +
         (RNN_BASE_SMEM + 1) * 4 + 1024;
}

template <int NB, int KBT>
static int cluster_config(int G, size_t smem, cudaStream_t stream, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr,
                          int grid) {
  auto kern = lstm_seq_cluster_kernel<NB, true, KBT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute(smem)");
  if (G > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute(non-portable cluster)");
  }
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = dim3(grid);
  cfg->blockDim = dim3(RNN_THREADS);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  return NNAM_OK;
}

// how many clusters of G CTAs can be resident at once (0: this device / configuration cannot run the variant)
template <int NB, int KBT>
static int cluster_max_groups(int G, int hidden) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  const size_t smem = rnn_cluster_smem_bytes(NB, hidden);
  if (cluster_config<NB, KBT>(G, smem, nullptr, &cfg, attr, G) != NNAM_OK) return 0;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, lstm_seq_cluster_kernel<NB, true, KBT>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <int NB, int KBT>
static int launch_cluster(const RnnTmaps& tm, const RnnParams& p, int G, int hidden, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  const size_t smem = rnn_cluster_smem_bytes(NB, hidden);
  int rc = cluster_config<NB, KBT>(G, smem, stream, &cfg, attr, p.n_groups * G);
  if (rc) return rc;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_seq_cluster_kernel<NB, true, KBT>, tm, p);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaLaunchKernelEx(cluster)");
  return NNAM_OK;
}

// Measured on B200 (profiles/r01_k3_phase_cycles.md): the DSMEM all-gather of the 32 KB h tile costs ~3.5 k cycles per
// step (st.shared::cluster sustains ~20 B/clk per producer SM) and only 7 clusters of 16 CTAs are resident, so this
// variant is SLOWER than the L2 exchange (cfg3: 31.3 ms vs 23.2 ms per pass).  It stays opt-in (NNAM_RNN_CLUSTER=1).
static bool cluster_disabled() {
  const char* v = getenv("NNAM_RNN_CLUSTER");
  return v == nullptr || v[0] != '1';
}

// Resident clusters for the DSMEM variant of this configuration, 0 if it does not apply.
static int rnn_cluster_groups(int cell, int hidden, int batch, int nsplit) {
  if (cluster_disabled() || cell != NNAM_CELL_LSTM || nsplit != 1 || batch != 32) return 0;
  const int G = 4 * hidden / 128;
  if (G < 2 || G > 16 || rnn_cluster_smem_bytes(batch, hidden) > 227 * 1024) return 0;
  static int cached[64][17] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (cached[dev][G] == 0) {
    int n;
    if (hidden == 512) n = cluster_max_groups<32, 8>(G, hidden);
    else if (hidden == 256) n = cluster_max_groups<32, 4>(G, hidden);
    else n = cluster_max_groups<32, 0>(G, hidden);
    cached[dev][G] = n > 0 ? n : -1;
  }
  return cached[dev][G] > 0 ? cached[dev][G] : 0;
}

size_t rnn_smem_bytes(int m_rows, int nb, int hidden, int nsplit) {
  const size_t kb = hidden / 64;
  const size_t mult = nsplit == 3 ? 2 : 1;
  return kb * (static_cast<size_t>(m_rows) * 128 + static_cast<size_t>(nb) * 128) * mult + 64 + nb * 4 +
         (RNN_BASE_SMEM + 1) * 4 + 1024;
}

// Pick the CTA slice height: 128 gate rows when the weights fit in shared memory, else 64.
int rnn_pick_m_rows(int gate_rows_total, int hidden, int nb, int nsplit) {
  for (int m : {128, 64}) {
    if (gate_rows_total % m) continue;
    if (rnn_smem_bytes(m, nb, hidden, nsplit) <= 227 * 1024) return m;
  }
  return 0;
}

int rnn_seq(const NnamRnnDesc* d, cudaStream_t stream) {
  if (d == nullptr) return set_error(NNAM_ERR_ARG, "rnn: NULL descriptor");
  if (d->cell != NNAM_CELL_LSTM && d->cell != NNAM_CELL_GRU)
    return set_error(NNAM_ERR_UNSUPPORTED, "rnn: cell kind %d not implemented", d->cell);
  const int H = d->hidden;
  if (H <= 0 || H % 64) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: hidden size must be a multiple of 64 (got %d)", H);
  if (d->n_dirs != 1 && d->n_dirs != 2) return set_error(NNAM_ERR_ARG, "rnn: n_dirs must be 1 or 2");
  if (d->batch != 32 && d->batch != 64) return set_error(NNAM_ERR_ARG, "rnn: batch must be 32 or 64");
  if (d->nsplit != 1 && d->nsplit != 3) return set_error(NNAM_ERR_ARG, "rnn: nsplit must be 1 or 3");
  if (d->n_items <= 0) return NNAM_OK;
  if (d->nsplit == 3 && (d->h_lo == nullptr)) return set_error(NNAM_ERR_ARG, "rnn: bf16x3 needs h_lo");
  if (d->h_ld % 8 || d->w_ld % 8) return set_error(NNAM_ERR_ARG, "rnn: h_ld and w_ld must be multiples of 8");
  const int gate_rows = 4 * H;
  const int m_rows = rnn_pick_m_rows(gate_rows, H, d->batch, d->nsplit);
  if (!m_rows)
    return set_error(NNAM_ERR_UNSUPPORTED,
                     "rnn: lateral weights of H=%d do not fit in shared memory in this precision mode", H);
  const int G = gate_rows / m_rows;
  const int max_groups = sm_count() / G;
  if (max_groups < 1) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: H=%d needs %d co-resident CTAs", H, G);
  if (d->n_groups < 1 || d->n_groups > max_groups)
    return set_error(NNAM_ERR_ARG, "rnn: n_groups %d outside [1, %d]", d->n_groups, max_groups);
  // DSMEM variant: groups are independent clusters (no carried state through this path)
  const bool use_cluster = m_rows == 128 && !d->h0_hi && !d->c0 && !d->c_out &&
                           rnn_cluster_groups(d->cell, H, d->batch, d->nsplit) > 0;

  RnnTmaps tm;
  int rc;
  for (int k = 0; k < d->n_dirs; ++k) {
    if ((rc = encode_tmap_bf16_2d(&tm.w_hi[k], d->w_hi[k], H, gate_rows, d->w_ld, 64, m_rows))) return rc;
    if (d->nsplit == 3) {
      if (!d->w_lo[k]) return set_error(NNAM_ERR_ARG, "rnn: bf16x3 needs w_lo");
      if ((rc = encode_tmap_bf16_2d(&tm.w_lo[k], d->w_lo[k], H, gate_rows, d->w_ld, 64, m_rows))) return rc;
    } else {
      tm.w_lo[k] = tm.w_hi[k];
    }
  }
  if (d->n_dirs == 1) {
    tm.w_hi[1] = tm.w_hi[0];
    tm.w_lo[1] = tm.w_lo[0];
  }
  RnnParams p;
  p.hidden = H;
  p.n_dirs = d->n_dirs;
  p.n_groups = d->n_groups;
  p.group_ctas = G;
  p.gx_ld = d->gx_ld;
  p.h_ld = d->h_ld;
  for (int k = 0; k < 2; ++k) {
    p.gx[k] = d->gx[k];
    p.u_bias[k] = d->u_bias[k];
  }
  p.h_hi = static_cast<__nv_bfloat16*>(d->h_hi);
  p.h_lo = static_cast<__nv_bfloat16*>(d->h_lo);
  p.aux_hi = static_cast<__nv_bfloat16*>(d->aux_hi);
  p.aux_lo = static_cast<__nv_bfloat16*>(d->aux_lo);
  if (d->cell == NNAM_CELL_GRU) {
    for (int k = 0; k < d->n_dirs; ++k)
      if (!d->u_bias[k]) return set_error(NNAM_ERR_ARG, "rnn: GRU cells need u_bias");
    if ((d->flags & 1) && (!d->aux_hi || (d->nsplit == 3 && !d->aux_lo)))
      return set_error(NNAM_ERR_ARG, "rnn: reset-gate GRU needs the aux (r*h) exchange buffer");
  }
  p.item_batch = d->item_batch;
  p.item_dir = d->item_dir;
  p.group_item_start = d->group_item_start;
  p.batch_row0 = d->batch_row0;
  p.batch_steps = d->batch_steps;
  p.batch_nutt = d->batch_nutt;
  p.batch_base_off = d->batch_base_off;
  p.base = d->base;
  p.utt_len = d->utt_len;
  p.h0_hi = static_cast<const __nv_bfloat16*>(d->h0_hi);
  p.h0_lo = static_cast<const __nv_bfloat16*>(d->h0_lo);
  p.c0 = d->c0;
  p.c_out = d->c_out;
  p.counters = d->counters;
  p.gru_flags = d->flags;
  p.prof = static_cast<long long*>(d->debug_cycles);
  if (p.h0_hi && d->nsplit == 3 && !p.h0_lo) return set_error(NNAM_ERR_ARG, "rnn: bf16x3 needs h0_lo with h0_hi");

  if (use_cluster) {
    if (H == 512) return launch_cluster<32, 8>(tm, p, G, H, stream);
    if (H == 256) return launch_cluster<32, 4>(tm, p, G, H, stream);
    return launch_cluster<32, 0>(tm, p, G, H, stream);
  }
  cudaError_t e = cudaMemsetAsync(d->counters, 0, sizeof(unsigned int) * d->n_groups, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaMemsetAsync");
  const int grid = d->n_groups * G;
  const size_t smem = rnn_smem_bytes(m_rows, d->batch, H, d->nsplit);
  const bool fast = d->nsplit == 1;  // bf16 mode: MUFU tanh; fp32-accurate mode: tanhf

  const int kbt = H / 64;
#define NNAM_RNN_CASE(M, NBV, NS, F, KBTV)                                                                 \
  if (m_rows == M && d->batch == NBV && d->nsplit == NS && fast == F && (KBTV == 0 || KBTV == kbt))      \
  return launch_rnn<M, NBV, NS, F, KBTV>(d->cell, tm, p, grid, smem, stream)
  // tuned instances (compile-time H): the BASELINE geometries
  NNAM_RNN_CASE(128, 32, 1, true, 8);   // H = 512, bf16
  NNAM_RNN_CASE(128, 64, 1, true, 8);
  NNAM_RNN_CASE(64, 32, 3, false, 8);   // H = 512, bf16x3
  NNAM_RNN_CASE(64, 32, 1, true, 16);   // H = 1024, bf16
  NNAM_RNN_CASE(128, 32, 1, true, 4);   // H = 256
  NNAM_RNN_CASE(128, 32, 3, false, 4);
  // generic instances (runtime H)
  NNAM_RNN_CASE(128, 32, 1, true, 0);
  NNAM_RNN_CASE(128, 64, 1, true, 0);
  NNAM_RNN_CASE(64, 32, 1, true, 0);
  NNAM_RNN_CASE(64, 64, 1, true, 0);
  NNAM_RNN_CASE(128, 32, 3, false, 0);
  NNAM_RNN_CASE(128, 64, 3, false, 0);
  NNAM_RNN_CASE(64, 32, 3, false, 0);
  NNAM_RNN_CASE(64, 64, 3, false, 0);
#undef NNAM_RNN_CASE
  return set_error(NNAM_ERR_UNSUPPORTED, "rnn: no kernel instance for this configuration");
}

int rnn_plan(int cell, int hidden, int batch, int nsplit, int* group_ctas, int* max_groups, int* step_cycles) {
  if (cell != NNAM_CELL_LSTM && cell != NNAM_CELL_GRU)
    return set_error(NNAM_ERR_UNSUPPORTED, "rnn: cell kind %d not implemented", cell);
  if (hidden <= 0 || hidden % 64) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: hidden size must be a multiple of 64");
  const int m_rows = rnn_pick_m_rows(4 * hidden, hidden, batch, nsplit);
  if (!m_rows)
    return set_error(NNAM_ERR_UNSUPPORTED,
                     "rnn: lateral weights of H=%d do not fit in shared memory in this precision mode", hidden);
  *group_ctas = 4 * hidden / m_rows;
  *max_groups = sm_count() / *group_ctas;
  if (*max_groups < 1) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: H=%d needs %d co-resident CTAs", hidden, *group_ctas);
  // measured SM cycles per recurrence step (profiles/r01_k3_phase_cycles.md), used by the host to choose the batch width
  int cycles = batch == 64 ? 11900 : 8200;
  if (nsplit == 3) cycles = cycles * 3 / 2;
  const int cl = m_rows == 128 ? rnn_cluster_groups(cell, hidden, batch, nsplit) : 0;
  if (cl > 0) {
    if (cl < *max_groups) *max_groups = cl;
    cycles = 5000;
  }
  if (step_cycles) *step_cycles = cycles;
  return NNAM_OK;
}

}  // namespace nnam
