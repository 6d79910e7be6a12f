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

#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

constexpr int RNN_THREADS = 128;

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

template <int M_ROWS, int NB, int NSPLIT, bool FAST_TANH>
__global__ void __launch_bounds__(RNN_THREADS, 1)
    lstm_seq_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int group = blockIdx.x / p.group_ctas;
  const int rank = blockIdx.x % p.group_ctas;
  if (group >= p.n_groups) return;

  const int H = p.hidden;
  const int KB = H >> 6;  // 64-element k-blocks
  constexpr int W_BLOCK = M_ROWS * 128;
  constexpr int H_BLOCK = NB * 128;
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

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
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

  // TMEM lane <-> gate row of this CTA's slice.  M_ROWS = 128: lane = tid.  M_ROWS = 64: rows 16q..16q+15 sit in
  // lanes 32q..32q+15 (the upper half of every subpartition is unused).
  const bool row_valid = (M_ROWS == 128) || (lane < 16);
  const int my_row = (M_ROWS == 128) ? tid : (warp * 16 + (lane & 15));
  const int gate = my_row & 3;                        // a, i, f, o
  const int unit = rank * (M_ROWS / 4) + (my_row >> 2);  // hidden unit index in [0, H)
  const int gate_col = rank * M_ROWS + my_row;        // column of gx / row of W_lat
  const uint32_t tmem_lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t idesc = make_idesc_bf16_f32(M_ROWS, NB);

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
    const float* gx = p.gx[d];
    const int h_col0 = d * H;
    __syncthreads();  // previous item's readers of s_len are done
    if (tid < NB) s_len[tid] = tid < nutt ? len[tid] : 0;
    __syncthreads();

    // cell state of (utterance 4m + gate, unit) lives in this thread for the whole item
    float c_reg[NB / 4];
#pragma unroll
    for (int m = 0; m < NB / 4; ++m) {
      const int u = 4 * m + gate;
      c_reg[m] = (p.c0 != nullptr && row_valid && u < nutt)
                     ? p.c0[(static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0 + unit]
                     : 0.0f;
    }
    const bool has_h0 = p.h0_hi != nullptr;

    for (int s = 0; s < T; ++s) {
      const int base_s = __ldg(base + s);
      const int n_s = __ldg(base + s + 1) - base_s;  // active utterances (prefix of the batch)

      // ---- prefetch the input projection of my gate row for every active utterance
      float gxr[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        gxr[u] = 0.0f;
        if (u < n_s) {
          const long long row = row0 + (bwd ? __ldg(base + (s_len[u] - 1 - s)) : base_s) + u;
          if (row_valid) gxr[u] = __ldg(gx + row * p.gx_ld + gate_col);
        }
      }

      const bool do_mma = (s > 0) || has_h0;
      float acc[NB];
      if (do_mma) {
        if (s > 0) {
          if (tid == 0) {
            const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
            while (ld_acquire_gpu(counter) < target) {
            }
          }
          __syncthreads();
        }
        // ---- h_{s-1} rows of the active utterances -> swizzled smem (B operand)
        const int chunks_per_row = H >> 3;
        const int prev_base = s > 0 ? (bwd ? 0 : __ldg(base + s - 1)) : 0;
        for (int idx = tid; idx < n_s * chunks_per_row; idx += RNN_THREADS) {
          const int u = idx / chunks_per_row;
          const int c = idx - u * chunks_per_row;
          const __nv_bfloat16 *src_hi, *src_lo = nullptr;
          if (s == 0) {
            const long long off = (static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0 + c * 8;
            src_hi = p.h0_hi + off;
            if (NSPLIT == 3) src_lo = p.h0_lo + off;
          } else {
            const long long row = row0 + (bwd ? __ldg(base + (s_len[u] - s)) : prev_base) + u;
            const long long off = row * p.h_ld + h_col0 + c * 8;
            src_hi = p.h_hi + off;
            if (NSPLIT == 3) src_lo = p.h_lo + off;
          }
          const uint32_t so = static_cast<uint32_t>((c >> 3) * H_BLOCK) + sw128_offset(u, c & 7);
          *reinterpret_cast<uint4*>(h_hi_s + so) = __ldcg(reinterpret_cast<const uint4*>(src_hi));
          if (NSPLIT == 3) *reinterpret_cast<uint4*>(h_lo_s + so) = __ldcg(reinterpret_cast<const uint4*>(src_lo));
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          uint32_t accum = 0;
#pragma unroll 1
          for (int pass = 0; pass < NSPLIT; ++pass) {
            const uint32_t wa = smem_u32(pass == 2 ? w_lo_s : w_hi_s);
            const uint32_t ha = smem_u32(pass == 1 ? h_lo_s : h_hi_s);
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tmem_base, make_sw128_kmajor_desc(wa + kb * W_BLOCK + k * 32),
                          make_sw128_kmajor_desc(ha + kb * H_BLOCK + k * 32), idesc, accum);
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
        for (int c0 = 0; c0 < NB; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(r[j]);
        }
        tc_fence_before();
      } else {
#pragma unroll
        for (int u = 0; u < NB; ++u) acc[u] = 0.0f;
      }

      // ---- gates, quad transpose, cell update, write h
#pragma unroll
      for (int m = 0; m < NB / 4; ++m) {
        if (4 * m >= n_s) break;  // warp-uniform
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = acc[4 * m + i] + gxr[4 * m + i];
          const float t = tanh_sel<FAST_TANH>(gate == 0 ? v : 0.5f * v);
          x[i] = gate == 0 ? t : fmaf(t, 0.5f, 0.5f);
        }
        quad_transpose(x, gate);  // x = {a, i, f, o} of utterance u = 4m + gate
        const int u = 4 * m + gate;
        const float c_new = fmaf(x[0], x[1], x[2] * c_reg[m]);
        const float h_new = x[3] * tanh_sel<FAST_TANH>(c_new);
        if (row_valid && u < n_s) {
          c_reg[m] = c_new;
          const int t_idx = bwd ? (s_len[u] - 1 - s) : s;
          const long long row = row0 + __ldg(base + t_idx) + u;
          const long long off = row * p.h_ld + h_col0 + unit;
          const __nv_bfloat16 hb = __float2bfloat16_rn(h_new);
          p.h_hi[off] = hb;
          if (p.h_lo != nullptr) p.h_lo[off] = __float2bfloat16_rn(h_new - __bfloat162float(hb));
          if (p.c_out != nullptr && s == s_len[u] - 1)
            p.c_out[(static_cast<long long>(b) * NB + u) * (H * p.n_dirs) + h_col0 + unit] = c_new;
        }
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) red_release_gpu_add(counter, 1u);
      ++steps_done;
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
template <int M_ROWS, int NB, int NSPLIT, bool FAST>
static int launch_lstm(const RnnTmaps& tm, const RnnParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = lstm_seq_kernel<M_ROWS, NB, NSPLIT, FAST>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute");
  void* args[] = {const_cast<RnnTmaps*>(&tm), const_cast<RnnParams*>(&p)};
  // cooperative launch: every CTA of a group must be co-resident (they wait on one another every step)
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(RNN_THREADS), args, smem, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaLaunchCooperativeKernel");
  return NNAM_OK;
}

size_t rnn_smem_bytes(int m_rows, int nb, int hidden, int nsplit) {
  const size_t kb = hidden / 64;
  const size_t mult = nsplit == 3 ? 2 : 1;
  return kb * (static_cast<size_t>(m_rows) * 128 + static_cast<size_t>(nb) * 128) * mult + 64 + nb * 4 + 1024;
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
  if (d->cell != NNAM_CELL_LSTM) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: cell kind %d not implemented", d->cell);
  const int H = d->hidden;
  if (H <= 0 || H % 64) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: hidden size must be a multiple of 64 (got %d)", H);
  if (d->n_dirs != 1 && d->n_dirs != 2) return set_error(NNAM_ERR_ARG, "rnn: n_dirs must be 1 or 2");
  if (d->batch != 16 && d->batch != 32 && d->batch != 64)
    return set_error(NNAM_ERR_ARG, "rnn: batch must be 16, 32 or 64");
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
  if (p.h0_hi && d->nsplit == 3 && !p.h0_lo) return set_error(NNAM_ERR_ARG, "rnn: bf16x3 needs h0_lo with h0_hi");

  cudaError_t e = cudaMemsetAsync(d->counters, 0, sizeof(unsigned int) * d->n_groups, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaMemsetAsync");
  const int grid = d->n_groups * G;
  const size_t smem = rnn_smem_bytes(m_rows, d->batch, H, d->nsplit);
  const bool fast = d->nsplit == 1;  // bf16 mode: MUFU tanh; fp32-accurate mode: tanhf

#define NNAM_RNN_CASE(M, NBV, NS, F)                                   \
  if (m_rows == M && d->batch == NBV && d->nsplit == NS && fast == F)  \
  return launch_lstm<M, NBV, NS, F>(tm, p, grid, smem, stream)
  NNAM_RNN_CASE(128, 16, 1, true);
  NNAM_RNN_CASE(128, 32, 1, true);
  NNAM_RNN_CASE(128, 64, 1, true);
  NNAM_RNN_CASE(64, 16, 1, true);
  NNAM_RNN_CASE(64, 32, 1, true);
  NNAM_RNN_CASE(64, 64, 1, true);
  NNAM_RNN_CASE(128, 16, 3, false);
  NNAM_RNN_CASE(128, 32, 3, false);
  NNAM_RNN_CASE(128, 64, 3, false);
  NNAM_RNN_CASE(64, 16, 3, false);
  NNAM_RNN_CASE(64, 32, 3, false);
  NNAM_RNN_CASE(64, 64, 3, false);
#undef NNAM_RNN_CASE
  return set_error(NNAM_ERR_UNSUPPORTED, "rnn: no kernel instance for this configuration");
}

int rnn_plan(int cell, int hidden, int batch, int nsplit, int* group_ctas, int* max_groups) {
  if (cell != NNAM_CELL_LSTM) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: cell kind %d not implemented", cell);
  if (hidden <= 0 || hidden % 64) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: hidden size must be a multiple of 64");
  const int m_rows = rnn_pick_m_rows(4 * hidden, hidden, batch, nsplit);
  if (!m_rows)
    return set_error(NNAM_ERR_UNSUPPORTED,
                     "rnn: lateral weights of H=%d do not fit in shared memory in this precision mode", hidden);
  *group_ctas = 4 * hidden / m_rows;
  *max_groups = sm_count() / *group_ctas;
  if (*max_groups < 1) return set_error(NNAM_ERR_UNSUPPORTED, "rnn: H=%d needs %d co-resident CTAs", hidden, *group_ctas);
  return NNAM_OK;
}

}  // namespace nnam
