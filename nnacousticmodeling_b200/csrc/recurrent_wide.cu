// K3, "wide" LSTM variant for bf16 mode: 128 utterances per batch, utterances on the M axis of the MMA.
//
// Same reference semantics and same group / exchange protocol as recurrent.cu (L.LSTM via
// scripts/common/chainer_networks.py:44-62, predict_folds.py:49-61); what changes is the shape of a step:
//
//   recurrent.cu      D[128 gate rows x NB utterances] = W_slice . h^T      a thread owns ONE gate row: the four gates
//                     of a unit sit in four TMEM lanes and meet through a shuffle transpose, gx is 16-32 scalar loads
//                     per thread, h leaves through a staging tile, and NB <= 64 because the whole h tile (NB x H) must
//                     sit in shared memory next to the 128 KB weight slice.
//   this kernel       D[128 utterances x 128 gate rows] = h . W_slice^T     a TMEM lane is an UTTERANCE and its columns
//                     are [unit-major, gate-minor]: a thread reads the four gates of a unit from four adjacent
//                     columns (no transpose), loads its gx as 4 x 128-bit, and stores its 8 new h values as ONE
//                     128-bit store per destination.  The h tile is the A operand and is STREAMED: k-block by k-block
//                     (128 rows x 128 B = 16 KB) through a 4-stage TMA ring straight into the MMAs, so NB = 128 fits
//                     and a step moves twice the utterances for about the same exchange latency.
//
// Two kernels live here.  lstm_seq_wide_kernel (one batch per CTA group, kept for A/B runs, NNAM_RNN_WIDE_STREAMS=1):
// roles per CTA (512 threads): warp 0 = exchange wait + TMA producer, warp 1 = MMA issuer (both with all lanes, one
// elected), then all 16 warps do the gate math: warp w owns TMEM lane quarter w & 3 (32 utterances) and columns
// [32 * (w >> 2), +32) = 8 units.  lstm_seq_wide2_kernel (the default, further down): S interleaved batches per group
// sharing one TMA ring -- its header comment has the protocol.
#include <cooperative_groups.h>

#include "recurrent_common.cuh"

namespace nnam {

constexpr int WIDE_NB = 128;
constexpr int WIDE_THREADS = 512;
constexpr int WIDE_STAGES = 4;
constexpr int WIDE_A_STAGE = WIDE_NB * 128;  // one k-block of the h tile: 128 rows x 128 B

__device__ __forceinline__ float wide_sigmoid(float v) { return fmaf(tanh_fast(0.5f * v), 0.5f, 0.5f); }

template <int KBT>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
    lstm_seq_wide_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int M_ROWS = 128;  // gate rows per CTA (the N of the MMA)
  constexpr int UNITS = 32;
  constexpr int NB = WIDE_NB;
  const int group = blockIdx.x / p.group_ctas;
  const int rank = blockIdx.x % p.group_ctas;
  announce_started(p);
  if (group >= p.n_groups) return;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int H = KBT > 0 ? KBT * 64 : p.hidden;
  const int KB = KBT > 0 ? KBT : (H >> 6);
  constexpr int W_BLOCK = M_ROWS * 128;
  uint8_t* w_s = smem;                          // KB blocks of [128 gate rows x 128 B], SWIZZLE_128B (B operand)
  uint8_t* a_s = w_s + KB * W_BLOCK;            // ring of WIDE_STAGES h k-blocks (A operand)
  uint8_t* tail = a_s + WIDE_STAGES * WIDE_A_STAGE;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;
  uint64_t* full_bar = bar_w + 2;               // [WIDE_STAGES]
  uint64_t* empty_bar = full_bar + WIDE_STAGES; // [WIDE_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_bar + WIDE_STAGES);
  int* s_len = reinterpret_cast<int*>(tmem_slot + 2);  // NB
  int* s_base = s_len + NB;                            // RNN_BASE_SMEM + 1

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int quarter = warp & 3;
  const int sub = warp >> 2;           // which 32 columns (8 units) of the CTA's 128
  const int u = quarter * 32 + lane;   // my utterance slot == my TMEM lane
  constexpr int TMEM_COLS = 128;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    for (int i = 0; i < WIDE_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_mine = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(sub * 32);
  const int f16 = p.f16;
  const uint32_t idesc = make_idesc_e16_f32(NB, M_ROWS, f16);  // M = utterances, N = gate rows

  const bool prof_on = p.prof != nullptr && tid == 64;  // a thread that is neither the producer nor the MMA issuer
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#define PROF_START() do { if (prof_on) prof_t = clock64(); } while (0)
#define PROF_MARK(i) do { if (prof_on) { const long long now = clock64(); prof_acc[i] += now - prof_t; prof_t = now; } } while (0)

  unsigned int steps_done = 0;
  uint32_t w_phase = 0, mma_phase = 0;
  int ring_stage = 0;            // producer and MMA issuer walk the ring in lock step (KB blocks per exchange)
  uint32_t ring_phase = 0;
  unsigned int* counter = p.counters + group;
  int it = p.group_item_start[group];
  const int it_end = p.group_item_start[group + 1];

  for (int d = 0; d < p.n_dirs; ++d) {
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar_w, static_cast<uint32_t>(KB * W_BLOCK));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * M_ROWS);
    }
    mbar_wait(bar_w, w_phase);
    w_phase ^= 1;
    const bool bwd = d == 1;
    const int h_col0 = d * H;

    for (; it < it_end && p.item_dir[it] == d; ++it) {
      const int b = p.item_batch[it];
      const long long row0 = p.batch_row0[b];
      const int T = p.batch_steps[b];
      const int nutt = p.batch_nutt[b];
      const int* base = p.base + p.batch_base_off[b];
      const int* len = p.utt_len + b * NB;
      // my 32 gate columns (8 units x [a, i, f, o]) of the bf16 input projection
      const __nv_bfloat16* gx = reinterpret_cast<const __nv_bfloat16*>(p.gx[d]) + rank * M_ROWS + sub * 32;
      // exchange slots are reused by the next item: wait until the whole group has finished the previous one
      if (steps_done > 0 && warp == 0) {
        const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
        while (ld_acquire_gpu(counter) < target) {
        }
      }
      __syncthreads();
      if (tid < NB) s_len[tid] = tid < nutt ? len[tid] : 0;
      const bool base_in_smem = T <= RNN_BASE_SMEM;
      if (base_in_smem)
        for (int i = tid; i <= T; i += WIDE_THREADS) s_base[i] = __ldg(base + i);
      __syncthreads();
      const int* bp = base_in_smem ? s_base : base;
      const int my_len = s_len[u];

      float c_reg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c_reg[j] = 0.0f;

      // 32 bf16 gate pre-activations of utterance u at step s (4 x 128-bit loads); inactive lanes load nothing
      auto load_gx = [&](int s, uint4 (&dst)[4]) {
        if (s < my_len) {
          const long long row = row0 + bp[bwd ? (my_len - 1 - s) : s] + u;
          const uint4* src = reinterpret_cast<const uint4*>(gx + row * p.gx_ld);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = __ldg(src + j);
        }
      };

      uint4 gxr[4] = {}, gxn[4] = {};
      load_gx(0, gxr);
      for (int s = 0; s < T; ++s) {
        PROF_START();
        const int n_s = bp[s + 1] - bp[s];  // active utterances = slots [0, n_s)
        const bool active = u < n_s;        // == s < my_len (utterances are sorted by length)
        float acc[32];
        if (s > 0) {
          // ---- the group's h of step s-1 streams from the exchange buffer through the ring into the MMAs
          if (warp == 0) {
            const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
            while (ld_acquire_gpu(counter) < target) {
            }
            if (elect_one()) fence_proxy_async_all();  // peers' generic-proxy stores -> this async-proxy (TMA) read
            __syncwarp();
            const int xrow = (group * 4 + ((s - 1) & 1)) * NB;
            int st = ring_stage;
            uint32_t ph = ring_phase;
            for (int kb = 0; kb < KB; ++kb) {
              mbar_wait(&empty_bar[st], ph ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&full_bar[st], WIDE_A_STAGE);
                tma_load_2d(a_s + st * WIDE_A_STAGE, &tmaps.x_hi, &full_bar[st], kb * 64, xrow);
              }
              __syncwarp();
              if (++st == WIDE_STAGES) {
                st = 0;
                ph ^= 1;
              }
            }
          } else if (warp == 1) {
            const uint64_t adesc0 = make_sw128_kmajor_desc(smem_u32(a_s));
            const uint64_t bdesc0 = make_sw128_kmajor_desc(smem_u32(w_s));
            int st = ring_stage;
            uint32_t ph = ring_phase;
            for (int kb = 0; kb < KB; ++kb) {
              mbar_wait(&full_bar[st], ph);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ad = adesc0 + static_cast<uint64_t>((st * WIDE_A_STAGE) >> 4);
                const uint64_t bd = bdesc0 + static_cast<uint64_t>((kb * W_BLOCK) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[st]);
                if (kb == KB - 1) umma_commit(bar_mma);
              }
              __syncwarp();
              if (++st == WIDE_STAGES) {
                st = 0;
                ph ^= 1;
              }
            }
          }
          // every thread advances its copy of the ring position by KB blocks
          {
            const int adv = ring_stage + KB;
            ring_phase ^= static_cast<uint32_t>((adv / WIDE_STAGES) & 1);
            ring_stage = adv % WIDE_STAGES;
          }
          PROF_MARK(1);
          if (s + 1 < T) load_gx(s + 1, gxn);  // lands while the tensor core works
          PROF_MARK(2);
          mbar_wait(bar_mma, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          PROF_MARK(3);
          uint32_t r[32];
          tmem_ld32(tmem_mine, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          tc_fence_before();
          PROF_MARK(4);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
          if (s + 1 < T) load_gx(s + 1, gxn);
        }

        // ---- gates and cell update for my 8 units (chainer F.lstm), h to the exchange slot and the layer output
        if (active) {
          const uint32_t* g2 = reinterpret_cast<const uint32_t*>(gxr);
          uint32_t hp[4];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 ga = e16x2_to_float2(g2[2 * j], f16);      // (a, i) pre-activations from the input projection
            const float2 gf = e16x2_to_float2(g2[2 * j + 1], f16);  // (f, o)
            const float a = tanh_fast(acc[4 * j] + ga.x);
            const float ig = wide_sigmoid(acc[4 * j + 1] + ga.y);
            const float fg = wide_sigmoid(acc[4 * j + 2] + gf.x);
            const float og = wide_sigmoid(acc[4 * j + 3] + gf.y);
            c_reg[j] = fmaf(a, ig, fg * c_reg[j]);
            const float h_new = og * tanh_fast(c_reg[j]);
            if (j & 1)
              hp[j >> 1] |= static_cast<uint32_t>(f32_to_e16(h_new, f16)) << 16;
            else
              hp[j >> 1] = static_cast<uint32_t>(f32_to_e16(h_new, f16));
          }
          const uint4 hv = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          const long long xoff = (static_cast<long long>(group * 4 + (s & 1)) * NB + u) * H + rank * UNITS + sub * 8;
          *reinterpret_cast<uint4*>(p.xchg_hi + xoff) = hv;
          const int t_idx = bwd ? (my_len - 1 - s) : s;
          const long long off = (row0 + bp[t_idx] + u) * p.h_ld + h_col0 + rank * UNITS + sub * 8;
          *reinterpret_cast<uint4*>(p.h_hi + off) = hv;
        }
        PROF_MARK(5);
        // ---- publish: the CTA barrier orders every thread's stores before thread 0's release
        __syncthreads();
        if (tid == 0) red_release_gpu_add(counter, 1u);
        ++steps_done;
        PROF_MARK(6);
#pragma unroll
        for (int j = 0; j < 4; ++j) gxr[j] = gxn[j];
      }
    }
  }

  if (prof_on) {
    prof_acc[7] = steps_done;
    for (int i = 0; i < 8; ++i) p.prof[blockIdx.x * 8 + i] = prof_acc[i];
  }
#undef PROF_START
#undef PROF_MARK
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Two batches ("streams") per CTA group, interleaved.  A step of the kernel above is a chain of latencies (exchange
// wait -> two rounds of the 4-slot TMA ring -> MMA -> gate math -> publish) that leaves every unit of the SM idle most of
// the time (profiles/r01_k3_phase_cycles.md).  Here the 16 warps split into two independent sets of 8; each set runs the
// whole recurrence of its own batch (own TMEM accumulator, own exchange slots and counter) against the SAME resident
// weight slice, and the two sets share one 5-slot TMA ring under a lock: a set takes the ring for the KB k-blocks of one
// step, so the ring keeps 80 KB in flight across both streams while the other stream does its gate math or waits for
// its peers.  A warp of a set owns TMEM lane quarter w & 3 and 64 columns (16 units), processed as two halves of 32.
// ------------------------------------------------------------------------------------------------------------------
constexpr int W2_STREAMS = 2;
constexpr int W2_TPS = 256;  // threads per stream
constexpr int W2_THREADS = W2_STREAMS * W2_TPS;
constexpr int W2_STAGES = 5;
constexpr int W2_BASE_SMEM = 1024;

template <int TPS>
__device__ __forceinline__ void w2_bar_sync(int stream) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + stream), "n"(TPS) : "memory");
}

// S streams of WPS warps each: (2, 8) -- a thread owns 64 columns (16 units) -- or (3, 4) -- a thread owns all 128
// columns (32 units) of its utterance, processed in four chunks of 32; three chains of latencies overlap instead of two.
template <int KBT, int S, int WPS>
__global__ void __launch_bounds__(S * WPS * 32, 1)
    lstm_seq_wide2_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int M_ROWS = 128;
  constexpr int UNITS = 32;
  constexpr int NB = WIDE_NB;
  const int group = blockIdx.x / p.group_ctas;
  const int rank = blockIdx.x % p.group_ctas;
  announce_started(p);
  if (group >= p.n_groups) return;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int H = KBT > 0 ? KBT * 64 : p.hidden;
  const int KB = KBT > 0 ? KBT : (H >> 6);
  constexpr int W_BLOCK = M_ROWS * 128;
  uint8_t* w_s = smem;
  uint8_t* a_s = w_s + KB * W_BLOCK;
  uint8_t* tail = a_s + W2_STAGES * WIDE_A_STAGE;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;                   // [2] a stream's accumulator is complete
  constexpr int TPS = WPS * 32;            // threads per stream
  constexpr int COLS = 128 / (WPS / 4);    // accumulator columns per thread
  constexpr int CHUNKS = COLS / 32;
  constexpr int BASE_SMEM = S == 2 ? W2_BASE_SMEM : 768;
  uint64_t* bar_grant = bar_mma + S;      // [2] producer -> MMA issuer: "your k-blocks start at s_f0[stream]"
  // [2][W2_STAGES] "slot filled", one set PER STREAM although the slots are shared: a parity wait on a barrier that
  // still carries the other stream's previous fill would see "the phase before" and pass at once.  With its own set
  // a stream's issuer only ever waits for its own fills, in order.  The "slot free" barriers are shared.
  uint64_t* full_bar_all = bar_grant + S;
  uint64_t* full_bar = full_bar_all + (static_cast<int>(threadIdx.x) / TPS) * W2_STAGES;
  uint64_t* empty_bar = full_bar_all + S * W2_STAGES;  // [W2_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_bar + W2_STAGES);
  unsigned int* ring_lock = tmem_slot + 2;
  volatile unsigned int* ring_fill = ring_lock + 1;  // k-blocks pushed through the ring so far (by either stream)
  volatile unsigned int* s_f0 = ring_lock + 2;       // [S] ring position of the first k-block of a stream's current step
  int* s_len_all = reinterpret_cast<int*>(ring_lock + 6);    // [S][NB]
  int* s_base_all = s_len_all + S * NB;                      // [S][BASE_SMEM + 1]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int stream = warp / WPS;
  const int sw = warp - stream * WPS;  // warp inside the stream
  const int stid = tid - stream * TPS;
  const int quarter = sw & 3;          // == warp & 3: the TMEM lane quarter this warp may read
  const int sub = sw >> 2;             // which COLS columns of the CTA's 128
  const int u = quarter * 32 + lane;   // my utterance slot == my TMEM lane
  constexpr int TMEM_COLS = S == 2 ? 256 : 512;
  int* s_len = s_len_all + stream * NB;
  int* s_base = s_base_all + stream * (BASE_SMEM + 1);

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&bar_mma[i], 1);
      mbar_init(&bar_grant[i], 1);
    }
    for (int i = 0; i < S * W2_STAGES; ++i) mbar_init(&full_bar_all[i], 1);
    for (int i = 0; i < W2_STAGES; ++i) mbar_init(&empty_bar[i], 1);
    *ring_lock = 0u;
    *ring_fill = 0u;
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(stream * 128);
  const uint32_t tmem_mine = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(sub * COLS);
  const int f16 = p.f16;
  const uint32_t idesc = make_idesc_e16_f32(NB, M_ROWS, f16);

  const bool prof_on = p.prof != nullptr && tid == 64;
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#define PROF_START() do { if (prof_on) prof_t = clock64(); } while (0)
#define PROF_MARK(i) do { if (prof_on) { const long long now = clock64(); prof_acc[i] += now - prof_t; prof_t = now; } } while (0)

  unsigned int steps_done = 0;
  uint32_t w_phase = 0, mma_phase = 0, grant_phase = 0;
  uint32_t full_parity = 0;  // bit st: parity of this stream's next fill of ring slot st
  const int lane_id_ = group * S + stream;  // (group, stream) = one lane of the schedule
  unsigned int* counter = p.counters + lane_id_;
  int it = p.group_item_start[lane_id_];
  const int it_end = p.group_item_start[lane_id_ + 1];

  for (int d = 0; d < p.n_dirs; ++d) {
    __syncthreads();  // both streams are done with the previous direction's weights
    if (tid == 0) {
      mbar_expect_tx(bar_w, static_cast<uint32_t>(KB * W_BLOCK));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * M_ROWS);
    }
    mbar_wait(bar_w, w_phase);
    w_phase ^= 1;
    const bool bwd = d == 1;
    const int h_col0 = d * H;

    for (; it < it_end && p.item_dir[it] == d; ++it) {
      const int b = p.item_batch[it];
      const long long row0 = p.batch_row0[b];
      const int T = p.batch_steps[b];
      const int nutt = p.batch_nutt[b];
      const int* base = p.base + p.batch_base_off[b];
      const int* len = p.utt_len + b * NB;
      const __nv_bfloat16* gx = reinterpret_cast<const __nv_bfloat16*>(p.gx[d]) + rank * M_ROWS + sub * COLS;
      // exchange slots are reused by the next item: wait until the whole group has finished the previous one
      if (steps_done > 0 && sw == 0) {
        const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
        while (ld_acquire_gpu(counter) < target) {
        }
      }
      w2_bar_sync<TPS>(stream);
      if (stid < NB) s_len[stid] = stid < nutt ? len[stid] : 0;
      const bool base_in_smem = T <= BASE_SMEM;
      if (base_in_smem)
        for (int i = stid; i <= T; i += TPS) s_base[i] = __ldg(base + i);
      w2_bar_sync<TPS>(stream);
      const int* bp = base_in_smem ? s_base : base;
      const int my_len = s_len[u];

      float c_reg[COLS / 4];
#pragma unroll
      for (int j = 0; j < COLS / 4; ++j) c_reg[j] = 0.0f;

      for (int s = 0; s < T; ++s) {
        PROF_START();
        const int n_s = bp[s + 1] - bp[s];
        const bool active = u < n_s;
        const int t_idx = bwd ? (my_len - 1 - s) : s;
        const long long my_row = active ? row0 + bp[t_idx] + u : 0;
        // 64 bf16 gate pre-activations of utterance u at this step.  Every lane reads its own row, so the four loads
        // cost ~1-2 k cycles of LSU issue per warp: the six plain warps issue them now (they land during the exchange);
        // the producer and the MMA issuer first start the h stream, which is the critical path of the step.
        uint32_t gxr[COLS / 16][8] = {};
        auto load_gx = [&]() {
          if (active) {
            const __nv_bfloat16* src = gx + my_row * p.gx_ld;
#pragma unroll
            for (int j = 0; j < COLS / 16; ++j) ldg_nc_256(src + 16 * j, gxr[j]);
          }
        };
        if (s == 0 || sw >= 2) load_gx();
        PROF_MARK(0);
        if (s > 0) {
          if (sw == 0) {
            // ---- producer: peers' h of step s-1 -> ring; the ring is taken for the KB blocks of this step
            const unsigned int target = steps_done * static_cast<unsigned int>(p.group_ctas);
            while (ld_acquire_gpu(counter) < target) {
            }
            unsigned int f0 = 0;
            if (lane == 0) {
              fence_proxy_async_all();  // peers' generic-proxy stores -> this async-proxy (TMA) read
              while (atomicCAS(ring_lock, 0u, 1u) != 0u) {
              }
              __threadfence_block();
              f0 = *ring_fill;
              s_f0[stream] = f0;
              __threadfence_block();
              mbar_arrive(&bar_grant[stream]);
            }
            f0 = __shfl_sync(0xffffffffu, f0, 0);
            const int xrow = (lane_id_ * 4 + ((s - 1) & 1)) * NB;

            for (int kb = 0; kb < KB; ++kb) {
              const unsigned int f = f0 + kb;
              const int st = static_cast<int>(f % W2_STAGES);
              const uint32_t ph = (f / W2_STAGES) & 1u;
              mbar_wait(&empty_bar[st], ph ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&full_bar[st], WIDE_A_STAGE);
                tma_load_2d(a_s + st * WIDE_A_STAGE, &tmaps.x_hi, &full_bar[st], kb * 64, xrow);
              }
              __syncwarp();
            }
            if (lane == 0) {
              *ring_fill = f0 + KB;
              __threadfence_block();
              atomicExch(ring_lock, 0u);
            }
            __syncwarp();
            load_gx();
          } else if (sw == 1) {
            // ---- MMA issuer of this stream
            mbar_wait(&bar_grant[stream], grant_phase);
            const unsigned int f0 = s_f0[stream];
            const uint64_t adesc0 = make_sw128_kmajor_desc(smem_u32(a_s));
            const uint64_t bdesc0 = make_sw128_kmajor_desc(smem_u32(w_s));
            for (int kb = 0; kb < KB; ++kb) {
              const unsigned int f = f0 + kb;
              const int st = static_cast<int>(f % W2_STAGES);
              mbar_wait(&full_bar[st], (full_parity >> st) & 1u);
              full_parity ^= 1u << st;
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ad = adesc0 + static_cast<uint64_t>((st * WIDE_A_STAGE) >> 4);
                const uint64_t bd = bdesc0 + static_cast<uint64_t>((kb * W_BLOCK) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[st]);
                if (kb == KB - 1) umma_commit(&bar_mma[stream]);
              }
              __syncwarp();
              if (kb == 0) load_gx();  // behind the first k-block: the next ones are still in flight
            }
          }
          grant_phase ^= 1;
          PROF_MARK(1);
          mbar_wait(&bar_mma[stream], mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          PROF_MARK(2);
        }

        // ---- gates and cell update for my units in chunks of 8 (chainer F.lstm); h leaves as one 256-bit store per
        //      16 units and destination (every lane writes its own row: the LSU cost is per instruction, not per byte)
        uint32_t hp[CHUNKS * 4];
#pragma unroll
        for (int half = 0; half < CHUNKS; ++half) {
          float acc[32];
          if (s > 0) {
            uint32_t r[32];
            tmem_ld32(tmem_mine + static_cast<uint32_t>(half * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
          }
          if (active) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // columns [unit-major, gate-minor]: 2 words (a, i), (f, o) per unit
            const float2 ga = e16x2_to_float2(gxr[2 * half + (j >> 2)][(2 * j) & 7], f16);
            const float2 gf = e16x2_to_float2(gxr[2 * half + (j >> 2)][(2 * j + 1) & 7], f16);
            const float a = tanh_fast(acc[4 * j] + ga.x);
            const float ig = wide_sigmoid(acc[4 * j + 1] + ga.y);
            const float fg = wide_sigmoid(acc[4 * j + 2] + gf.x);
            const float og = wide_sigmoid(acc[4 * j + 3] + gf.y);
            float& c = c_reg[half * 8 + j];
            c = fmaf(a, ig, fg * c);
            const float h_new = og * tanh_fast(c);
            const uint32_t hb = static_cast<uint32_t>(f32_to_e16(h_new, f16));
            if (j & 1)
              hp[half * 4 + (j >> 1)] |= hb << 16;
            else
              hp[half * 4 + (j >> 1)] = hb;
          }
          }
        }
        if (active) {
          const int col = rank * UNITS + sub * (COLS / 4);
          const long long xoff = (static_cast<long long>(lane_id_ * 4 + (s & 1)) * NB + u) * H + col;
#pragma unroll
          for (int q = 0; q < CHUNKS / 2; ++q) {
            uint32_t v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v8[j] = hp[8 * q + j];
            stg_256(p.xchg_hi + xoff + 16 * q, v8);
            stg_256(p.h_hi + my_row * p.h_ld + h_col0 + col + 16 * q, v8);
          }
        }
        tc_fence_before();
        PROF_MARK(3);
        // ---- publish: the stream barrier orders every thread's stores before one thread's release
        w2_bar_sync<TPS>(stream);
        if (stid == 0) red_release_gpu_add(counter, 1u);
        ++steps_done;
        PROF_MARK(4);
      }
    }
  }

  if (prof_on) {
    prof_acc[7] = steps_done;
    for (int i = 0; i < 8; ++i) p.prof[blockIdx.x * 8 + i] = prof_acc[i];
  }
#undef PROF_START
#undef PROF_MARK
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

size_t rnn_wide2_smem_bytes(int hidden, int streams) {
  const size_t kb = hidden / 64;
  const size_t base_smem = streams == 2 ? W2_BASE_SMEM : 768;
  return kb * 128 * 128 + static_cast<size_t>(W2_STAGES) * WIDE_A_STAGE + 8 * (1 + 2 * streams + (streams + 1) * W2_STAGES) +
         8 * 4 + streams * (WIDE_NB + base_smem + 1) * 4;
}

// streams per CTA group of the 128-slot LSTM kernel: 2 (default); NNAM_RNN_WIDE_STREAMS=1 selects the single-batch
// kernel, =3 three streams of four warps (27 % more utterance-steps per cycle when every stream is busy, but a stream's
// own step gets longer -- 14.7 k cycles busy, 11.2 k alone -- and the BASELINE sets are bound by their longest batch:
// cfg3t 5.2 ms per layer against 4.5 ms with two streams; profiles/r01_k3_phase_cycles.md)
int rnn_wide_streams() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NNAM_RNN_WIDE_STREAMS");
    v = (e && atoi(e) >= 1 && atoi(e) <= 3) ? atoi(e) : 2;
  }
  return v;
}

size_t rnn_wide_smem_bytes(int hidden) {
  const size_t kb = hidden / 64;
  return kb * 128 * 128 + static_cast<size_t>(WIDE_STAGES) * WIDE_A_STAGE + 8 * (2 + 2 * WIDE_STAGES) + 16 +
         (WIDE_NB + RNN_BASE_SMEM + 1) * 4;
}

// The wide kernel applies to: LSTM, bf16 mode, 128 slots per batch, no carried state, H a multiple of 64 whose slice fits.
bool rnn_wide_applies(int cell, int hidden, int batch, int nsplit) {
  if (cell != NNAM_CELL_LSTM || nsplit != 1 || batch != WIDE_NB) return false;
  if (hidden % 64 || (4 * hidden) % 128) return false;
  if ((rnn_wide_streams() >= 2 ? rnn_wide2_smem_bytes(hidden, rnn_wide_streams()) : rnn_wide_smem_bytes(hidden)) > 227 * 1024)
    return false;
  return sm_count() >= 4 * hidden / 128;
}

template <int KBT>
static int launch_wide(const RnnTmaps& tm, const RnnParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = lstm_seq_wide_kernel<KBT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute(wide)");
  void* args[] = {const_cast<RnnTmaps*>(&tm), const_cast<RnnParams*>(&p)};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(WIDE_THREADS), args, smem, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaLaunchCooperativeKernel(wide)");
  return NNAM_OK;
}

template <int KBT, int S, int WPS>
static int launch_wide2(const RnnTmaps& tm, const RnnParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = lstm_seq_wide2_kernel<KBT, S, WPS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute(wide2)");
  void* args[] = {const_cast<RnnTmaps*>(&tm), const_cast<RnnParams*>(&p)};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(S * WPS * 32), args, smem, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaLaunchCooperativeKernel(wide2)");
  return NNAM_OK;
}

int rnn_wide_launch(const RnnTmaps& tm, const RnnParams& p, int hidden, cudaStream_t stream) {
  const int grid = p.n_groups * p.group_ctas;
  if (rnn_wide_streams() == 3) {
    const size_t smem3 = rnn_wide2_smem_bytes(hidden, 3);
    if (hidden == 512) return launch_wide2<8, 3, 4>(tm, p, grid, smem3, stream);
    return launch_wide2<0, 3, 4>(tm, p, grid, smem3, stream);
  }
  if (rnn_wide_streams() == 2) {
    const size_t smem2 = rnn_wide2_smem_bytes(hidden, 2);
    if (hidden == 512) return launch_wide2<8, 2, 8>(tm, p, grid, smem2, stream);
    return launch_wide2<0, 2, 8>(tm, p, grid, smem2, stream);
  }
  const size_t smem = rnn_wide_smem_bytes(hidden);
  if (hidden == 512) return launch_wide<8>(tm, p, grid, smem, stream);
  return launch_wide<0>(tm, p, grid, smem, stream);
}

}  // namespace nnam
