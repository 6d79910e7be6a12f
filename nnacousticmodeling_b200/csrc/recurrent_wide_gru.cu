// K3, "wide" variant for the GRU family in bf16 mode: 128 utterances per batch on the MMA's M axis, two interleaved
// batches (streams) per CTA group -- the structure of lstm_seq_wide2_kernel (recurrent_wide.cu) applied to
// MGRU.py:67-85 / L.GRU (scripts/common/chainer_networks.py:64-101 via predict_folds.py:49-61):
//
//     z = s(W_z x + U_z h)      r = s(W_r x + U_r h)      hb = act(W x + U (r * h))      h' = (1 - z) h + z hb
//
// With the reset gate the product r * h has to go round the group before the candidate can be formed: two exchanges
// and two tensor-core passes per step.  To keep the second pass small the rows of this kernel's weight slice are
// GATE-BLOCKED instead of unit-interleaved: a CTA owns 32 units and holds the rows [z(32) | r(32) | cand(32)]
// (NG = 3, 96 rows) or [z(32) | cand(32)] without reset gate (NG = 2, 64 rows).  Pass 1 multiplies the first 64 rows
// (N = 64), pass 2 only the candidate block (N = 32).  The input projection gx uses the same column order, so a
// thread reads the 16 units it owns as ONE 256-bit load per gate, and there are no padding rows (the narrow kernel
// carries a zero fourth row per unit).  The U biases are folded into the projection bias by the host; step 0, where
// MGRU has no U terms at all (MGRU.py:70-83), subtracts them again.
#include <cooperative_groups.h>

#include "recurrent_common.cuh"

namespace nnam {

constexpr int GW_NB = 128;
constexpr int GW_STREAMS = 2;
constexpr int GW_TPS = 256;
constexpr int GW_THREADS = GW_STREAMS * GW_TPS;
constexpr int GW_MAX_STAGES = 8;
constexpr int GW_A_STAGE = GW_NB * 128;
constexpr int GW_BASE_SMEM = 1024;

__device__ __forceinline__ float gw_sigmoid(float v) { return fmaf(tanh_fast(0.5f * v), 0.5f, 0.5f); }
__device__ __forceinline__ void gw_bar_sync(int stream) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + stream), "n"(GW_TPS) : "memory");
}
__device__ __forceinline__ float gw_act(int kind, float v) {
  return kind == NNAM_ACT_RELU ? fmaxf(v, 0.0f)
                               : (kind == NNAM_ACT_SIGMOID ? gw_sigmoid(v) : (kind == NNAM_ACT_TANH ? tanh_fast(v) : v));
}

template <int NG, int KBT>
__global__ void __launch_bounds__(GW_THREADS, 1)
    gru_seq_wide2_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p, const int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int ROWS = 32 * NG;  // weight rows per CTA
  constexpr int UNITS = 32;
  constexpr int NB = GW_NB;
  constexpr bool RESET = NG == 3;
  const int group = blockIdx.x / p.group_ctas;
  const int rank = blockIdx.x % p.group_ctas;
  announce_started(p);
  if (group >= p.n_groups) return;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int H = KBT > 0 ? KBT * 64 : p.hidden;
  const int KB = KBT > 0 ? KBT : (H >> 6);
  constexpr int W_BLOCK = ROWS * 128;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int stream = warp >> 3;
  const int sw = warp & 7;
  const int stid = tid & (GW_TPS - 1);
  const int quarter = sw & 3;
  const int sub = sw >> 2;             // which 16 of the CTA's 32 units
  const int u = quarter * 32 + lane;   // my utterance slot == my TMEM lane

  uint8_t* w_s = smem;                      // KB blocks of [ROWS x 128 B], SWIZZLE_128B (B operand)
  uint8_t* a_s = w_s + KB * W_BLOCK;        // ring of `stages` k-blocks of the h (or r*h) tile (A operand)
  uint8_t* tail = a_s + stages * GW_A_STAGE;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;                                        // [2]
  uint64_t* bar_grant = bar_mma + GW_STREAMS;                           // [2]
  uint64_t* full_bar_all = bar_grant + GW_STREAMS;                      // [2][GW_MAX_STAGES], per stream (see wide2)
  uint64_t* full_bar = full_bar_all + stream * GW_MAX_STAGES;
  uint64_t* empty_bar = full_bar_all + GW_STREAMS * GW_MAX_STAGES;      // [GW_MAX_STAGES], shared
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_bar + GW_MAX_STAGES);
  unsigned int* ring_lock = tmem_slot + 2;
  volatile unsigned int* ring_fill = ring_lock + 1;
  volatile unsigned int* s_f0 = ring_lock + 2;  // [2]
  int* s_len = reinterpret_cast<int*>(ring_lock + 6) + stream * NB;
  int* s_base = reinterpret_cast<int*>(ring_lock + 6) + GW_STREAMS * NB + stream * (GW_BASE_SMEM + 1);
  constexpr int TMEM_COLS = GW_STREAMS * 128;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < GW_STREAMS; ++i) {
      mbar_init(&bar_mma[i], 1);
      mbar_init(&bar_grant[i], 1);
    }
    for (int i = 0; i < GW_STREAMS * GW_MAX_STAGES; ++i) mbar_init(&full_bar_all[i], 1);
    for (int i = 0; i < GW_MAX_STAGES; ++i) mbar_init(&empty_bar[i], 1);
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
  // my lane quarter, my 16 units inside a 32-column gate block
  const uint32_t tmem_mine = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(sub * 16);
  const int f16 = p.f16;
  const uint32_t idesc1 = make_idesc_e16_f32(NB, 64, f16);  // pass 1: [z | r] or [z | cand]
  const uint32_t idesc2 = make_idesc_e16_f32(NB, 32, f16);  // pass 2: cand block on r*h
  const int act_kind = (p.gru_flags >> 1) & 3;

  const bool prof_on = p.prof != nullptr && tid == 64;
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#define PROF_START() do { if (prof_on) prof_t = clock64(); } while (0)
#define PROF_MARK(i) do { if (prof_on) { const long long now = clock64(); prof_acc[i] += now - prof_t; prof_t = now; } } while (0)

  unsigned int pubs_done = 0;  // publishes of this stream so far (the group counter reaches pubs_done * G)
  unsigned int rec_steps = 0;
  uint32_t w_phase = 0, mma_phase = 0, grant_phase = 0;
  uint32_t full_parity = 0;
  const int lane_id_ = group * GW_STREAMS + stream;
  unsigned int* counter = p.counters + lane_id_;
  int it = p.group_item_start[lane_id_];
  const int it_end = p.group_item_start[lane_id_ + 1];

  // One tile (exchange slot `xslot`) through the shared ring into `n_mma`-wide MMAs against weight rows
  // [brow, brow + N) of the slice; accumulator columns [dcol, dcol + N).  Producer = warp 0, issuer = warp 1 of
  // the stream; `hook` runs once in each of the two warps at a point where it delays nothing (their gx loads).
  auto ring_pass = [&](int xslot, int brow, uint32_t idesc, uint32_t dcol, auto&& hook) {
    if (sw == 0) {
      const unsigned int target = pubs_done * static_cast<unsigned int>(p.group_ctas);
      while (ld_acquire_gpu(counter) < target) {
      }
      unsigned int f0 = 0;
      if (lane == 0) {
        fence_proxy_async_all();
        while (atomicCAS(ring_lock, 0u, 1u) != 0u) {
        }
        __threadfence_block();
        f0 = *ring_fill;
        s_f0[stream] = f0;
        __threadfence_block();
        mbar_arrive(&bar_grant[stream]);
      }
      f0 = __shfl_sync(0xffffffffu, f0, 0);
      const int xrow = (lane_id_ * 4 + xslot) * NB;
      for (int kb = 0; kb < KB; ++kb) {
        const unsigned int f = f0 + kb;
        const int st = static_cast<int>(f % static_cast<unsigned int>(stages));
        const uint32_t ph = (f / static_cast<unsigned int>(stages)) & 1u;
        mbar_wait(&empty_bar[st], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[st], GW_A_STAGE);
          tma_load_2d(a_s + st * GW_A_STAGE, &tmaps.x_hi, &full_bar[st], kb * 64, xrow);
        }
        __syncwarp();
      }
      if (lane == 0) {
        *ring_fill = f0 + KB;
        __threadfence_block();
        atomicExch(ring_lock, 0u);
      }
      __syncwarp();
      hook();
    } else if (sw == 1) {
      mbar_wait(&bar_grant[stream], grant_phase);
      const unsigned int f0 = s_f0[stream];
      const uint64_t adesc0 = make_sw128_kmajor_desc(smem_u32(a_s));
      const uint64_t bdesc0 = make_sw128_kmajor_desc(smem_u32(w_s) + static_cast<uint32_t>(brow * 128));
      for (int kb = 0; kb < KB; ++kb) {
        const unsigned int f = f0 + kb;
        const int st = static_cast<int>(f % static_cast<unsigned int>(stages));
        mbar_wait(&full_bar[st], (full_parity >> st) & 1u);
        full_parity ^= 1u << st;
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>((st * GW_A_STAGE) >> 4);
          const uint64_t bd = bdesc0 + static_cast<uint64_t>((kb * W_BLOCK) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_acc + dcol, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[st]);
          if (kb == KB - 1) umma_commit(&bar_mma[stream]);
        }
        __syncwarp();
        if (kb == 0) hook();
      }
    }
    grant_phase ^= 1;
    mbar_wait(&bar_mma[stream], mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };
  auto publish = [&]() {
    gw_bar_sync(stream);
    if (stid == 0) red_release_gpu_add(counter, 1u);
    ++pubs_done;
  };

  for (int d = 0; d < p.n_dirs; ++d) {
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar_w, static_cast<uint32_t>(KB * W_BLOCK));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * ROWS);
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
      // my 16 units of gate block g start at column rank * ROWS + g * 32 + sub * 16 of the projection
      const __nv_bfloat16* gx = reinterpret_cast<const __nv_bfloat16*>(p.gx[d]) + rank * ROWS + sub * 16;
      const float* ubias = p.u_bias[d] + rank * ROWS + sub * 16;
      if (pubs_done > 0 && sw == 0) {
        const unsigned int target = pubs_done * static_cast<unsigned int>(p.group_ctas);
        while (ld_acquire_gpu(counter) < target) {
        }
      }
      gw_bar_sync(stream);
      if (stid < NB) s_len[stid] = stid < nutt ? len[stid] : 0;
      const bool base_in_smem = T <= GW_BASE_SMEM;
      if (base_in_smem)
        for (int i = stid; i <= T; i += GW_TPS) s_base[i] = __ldg(base + i);
      gw_bar_sync(stream);
      const int* bp = base_in_smem ? s_base : base;
      const int my_len = s_len[u];

      float h_reg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) h_reg[j] = 0.0f;

      for (int s = 0; s < T; ++s) {
        PROF_START();
        const int n_s = bp[s + 1] - bp[s];
        const bool active = u < n_s;
        const int t_idx = bwd ? (my_len - 1 - s) : s;
        const long long my_row = active ? row0 + bp[t_idx] + u : 0;
        uint32_t gxr[NG][8] = {};
        auto load_gx = [&]() {
          if (active) {
            const __nv_bfloat16* src = gx + my_row * p.gx_ld;
#pragma unroll
            for (int g = 0; g < NG; ++g) ldg_nc_256(src + 32 * g, gxr[g]);
          }
        };
        auto gxv = [&](int g, int j) -> float {  // pre-activation of gate block g, my unit j, from the projection
          const float2 v = e16x2_to_float2(gxr[g][j >> 1], f16);
          return (j & 1) ? v.y : v.x;
        };
        if (s == 0 || sw >= 2) load_gx();
        PROF_MARK(0);

        float z[16];
        float hb[16];
        if (s == 0) {
          // no U terms and no U biases at the first step (MGRU.py:70-83): take the folded biases out again
          if (active) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              z[j] = gw_sigmoid(gxv(0, j) - __ldg(ubias + j));
              hb[j] = gw_act(act_kind, gxv(NG - 1, j) - __ldg(ubias + (NG - 1) * 32 + j));
            }
          }
        } else {
          ring_pass((s - 1) & 1, 0, idesc1, 0u, load_gx);  // D[:, 0:64) = h_{s-1} . [U_z | U_r or U]^T
          PROF_MARK(1);
          uint32_t r16[16];
          tmem_ld16(tmem_mine, r16);
          tmem_ld_wait();
          if (active) {
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = gw_sigmoid(__uint_as_float(r16[j]) + gxv(0, j));
          }
          tmem_ld16(tmem_mine + 32u, r16);
          tmem_ld_wait();
          tc_fence_before();
          if (RESET) {
            uint32_t rh[8];
            if (active) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float r = gw_sigmoid(__uint_as_float(r16[j]) + gxv(1, j));
                const uint32_t v = static_cast<uint32_t>(f32_to_e16(r * h_reg[j], f16));
                if (j & 1)
                  rh[j >> 1] |= v << 16;
                else
                  rh[j >> 1] = v;
              }
              const long long xoff =
                  (static_cast<long long>(lane_id_ * 4 + 2 + (s & 1)) * NB + u) * H + rank * UNITS + sub * 16;
              stg_256(p.xchg_hi + xoff, rh);
            }
            PROF_MARK(2);
            publish();  // r * h slices are out
            PROF_MARK(3);
            ring_pass(2 + (s & 1), 64, idesc2, 64u, [] {});  // D[:, 64:96) = (r * h) . U^T
            PROF_MARK(4);
            tmem_ld16(tmem_mine + 64u, r16);
            tmem_ld_wait();
            tc_fence_before();
          }
          if (active) {
#pragma unroll
            for (int j = 0; j < 16; ++j) hb[j] = gw_act(act_kind, __uint_as_float(r16[j]) + gxv(NG - 1, j));
          }
        }

        // ---- h' = (1 - z) h + z hb  (h = 0 at the first step), out through the exchange slot and the layer output
        if (active) {
          uint32_t hp[8];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float h_new = fmaf(z[j], hb[j], (1.0f - z[j]) * h_reg[j]);
            h_reg[j] = h_new;
            const uint32_t v = static_cast<uint32_t>(f32_to_e16(h_new, f16));
            if (j & 1)
              hp[j >> 1] |= v << 16;
            else
              hp[j >> 1] = v;
          }
          const int col = rank * UNITS + sub * 16;
          const long long xoff = (static_cast<long long>(lane_id_ * 4 + (s & 1)) * NB + u) * H + col;
          stg_256(p.xchg_hi + xoff, hp);
          stg_256(p.h_hi + my_row * p.h_ld + h_col0 + col, hp);
        }
        PROF_MARK(5);
        publish();
        ++rec_steps;
        PROF_MARK(6);
      }
    }
  }

  if (prof_on) {
    prof_acc[7] = rec_steps;
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

static size_t gru_wide_fixed_smem() {
  return 8 * (1 + 2 * GW_STREAMS + (GW_STREAMS + 1) * GW_MAX_STAGES) + 8 * 4 +
         GW_STREAMS * (GW_NB + GW_BASE_SMEM + 1) * 4;
}

// ring slots that fit next to the weight slice (0: the configuration does not fit)
int rnn_wide_gru_stages(int hidden, int n_gates) {
  const size_t w = static_cast<size_t>(hidden / 64) * (32 * n_gates) * 128;
  const size_t budget = 227 * 1024;
  if (w + gru_wide_fixed_smem() + 2 * GW_A_STAGE > budget) return 0;
  const size_t n = (budget - w - gru_wide_fixed_smem()) / GW_A_STAGE;
  return static_cast<int>(n > GW_MAX_STAGES ? GW_MAX_STAGES : n);
}

bool rnn_wide_gru_applies(int hidden, int nsplit, int n_gates) {
  if (nsplit != 1 || hidden % 64 || hidden % 32) return false;
  if (rnn_wide_gru_stages(hidden, n_gates) < 2) return false;
  return sm_count() >= hidden / 32;
}

template <int NG, int KBT>
static int launch_gru_wide(const RnnTmaps& tm, const RnnParams& p, int grid, int stages, size_t smem,
                           cudaStream_t stream) {
  auto kern = gru_seq_wide2_kernel<NG, KBT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaFuncSetAttribute(gru wide)");
  void* args[] = {const_cast<RnnTmaps*>(&tm), const_cast<RnnParams*>(&p), &stages};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(GW_THREADS), args, smem, stream);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn: cudaLaunchCooperativeKernel(gru wide)");
  return NNAM_OK;
}

int rnn_wide_gru_launch(const RnnTmaps& tm, const RnnParams& p, int hidden, int n_gates, cudaStream_t stream) {
  const int grid = p.n_groups * p.group_ctas;
  const int stages = rnn_wide_gru_stages(hidden, n_gates);
  const size_t smem = static_cast<size_t>(hidden / 64) * (32 * n_gates) * 128 +
                      static_cast<size_t>(stages) * GW_A_STAGE + gru_wide_fixed_smem();
  if (n_gates == 3) {
    if (hidden == 512) return launch_gru_wide<3, 8>(tm, p, grid, stages, smem, stream);
    return launch_gru_wide<3, 0>(tm, p, grid, stages, smem, stream);
  }
  if (hidden == 512) return launch_gru_wide<2, 8>(tm, p, grid, stages, smem, stream);
  return launch_gru_wide<2, 0>(tm, p, grid, stages, smem, stream);
}

}  // namespace nnam
