// K3, experimental variant: LSTM recurrence with the group mapped onto a thread-block cluster and the hidden state
// exchanged through distributed shared memory.  Opt-in (NNAM_RNN_CLUSTER=1): measured slower than the default L2
// exchange on B200, see profiles/r01_k3_phase_cycles.md.  Reference semantics as in recurrent.cu
// (scripts/common/chainer_networks.py:44-62 via predict_folds.py:49-61).
// Not part of the default build: add -DNNAM_WITH_CLUSTER_EXPERIMENT and this file (NNAM_WITH_CLUSTER_EXPERIMENT=1 in the
// environment of nnacousticmodeling_b200._native.build) to compile it in.
#include "../recurrent_common.cuh"

namespace nnam {

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
      const __nv_bfloat16* gx = reinterpret_cast<const __nv_bfloat16*>(p.gx[d]) + gate_col;  // bf16 mode only
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
          gxr[j] = __bfloat162float(__ldg(gx + row * p.gx_ld));
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
int rnn_cluster_groups(int cell, int hidden, int batch, int nsplit) {
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


int rnn_cluster_launch(const RnnTmaps& tm, const RnnParams& p, int G, int hidden, cudaStream_t stream) {
  if (hidden == 512) return launch_cluster<32, 8>(tm, p, G, hidden, stream);
  if (hidden == 256) return launch_cluster<32, 4>(tm, p, G, hidden, stream);
  return launch_cluster<32, 0>(tm, p, G, hidden, stream);
}

}  // namespace nnam
