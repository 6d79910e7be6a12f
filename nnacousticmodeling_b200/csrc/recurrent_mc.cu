// K3, low-latency variant for the critical path: the CTA group is a thread-block CLUSTER and the hidden state is
// exchanged by TMA MULTICAST -- no release/acquire counter, no polling, no per-CTA pull of the whole tile.
//
// Reference semantics as in recurrent.cu (L.LSTM / F.lstm per Python loop iteration, scripts/common/chainer_networks.py:
// 44-62 driven by predict_folds.py:49-61).  LSTM family only: the GRU step with a reset gate is TWO dependent exchanges,
// and there this exchange (one L2 write + multicast read per phase) measured slower than the counter-based one of
// rnn_seq_kernel (8.0 k against 6.0 k cycles per step, profiles/r02_k3_phase_cycles.md), so the GRU family stays there.
//
// Why: a step of rnn_seq_kernel is a chain of latencies (profiles/r01_k3_phase_cycles.md): slice stores -> red.release.gpu
// (~1.1 k cycles) -> the peers' ld.acquire poll (~1 k) -> each of the 16 CTAs pulls the WHOLE h tile back from L2 by TMA
// (~1.5-2.5 k) -> MMA -> gates.  On sets whose longest utterance is the critical path (cfg3 / cfg4: 785 steps per layer
// and direction) that chain IS the run time.  Here every CTA, right after its gate math,
//   1. writes its slice of the new h (NB utterances x 32 units) to a private 2 KB spot of the exchange buffer,
//   2. and has ONE thread issue ONE `cp.async.bulk ... .multicast::cluster`: L2 -> the operand tile of ALL CTAs of the
//      cluster, completing `tx` bytes on the mbarrier at the same offset in every CTA.
// A CTA starts step s+1 when its own mbarrier has collected the 16 slices: one L2 round trip, no counter, and the L2 ->
// SM traffic of a group falls 16x.  (The direct CTA -> CTA push of round 1's cluster experiment moved the same bytes at
// the ~20 B/clk an SM can push into distributed shared memory: 3.5 k cycles per step.)
//
// The slice must be contiguous in the operand tile for a 1-D bulk copy, so the h tile (B operand, N = NB utterances,
// K = H) uses the canonical NO-SWIZZLE K-major layout with the 16-byte k-chunks outermost:
//     byte offset of (utterance r, unit k) = (k / 8) * NB * 16 + r * 16 + (k % 8) * 2
// (core matrix = 8 rows x 16 B contiguous; LBO = NB * 16 between k-chunks, SBO = 128 between 8-row groups).  CTA `rank`
// owns units [32 rank, 32 rank + 32) = k-chunks [4 rank, 4 rank + 4) = bytes [rank * NB * 64, (rank + 1) * NB * 64).
//
// Tiles and mbarriers are double-buffered by the parity of the cluster's global step counter g: step g reads tile g & 1
// and its slices land in tile (g + 1) & 1.  WAR-safe without a second handshake: a peer multicasts the slice of step g
// only after it has received EVERY slice of step g - 1, and a CTA sends its slice of step g - 1 after the MMAs of that
// step -- the last readers of tile (g + 1) & 1 -- have completed.
#include <cooperative_groups.h>

#include "recurrent_common.cuh"

namespace nnam {

// K-major operand without swizzle: LBO = byte distance of neighbouring 16-byte k-chunks, SBO = of 8-row groups
__device__ __forceinline__ uint64_t make_nosw_kmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  return d;                             // layout type 0 = SWIZZLE_NONE
}

// global -> shared memory of every CTA in `mask`, same CTA-relative destination and mbarrier offsets everywhere
__device__ __forceinline__ void bulk_load_1d_multicast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                                       uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

constexpr int MC_THREADS = 512;  // 16 warps: 4 warps per TMEM lane quarter, each owning NB / 4 utterance slots
constexpr int MC_ROWS = 128;     // gate rows per CTA (32 whole units)
constexpr int MC_UNITS = 32;

// NB: utterance slots per batch (32).  KBT: H / 64.
template <int NB, int KBT>
__global__ void __launch_bounds__(MC_THREADS, 1)
    rnn_seq_mc_kernel(const __grid_constant__ RnnTmaps tmaps, const RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int H = KBT * 64;
  constexpr int W_BLOCK = MC_ROWS * 128;
  constexpr int NBT = NB / 4;                 // slots per thread
  constexpr int SLICE = NB * MC_UNITS * 2;    // bytes of one CTA's slice of the tile
  constexpr int TILE = NB * H * 2;
  constexpr int LBO = NB * 16;
  constexpr int TILES = 2;
  constexpr int BASE_SMEM = RNN_BASE_SMEM;
  const int G = p.group_ctas;                 // == cluster size
  const int group = blockIdx.x / G;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool active = group < p.n_groups;     // uniform over the cluster
  announce_started(p);

  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // SWIZZLE_128B weight blocks need 1024-byte alignment
  uint8_t* w_s = smem_raw;                    // KBT * W_BLOCK, SWIZZLE_128B (TMA)
  uint8_t* tiles = w_s + KBT * W_BLOCK;       // TILES operand tiles
  uint8_t* stage = tiles + TILES * TILE;      // my freshly computed slice, in tile layout (SLICE bytes)
  uint8_t* tail = stage + SLICE;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_mma = bar_w + 1;
  uint64_t* bar_x = bar_mma + 1;              // [TILES]: "every slice of this tile has landed"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_x + TILES);
  int* s_len = reinterpret_cast<int*>(tmem_slot + 2);
  int* s_base = s_len + NB;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int quarter = warp & 3;
  const int u_lo = (warp >> 2) * NBT;
  constexpr int TMEM_COLS = NB < 32 ? 32 : NB;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    for (int i = 0; i < TILES; ++i) mbar_init(&bar_x[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // every CTA's mbarriers are initialised before any multicast can signal them
  cluster_arrive_release();
  cluster_wait_acquire();

  const int my_row = quarter * 32 + lane;
  const int gate = my_row & 3;          // a, i, f, o
  const int ul = my_row >> 2;           // unit inside the CTA's slice
  const int gate_col = rank * MC_ROWS + my_row;
  const uint32_t tmem_lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + u_lo;
  const int f16 = p.f16;
  const uint32_t idesc = make_idesc_e16_f32(MC_ROWS, NB, f16);
  const uint32_t stage_off = static_cast<uint32_t>((ul >> 3) * LBO + (ul & 7) * 2);  // + u * 16
  const uint16_t mask = static_cast<uint16_t>((1u << G) - 1u);
  // my private spots in the exchange buffer: [group][tile][rank] slices
  uint8_t* xg = reinterpret_cast<uint8_t*>(p.xchg_hi) + (static_cast<size_t>(group) * TILES * G + rank) * SLICE;

  const bool prof_on = p.prof != nullptr && tid == 64;
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#define PROF_START() do { if (prof_on) prof_t = clock64(); } while (0)
#define PROF_MARK(i) do { if (prof_on) { const long long now = clock64(); prof_acc[i] += now - prof_t; prof_t = now; } } while (0)

  unsigned int g = 0;            // h exchanges so far (tile / barrier parity); runs across work items
  uint32_t w_phase = 0, mma_phase = 0;
  int cur_dir = -1;

  // D[128 x NB] = W_slice . tile^T : issued by one elected lane of warp 0 once the tile's barrier has completed
  auto mma_issue = [&](int tile_idx, uint32_t parity) {
    if (warp == 0) {
      mbar_wait(&bar_x[tile_idx], parity);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t wd = make_sw128_kmajor_desc(smem_u32(w_s));
        const uint64_t hd = make_nosw_kmajor_desc(smem_u32(tiles + tile_idx * TILE), LBO, 128);
        uint32_t accum = 0;
#pragma unroll
        for (int kb = 0; kb < KBT; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // A: 32 bytes further inside the 128-byte swizzle row; B: two k-chunks (2 * LBO bytes) further
            umma_bf16(tmem_base, wd + ((kb * W_BLOCK + k * 32) >> 4), hd + (((kb * 4 + k) * 2 * LBO) >> 4), idesc, accum);
            accum = 1;
          }
        umma_commit(bar_mma);
      }
      __syncwarp();
    }
  };
  auto mma_collect = [&](float (&acc)[NBT]) {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    uint32_t r[8];
#pragma unroll
    for (int c0 = 0; c0 < NBT; c0 += 8) {
      tmem_ld8(tmem_lane_addr + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  };
  // My slice is complete in `stage` (tile layout).  Critical path: warp 0 alone copies the 2 KB to my spot of exchange tile
  // `tile_idx` (4 x 16 bytes per lane) and one of its lanes multicasts it from there into tile `tile_idx` of every CTA of
  // the cluster -- no second CTA barrier.  The other warps meanwhile store the slice into the layer output rows (if any).
  auto publish = [&](int tile_idx, __nv_bfloat16* dst_hi, int n_act, int s, const int* bp, long long row0, bool bwd,
                     int h_col0) {
    __syncthreads();  // staging writes of all warps are visible
    if (warp == 0) {
      uint8_t* spot = xg + static_cast<size_t>(tile_idx) * G * SLICE;
      constexpr int CHUNKS = SLICE / 16;  // NB * 4
#pragma unroll
      for (int q = lane; q < CHUNKS; q += 32)
        *reinterpret_cast<uint4*>(spot + q * 16) = *reinterpret_cast<const uint4*>(stage + q * 16);
      __syncwarp();  // orders the lanes' global stores before the elected lane's proxy fence + bulk copy
      if (elect_one()) {
        fence_proxy_async_all();  // generic-proxy global writes -> visible to the async proxy (the bulk copy's read)
        bulk_load_1d_multicast(tiles + tile_idx * TILE + rank * SLICE, spot, SLICE, &bar_x[tile_idx], mask);
      }
      __syncwarp();
    } else if (dst_hi != nullptr) {
      for (int q = tid - 32; q < n_act * 4; q += MC_THREADS - 32) {
        const int u = q >> 2, j = q & 3;  // 16-byte chunk j (8 units) of utterance u's 64-byte slice
        const uint4 v = *reinterpret_cast<const uint4*>(stage + j * LBO + u * 16);
        const int t_idx = bwd ? (s_len[u] - 1 - s) : s;
        *reinterpret_cast<uint4*>(dst_hi + (row0 + bp[t_idx] + u) * p.h_ld + h_col0 + rank * MC_UNITS + j * 8) = v;
      }
    }
  };

  if (active) {
    // arm the barriers of the first exchanges
    if (tid == 0) {
      mbar_expect_tx(&bar_x[1], static_cast<uint32_t>(TILE));  // slices of h step 0 land in tile 1
    }
    for (int it = p.group_item_start[group]; it < p.group_item_start[group + 1]; ++it) {
      const int b = p.item_batch[it];
      const int d = p.item_dir[it];
      const bool bwd = d == 1;
      if (d != cur_dir) {
        __syncthreads();
        if (tid == 0) {
          mbar_expect_tx(bar_w, static_cast<uint32_t>(KBT * W_BLOCK));
          for (int kb = 0; kb < KBT; ++kb) tma_load_2d(w_s + kb * W_BLOCK, &tmaps.w_hi[d], bar_w, kb * 64, rank * MC_ROWS);
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
      const uint16_t* gx = reinterpret_cast<const uint16_t*>(p.gx[d]) + gate_col;
      const int h_col0 = d * H;
      __syncthreads();
      if (tid < NB) s_len[tid] = tid < nutt ? len[tid] : 0;
      const bool base_in_smem = BASE_SMEM > 0 && T <= BASE_SMEM;
      if (base_in_smem)
        for (int i = tid; i <= T; i += MC_THREADS) s_base[i] = __ldg(base + i);
      __syncthreads();
      const int* bp = base_in_smem ? s_base : base;

      float st_reg[NBT / 4];  // cell state c of (utterance u_lo + 4m + gate, my unit)
#pragma unroll
      for (int m = 0; m < NBT / 4; ++m) st_reg[m] = 0.0f;

      // raw 16-bit loads; the conversion waits until the values are used, one step later, so the (DRAM-latency) loads
      // never stall the thread that issued them
      auto load_gx = [&](int s, uint16_t (&dst)[NBT]) {
        const int base_s = bp[s];
        const int n_act = bp[s + 1] - base_s;
#pragma unroll
        for (int j = 0; j < NBT; ++j) {
          const int u = u_lo + j;
          const int uu = u < n_act ? u : n_act - 1;
          const long long row = row0 + (bwd ? bp[s_len[uu] - 1 - s] : base_s) + uu;
          dst[j] = __ldg(gx + row * p.gx_ld);
        }
      };
      auto stage_put = [&](int u, float v) {
        *reinterpret_cast<uint16_t*>(stage + stage_off + u * 16) = f32_to_e16(v, f16);
      };

      uint16_t gxr[NBT], gxn[NBT];
      load_gx(0, gxr);
      for (int s = 0; s < T; ++s, ++g) {
        PROF_START();
        const int n_s = bp[s + 1] - bp[s];
        float acc[NBT];
        // slices of h step g-1 are in tile g & 1, announced on bar_x[g & 1]; its use before was step g-3, consumed at g-2
        const int cur = static_cast<int>(g & 1u);
        if (g > 0) {
          if (s > 0) {
            mma_issue(cur, ((g - 1) >> 1) & 1u);
          } else if (warp == 0) {
            mbar_wait(&bar_x[cur], ((g - 1) >> 1) & 1u);  // last step of the previous item: keep the phases in step
          }
        }
        // arm the barrier the slices of THIS step will complete (tile / barrier (g + 1) & 1; previous use consumed above
        // one step ago).  Step 0 of the launch was armed before the loop.
        if (tid == 0 && g > 0) mbar_expect_tx(&bar_x[cur ^ 1], static_cast<uint32_t>(TILE));
        PROF_MARK(0);
        if (s + 1 < T) load_gx(s + 1, gxn);
        PROF_MARK(1);
        if (s > 0) {
          mma_collect(acc);
        } else {
#pragma unroll
          for (int j = 0; j < NBT; ++j) acc[j] = 0.0f;
        }
        PROF_MARK(2);
        __syncthreads();  // every warp has finished reading last step's staging slice (output rows / exchange spot)

        {
          // ---- chainer F.lstm: c = tanh(a) s(i) + s(f) c; h = s(o) tanh(c); s(x) = tanh(x/2)/2 + 1/2
#pragma unroll
          for (int m = 0; m < NBT / 4; ++m) {
            if (u_lo + 4 * m >= n_s) break;  // warp-uniform
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v = acc[4 * m + i] + e16_to_f32(gxr[4 * m + i], f16);
              const float t = tanh_fast(gate == 0 ? v : 0.5f * v);
              x[i] = gate == 0 ? t : fmaf(t, 0.5f, 0.5f);
            }
            quad_transpose(x, gate);  // x = {a, i, f, o} of utterance u = u_lo + 4m + gate
            const int u = u_lo + 4 * m + gate;
            const float c_new = fmaf(x[0], x[1], x[2] * st_reg[m]);
            const float h_new = x[3] * tanh_fast(c_new);
            if (u < n_s) {
              st_reg[m] = c_new;
              stage_put(u, h_new);
            }
          }
          PROF_MARK(3);
          publish(cur ^ 1, p.h_hi, n_s, s, bp, row0, bwd, h_col0);
          PROF_MARK(4);
        }
#pragma unroll
        for (int j = 0; j < NBT; ++j) gxr[j] = gxn[j];
      }
    }
    // the slices of the last step must have landed before anyone leaves: peers multicast into this CTA's shared memory
    if (g > 0 && warp == 0) mbar_wait(&bar_x[g & 1u], ((g - 1) >> 1) & 1u);
  }
  if (prof_on) {
    prof_acc[7] = g;
    for (int i = 0; i < 8; ++i) p.prof[blockIdx.x * 8 + i] = prof_acc[i];
  }
#undef PROF_START
#undef PROF_MARK
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
static size_t mc_smem_bytes(int nb, int hidden) {
  return static_cast<size_t>(hidden / 64) * MC_ROWS * 128 + 2 * static_cast<size_t>(nb) * hidden * 2 +
         static_cast<size_t>(nb) * 64 + 128 + static_cast<size_t>(nb) * 4 + (RNN_BASE_SMEM + 1) * 4 + 16;
}

template <int NB, int KBT>
static int mc_config(int G, cudaStream_t stream, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid) {
  auto kern = rnn_seq_mc_kernel<NB, KBT>;
  const size_t smem = mc_smem_bytes(NB, KBT * 64);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "rnn(mc): cudaFuncSetAttribute(smem)");
  if (G > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return set_cuda_error(e, "rnn(mc): cudaFuncSetAttribute(non-portable cluster)");
  }
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = dim3(grid);
  cfg->blockDim = dim3(MC_THREADS);
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

template <int NB, int KBT>
static int mc_max_groups(int G) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  if (mc_config<NB, KBT>(G, nullptr, &cfg, attr, G) != NNAM_OK) return 0;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rnn_seq_mc_kernel<NB, KBT>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <int NB, int KBT>
static int mc_launch(const RnnTmaps& tm, const RnnParams& p, int G, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  const int rc = mc_config<NB, KBT>(G, stream, &cfg, attr, p.n_groups * G);
  if (rc) return rc;
  cudaError_t e = cudaLaunchKernelEx(&cfg, rnn_seq_mc_kernel<NB, KBT>, tm, p);
  if (e != cudaSuccess) return set_cuda_error(e, "rnn(mc): cudaLaunchKernelEx");
  return NNAM_OK;
}

#define NNAM_MC_DISPATCH(FN, ...)                               \
  do {                                                          \
    if (hidden == 512) return FN<32, 8>(__VA_ARGS__);           \
    if (hidden == 256) return FN<32, 4>(__VA_ARGS__);           \
  } while (0)

static int mc_max_groups_dispatch(int hidden, int G) {
  NNAM_MC_DISPATCH(mc_max_groups, G);
  return 0;
}

// Resident clusters of the multicast variant for (cell, hidden, slots, precision), 0 if it does not apply:
// LSTM family, single-pass 16-bit operands, 32 slots, H in {256, 512} (cluster of 8 / 16 CTAs), no carried state.
// NNAM_RNN_MC=0 switches it off (A/B measurements).
int rnn_mc_groups(int cell, int hidden, int batch, int nsplit) {
  static const int enabled = [] {
    const char* v = getenv("NNAM_RNN_MC");
    return (v != nullptr && v[0] == '0') ? 0 : 1;
  }();
  if (!enabled || nsplit != 1 || batch != 32 || (hidden != 256 && hidden != 512)) return 0;
  if (cell != NNAM_CELL_LSTM) return 0;
  if (mc_smem_bytes(batch, hidden) > 227 * 1024) return 0;
  const int G = 4 * hidden / MC_ROWS;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  static int cached[64][2] = {};  // [device][hidden == 512]; racing writers store the same value
  int& c = cached[dev][hidden == 512 ? 1 : 0];
  if (c == 0) {
    const int n = mc_max_groups_dispatch(hidden, G);
    c = n > 0 ? n : -1;
  }
  return c > 0 ? c : 0;
}

int rnn_mc_launch(int cell, const RnnTmaps& tm, const RnnParams& p, int G, int hidden, cudaStream_t stream) {
  if (cell != NNAM_CELL_LSTM) return set_error(NNAM_ERR_UNSUPPORTED, "rnn(mc): LSTM family only");
  NNAM_MC_DISPATCH(mc_launch, tm, p, G, stream);
  return set_error(NNAM_ERR_UNSUPPORTED, "rnn(mc): no instance for this configuration");
}

}  // namespace nnam
