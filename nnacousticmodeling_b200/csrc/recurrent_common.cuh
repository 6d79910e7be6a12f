// Shared declarations of the K3 recurrence kernels (recurrent.cu, recurrent_wide.cu, recurrent_cluster.cu).
#pragma once
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

constexpr int RNN_THREADS = 256;  // cluster variant; the default kernels use S * WPS * 32 threads

struct RnnTmaps {
  CUtensorMap w_hi[2];
  CUtensorMap w_lo[2];
  CUtensorMap x_hi;  // exchange buffer (lanes * 4 * NB rows, H columns): box {64, NB}
  CUtensorMap x_lo;
};

struct RnnParams {
  int hidden;     // H
  int n_dirs;
  int n_groups;   // groups that have work
  int group_ctas; // G
  long long gx_ld, h_ld;
  const void* gx[2];      // per direction: (rows, gx_ld) gate-interleaved columns; fp32 (nsplit 3) or bf16 (nsplit 1)
  const float* u_bias[2]; // GRU family only
  __nv_bfloat16* h_hi;    // (rows, h_ld); direction d owns columns [d*H, (d+1)*H)
  __nv_bfloat16* h_lo;
  __nv_bfloat16* xchg_hi;  // (lanes * 4 * NB, H): per lane 4 slots of NB rows -- h parity 0/1, r*h parity 0/1
  __nv_bfloat16* xchg_lo;
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
  int f16;                     // 16-bit element type of w / h / xchg / h0 / (nsplit 1) gx: 0 = bf16, 1 = fp16
  long long* prof;             // optional: 8 cycle accumulators per CTA (thread 0), phases of a step
  unsigned int* started;       // optional (pinned host memory): started[blockIdx.x] = started_tag when the CTA starts
  unsigned int started_tag;
};

// "this CTA is resident": one store into device-accessible host memory (NnamRnnDesc.started)
__device__ __forceinline__ void announce_started(const RnnParams& p) {
  if (p.started != nullptr && threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned int*>(p.started + blockIdx.x) = p.started_tag;
    __threadfence_system();
  }
}

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


// host-side pieces shared by the translation units
#ifdef NNAM_WITH_CLUSTER_EXPERIMENT  // csrc/experimental/recurrent_cluster.cu (DSMEM all-gather; measured slower, not built by default)
size_t rnn_cluster_smem_bytes(int nb, int hidden);
int rnn_cluster_groups(int cell, int hidden, int batch, int nsplit);
int rnn_cluster_launch(const RnnTmaps& tm, const RnnParams& p, int G, int hidden, cudaStream_t stream);
#else
inline int rnn_cluster_groups(int, int, int, int) { return 0; }
inline int rnn_cluster_launch(const RnnTmaps&, const RnnParams&, int, int, cudaStream_t) {
  return set_error(NNAM_ERR_UNSUPPORTED, "rnn: built without the cluster experiment");
}
#endif
// recurrent_mc.cu: cluster + TMA-multicast exchange, 32 slots per batch (the critical-path kernel)
int rnn_mc_groups(int cell, int hidden, int batch, int nsplit);
int rnn_mc_launch(int cell, const RnnTmaps& tm, const RnnParams& p, int G, int hidden, cudaStream_t stream);
// recurrent_wide.cu: LSTM / bf16 / 128 slots per batch
bool rnn_wide_applies(int cell, int hidden, int batch, int nsplit);
int rnn_wide_launch(const RnnTmaps& tm, const RnnParams& p, int hidden, cudaStream_t stream);
int rnn_wide_streams();
// recurrent_wide_gru.cu: GRU family / bf16 / 128 slots per batch, gate-blocked weight rows (n_gates = 3 with reset gate)
bool rnn_wide_gru_applies(int hidden, int nsplit, int n_gates);
int rnn_wide_gru_launch(const RnnTmaps& tm, const RnnParams& p, int hidden, int n_gates, cudaStream_t stream);  // batches a CTA group of the 128-slot kernel runs concurrently (2, or 1 with NNAM_RNN_WIDE_STREAMS=1)

}  // namespace nnam
