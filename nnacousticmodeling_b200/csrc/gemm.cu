// K2 -- Linear layer as a tcgen05/TMEM GEMM fed by TMA, bias + activation fused in the epilogue.
//
//   out[M, N] = act( A[M, K] . W[N, K]^T + bias[N] )
//
// Replaces every ``L.Linear`` call of the reference model specs (scripts/common/chainer_networks.py:16-22,
// 51-52, 61-62 ...; Chainer F.linear = x.dot(W.T) + b) and the batched input-to-hidden projections of the
// recurrent cells.  A and W are bf16, K-major (row-major with K contiguous), exactly the Chainer (out, in)
// weight layout, so no transposition is ever needed.
//
// Precision modes (nsplit):
//   1  "bf16"   : one pass, A_hi . W_hi
//   3  "bf16x3" : fp32-accurate.  A = A_hi + A_lo, W = W_hi + W_lo (each bf16); the kernel accumulates
//                 A_hi.W_hi + A_hi.W_lo + A_lo.W_hi into the same fp32 TMEM accumulator (the dropped
//                 A_lo.W_lo term is ~2^-18 relative).  Costs 3 bf16 passes = 1.5 TF32 passes but carries
//                 16 mantissa bits instead of TF32's 10.
//
// Structure: persistent, warp-specialised, one CTA per SM.
//   warp 0 : TMA producer (one lane)          -- smem ring of STAGES x (A 128x64 + W bn x64) bf16, SWIZZLE_128B
//   warp 1 : tcgen05.mma issuer (one lane)    -- UMMA 128 x bn x 16, fp32 accumulators in TMEM, double-buffered
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue -- tcgen05.ld 32x32b, + bias, activation, (hi/lo split), vector stores
// Tile order is n-fastest so that CTAs running concurrently share the A tile in L2 while W stays L2-resident.
#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per k-block = 128 B = one swizzle row
constexpr int UK = 16;  // K of one tcgen05.mma (kind::f16)
constexpr int MAX_BN = 256;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;      // 16 KiB
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;  // 32 KiB
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + 1024;  // + barriers + align slack
constexpr int TMEM_COLS = 512;

struct GemmParams {
  int M, N, K;
  int bn;         // tile width, multiple of 16, <= 256
  int tiles_m, tiles_n;
  int k_blocks;   // ceil(K / 64)
  int nsplit;     // 1 or 3
  int act;        // NNAM_ACT_*
  int out_kind;   // NNAM_OUT_*
  const float* bias;
  void* out_hi;
  void* out_lo;
  long long ldo;  // elements
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case NNAM_ACT_RELU: return fmaxf(v, 0.0f);
    case NNAM_ACT_SIGMOID: return tanhf(v * 0.5f) * 0.5f + 0.5f;  // Chainer's sigmoid formulation
    case NNAM_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ void store_chunk16(const GemmParams& p, long long row, int col, const float (&v)[16]) {
  if (p.out_kind == NNAM_OUT_F32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out_hi) + row * p.ldo + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    uint4* dh = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_hi) + row * p.ldo + col);
    dh[0] = make_uint4(h[0], h[1], h[2], h[3]);
    dh[1] = make_uint4(h[4], h[5], h[6], h[7]);
    if (p.out_kind == NNAM_OUT_BF16_SPLIT) {
      uint32_t l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        l[j] = pack_bf16x2(v[2 * j] - bf16_round(v[2 * j]), v[2 * j + 1] - bf16_round(v[2 * j + 1]));
      uint4* dl = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_lo) + row * p.ldo + col);
      dl[0] = make_uint4(l[0], l[1], l[2], l[3]);
      dl[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_bias_act_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                         const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                         const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024 B alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const int k_iters = p.k_blocks * p.nsplit;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_w_hi);
    if (p.nsplit > 1) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_w_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
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

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>((BM + p.bn) * BK * 2);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n;
        const int n_blk = tile % p.tiles_n;
        for (int pass = 0; pass < p.nsplit; ++pass) {
          const CUtensorMap* ma = (pass == 2) ? &tm_a_lo : &tm_a_hi;
          const CUtensorMap* mw = (pass == 1) ? &tm_w_lo : &tm_w_hi;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_2d(smem_a + stage * A_STAGE_BYTES, ma, &full_bar[stage], kb * BK, m_blk * BM);
            tma_load_2d(smem_b + stage * B_STAGE_BYTES, mw, &full_bar[stage], kb * BK, n_blk * p.bn);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc_bf16_f32(BM, static_cast<uint32_t>(p.bn));
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * MAX_BN);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t adesc = make_sw128_kmajor_desc(a_addr + k * UK * 2);
            const uint64_t bdesc = make_sw128_kmajor_desc(b_addr + k * UK * 2);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (it == k_iters - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------- epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n;
      const int n_blk = tile % p.tiles_n;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * MAX_BN);
      const long long row = static_cast<long long>(m_blk) * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t r0[16], r1[16];
        const bool second = (c0 + 16) < p.bn;  // warp-uniform
        tmem_ld16(taddr + c0, r0);
        if (second) tmem_ld16(taddr + c0 + 16, r1);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && !second) break;
          const int col = n_blk * p.bn + c0 + half * 16;
          if (col >= p.N) break;  // warp-uniform
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(half == 0 ? r0[j] : r1[j]);
          if (p.bias != nullptr) {
            if (col + 16 <= p.N) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b = __ldg(b4 + j);
                v[4 * j] += b.x;
                v[4 * j + 1] += b.y;
                v[4 * j + 2] += b.z;
                v[4 * j + 3] += b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col + j < p.N) v[j] += __ldg(p.bias + col + j);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
          if (row_ok) store_chunk16(p, row, col, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
static int pick_bn(int n) {
  const int tiles = (n + MAX_BN - 1) / MAX_BN;
  int bn = (n + tiles - 1) / tiles;
  bn = (bn + 15) / 16 * 16;
  return bn < 16 ? 16 : bn;
}

int gemm_bias_act(const void* a_hi, const void* a_lo, long long lda, const void* w_hi, const void* w_lo,
                  long long ldw, const float* bias, void* out_hi, void* out_lo, long long ldo, int M, int N, int K,
                  int act, int out_kind, int nsplit, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(NNAM_ERR_ARG, "gemm: empty problem");
  if (nsplit != 1 && nsplit != 3) return set_error(NNAM_ERR_ARG, "gemm: nsplit must be 1 or 3");
  if (lda % 8 || ldw % 8) return set_error(NNAM_ERR_ARG, "gemm: lda/ldw must be multiples of 8 elements (16 B)");
  if (lda < K || ldw < K) return set_error(NNAM_ERR_ARG, "gemm: leading dimension smaller than K");
  if (nsplit == 3 && (!a_lo || !w_lo)) return set_error(NNAM_ERR_ARG, "gemm: bf16x3 needs lo operands");
  if (out_kind == NNAM_OUT_BF16_SPLIT && !out_lo) return set_error(NNAM_ERR_ARG, "gemm: split output needs out_lo");
  const int n16 = (N + 15) / 16 * 16;
  if (ldo < n16) return set_error(NNAM_ERR_ARG, "gemm: ldo must be >= N rounded up to 16");
  if (out_kind == NNAM_OUT_F32 ? (ldo % 4) : (ldo % 8))
    return set_error(NNAM_ERR_ARG, "gemm: ldo must keep rows 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(a_lo) |
       reinterpret_cast<uintptr_t>(w_lo) | reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo) |
       reinterpret_cast<uintptr_t>(bias)) & 15)
    return set_error(NNAM_ERR_ARG, "gemm: pointers must be 16-byte aligned");

  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.bn = pick_bn(N);
  p.tiles_m = (M + BM - 1) / BM;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.k_blocks = (K + BK - 1) / BK;
  p.nsplit = nsplit;
  p.act = act;
  p.out_kind = out_kind;
  p.bias = bias;
  p.out_hi = out_hi;
  p.out_lo = out_lo;
  p.ldo = ldo;

  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  int rc;
  if ((rc = encode_tmap_bf16_2d(&ta_hi, a_hi, K, M, lda, BK, BM))) return rc;
  if ((rc = encode_tmap_bf16_2d(&tw_hi, w_hi, K, N, ldw, BK, p.bn))) return rc;
  if (nsplit == 3) {
    if ((rc = encode_tmap_bf16_2d(&ta_lo, a_lo, K, M, lda, BK, BM))) return rc;
    if ((rc = encode_tmap_bf16_2d(&tw_lo, w_lo, K, N, ldw, BK, p.bn))) return rc;
  } else {
    ta_lo = ta_hi;
    tw_lo = tw_hi;
  }

  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev]) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_bias_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "gemm: cudaFuncSetAttribute");
    attr_set[dev] = true;
  }
  const int total = p.tiles_m * p.tiles_n;
  const int grid = total < sm_count() ? total : sm_count();
  gemm_bias_act_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
  return check_launch("gemm_bias_act_kernel");
}

}  // namespace nnam
