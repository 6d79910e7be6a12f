// Peephole-LSTM cell arithmetic (L.StatefulPeepholeLSTM, used by PeepholeLSTM in
// scripts/common/chainer_networks.py:103-121): full-matrix peepholes peep_i / peep_f on c_prev and peep_o on c_new.
//
//   g = upward(x) + lateral(h);  a = tanh(g_a);  i = s(g_i + P_i c);  f = s(g_f + P_f c);  c' = a i + f c;
//   o = s(g_o + P_o c');  h' = o tanh(c')                      s(x) = tanh(x/2)/2 + 1/2  (Chainer's formulation)
//
// Unlike LSTM/GRU this cell is NOT in the persistent K3 kernel yet: the lateral + peephole weights of one unit slice
// (7 H-long rows per unit) do not fit next to the operand tiles in shared memory at H = 512.  It runs time step by
// time step on the packed rows: K2 computes  [h | c] . [W_lat | P_if]^T  and  c' . P_o^T, and this kernel fuses the
// gate arithmetic around them.  Element (row, unit) per thread, float4 gate loads (gates are interleaved per unit).
#include "ptx.cuh"
#include "nnam_internal.h"

namespace nnam {

__device__ __forceinline__ float pp_tanh(float x, int fast) {
  if (fast) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  return tanhf(x);
}
__device__ __forceinline__ float pp_sigmoid(float x, int fast) { return fmaf(pp_tanh(0.5f * x, fast), 0.5f, 0.5f); }

struct PeepParams {
  const float* gx;   // (n, gx_ld): upward(x) + b, gate-interleaved [a, i, f, o] per unit
  long long gx_ld;
  const float* g1;   // (n, g1_ld) or NULL: lateral(h) with P_i c / P_f c folded into the i / f columns
  long long g1_ld;
  const float* p2;   // phase 1: (n, p2_ld) = P_o c'
  long long p2_ld;
  const float* c_prev;  // (n, H) fp32 or NULL (first step: c = 0)
  float* c_new;         // (n, H) fp32: written in phase 0, read in phase 1
  __nv_bfloat16* out_hi;  // phase 0: c' -> columns [H, 2H) of the [h | c] rows; phase 1: h' -> columns [0, H)
  __nv_bfloat16* out_lo;  // low halves (bf16x3 mode) or NULL
  long long out_ld;
  int n, H, fast, f16;
};

template <int PHASE>
__global__ void peephole_cell_kernel(const PeepParams p) {
  const long long total = static_cast<long long>(p.n) * p.H;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = e / p.H;
    const int j = static_cast<int>(e - r * p.H);
    float4 g = __ldg(reinterpret_cast<const float4*>(p.gx + r * p.gx_ld) + j);
    if (p.g1 != nullptr) {
      const float4 l = __ldg(reinterpret_cast<const float4*>(p.g1 + r * p.g1_ld) + j);
      g.x += l.x;
      g.y += l.y;
      g.z += l.z;
      g.w += l.w;
    }
    float v;
    if (PHASE == 0) {
      const float c = p.c_prev != nullptr ? p.c_prev[r * p.H + j] : 0.0f;
      v = fmaf(pp_tanh(g.x, p.fast), pp_sigmoid(g.y, p.fast), pp_sigmoid(g.z, p.fast) * c);
      p.c_new[r * p.H + j] = v;
    } else {
      const float c = p.c_new[r * p.H + j];
      v = pp_sigmoid(g.w + __ldg(p.p2 + r * p.p2_ld + j), p.fast) * pp_tanh(c, p.fast);
    }
    const long long o = r * p.out_ld + (PHASE == 0 ? p.H : 0) + j;
    const uint16_t hb = f32_to_e16(v, p.f16);
    reinterpret_cast<uint16_t*>(p.out_hi)[o] = hb;
    if (p.out_lo != nullptr) p.out_lo[o] = __float2bfloat16_rn(v - e16_to_f32(hb, 0));
  }
}

int peephole_cell(int phase, const float* gx, long long gx_ld, const float* g1, long long g1_ld, const float* p2,
                  long long p2_ld, const float* c_prev, float* c_new, void* out_hi, void* out_lo, long long out_ld, int n,
                  int H, int fast, int elem, cudaStream_t stream) {
  if (phase != 0 && phase != 1) return set_error(NNAM_ERR_ARG, "peephole: phase must be 0 or 1");
  if (n < 0 || H <= 0) return set_error(NNAM_ERR_ARG, "peephole: bad shape");
  if (n == 0) return NNAM_OK;
  if (!gx || !c_new || !out_hi) return set_error(NNAM_ERR_ARG, "peephole: NULL buffer");
  if (phase == 1 && !p2) return set_error(NNAM_ERR_ARG, "peephole: phase 1 needs P_o c'");
  if (elem != NNAM_ELEM_BF16 && elem != NNAM_ELEM_F16) return set_error(NNAM_ERR_ARG, "peephole: unknown element type");
  if (elem == NNAM_ELEM_F16 && out_lo) return set_error(NNAM_ERR_ARG, "peephole: the hi/lo split is bf16 only");
  if (gx_ld % 4 || (g1 && g1_ld % 4) || (reinterpret_cast<uintptr_t>(gx) & 15) || (reinterpret_cast<uintptr_t>(g1) & 15))
    return set_error(NNAM_ERR_ARG, "peephole: gate rows must be 16-byte aligned");
  if (gx_ld < 4LL * H || (g1 && g1_ld < 4LL * H) || out_ld < 2LL * H)
    return set_error(NNAM_ERR_ARG, "peephole: leading dimension too small");
  PeepParams p{gx, gx_ld, g1, g1_ld, p2, p2_ld, c_prev, c_new, static_cast<__nv_bfloat16*>(out_hi),
               static_cast<__nv_bfloat16*>(out_lo), out_ld, n, H, fast, elem == NNAM_ELEM_F16};
  const long long total = static_cast<long long>(n) * H;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (phase == 0)
    peephole_cell_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  else
    peephole_cell_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  return check_launch("peephole_cell_kernel");
}

}  // namespace nnam
