// 3xTF32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[M,N] = act( A[M,K] * B[N,K]^T + bias ),   fp32 in, fp32 out, fp32-class accuracy.
//
// A single TF32 pass cannot meet the 1e-4 RMS parity gate (rounding the
// weights alone to TF32 costs 1.1e-4 .. 5e-4), so every operand is split once into two TF32
// numbers, x = x_hi + x_lo (x_hi = rna_tf32(x), x_lo = rna_tf32(x - x_hi), 21+ mantissa bits
// together) and the product is accumulated in fp32 in tensor memory as
//        A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          (A_lo*B_lo ~ 2^-22 is dropped)
// i.e. three tcgen05.mma.kind::tf32 per K-slice on four TMA-staged operand tiles.
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d, SWIZZLE_128B tiles, mbarrier tx
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
//   warps 2-9: epilogue      -- tcgen05.ld (32 lanes x 64 columns each: 4 lane quarters x 2 column
//                              halves), fp32 promotion, bias + activation / LSTM cell, st.global
//   smem ring: STAGES x {A_hi, A_lo, B_hi, B_lo} tiles of 128 rows x 32 fp32 (128-byte rows)
//   TMEM     : a ring of 4 accumulators of 128 lanes x 128 columns, one per K-CHUNK in flight:
//              the tensor core's fp32 accumulate truncates (measured: error grows linearly with
//              the number of accumulation steps, 3.4e-5 at K=1024), so every TC_CHUNK_KB k-blocks
//              the partial sum is promoted to fp32 registers of the epilogue warps (round-to-
//              nearest adds on the CUDA cores) while the MMA warp fills the other accumulator.
#include <stdlib.h>

#include "tc_common.cuh"

namespace se {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;
constexpr int TC_BK16 = 64;                                // fp16-pair operands: the same 128-byte tile rows hold 64 elements
constexpr int TC_STAGES = 3;
constexpr int TC_CHUNK_KB = 4;                              // k-blocks (of 32) per TMEM accumulation chunk
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;           // 16 KB (same for A and B tiles)
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;          // A_hi, A_lo, B_hi, B_lo
constexpr int TC_THREADS = 320;                            // 2 control warps + 8 epilogue warps
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_EPI_COLS = TC_BN / 2;                     // columns per epilogue warp
constexpr int TC_NACC = 4;                                 // TMEM accumulators in the ring: the MMA thread runs up to 4 chunks
                                                           // ahead of the epilogue warps (which also run the store epilogue)
constexpr int TC_TMEM_COLS = TC_NACC * TC_BN;              // 4 x 128 columns: all of tensor memory (one CTA per SM)
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum { EPI_BIAS_ACT = 0, EPI_LSTM_CELL = 1 };

struct TcParams {
  int M, N;
  int kb0, kb1;  // k-blocks (of 32) taken from A source 0 / A source 1  (A = [A0 | A1] along K)
  const float* bias;
  int act;
  float* C;      // EPI_BIAS_ACT: output [M, ldc] (may be NULL when only the split is wanted)
  long long ldc;
  float act_param, alpha;   // C = alpha * act(acc + bias, act_param) + res
  const float* res;
  float *c_hi, *c_lo;
  int panel_m;   // m-blocks per L2 panel
  // EPI_LSTM_CELL: N = 4H, column tile j holds units [32j, 32j+32): two 64-column halves, each
  // [i | f | g | o] x 16 units  (column j*128 + hf*64 + g*16 + u  <-  gate g of unit 32j + 16hf + u)
  float* c_state;   // [M, H] in/out
  float* h_hi;      // [M, H] out: rna_tf32(h)
  float* h_lo;      // [M, H] out: rna_tf32(h - h_hi)
  float* h_out;     // [M, H] out (plain fp32) or NULL
  int H;
  long long ld_hout;   // row stride of h_hi / h_lo / h_out (>= H; c_state is always [M, H] contiguous)
  int first_step;      // 1: h_{-1} = c_{-1} = 0 -- no recurrent k-blocks, c_state is not read
  // fp16-pair operands (F16 kernels): A and B arrive scaled by powers of two, out_scale = 1 / (scale_A * scale_B) puts
  // the sums back (exact); 1.0f on the TF32 path.  c16_hi / c16_lo: optional fp16-pair copy of the output for the next
  // f16 GEMM, scaled by c16_scale; in the F16 LSTM cell h_hi / h_lo point at fp16 data scaled by c16_scale.
  float out_scale;
  unsigned short *c16_hi, *c16_lo;
  float c16_scale;
};

// Final epilogue of one thread: NC consecutive columns [n0, n0 + NC) of output row `row`, fp32 sums in registers.
//   EPI_BIAS_ACT : C = alpha * act(sum + bias) + res, optional TF32 split of the result
//   EPI_LSTM_CELL: every 64-column group g64 = n / 64 holds [i | f | g | o] x 16 hidden units [16 g64, 16 g64 + 16)
//                  of sequence `row`: gates -> c, h, and the TF32 split of h for the next step's GEMM
// The 8 epilogue warps are also the ones that drain the TMEM chunk sums, and the MMA thread can only run as far ahead
// as there are free accumulators, so the instruction count here bounds the whole kernel (ncu, round 1: 57 % of all
// samples sat in a per-element version of this code -- activation switch, scalar bias loads and range checks for each
// of the 64 columns -- while the tensor pipe idled at 36 %).  Hence: the activation is a template parameter, the
// in-range aligned case moves bias / residual as float4, and the gates use ex2.approx + fast divide (|err| ~ 2e-7,
// the forms the recurrence kernel csrc/lstm_tc.cu has been parity-tested with).
template <int NC, int ACT>
__device__ __forceinline__ void tc_store_bias_act(const TcParams& p, const float (&sum)[NC], int row, int n0) {
  const long long roff = (long long)row * p.ldc;
  const bool vec = (p.ldc & 3) == 0;
  const float* bias = p.bias;
  const float* res = p.res ? p.res + roff : nullptr;
  if (vec && n0 + NC <= p.N && ((reinterpret_cast<uintptr_t>(bias) & 15) == 0)) {
    // whole span in range, 16-byte aligned rows: float4 traffic only, no per-element checks
    const bool v8 = (p.ldc & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.C) | reinterpret_cast<uintptr_t>(p.c_hi) |
                                         reinterpret_cast<uintptr_t>(p.c_lo) | reinterpret_cast<uintptr_t>(bias) |
                                         reinterpret_cast<uintptr_t>(p.res)) & 31) == 0;
    if (v8) {   // 32-byte rows: eight columns per step, whole-sector stores (st_global_v8)
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (bias) {
          b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + j));
          b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + j + 4));
        }
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float os = p.out_scale;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = tc_act<ACT>(fmaf(sum[j + e], os, bb[e]), p.act_param) * p.alpha;
        if (res) {
          float r[8];
          ldg_v8(res + n0 + j, r);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += r[e];
        }
        if (p.C) st_global_v8(p.C + roff + n0 + j, o);
        if (p.c_hi) {
          float hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_tf32_dev(o[e], hi[e], lo[e]);
          st_global_v8(p.c_hi + roff + n0 + j, hi);
          st_global_v8(p.c_lo + roff + n0 + j, lo);
        }
        if (p.c16_hi) {
          unsigned short hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_f16_dev(o[e], p.c16_scale, hi[e], lo[e]);
          st_global_h8(p.c16_hi + roff + n0 + j, hi);
          st_global_h8(p.c16_lo + roff + n0 + j, lo);
        }
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      const float4 bb = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float os = p.out_scale;   // 1.0f on the TF32 path: fmaf(s, 1, b) == s + b
      float o[4] = {tc_act<ACT>(fmaf(sum[j], os, bb.x), p.act_param) * p.alpha,
                    tc_act<ACT>(fmaf(sum[j + 1], os, bb.y), p.act_param) * p.alpha,
                    tc_act<ACT>(fmaf(sum[j + 2], os, bb.z), p.act_param) * p.alpha,
                    tc_act<ACT>(fmaf(sum[j + 3], os, bb.w), p.act_param) * p.alpha};
      if (res) {
        const float4 rr = __ldg(reinterpret_cast<const float4*>(res + n0 + j));
        o[0] += rr.x;
        o[1] += rr.y;
        o[2] += rr.z;
        o[3] += rr.w;
      }
      if (p.C) *reinterpret_cast<float4*>(p.C + roff + n0 + j) = make_float4(o[0], o[1], o[2], o[3]);
      if (p.c_hi) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32_dev(o[e], hi[e], lo[e]);
        *reinterpret_cast<float4*>(p.c_hi + roff + n0 + j) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(p.c_lo + roff + n0 + j) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      if (p.c16_hi) {
        unsigned short hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_f16_dev(o[e], p.c16_scale, hi[e], lo[e]);
        *reinterpret_cast<uint2*>(p.c16_hi + roff + n0 + j) = make_uint2(hi[0] | ((unsigned)hi[1] << 16), hi[2] | ((unsigned)hi[3] << 16));
        *reinterpret_cast<uint2*>(p.c16_lo + roff + n0 + j) = make_uint2(lo[0] | ((unsigned)lo[1] << 16), lo[2] | ((unsigned)lo[3] << 16));
      }
    }
    return;
  }
  // ragged edge (N not a multiple of the span, or rows that are not 16-byte aligned): element by element
#pragma unroll   // fully unrolled: a dynamic index would move sum[] to local memory for the whole kernel
  for (int j = 0; j < NC; ++j) {
    const int n = n0 + j;
    if (n < p.N) {
      float o = tc_act<ACT>(fmaf(sum[j], p.out_scale, bias ? __ldg(bias + n) : 0.f), p.act_param) * p.alpha;
      if (res) o += __ldg(res + n);
      if (p.C) p.C[roff + n] = o;
      if (p.c_hi) split_tf32_dev(o, p.c_hi[roff + n], p.c_lo[roff + n]);
      if (p.c16_hi) split_f16_dev(o, p.c16_scale, p.c16_hi[roff + n], p.c16_lo[roff + n]);
    }
  }
}

template <int EPI, int NC, bool F16 = false>
__device__ __forceinline__ void tc_epilogue_store(const TcParams& p, const float (&sum)[NC], int row, int n0) {
  if constexpr (EPI == EPI_BIAS_ACT) {
    switch (p.act) {   // uniform across the grid: one predictable branch per tile instead of one per element
      case SE_ACT_PRELU: tc_store_bias_act<NC, SE_ACT_PRELU>(p, sum, row, n0); break;
      case SE_ACT_ELU: tc_store_bias_act<NC, SE_ACT_ELU>(p, sum, row, n0); break;
      case SE_ACT_SOFTPLUS: tc_store_bias_act<NC, SE_ACT_SOFTPLUS>(p, sum, row, n0); break;
      case SE_ACT_RELU: tc_store_bias_act<NC, SE_ACT_RELU>(p, sum, row, n0); break;
      case SE_ACT_SIGMOID: tc_store_bias_act<NC, SE_ACT_SIGMOID>(p, sum, row, n0); break;
      case SE_ACT_TANH: tc_store_bias_act<NC, SE_ACT_TANH>(p, sum, row, n0); break;
      default: tc_store_bias_act<NC, SE_ACT_NONE>(p, sum, row, n0); break;
    }
  } else {
    static_assert(NC % 64 == 0, "the fused LSTM cell works on whole 64-column gate groups");
#pragma unroll
    for (int q = 0; q < NC / 64; ++q) {
      const int nq = n0 + q * 64;
      if (nq >= p.N) break;                 // N = 4H, H % 32 == 0 (checked on the host): whole groups only
      const int unit0 = (nq >> 6) * 16;
      const long long off = (long long)row * p.H + unit0;
      const long long hoff = (long long)row * p.ld_hout + unit0;
      const float4* bias4 = reinterpret_cast<const float4*>(p.bias + nq);   // 16-byte aligned (checked on the host)
      const bool v8 = (p.ld_hout & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.c_state) | reinterpret_cast<uintptr_t>(p.h_out) |
                                                 (F16 ? 0 : (reinterpret_cast<uintptr_t>(p.h_hi) | reinterpret_cast<uintptr_t>(p.h_lo)))) & 31) == 0;
      if (v8) {   // eight units per step: 32-byte state loads / stores (whole sectors), 16-byte fp16 pairs
#pragma unroll
        for (int u = 0; u < 16; u += 8) {
          float co[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (!p.first_step) ld_global_v8(p.c_state + off + u, co);
          float vb[4][8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b0 = __ldg(bias4 + 4 * g + (u >> 2)), b1 = __ldg(bias4 + 4 * g + (u >> 2) + 1);
            vb[g][0] = b0.x, vb[g][1] = b0.y, vb[g][2] = b0.z, vb[g][3] = b0.w;
            vb[g][4] = b1.x, vb[g][5] = b1.y, vb[g][6] = b1.z, vb[g][7] = b1.w;
          }
          float cn[8], hn[8], hh[8], hl[8];
          unsigned short h16[8], l16[8];
          const float os = p.out_scale;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float ig = fast_sigmoid(fmaf(sum[q * 64 + u + e], os, vb[0][e]));
            const float fg = fast_sigmoid(fmaf(sum[q * 64 + 16 + u + e], os, vb[1][e]));
            const float gg = fast_tanh(fmaf(sum[q * 64 + 32 + u + e], os, vb[2][e]));
            const float og = fast_sigmoid(fmaf(sum[q * 64 + 48 + u + e], os, vb[3][e]));
            cn[e] = fg * co[e] + ig * gg;
            hn[e] = og * fast_tanh(cn[e]);
            if constexpr (F16) split_f16_dev(hn[e], p.c16_scale, h16[e], l16[e]);
            else split_tf32_dev(hn[e], hh[e], hl[e]);
          }
          st_global_v8(p.c_state + off + u, cn);
          if constexpr (F16) {
            st_global_h8(reinterpret_cast<unsigned short*>(p.h_hi) + hoff + u, h16);
            st_global_h8(reinterpret_cast<unsigned short*>(p.h_lo) + hoff + u, l16);
          } else {
            st_global_v8(p.h_hi + hoff + u, hh);
            st_global_v8(p.h_lo + hoff + u, hl);
          }
          if (p.h_out) st_global_v8(p.h_out + hoff + u, hn);
        }
        continue;
      }
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        const float4 cold = p.first_step ? make_float4(0.f, 0.f, 0.f, 0.f)
                                         : *reinterpret_cast<const float4*>(p.c_state + off + u);
        const float4 bi = __ldg(bias4 + (u >> 2)), bf = __ldg(bias4 + 4 + (u >> 2)), bg = __ldg(bias4 + 8 + (u >> 2)),
                     bo = __ldg(bias4 + 12 + (u >> 2));
        const float co[4] = {cold.x, cold.y, cold.z, cold.w};
        const float vbi[4] = {bi.x, bi.y, bi.z, bi.w}, vbf[4] = {bf.x, bf.y, bf.z, bf.w}, vbg[4] = {bg.x, bg.y, bg.z, bg.w},
                    vbo[4] = {bo.x, bo.y, bo.z, bo.w};
        float cn[4], hn[4], hh[4], hl[4];
        unsigned short h16[4], l16[4];
        const float os = p.out_scale;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float ig = fast_sigmoid(fmaf(sum[q * 64 + u + e], os, vbi[e]));
          const float fg = fast_sigmoid(fmaf(sum[q * 64 + 16 + u + e], os, vbf[e]));
          const float gg = fast_tanh(fmaf(sum[q * 64 + 32 + u + e], os, vbg[e]));
          const float og = fast_sigmoid(fmaf(sum[q * 64 + 48 + u + e], os, vbo[e]));
          cn[e] = fg * co[e] + ig * gg;
          hn[e] = og * fast_tanh(cn[e]);
          if constexpr (F16) split_f16_dev(hn[e], p.c16_scale, h16[e], l16[e]);
          else split_tf32_dev(hn[e], hh[e], hl[e]);
        }
        *reinterpret_cast<float4*>(p.c_state + off + u) = make_float4(cn[0], cn[1], cn[2], cn[3]);
        if constexpr (F16) {
          unsigned short* o_hi = reinterpret_cast<unsigned short*>(p.h_hi);
          unsigned short* o_lo = reinterpret_cast<unsigned short*>(p.h_lo);
          *reinterpret_cast<uint2*>(o_hi + hoff + u) = make_uint2(h16[0] | ((unsigned)h16[1] << 16), h16[2] | ((unsigned)h16[3] << 16));
          *reinterpret_cast<uint2*>(o_lo + hoff + u) = make_uint2(l16[0] | ((unsigned)l16[1] << 16), l16[2] | ((unsigned)l16[3] << 16));
        } else {
          *reinterpret_cast<float4*>(p.h_hi + hoff + u) = make_float4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<float4*>(p.h_lo + hoff + u) = make_float4(hl[0], hl[1], hl[2], hl[3]);
        }
        if (p.h_out) *reinterpret_cast<float4*>(p.h_out + hoff + u) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
  }
}

// CM x CN > 1: a thread-block CLUSTER of CM x CN CTAs computes a (128 CM) x (128 CN) super-tile, every CTA its own
// 128 x 128 tile with the one-CTA MMA.  The operand tiles the CTAs of a cluster row / column have in common are read
// from L2 ONCE and TMA-multicast: with CN = 2 the CTA with rn = 0 loads A_hi and the one with rn = 1 loads A_lo of the
// row's A tile, each into both CTAs; with CM = 2 likewise B_hi / B_lo down a column.  The kernel is bound by L2 -> SM
// operand traffic (~8 TB/s aggregate: profiles/), which this cuts from 64 KB to 48 KB (1 x 2) or 32 KB (2 x 2) per
// k-block and CTA.  A stage may be refilled only when every CTA that receives this CTA's multicast has consumed it:
// `empty` counts CM + CN - 1 arrivals, each MMA thread commits to its row and column peers.
template <int EPI, int CM, int CN, bool F16 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a0hi, const __grid_constant__ CUtensorMap map_a0lo,
                   const __grid_constant__ CUtensorMap map_a1hi, const __grid_constant__ CUtensorMap map_a1lo,
                   const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                   const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = base;                                        // STAGES x 4 x 16 KB, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES]
  uint64_t* empty = bars + TC_STAGES;      // [STAGES]
  uint64_t* tfull = bars + 2 * TC_STAGES;  // [TC_NACC]
  uint64_t* tempty = tfull + TC_NACC;      // [TC_NACC]
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + TC_NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CSIZE = CM * CN;
  constexpr int BK = F16 ? TC_BK16 : TC_BK;                       // elements per k-block (128 bytes either way)
  const unsigned rank = CSIZE > 1 ? cluster_ctarank() : 0u;
  const int rm = (int)rank / CN, rn = (int)rank % CN;             // position of this CTA in the cluster
  const int cl = (int)blockIdx.x / CSIZE, ncl = (int)gridDim.x / CSIZE;
  const int mblocks = ceil_div(p.M, TC_BM * CM), nblocks = ceil_div(p.N, TC_BN * CN);   // in super-tiles
  const int ntiles = mblocks * nblocks;
  const int kblocks = p.kb0 + p.kb1;
  unsigned short row_mask = 0, col_mask = 0;                      // CTAs sharing my A tile / my B tile
#pragma unroll
  for (int j = 0; j < CN; ++j) row_mask |= (unsigned short)(1u << (rm * CN + j));
#pragma unroll
  for (int i = 0; i < CM; ++i) col_mask |= (unsigned short)(1u << (i * CN + rn));

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CM + CN - 1);
    }
    for (int a = 0; a < TC_NACC; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_a0hi);
    tma_prefetch_desc(&map_a0lo);
    tma_prefetch_desc(&map_bhi);
    tma_prefetch_desc(&map_blo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if constexpr (CSIZE > 1) {
    __syncwarp();
    cluster_sync_all();   // every CTA's barriers are initialised before a peer multicasts into it / signals them
  }
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;

  // tile order: panels of `panel_m` m-blocks, n fastest inside a panel (A panel + all of B stay in L2)
  auto tile_coords = [&](int tile, int& mb, int& nb) {
    const int per_panel = p.panel_m * nblocks;
    const int panel = tile / per_panel;
    const int r = tile - panel * per_panel;
    const int pm = min(p.panel_m, mblocks - panel * p.panel_m);
    nb = (r / pm) * CN + rn;                                       // this CTA's 128 x 128 tile of the super-tile
    mb = (panel * p.panel_m + (r - (r / pm) * pm)) * CM + rm;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      unsigned phase = 0;
      for (int tile = cl; tile < ntiles; tile += ncl) {
        int mb, nb;
        tile_coords(tile, mb, nb);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_parity(&empty[stage], phase ^ 1);
          unsigned char* st = tiles + stage * TC_STAGE_BYTES;
          mbar_expect_tx(&full[stage], TC_STAGE_BYTES);
          const bool src0 = kb < p.kb0;
          const int ak = (src0 ? kb : kb - p.kb0) * BK;
          if constexpr (CN == 1) {
            tma_load_2d(src0 ? &map_a0hi : &map_a1hi, &full[stage], st + 0 * TC_TILE_BYTES, ak, mb * TC_BM);
            tma_load_2d(src0 ? &map_a0lo : &map_a1lo, &full[stage], st + 1 * TC_TILE_BYTES, ak, mb * TC_BM);
          } else if (rn == 0) {   // my row's A_hi, into every CTA of the row (the rn = 1 CTA sends A_lo)
            tma_load_2d_mc(src0 ? &map_a0hi : &map_a1hi, &full[stage], st + 0 * TC_TILE_BYTES, ak, mb * TC_BM, row_mask);
          } else {
            tma_load_2d_mc(src0 ? &map_a0lo : &map_a1lo, &full[stage], st + 1 * TC_TILE_BYTES, ak, mb * TC_BM, row_mask);
          }
          if constexpr (CM == 1) {
            tma_load_2d(&map_bhi, &full[stage], st + 2 * TC_TILE_BYTES, kb * BK, nb * TC_BN);
            tma_load_2d(&map_blo, &full[stage], st + 3 * TC_TILE_BYTES, kb * BK, nb * TC_BN);
          } else if (rm == 0) {   // my column's B_hi, into every CTA of the column (the rm = 1 CTA sends B_lo)
            tma_load_2d_mc(&map_bhi, &full[stage], st + 2 * TC_TILE_BYTES, kb * BK, nb * TC_BN, col_mask);
          } else {
            tma_load_2d_mc(&map_blo, &full[stage], st + 3 * TC_TILE_BYTES, kb * BK, nb * TC_BN, col_mask);
          }
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (elect_one()) {
      constexpr unsigned idesc = F16 ? make_idesc_f16(TC_BM, TC_BN) : make_idesc_tf32(TC_BM, TC_BN);
      int stage = 0;
      unsigned phase = 0;
      int acc = 0;
      unsigned acc_phase = 0;
      for (int tile = cl; tile < ntiles; tile += ncl) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const bool chunk_start = (kb % TC_CHUNK_KB) == 0;
          if (chunk_start) {
            mbar_wait_parity(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
          }
          const unsigned d_tmem = tmem_base + (unsigned)(acc * TC_BN);
          mbar_wait_parity(&full[stage], phase);
          tc_fence_after();
          unsigned char* st = tiles + stage * TC_STAGE_BYTES;
          const uint64_t d_ahi = make_smem_desc(st + 0 * TC_TILE_BYTES);
          const uint64_t d_alo = make_smem_desc(st + 1 * TC_TILE_BYTES);
          const uint64_t d_bhi = make_smem_desc(st + 2 * TC_TILE_BYTES);
          const uint64_t d_blo = make_smem_desc(st + 3 * TC_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 32 bytes per MMA K slice (8 tf32 / 16 fp16), in 16-byte units
            if constexpr (F16) {
              umma_f16(d_tmem, d_alo + adv, d_bhi + adv, idesc, (!chunk_start || k > 0) ? 1u : 0u);
              umma_f16(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_f16(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            } else {
              umma_tf32(d_tmem, d_alo + adv, d_bhi + adv, idesc, (!chunk_start || k > 0) ? 1u : 0u);
              umma_tf32(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_tf32(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            }
          }
          // smem slot reusable once these MMAs have read it: tell every CTA that multicasts into this one
          if constexpr (CSIZE > 1) umma_commit_mc(&empty[stage], (unsigned short)(row_mask | col_mask));
          else umma_commit(&empty[stage]);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if ((kb % TC_CHUNK_KB) == TC_CHUNK_KB - 1 || kb == kblocks - 1) {
            umma_commit(&tfull[acc]);  // chunk accumulator complete
            if (++acc == TC_NACC) {
              acc = 0;
              acc_phase ^= 1;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quarter = warp & 3;        // tcgen05.ld: warp w may touch lanes 32*(w%4) .. +31
    const int half = (warp - 2) >> 2;    // which 64 columns of the 128-column accumulator
    int acc = 0;
    unsigned acc_phase = 0;
    for (int tile = cl; tile < ntiles; tile += ncl) {
      int mb, nb;
      tile_coords(tile, mb, nb);
      const int row = mb * TC_BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      float sum[TC_EPI_COLS];
#pragma unroll
      for (int j = 0; j < TC_EPI_COLS; ++j) sum[j] = 0.f;
      const int nchunks = (kblocks + TC_CHUNK_KB - 1) / TC_CHUNK_KB;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait_parity(&tfull[acc], acc_phase);
        tc_fence_after();
        {
          float v[TC_EPI_COLS];
          const unsigned taddr =
              tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(acc * TC_BN + half * TC_EPI_COLS);
          tmem_ld_32x64(taddr, v);
#pragma unroll
          for (int j = 0; j < TC_EPI_COLS; ++j) sum[j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (++acc == TC_NACC) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (!row_ok) continue;
      tc_epilogue_store<EPI, TC_EPI_COLS, F16>(p, sum, row, nb * TC_BN + half * TC_EPI_COLS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CSIZE > 1) {
    __syncwarp();
    cluster_sync_all();   // no CTA leaves while a peer may still multicast into its smem or signal its barriers
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// =================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): one 256 x 256 output tile per pair of CTAs (two SMs of a TPC).
//
// Why: the 128 x 128 kernel above is bound by L2 -> SM operand traffic, not by the tensor pipe (four 16 KB tiles per
// 12 MMAs: 69 B/clk/SM wanted, ~37 delivered; profiles/ncu_full_r01_summary.txt).  A pair multiplies M = 256 rows
// (128 per CTA) by N = 256 columns with each CTA staging only ITS 128 A rows and ITS 128 of the 256 B rows: the same
// 64 KB per k-block per SM now feeds twice the MMA work.
//
//   * both CTAs: TMA producer warp loads the CTA's own tiles into its own smem; the transaction bytes of BOTH CTAs are
//     counted on the LEADER's (cluster rank 0) `full` barrier (.cta_group::2 TMA, barrier address with the peer bit
//     cleared -- the addressing CUTLASS' SM100_TMA_2SM_LOAD uses);
//   * leader only: one thread issues tcgen05.mma.cta_group::2 (reads A/B from both CTAs' smem at the same offsets,
//     writes rows 0-127 to its own TMEM and rows 128-255 to the peer's); tcgen05.commit ... multicast::cluster arrives
//     on the `empty` / `tfull` barriers of both CTAs;
//   * both CTAs: 8 epilogue warps drain their own TMEM (32 lanes x 128 columns each), promote chunk sums to fp32
//     registers as above, and arrive REMOTELY on the leader's `tempty` barrier (count 16).
constexpr int T2_BN = 256;
constexpr int T2_EPI_COLS = T2_BN / 2;                     // 128 columns per epilogue warp (EW = 8)
// EW = 16 epilogue warps (four per TMEM lane quarter, 64 columns each): an experiment for the fused LSTM cell on fp16 pairs,
// where the MMA stream of a tile (12 k-blocks at twice the TF32 rate) is about as long as the cell epilogue of eight warps
// (8192 cell updates, ten MUFU each, two warps per scheduler).  Measured on B200: slower (FullSubNet 118.7 vs 101.2 ms per
// batch; 576 threads leave 96 registers per thread and the cell epilogue spills) -- not the default, SE_CELL_EPI_WARPS=16.
constexpr int T2_TMEM_COLS = 512;                          // 2 accumulators x 256 columns: all of tensor memory
template <int EPI, bool F16 = false, int EW = 8>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a0hi, const __grid_constant__ CUtensorMap map_a0lo,
                        const __grid_constant__ CUtensorMap map_a1hi, const __grid_constant__ CUtensorMap map_a1lo,
                        const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                        const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = base;                                        // STAGES x {A_hi, A_lo, B_hi, B_lo} x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES]  used in the leader CTA only
  uint64_t* empty = bars + TC_STAGES;      // [STAGES]  one per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * TC_STAGES;  // [2]       one per CTA (multicast commit)
  uint64_t* tempty = tfull + 2;            // [2]       used in the leader CTA only
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();            // 0 = leader
  constexpr int BK = F16 ? TC_BK16 : TC_BK;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int mblocks = ceil_div(p.M, 2 * TC_BM), nblocks = ceil_div(p.N, T2_BN);
  const int ntiles = mblocks * nblocks;
  const int kblocks = p.kb0 + p.kb1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * EW);  // one arrive per epilogue warp of either CTA
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_a0hi);
    tma_prefetch_desc(&map_a0lo);
    tma_prefetch_desc(&map_bhi);
    tma_prefetch_desc(&map_blo);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, T2_TMEM_COLS);   // the same warp of both CTAs, same smem slot
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();     // barriers of both CTAs initialised and both allocations done before any remote signal
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int& mb, int& nb) {
    const int per_panel = p.panel_m * nblocks;
    const int panel = tile / per_panel;
    const int r = tile - panel * per_panel;
    const int pm = min(p.panel_m, mblocks - panel * p.panel_m);
    nb = r / pm;
    mb = panel * p.panel_m + (r - nb * pm);
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of the B rows) =====================
    if (elect_one()) {
      int stage = 0;
      unsigned phase = 0;
      for (int tile = pair; tile < ntiles; tile += npairs) {
        int mb, nb;
        tile_coords(tile, mb, nb);
        const int arow = mb * 2 * TC_BM + (int)rank * TC_BM;
        const int brow = nb * T2_BN + (int)rank * (T2_BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_parity(&empty[stage], phase ^ 1);
          unsigned char* st = tiles + stage * TC_STAGE_BYTES;
          const unsigned lbar = smem_u32(&full[stage]) & T2_PEER_BIT_MASK;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * TC_STAGE_BYTES);   // both CTAs' bytes land on this barrier
          if (kb < p.kb0) {
            tma_load_2d_pair(&map_a0hi, lbar, st + 0 * TC_TILE_BYTES, kb * BK, arow);
            tma_load_2d_pair(&map_a0lo, lbar, st + 1 * TC_TILE_BYTES, kb * BK, arow);
          } else {
            tma_load_2d_pair(&map_a1hi, lbar, st + 0 * TC_TILE_BYTES, (kb - p.kb0) * BK, arow);
            tma_load_2d_pair(&map_a1lo, lbar, st + 1 * TC_TILE_BYTES, (kb - p.kb0) * BK, arow);
          }
          tma_load_2d_pair(&map_bhi, lbar, st + 2 * TC_TILE_BYTES, kb * BK, brow);
          tma_load_2d_pair(&map_blo, lbar, st + 3 * TC_TILE_BYTES, kb * BK, brow);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (rank == 0 && elect_one()) {
      constexpr unsigned idesc = F16 ? make_idesc_f16(2 * TC_BM, T2_BN) : make_idesc_tf32(2 * TC_BM, T2_BN);
      int stage = 0;
      unsigned phase = 0;
      int acc = 0;
      unsigned acc_phase = 0;
      for (int tile = pair; tile < ntiles; tile += npairs) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const bool chunk_start = (kb % TC_CHUNK_KB) == 0;
          if (chunk_start) {
            mbar_wait_parity(&tempty[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
            tc_fence_after();
          }
          const unsigned d_tmem = tmem_base + (unsigned)(acc * T2_BN);
          mbar_wait_parity(&full[stage], phase);
          tc_fence_after();
          unsigned char* st = tiles + stage * TC_STAGE_BYTES;
          const uint64_t d_ahi = make_smem_desc(st + 0 * TC_TILE_BYTES);
          const uint64_t d_alo = make_smem_desc(st + 1 * TC_TILE_BYTES);
          const uint64_t d_bhi = make_smem_desc(st + 2 * TC_TILE_BYTES);
          const uint64_t d_blo = make_smem_desc(st + 3 * TC_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            if constexpr (F16) {
              umma_f16_pair(d_tmem, d_alo + adv, d_bhi + adv, idesc, (!chunk_start || k > 0) ? 1u : 0u);
              umma_f16_pair(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_f16_pair(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            } else {
              umma_tf32_pair(d_tmem, d_alo + adv, d_bhi + adv, idesc, (!chunk_start || k > 0) ? 1u : 0u);
              umma_tf32_pair(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_tf32_pair(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            }
          }
          umma_commit_pair(&empty[stage]);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if ((kb % TC_CHUNK_KB) == TC_CHUNK_KB - 1 || kb == kblocks - 1) {
            umma_commit_pair(&tfull[acc]);
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own TMEM rows -> registers -> global =====================
    constexpr int ECOLS = T2_BN / (EW / 4);   // columns per epilogue warp: 128 (EW = 8) or 64 (EW = 16)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;    // which ECOLS columns of the 256-column accumulator
    int acc = 0;
    unsigned acc_phase = 0;
    const int nchunks = (kblocks + TC_CHUNK_KB - 1) / TC_CHUNK_KB;
    for (int tile = pair; tile < ntiles; tile += npairs) {
      int mb, nb;
      tile_coords(tile, mb, nb);
      const int row = mb * 2 * TC_BM + (int)rank * TC_BM + quarter * 32 + lane;
      float sum[ECOLS];
#pragma unroll
      for (int j = 0; j < ECOLS; ++j) sum[j] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait_parity(&tfull[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int piece = 0; piece < ECOLS / 32; ++piece) {
          float v[32];
          const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16) +
                                 (unsigned)(acc * T2_BN + half * ECOLS + piece * 32);
          tmem_ld_32x32(taddr, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[piece * 32 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (row >= p.M) continue;
      tc_epilogue_store<EPI, ECOLS, F16>(p, sum, row, nb * T2_BN + half * ECOLS);
    }
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();     // the peer's TMEM / barriers stay alive until the leader's last MMA and commit have landed
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, T2_TMEM_COLS);
  }
}

// x -> (hi, lo):  hi = rna_tf32(x), lo = rna_tf32(x - hi)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                                        float4* __restrict__ lo, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    float in[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      unsigned hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hb) : "f"(in[e]));
      h[e] = __uint_as_float(hb);
      const float r = in[e] - h[e];
      asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lb) : "f"(r));
      l[e] = __uint_as_float(lb);
    }
    hi[i] = make_float4(h[0], h[1], h[2], h[3]);
    lo[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// rows of K floats -> rows of Kpad >= K floats (zero tail), split into (hi, lo): lets a K that is not a multiple of 32
// (the 161 bins of LSTM/LSTM.py:17) run on the tensor-core GEMM against weights padded the same way
__global__ void __launch_bounds__(256) pad_split_tf32_kernel(const float* __restrict__ x, long long rows, int K, int Kpad,
                                                            float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = rows * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Kpad;
    const int c = (int)(i - r * Kpad);
    float h = 0.f, l = 0.f;
    if (c < K) split_tf32_dev(__ldg(x + r * K + c), h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// rows of K floats (row stride ldx) -> rows of Kpad >= K fp16 pairs (zero tail): x * scale = hi + lo
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, long long rows, int K, long long ldx,
                                                       int Kpad, float scale, unsigned short* __restrict__ hi,
                                                       unsigned short* __restrict__ lo) {
  const int kq = Kpad >> 2;                                   // 4 elements per thread
  const long long total = rows * kq;
  const bool vec = (ldx & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / kq;
    const int c = (int)(i - r * kq) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec && c + 4 <= K) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c + e < K) v[e] = __ldg(x + r * ldx + c + e);
    }
    unsigned short h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16_dev(v[e], scale, h[e], l[e]);
    *reinterpret_cast<uint2*>(hi + r * Kpad + c) = make_uint2(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16));
    *reinterpret_cast<uint2*>(lo + r * Kpad + c) = make_uint2(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16));
  }
}

// ---- host: tensor maps through the driver entry point (no link-time libcuda dependency) ----------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* ptr, int rows, int K, long long ld, bool f16 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SE_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(f16 ? TC_BK16 : TC_BK), (cuuint32_t)TC_BM};   // 128 bytes x 128 rows either way
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d K=%d ld=%lld)", (int)r, rows, K, ld);
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

}  // namespace se

using namespace se;

extern "C" int se_split_tf32(const float* x, float* hi, float* lo, long long n, se_stream_t stream) {
  SE_REQUIRE(x && hi && lo && n > 0 && (n & 3) == 0, "se_split_tf32: n=%lld must be a positive multiple of 4", n);
  SE_REQUIRE(((((uintptr_t)x) | ((uintptr_t)hi) | ((uintptr_t)lo)) & 15) == 0, "se_split_tf32: unaligned");
  const long long n4 = n / 4;
  const int blocks = (int)min((long long)148 * 8, ceil_div_ll(n4, 256));
  split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x),
                                                              reinterpret_cast<float4*>(hi),
                                                              reinterpret_cast<float4*>(lo), n4);
  return check_launch("se_split_tf32");
}

extern "C" int se_pad_split_tf32(const float* x, long long rows, int K, int Kpad, float* hi, float* lo,
                                 se_stream_t stream) {
  SE_REQUIRE(x && hi && lo && rows > 0 && K > 0 && Kpad >= K && (Kpad & 3) == 0, "se_pad_split_tf32: rows=%lld K=%d Kpad=%d",
             rows, K, Kpad);
  const int blocks = (int)min((long long)148 * 8, ceil_div_ll(rows * Kpad, 256));
  pad_split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, K, Kpad, hi, lo);
  return check_launch("se_pad_split_tf32");
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// 0: gemm_tf32x3_kernel<.., 1, 1> for every shape; 1: gemm_tf32x3_pair_kernel where M >= 256 and N >= 256;
// 2 / 3 / 4: gemm_tf32x3_kernel in clusters of 2 x 2 / 1 x 2 / 2 x 1 CTAs with TMA-multicast operand tiles where the
// problem has at least one full super-tile; 5 (default): the pair kernel for GEMMs with K >= 384 (12 k-blocks), M >= 256
// and N >= 256, the one-CTA kernel elsewhere -- what measured fastest per shape on B200 after the epilogue rewrite
// (profiles/kbench_r01f.json: 0.83 vs 0.89 ms on the CRN projection, FullSubNet +7 %; short-K GEMMs and the convs of the
// TCM models were 2-6 % slower on pairs).  SE_GEMM_ENGINE in the environment overrides the default;
// se_set_gemm_engine() overrides both.
constexpr int kDefaultGemmEngine = 5;
constexpr int kMaxGemmEngine = 5;
constexpr int kAutoPairMinKBlocks = 12;
constexpr int kAutoPairMinKBlocksF16 = 6;     // the same K = 384 in 64-element k-blocks
static int g_gemm_engine = -1;
static int gemm_engine() {
  if (g_gemm_engine < 0) {
    g_gemm_engine = kDefaultGemmEngine;
    if (const char* e = getenv("SE_GEMM_ENGINE")) {
      const int v = atoi(e);
      if (v >= 0 && v <= kMaxGemmEngine) g_gemm_engine = v;
    }
  }
  return g_gemm_engine;
}

int se::gemm_engine_is_pair() { return gemm_engine() == 1; }     // conv_tc.cu follows the same switch

template <int EPI, bool F16 = false, int EW = 8>
static int launch_pair_kernel(const CUtensorMap& a0hi, const CUtensorMap& a0lo, const CUtensorMap& a1hi,
                              const CUtensorMap& a1lo, const CUtensorMap& bhi, const CUtensorMap& blo, const TcParams& p,
                              int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_pair_kernel<EPI, F16, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e != cudaSuccess) {
    set_error("tcgen05 pair gemm: smem attribute: %s", cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(64 + 32 * EW);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, gemm_tf32x3_pair_kernel<EPI, F16, EW>, a0hi, a0lo, a1hi, a1lo, bhi, blo, p);
  if (e != cudaSuccess) {
    set_error("tcgen05 pair gemm: cluster launch: %s", cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

static int launch_tc_pair(int epi, const CUtensorMap& a0hi, const CUtensorMap& a0lo, const CUtensorMap& a1hi,
                          const CUtensorMap& a1lo, const CUtensorMap& bhi, const CUtensorMap& blo, TcParams p, int sms,
                          cudaStream_t stream, bool f16 = false) {
  const int mblocks = ceil_div(p.M, 2 * TC_BM), nblocks = ceil_div(p.N, T2_BN);
  p.panel_m = min(8, mblocks);                    // 8 x 256 rows: the same A panel as 16 x 128
  const int grid = 2 * min(sms / 2, mblocks * nblocks);
  if (f16) {
    static const bool wide = []() {            // SE_CELL_EPI_WARPS=16: measured SLOWER (FullSubNet 118.7 vs 101.2 ms per batch:
      const char* e = getenv("SE_CELL_EPI_WARPS");   // 96 registers per thread with spills) -- kept for A/B runs only
      return e && atoi(e) == 16;
    }();
    if (epi == EPI_BIAS_ACT) return launch_pair_kernel<EPI_BIAS_ACT, true>(a0hi, a0lo, a1hi, a1lo, bhi, blo, p, grid, stream);
    return wide ? launch_pair_kernel<EPI_LSTM_CELL, true, 16>(a0hi, a0lo, a1hi, a1lo, bhi, blo, p, grid, stream)
                : launch_pair_kernel<EPI_LSTM_CELL, true>(a0hi, a0lo, a1hi, a1lo, bhi, blo, p, grid, stream);
  }
  return epi == EPI_BIAS_ACT ? launch_pair_kernel<EPI_BIAS_ACT>(a0hi, a0lo, a1hi, a1lo, bhi, blo, p, grid, stream)
                             : launch_pair_kernel<EPI_LSTM_CELL>(a0hi, a0lo, a1hi, a1lo, bhi, blo, p, grid, stream);
}

// gemm_tf32x3_kernel<EPI, CM, CN>: persistent grid of as many CM x CN clusters as are co-resident on the device (a
// cluster that had to wait for a second wave would double the run time).
template <int EPI, int CM, int CN, bool F16 = false>
static int launch_tc_cluster(const CUtensorMap* const* m, TcParams p, int sms, cudaStream_t stream) {
  constexpr int CSIZE = CM * CN;
  auto kernel = gemm_tf32x3_kernel<EPI, CM, CN, F16>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e != cudaSuccess) {
    set_error("tcgen05 gemm: smem attribute: %s", cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  const int mblocks = ceil_div(p.M, TC_BM * CM), nblocks = ceil_div(p.N, TC_BN * CN);
  p.panel_m = min(16 / CM, mblocks);
  if constexpr (CSIZE == 1) {
    const int grid = min(sms, mblocks * nblocks);
    kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(*m[0], *m[1], *m[2], *m[3], *m[4], *m[5], p);
    return SE_OK;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CSIZE;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  static int max_clusters = -1;           // per instantiation: co-resident clusters of this kernel on this device
  if (max_clusters < 0) {
    cfg.gridDim = dim3((unsigned)(sms / CSIZE * CSIZE));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)kernel, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = sms / CSIZE;
    }
    max_clusters = n;
  }
  cfg.gridDim = dim3((unsigned)(CSIZE * min(max_clusters, mblocks * nblocks)));
  e = cudaLaunchKernelEx(&cfg, kernel, *m[0], *m[1], *m[2], *m[3], *m[4], *m[5], p);
  if (e != cudaSuccess) {
    set_error("tcgen05 gemm (%d x %d cluster): launch: %s", CM, CN, cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

// f16: operands are fp16 pairs; K0 / K1 need not be multiples of the 64-element k-block (the TMA box is zero-filled past
// the end of a row), but B is laid out with source 0 padded to whole k-blocks: B = [K0 rounded up to 64 | K1].
static int launch_tc(int epi, const void* a0_hi, const void* a0_lo, long long lda0, int K0, const void* a1_hi,
                     const void* a1_lo, long long lda1, int K1, const void* b_hi, const void* b_lo, long long ldb,
                     TcParams p, cudaStream_t stream, bool f16 = false) {
  CUtensorMap m_a0hi, m_a0lo, m_a1hi, m_a1lo, m_bhi, m_blo;
  int rc;
  const int bk = f16 ? TC_BK16 : TC_BK;
  if ((rc = make_map(&m_a0hi, a0_hi, p.M, K0, lda0, f16))) return rc;
  if ((rc = make_map(&m_a0lo, a0_lo, p.M, K0, lda0, f16))) return rc;
  if (K1 > 0) {
    if ((rc = make_map(&m_a1hi, a1_hi, p.M, K1, lda1, f16))) return rc;
    if ((rc = make_map(&m_a1lo, a1_lo, p.M, K1, lda1, f16))) return rc;
  } else {
    m_a1hi = m_a0hi;
    m_a1lo = m_a0lo;
  }
  p.kb0 = ceil_div(K0, bk);
  p.kb1 = ceil_div(K1, bk);
  const int kb_total = K1 > 0 ? p.kb0 * bk + K1 : K0;      // columns of B that hold data
  if ((rc = make_map(&m_bhi, b_hi, p.N, kb_total, ldb, f16))) return rc;
  if ((rc = make_map(&m_blo, b_lo, p.N, kb_total, ldb, f16))) return rc;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int engine = gemm_engine();
  if (f16) {
    if ((engine == 1 || (engine == 5 && p.kb0 + p.kb1 >= kAutoPairMinKBlocksF16)) && p.M >= 2 * TC_BM && p.N >= T2_BN)
      return launch_tc_pair(epi, m_a0hi, m_a0lo, m_a1hi, m_a1lo, m_bhi, m_blo, p, sms, stream, true);
    const CUtensorMap* maps16[6] = {&m_a0hi, &m_a0lo, &m_a1hi, &m_a1lo, &m_bhi, &m_blo};
    return epi == EPI_BIAS_ACT ? launch_tc_cluster<EPI_BIAS_ACT, 1, 1, true>(maps16, p, sms, stream)
                               : launch_tc_cluster<EPI_LSTM_CELL, 1, 1, true>(maps16, p, sms, stream);
  }
  if ((engine == 1 || (engine == 5 && p.kb0 + p.kb1 >= kAutoPairMinKBlocks)) && p.M >= 2 * TC_BM && p.N >= T2_BN)
    return launch_tc_pair(epi, m_a0hi, m_a0lo, m_a1hi, m_a1lo, m_bhi, m_blo, p, sms, stream);
  const CUtensorMap* maps[6] = {&m_a0hi, &m_a0lo, &m_a1hi, &m_a1lo, &m_bhi, &m_blo};
  if (engine == 2 && p.M >= 2 * TC_BM && p.N >= 2 * TC_BN)
    return epi == EPI_BIAS_ACT ? launch_tc_cluster<EPI_BIAS_ACT, 2, 2>(maps, p, sms, stream) : launch_tc_cluster<EPI_LSTM_CELL, 2, 2>(maps, p, sms, stream);
  if (engine == 3 && p.N >= 2 * TC_BN)
    return epi == EPI_BIAS_ACT ? launch_tc_cluster<EPI_BIAS_ACT, 1, 2>(maps, p, sms, stream) : launch_tc_cluster<EPI_LSTM_CELL, 1, 2>(maps, p, sms, stream);
  if (engine == 4 && p.M >= 2 * TC_BM)
    return epi == EPI_BIAS_ACT ? launch_tc_cluster<EPI_BIAS_ACT, 2, 1>(maps, p, sms, stream) : launch_tc_cluster<EPI_LSTM_CELL, 2, 1>(maps, p, sms, stream);
  return epi == EPI_BIAS_ACT ? launch_tc_cluster<EPI_BIAS_ACT, 1, 1>(maps, p, sms, stream) : launch_tc_cluster<EPI_LSTM_CELL, 1, 1>(maps, p, sms, stream);
}

extern "C" int se_set_gemm_engine(int engine) {
  SE_REQUIRE(engine >= 0 && engine <= kMaxGemmEngine,
             "se_set_gemm_engine: 0 (one CTA per tile), 1 (CTA pairs), 2 / 3 / 4 (multicast clusters 2x2 / 1x2 / 2x1), 5 (auto)");
  g_gemm_engine = engine;
  return SE_OK;
}

extern "C" int se_gemm_tf32x3_ex(const float* a_hi, const float* a_lo, long long lda, const float* b_hi,
                                 const float* b_lo, long long ldb, int M, int N, int K, const float* bias, int act,
                                 float act_param, float alpha, const float* res, float* C, float* c_hi, float* c_lo,
                                 long long ldc, se_stream_t stream) {
  SE_REQUIRE(a_hi && a_lo && b_hi && b_lo && (C || c_hi), "se_gemm_tf32x3: null pointer");
  SE_REQUIRE((c_hi == nullptr) == (c_lo == nullptr), "se_gemm_tf32x3: c_hi/c_lo go together");
  SE_REQUIRE(M > 0 && N > 0 && K > 0 && K % TC_BK == 0, "se_gemm_tf32x3: K=%d must be a multiple of %d", K, TC_BK);
  SE_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0, "se_gemm_tf32x3: operand leading dims must be %% 4");
  SE_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && (((uintptr_t)C) & 3) == 0,
             "se_gemm_tf32x3: pointers must be 16-byte aligned");
  SE_REQUIRE((ldc & 3) != 0 || ((!C || aligned16(C)) && (!c_hi || (aligned16(c_hi) && aligned16(c_lo))) &&
                                (!res || aligned16(res))),
             "se_gemm_tf32x3: outputs must be 16-byte aligned when ldc %% 4 == 0");
  TcParams p{};
  p.M = M;
  p.N = N;
  p.bias = bias;
  p.act = act;
  p.act_param = act_param;
  p.alpha = alpha;
  p.res = res;
  p.C = C;
  p.c_hi = c_hi;
  p.c_lo = c_lo;
  p.ldc = ldc;
  p.out_scale = 1.0f;
  int rc = launch_tc(EPI_BIAS_ACT, a_hi, a_lo, lda, K, nullptr, nullptr, 0, 0, b_hi, b_lo, ldb, p, (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("se_gemm_tf32x3");
}

extern "C" int se_gemm_tf32x3(const float* a_hi, const float* a_lo, long long lda, const float* b_hi,
                              const float* b_lo, long long ldb, int M, int N, int K, const float* bias, int act,
                              float* C, long long ldc, se_stream_t stream) {
  return se_gemm_tf32x3_ex(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, bias, act, 0.f, 1.f, nullptr, C, nullptr, nullptr,
                           ldc, stream);
}

extern "C" int se_lstm_cell_tf32x3_ex(const float* x_hi, const float* x_lo, long long ldx, int Kx, const float* h_hi,
                                      const float* h_lo, long long ldh, int H, const float* w_hi, const float* w_lo,
                                      long long ldw, const float* bias, int M, float* c_state, float* h_hi_out,
                                      float* h_lo_out, float* h_out, long long ld_hout, int first_step,
                                      se_stream_t stream) {
  SE_REQUIRE(x_hi && x_lo && w_hi && w_lo && bias && c_state && h_hi_out && h_lo_out, "se_lstm_cell_tf32x3: null pointer");
  SE_REQUIRE(first_step || (h_hi && h_lo), "se_lstm_cell_tf32x3: state pointers are required after the first step");
  SE_REQUIRE(M > 0 && Kx > 0 && Kx % TC_BK == 0 && H > 0 && H % 32 == 0, "se_lstm_cell_tf32x3: Kx=%d H=%d (%%32)", Kx, H);
  SE_REQUIRE((ldx & 3) == 0 && (ldh & 3) == 0 && (ldw & 3) == 0 && (ld_hout & 3) == 0 && ld_hout >= H,
             "se_lstm_cell_tf32x3: leading dims must be %% 4 (ld_hout=%lld)", ld_hout);
  SE_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(h_hi) && aligned16(h_lo) && aligned16(w_hi) &&
                 aligned16(w_lo) && aligned16(c_state) && aligned16(h_hi_out) && aligned16(h_lo_out) &&
                 (!h_out || aligned16(h_out)) && aligned16(bias),
             "se_lstm_cell_tf32x3: pointers must be 16-byte aligned");
  SE_REQUIRE(first_step || (h_hi != h_hi_out && h_lo != h_lo_out), "se_lstm_cell_tf32x3: state must be double buffered");
  TcParams p{};
  p.M = M;
  p.N = 4 * H;
  p.bias = bias;
  p.c_state = c_state;
  p.h_hi = h_hi_out;
  p.h_lo = h_lo_out;
  p.h_out = h_out;
  p.H = H;
  p.ld_hout = ld_hout;
  p.first_step = first_step ? 1 : 0;
  p.out_scale = 1.0f;
  int rc = launch_tc(EPI_LSTM_CELL, x_hi, x_lo, ldx, Kx, h_hi, h_lo, ldh, first_step ? 0 : H, w_hi, w_lo, ldw, p,
                     (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("se_lstm_cell_tf32x3");
}

extern "C" int se_lstm_cell_tf32x3(const float* x_hi, const float* x_lo, long long ldx, int Kx, const float* h_hi,
                                   const float* h_lo, long long ldh, int H, const float* w_hi, const float* w_lo,
                                   long long ldw, const float* bias, int M, float* c_state, float* h_hi_out,
                                   float* h_lo_out, float* h_out, se_stream_t stream) {
  return se_lstm_cell_tf32x3_ex(x_hi, x_lo, ldx, Kx, h_hi, h_lo, ldh, H, w_hi, w_lo, ldw, bias, M, c_state, h_hi_out,
                                h_lo_out, h_out, H, 0, stream);
}

// ---- fp16-pair operands (kind::f16): half the MMAs and half the operand bytes of 3xTF32 for the same 22-bit products --
extern "C" int se_split_f16(const float* x, long long rows, int K, long long ldx, int Kpad, int scale_log2,
                            unsigned short* hi, unsigned short* lo, se_stream_t stream) {
  SE_REQUIRE(x && hi && lo && rows > 0 && K > 0 && ldx >= K && Kpad >= K && (Kpad & 7) == 0,
             "se_split_f16: rows=%lld K=%d ldx=%lld Kpad=%d (Kpad %% 8)", rows, K, ldx, Kpad);
  SE_REQUIRE(aligned16(hi) && aligned16(lo) && scale_log2 >= -14 && scale_log2 <= 15, "se_split_f16: alignment / scale_log2=%d",
             scale_log2);
  const int blocks = (int)min((long long)148 * 8, ceil_div_ll(rows * (Kpad / 4), 256));
  split_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, K, ldx, Kpad, ldexpf(1.0f, scale_log2), hi, lo);
  return check_launch("se_split_f16");
}

extern "C" int se_gemm_f16x3(const unsigned short* a_hi, const unsigned short* a_lo, long long lda,
                             const unsigned short* b_hi, const unsigned short* b_lo, long long ldb, int M, int N, int K,
                             int scale_log2_ab, const float* bias, int act, float act_param, float alpha,
                             const float* res, float* C, float* c_hi, float* c_lo, unsigned short* c16_hi,
                             unsigned short* c16_lo, int c16_scale_log2, long long ldc, se_stream_t stream) {
  SE_REQUIRE(a_hi && a_lo && b_hi && b_lo && (C || c_hi || c16_hi), "se_gemm_f16x3: null pointer");
  SE_REQUIRE((c_hi == nullptr) == (c_lo == nullptr) && (c16_hi == nullptr) == (c16_lo == nullptr),
             "se_gemm_f16x3: hi / lo outputs go together");
  SE_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "se_gemm_f16x3: K=%d must be a multiple of 8", K);
  SE_REQUIRE((lda & 7) == 0 && (ldb & 7) == 0, "se_gemm_f16x3: operand leading dims must be %% 8");
  SE_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && (((uintptr_t)C) & 3) == 0,
             "se_gemm_f16x3: pointers must be 16-byte aligned");
  SE_REQUIRE((ldc & 3) != 0 || ((!C || aligned16(C)) && (!c_hi || (aligned16(c_hi) && aligned16(c_lo))) &&
                                (!c16_hi || (aligned16(c16_hi) && aligned16(c16_lo))) && (!res || aligned16(res))),
             "se_gemm_f16x3: outputs must be 16-byte aligned when ldc %% 4 == 0");
  TcParams p{};
  p.M = M;
  p.N = N;
  p.bias = bias;
  p.act = act;
  p.act_param = act_param;
  p.alpha = alpha;
  p.res = res;
  p.C = C;
  p.c_hi = c_hi;
  p.c_lo = c_lo;
  p.c16_hi = c16_hi;
  p.c16_lo = c16_lo;
  p.c16_scale = ldexpf(1.0f, c16_scale_log2);
  p.ldc = ldc;
  p.out_scale = ldexpf(1.0f, -scale_log2_ab);
  int rc = launch_tc(EPI_BIAS_ACT, a_hi, a_lo, lda, K, nullptr, nullptr, 0, 0, b_hi, b_lo, ldb, p, (cudaStream_t)stream, true);
  if (rc) return rc;
  return check_launch("se_gemm_f16x3");
}

extern "C" int se_lstm_cell_f16x3(const unsigned short* x_hi, const unsigned short* x_lo, long long ldx, int Kx,
                                  const unsigned short* h_hi, const unsigned short* h_lo, long long ldh, int H,
                                  const unsigned short* w_hi, const unsigned short* w_lo, long long ldw,
                                  int scale_log2_a, int scale_log2_w, const float* bias, int M, float* c_state,
                                  unsigned short* h_hi_out, unsigned short* h_lo_out, float* h_out, long long ld_hout,
                                  int first_step, se_stream_t stream) {
  SE_REQUIRE(x_hi && x_lo && w_hi && w_lo && bias && c_state && h_hi_out && h_lo_out, "se_lstm_cell_f16x3: null pointer");
  SE_REQUIRE(first_step || (h_hi && h_lo), "se_lstm_cell_f16x3: state pointers are required after the first step");
  SE_REQUIRE(M > 0 && Kx > 0 && Kx % 8 == 0 && H > 0 && H % 32 == 0, "se_lstm_cell_f16x3: Kx=%d (%%8) H=%d (%%32)", Kx, H);
  SE_REQUIRE((ldx & 7) == 0 && (ldh & 7) == 0 && (ldw & 7) == 0 && (ld_hout & 7) == 0 && ld_hout >= H,
             "se_lstm_cell_f16x3: leading dims must be %% 8 (ld_hout=%lld)", ld_hout);
  SE_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(h_hi) && aligned16(h_lo) && aligned16(w_hi) &&
                 aligned16(w_lo) && aligned16(c_state) && aligned16(h_hi_out) && aligned16(h_lo_out) &&
                 (!h_out || aligned16(h_out)) && aligned16(bias),
             "se_lstm_cell_f16x3: pointers must be 16-byte aligned");
  SE_REQUIRE(first_step || (h_hi != h_hi_out && h_lo != h_lo_out), "se_lstm_cell_f16x3: state must be double buffered");
  TcParams p{};
  p.M = M;
  p.N = 4 * H;
  p.bias = bias;
  p.c_state = c_state;
  p.h_hi = reinterpret_cast<float*>(h_hi_out);     // fp16 data in the F16 kernels
  p.h_lo = reinterpret_cast<float*>(h_lo_out);
  p.h_out = h_out;
  p.H = H;
  p.ld_hout = ld_hout;
  p.first_step = first_step ? 1 : 0;
  p.out_scale = ldexpf(1.0f, -(scale_log2_a + scale_log2_w));
  p.c16_scale = ldexpf(1.0f, scale_log2_a);        // h leaves with the scale the next step's A operand expects
  int rc = launch_tc(EPI_LSTM_CELL, x_hi, x_lo, ldx, Kx, h_hi, h_lo, ldh, first_step ? 0 : H, w_hi, w_lo, ldw, p,
                     (cudaStream_t)stream, true);
  if (rc) return rc;
  return check_launch("se_lstm_cell_f16x3");
}
