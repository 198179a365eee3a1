// Causal Conv2d / ConvTranspose2d as an implicit GEMM on the tensor cores (tcgen05, 3xTF32).
//
// Same contract as se_conv_gemm (gemm.cu), for layers whose channel counts are multiples of 32:
//   out[b, t, dst_f0 + fo*dst_fstep, co] = act(bias[co] + sum_{tap, ci} in[b, t+dt, fo*sf+df, ci] W[co, tap, ci])
// The im2col matrix is never formed: for each (tap, 32-channel slice) the A tile -- Tbox frames x
// Fout output columns = up to 128 rows of 32 fp32 -- is ONE 4-D TMA box over the channels-last
// activation [B, T, F, C] with traversal stride sf on F, started at (c0, df, t0+dt, b).  Taps that
// fall outside the tensor (causal / look-ahead time pad, frequency pad, transposed-conv borders)
// are zero-filled by the TMA unit.  Skip connections (torch.cat) are a second tensor map.
// Operands are TF32 hi/lo pairs (see gemm_tc.cu); the epilogue can emit the split pair of its
// output so the next layer needs no separate split pass.
//
// F16 = true: the same kernel on fp16 operand pairs (kind::f16, see gemm_tc.cu / include/se_b200.h): a k-block is a 64-channel
// slice (the same 128-byte tile rows), channel counts that are not multiples of 64 are zero-filled by the TMA unit on the
// activation side and zero-padded in the packed weights ([Cout][ntaps * (pad64(C0) + pad64(C1))]).
#pragma once
#include "tc_common.cuh"

namespace se {

constexpr int CT_BM = 128, CT_BK = 32, CT_BK16 = 64;
constexpr int CT_CHUNK_KB = 4;
constexpr int CT_A_BYTES = CT_BM * CT_BK * 4;  // 16 KB per A tile (hi or lo)
// PAIR = 1: CTA pairs (tcgen05 cta_group::2, see tc_common.cuh and gemm_tc.cu): the pair multiplies TWO activation
// tiles (256 rows) by BN output channels, each CTA staging its own activation tile and HALF of the weight rows, so the
// weight traffic per SM halves and Cout = 256 layers get a 256-wide tile (otherwise two 128-wide passes over the
// activation).
template <int BN, int PAIR>
struct CtCfg {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;                   // weight rows this CTA stages
  static constexpr int B_BYTES = B_ROWS * CT_BK * 4;
  static constexpr int B_SLOT = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = 2 * CT_A_BYTES + 2 * B_SLOT;
  // Measured and dropped: TWO CTAs per SM for the narrow tiles (BN <= 64: 16 epilogue warps, two stages each) -- no gain
  // at BN = 16 / 32 (0.204 / 0.193 vs 0.201 / 0.192 ms), slower at BN = 64 (96 registers per thread: 0.180 vs 0.162 ms).
  static constexpr int CTAS_PER_SM = 1;
  static constexpr int STAGES = CTAS_PER_SM == 2 ? 2 : (STAGE_BYTES >= 64 * 1024 ? 3 : 4);     // 227 KB smem limit
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};
constexpr int CT_THREADS = 320;
constexpr int CT_EPI_WARPS = 8;

struct ConvTcParams {
  int B, T, Fout, Tbox;       // tile = Tbox frames x Fout columns (<= 128 rows)
  int ntaps;
  int dt[SE_MAX_TAPS], df[SE_MAX_TAPS];
  int kb0, kb1;               // 32-channel slices per tap from source 0 / source 1
  int Cout;
  const float* bias;
  int act;
  float act_param;
  float *out, *out_hi, *out_lo;  // any may be NULL
  // Two output-column parity classes of a transposed conv in ONE launch (cls_cols > 0): GEMM columns [0, cls_cols) are the
  // class written at output column dst_f0 + fo * dst_fstep, columns [cls_cols, 2 cls_cols) the class at dst_f0 + 1 +
  // fo * dst_fstep for fo < fout1; both have cls_cols channels and share the bias.  The tap list is the union (the second
  // class has zero weights on the taps it does not use), so the activation tiles are read once instead of twice.
  int cls_cols, fout1;
  float out_scale;               // fp32 sums -> values: 1 / (scale_A * scale_W) with fp16 pairs, 1 otherwise
  unsigned short *out16_hi, *out16_lo;   // fp16-pair copy of the output (scaled by out16_scale) or NULL
  float out16_scale;
  int dstF, dst_f0, dst_fstep;
  int a_bytes;                // bytes one A box writes (Tbox*Fout*128)
  // gated conv (GCRN/GCRN_noncprs.py:42-57): GEMM columns (2j, 2j+1) = (conv1, conv2) of output channel j; the epilogue
  // writes act((a * sigmoid(b)) * glu_scale[j] + glu_shift[j]) into a tensor of Cout / 2 channels
  int glu;
  const float *glu_scale, *glu_shift;
};

__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, unsigned leader_bar, void* dst, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int N>
__device__ __forceinline__ void tmem_ld_cols(unsigned taddr, float (&v)[N]);

template <>
__device__ __forceinline__ void tmem_ld_cols<64>(unsigned taddr, float (&v)[64]) {
  tmem_ld_32x64(taddr, v);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(unsigned taddr, float (&v)[32]) {
  tmem_ld_32x32(taddr, v);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(unsigned taddr, float (&v)[8]) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Store epilogue of one thread: NC consecutive output channels [n0, n0 + NC) of one output position (row offset
// `orow`): bias + activation, fp32 and / or the TF32 split for the next layer.  Cout % 4 == 0 (host check), so the
// channels go four at a time; like the GEMM epilogue (gemm_tc.cu) this code runs on the warps that drain the TMEM chunk
// sums, so it is kept short: activation as a template parameter, float4 bias loads.
template <int NC, int ACT>
__device__ __forceinline__ void conv_tc_store(const ConvTcParams& p, const float (&sum)[NC], long long orow, int n0, int cout) {
  const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  // the bias loads of 32 columns in one batch (one exposed latency instead of eight: two epilogue warps per scheduler
  // cannot hide a load per four columns -- the FFMAs behind these loads were the top stall sites of the kernel)
  constexpr int G = NC < 32 ? NC : 32;
  // rows of cout floats: 32-byte aligned pieces when cout % 8 == 0 and the bases are (every epilogue span starts at a
  // multiple of 8 columns)
  const bool v8 = (cout & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.out_hi) |
                                       reinterpret_cast<uintptr_t>(p.out_lo)) & 31) == 0 &&
                  ((reinterpret_cast<uintptr_t>(p.out16_hi) | reinterpret_cast<uintptr_t>(p.out16_lo)) & 15) == 0;
#pragma unroll
  for (int g0 = 0; g0 < NC; g0 += G) {
  float4 bias4[G / 4];
#pragma unroll
  for (int j = 0; j < G; j += 4) {
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + g0 + j < cout) {
      if (bias_vec) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + g0 + j));
      else if (p.bias) bb = make_float4(__ldg(p.bias + n0 + g0 + j), __ldg(p.bias + n0 + g0 + j + 1),
                                        __ldg(p.bias + n0 + g0 + j + 2), __ldg(p.bias + n0 + g0 + j + 3));
    }
    bias4[j / 4] = bb;
  }
  if (v8) {   // eight columns per step: 32-byte fp32 stores (whole sectors), 16-byte fp16 stores
#pragma unroll
    for (int jj = 0; jj < G; jj += 8) {
      const int j = g0 + jj;
      if (n0 + j >= cout) break;
      const float4 b0 = bias4[jj / 4], b1 = bias4[jj / 4 + 1];
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float os = p.out_scale;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = tc_act<ACT>(fmaf(sum[j + e], os, bb[e]), p.act_param);
      if (p.out) st_global_v8(p.out + orow + n0 + j, o);
      if (p.out_hi) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_tf32_dev(o[e], hi[e], lo[e]);
        st_global_v8(p.out_hi + orow + n0 + j, hi);
        st_global_v8(p.out_lo + orow + n0 + j, lo);
      }
      if (p.out16_hi) {
        unsigned short hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_f16_dev(o[e], p.out16_scale, hi[e], lo[e]);
        st_global_h8(p.out16_hi + orow + n0 + j, hi);
        st_global_h8(p.out16_lo + orow + n0 + j, lo);
      }
    }
    continue;
  }
#pragma unroll
  for (int jj = 0; jj < G; jj += 4) {
    const int j = g0 + jj;
    if (n0 + j >= cout) break;
    const float4 bb = bias4[jj / 4];
    const float os = p.out_scale;   // 1.0f on the TF32 path: fmaf(s, 1, b) == s + b
    const float o[4] = {tc_act<ACT>(fmaf(sum[j], os, bb.x), p.act_param), tc_act<ACT>(fmaf(sum[j + 1], os, bb.y), p.act_param),
                        tc_act<ACT>(fmaf(sum[j + 2], os, bb.z), p.act_param), tc_act<ACT>(fmaf(sum[j + 3], os, bb.w), p.act_param)};
    if (p.out) *reinterpret_cast<float4*>(p.out + orow + n0 + j) = make_float4(o[0], o[1], o[2], o[3]);
    if (p.out_hi) {
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32_dev(o[e], hi[e], lo[e]);
      *reinterpret_cast<float4*>(p.out_hi + orow + n0 + j) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(p.out_lo + orow + n0 + j) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (p.out16_hi) {
      unsigned short hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_f16_dev(o[e], p.out16_scale, hi[e], lo[e]);
      *reinterpret_cast<uint2*>(p.out16_hi + orow + n0 + j) = make_uint2(hi[0] | ((unsigned)hi[1] << 16), hi[2] | ((unsigned)hi[3] << 16));
      *reinterpret_cast<uint2*>(p.out16_lo + orow + n0 + j) = make_uint2(lo[0] | ((unsigned)lo[1] << 16), lo[2] | ((unsigned)lo[3] << 16));
    }
  }
  }
}

// Gated variant: the thread's NC columns are NC / 2 (conv1, conv2) pairs -> NC / 2 output channels starting at n0 / 2.
// `orow` is the row offset in the OUTPUT tensor (Cout / 2 channels).
template <int NC, int ACT>
__device__ __forceinline__ void conv_tc_store_glu(const ConvTcParams& p, const float (&sum)[NC], long long orow, int n0) {
  if constexpr (NC >= 16) {
    // 16 columns = 8 output channels per step: 32-byte stores (whole sectors) where the rows allow it
    const bool v8 = (p.Cout & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.out_hi) |
                                            reinterpret_cast<uintptr_t>(p.out_lo)) & 31) == 0 &&
                    ((reinterpret_cast<uintptr_t>(p.out16_hi) | reinterpret_cast<uintptr_t>(p.out16_lo)) & 15) == 0;
    if (v8) {
#pragma unroll
      for (int j = 0; j < NC; j += 16) {
        if (n0 + j >= p.Cout) break;
        const int co = (n0 + j) >> 1;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float ba = p.bias ? __ldg(p.bias + n0 + j + 2 * e) : 0.f, bg = p.bias ? __ldg(p.bias + n0 + j + 2 * e + 1) : 0.f;
          const float a = fmaf(sum[j + 2 * e], p.out_scale, ba), g = fmaf(sum[j + 2 * e + 1], p.out_scale, bg);
          float v = a * fast_sigmoid(g);
          v = v * (p.glu_scale ? __ldg(p.glu_scale + co + e) : 1.f) + (p.glu_shift ? __ldg(p.glu_shift + co + e) : 0.f);
          o[e] = tc_act<ACT>(v, p.act_param);
        }
        if (p.out) st_global_v8(p.out + orow + co, o);
        if (p.out_hi) {
          float hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_tf32_dev(o[e], hi[e], lo[e]);
          st_global_v8(p.out_hi + orow + co, hi);
          st_global_v8(p.out_lo + orow + co, lo);
        }
        if (p.out16_hi) {
          unsigned short hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_f16_dev(o[e], p.out16_scale, hi[e], lo[e]);
          st_global_h8(p.out16_hi + orow + co, hi);
          st_global_h8(p.out16_lo + orow + co, lo);
        }
      }
      return;
    }
  }
  if constexpr (NC >= 8) {
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
      if (n0 + j >= p.Cout) break;          // Cout % 8 == 0 (host check)
      float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.bias) {
#pragma unroll
        for (int e = 0; e < 8; ++e) bb[e] = __ldg(p.bias + n0 + j + e);
      }
      const int co = (n0 + j) >> 1;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = fmaf(sum[j + 2 * e], p.out_scale, bb[2 * e]), g = fmaf(sum[j + 2 * e + 1], p.out_scale, bb[2 * e + 1]);
        float v = a * fast_sigmoid(g);     // ex2.approx + fast divide (|error| ~ 2e-7), as in the fused LSTM cell epilogue
        v = v * (p.glu_scale ? __ldg(p.glu_scale + co + e) : 1.f) + (p.glu_shift ? __ldg(p.glu_shift + co + e) : 0.f);
        o[e] = tc_act<ACT>(v, p.act_param);
      }
      if (p.out) *reinterpret_cast<float4*>(p.out + orow + co) = make_float4(o[0], o[1], o[2], o[3]);
      if (p.out_hi) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32_dev(o[e], hi[e], lo[e]);
        *reinterpret_cast<float4*>(p.out_hi + orow + co) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(p.out_lo + orow + co) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      if (p.out16_hi) {
        unsigned short hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_f16_dev(o[e], p.out16_scale, hi[e], lo[e]);
        *reinterpret_cast<uint2*>(p.out16_hi + orow + co) = make_uint2(hi[0] | ((unsigned)hi[1] << 16), hi[2] | ((unsigned)hi[3] << 16));
        *reinterpret_cast<uint2*>(p.out16_lo + orow + co) = make_uint2(lo[0] | ((unsigned)lo[1] << 16), lo[2] | ((unsigned)lo[3] << 16));
      }
    }
  }
}

template <int BN, int PAIR, bool F16>
__global__ void __launch_bounds__(CT_THREADS, CtCfg<BN, PAIR>::CTAS_PER_SM)
conv_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a0hi, const __grid_constant__ CUtensorMap map_a0lo,
                   const __grid_constant__ CUtensorMap map_a1hi, const __grid_constant__ CUtensorMap map_a1lo,
                   const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                   const ConvTcParams p) {
  using Cfg = CtCfg<BN, PAIR>;
  constexpr int B_ROWS = Cfg::B_ROWS;
  constexpr int B_BYTES = Cfg::B_BYTES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int B_SLOT = Cfg::B_SLOT;
  constexpr int CT_STAGES = Cfg::STAGES;
  constexpr int EPI_COLS = BN / 2;
  constexpr int BK = F16 ? CT_BK16 : CT_BK;                     // channels per k-block (128 bytes either way)
  constexpr int NACC = BN >= 256 ? 2 : 4;                       // TMEM accumulator ring (512 columns in all)
  constexpr int TMEM_COLS = (NACC * BN < 32) ? 32 : NACC * BN;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + CT_STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + CT_STAGES;
  uint64_t* tfull = bars + 2 * CT_STAGES;
  uint64_t* tempty = tfull + NACC;
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs of the pair)
  const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nctas = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int ttiles = ceil_div(p.T, p.Tbox);
  const int mtiles = p.B * ttiles;                              // activation tiles (Tbox frames of one clip)
  const int munits = PAIR ? ceil_div(mtiles, 2) : mtiles;       // a pair takes tiles 2u (leader) and 2u + 1 (peer)
  const int ntiles_n = ceil_div(p.Cout, BN);
  const int ntiles = munits * ntiles_n;
  const int kb_per_tap = p.kb0 + p.kb1;
  const int kblocks = p.ntaps * kb_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < CT_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], (PAIR ? 2 : 1) * CT_EPI_WARPS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_a0hi);
    tma_prefetch_desc(&map_a0lo);
    tma_prefetch_desc(&map_bhi);
    tma_prefetch_desc(&map_blo);
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS);   // the same warp of both CTAs, same smem slot
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    __syncwarp();
    cluster_sync_all();   // barriers of both CTAs initialised, both allocations done, before any remote signal
  }
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;

  // n fastest: consecutive CTAs share the same activation tile (L2) across output-channel tiles.  In pair mode an odd
  // tile count leaves the last peer without a tile: b == B then, its TMA boxes are out of bounds (zero filled) and its
  // epilogue skips the store.
  auto tile_coords = [&](int tile, int& b, int& t0, int& nb) {
    nb = tile % ntiles_n;
    const int mt = (tile / ntiles_n) * (PAIR ? 2 : 1) + (int)rank;
    b = mt / ttiles;
    t0 = (mt - b * ttiles) * p.Tbox;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      unsigned phase = 0;
      for (int tile = cta; tile < ntiles; tile += nctas) {
        int b, t0, nb;
        tile_coords(tile, b, t0, nb);
        for (int kb = 0; kb < kblocks; ++kb) {
          const int tap = kb / kb_per_tap;
          const int r = kb - tap * kb_per_tap;
          mbar_wait_parity(&empty[stage], phase ^ 1);
          unsigned char* st = tiles + stage * STAGE_BYTES;
          const unsigned stage_tx = 2u * (unsigned)p.a_bytes + 2u * (unsigned)B_BYTES;
          const int tt = t0 + p.dt[tap], ff = p.df[tap];
          if constexpr (PAIR) {
            const unsigned lbar = smem_u32(&full[stage]) & T2_PEER_BIT_MASK;
            if (rank == 0) mbar_expect_tx(&full[stage], 2u * stage_tx);   // both CTAs' bytes land on this barrier
            if (r < p.kb0) {
              tma_load_4d_pair(&map_a0hi, lbar, st, r * BK, ff, tt, b);
              tma_load_4d_pair(&map_a0lo, lbar, st + CT_A_BYTES, r * BK, ff, tt, b);
            } else {
              tma_load_4d_pair(&map_a1hi, lbar, st, (r - p.kb0) * BK, ff, tt, b);
              tma_load_4d_pair(&map_a1lo, lbar, st + CT_A_BYTES, (r - p.kb0) * BK, ff, tt, b);
            }
            const int brow = nb * BN + (int)rank * B_ROWS;
            tma_load_2d_pair(&map_bhi, lbar, st + 2 * CT_A_BYTES, kb * BK, brow);
            tma_load_2d_pair(&map_blo, lbar, st + 2 * CT_A_BYTES + B_SLOT, kb * BK, brow);
          } else {
            mbar_expect_tx(&full[stage], stage_tx);
            if (r < p.kb0) {
              tma_load_4d(&map_a0hi, &full[stage], st, r * BK, ff, tt, b);
              tma_load_4d(&map_a0lo, &full[stage], st + CT_A_BYTES, r * BK, ff, tt, b);
            } else {
              tma_load_4d(&map_a1hi, &full[stage], st, (r - p.kb0) * BK, ff, tt, b);
              tma_load_4d(&map_a1lo, &full[stage], st + CT_A_BYTES, (r - p.kb0) * BK, ff, tt, b);
            }
            tma_load_2d(&map_bhi, &full[stage], st + 2 * CT_A_BYTES, kb * BK, nb * BN);
            tma_load_2d(&map_blo, &full[stage], st + 2 * CT_A_BYTES + B_SLOT, kb * BK, nb * BN);
          }
          if (++stage == CT_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr unsigned idesc = F16 ? make_idesc_f16(PAIR ? 2 * CT_BM : CT_BM, BN) : make_idesc_tf32(PAIR ? 2 * CT_BM : CT_BM, BN);
      int stage = 0;
      unsigned phase = 0;
      int acc = 0;
      unsigned acc_phase = 0;
      for (int tile = cta; tile < ntiles; tile += nctas) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const bool chunk_start = (kb % CT_CHUNK_KB) == 0;
          if (chunk_start) {
            mbar_wait_parity(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
          }
          const unsigned d_tmem = tmem_base + (unsigned)(acc * BN);
          mbar_wait_parity(&full[stage], phase);
          tc_fence_after();
          unsigned char* st = tiles + stage * STAGE_BYTES;
          const uint64_t d_ahi = make_smem_desc(st);
          const uint64_t d_alo = make_smem_desc(st + CT_A_BYTES);
          const uint64_t d_bhi = make_smem_desc(st + 2 * CT_A_BYTES);
          const uint64_t d_blo = make_smem_desc(st + 2 * CT_A_BYTES + B_SLOT);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);   // 32 bytes per MMA K slice (8 tf32 / 16 fp16)
            const unsigned accum = (!chunk_start || k > 0) ? 1u : 0u;
            if constexpr (PAIR && F16) {
              umma_f16_pair(d_tmem, d_alo + adv, d_bhi + adv, idesc, accum);
              umma_f16_pair(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_f16_pair(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            } else if constexpr (PAIR) {
              umma_tf32_pair(d_tmem, d_alo + adv, d_bhi + adv, idesc, accum);
              umma_tf32_pair(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_tf32_pair(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            } else if constexpr (F16) {
              umma_f16(d_tmem, d_alo + adv, d_bhi + adv, idesc, accum);
              umma_f16(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_f16(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            } else {
              umma_tf32(d_tmem, d_alo + adv, d_bhi + adv, idesc, accum);
              umma_tf32(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              umma_tf32(d_tmem, d_ahi + adv, d_bhi + adv, idesc, 1u);
            }
          }
          if constexpr (PAIR) umma_commit_pair(&empty[stage]);
          else umma_commit(&empty[stage]);
          if (++stage == CT_STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if ((kb % CT_CHUNK_KB) == CT_CHUNK_KB - 1 || kb == kblocks - 1) {
            if constexpr (PAIR) umma_commit_pair(&tfull[acc]);
            else umma_commit(&tfull[acc]);
            if (++acc == NACC) {
              acc = 0;
              acc_phase ^= 1;
            }
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    unsigned acc_phase = 0;
    const int nchunks = (kblocks + CT_CHUNK_KB - 1) / CT_CHUNK_KB;
    for (int tile = cta; tile < ntiles; tile += nctas) {
      int b, t0, nb;
      tile_coords(tile, b, t0, nb);
      float sum[EPI_COLS];
#pragma unroll
      for (int j = 0; j < EPI_COLS; ++j) sum[j] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait_parity(&tfull[acc], acc_phase);
        tc_fence_after();
        {
          constexpr int PIECE = EPI_COLS > 64 ? 32 : EPI_COLS;   // 128 columns go through registers 32 at a time
#pragma unroll
          for (int piece = 0; piece < EPI_COLS / PIECE; ++piece) {
            float v[PIECE];
            const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16) +
                                   (unsigned)(acc * BN + half * EPI_COLS + piece * PIECE);
            tmem_ld_cols<PIECE>(taddr, v);
#pragma unroll
            for (int j = 0; j < PIECE; ++j) sum[piece * PIECE + j] += v[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(&tempty[acc], 0);   // the leader's MMA thread waits for both CTAs
          else mbar_arrive(&tempty[acc]);
        }
        if (++acc == NACC) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      const int r = quarter * 32 + lane;  // row of the tile = (frame, column)
      const int tl = r / p.Fout;
      const int fo = r - tl * p.Fout;
      const int t = t0 + tl;
      if (tl >= p.Tbox || t >= p.T || b >= p.B) continue;
      int n0 = nb * BN + half * EPI_COLS;
      int cout = p.Cout, cls = 0;
      if (p.cls_cols) {                   // EPI_COLS divides cls_cols (host check): a thread's columns are of one class
        cls = n0 >= p.cls_cols ? 1 : 0;
        n0 -= cls * p.cls_cols;
        cout = p.cls_cols;
        if (cls && fo >= p.fout1) continue;
      }
      const long long opos = ((long long)b * p.T + t) * p.dstF + p.dst_f0 + cls + (long long)fo * p.dst_fstep;
      const long long orow = opos * (long long)cout;
      if (p.glu) {
        const long long grow = opos * (long long)(p.Cout >> 1);
        if (p.act == SE_ACT_ELU) conv_tc_store_glu<EPI_COLS, SE_ACT_ELU>(p, sum, grow, n0);
        else conv_tc_store_glu<EPI_COLS, SE_ACT_NONE>(p, sum, grow, n0);
        continue;
      }
      switch (p.act) {   // uniform: one branch per tile, the activation itself is a template parameter
        case SE_ACT_PRELU: conv_tc_store<EPI_COLS, SE_ACT_PRELU>(p, sum, orow, n0, cout); break;
        case SE_ACT_ELU: conv_tc_store<EPI_COLS, SE_ACT_ELU>(p, sum, orow, n0, cout); break;
        case SE_ACT_SOFTPLUS: conv_tc_store<EPI_COLS, SE_ACT_SOFTPLUS>(p, sum, orow, n0, cout); break;
        case SE_ACT_RELU: conv_tc_store<EPI_COLS, SE_ACT_RELU>(p, sum, orow, n0, cout); break;
        case SE_ACT_SIGMOID: conv_tc_store<EPI_COLS, SE_ACT_SIGMOID>(p, sum, orow, n0, cout); break;
        case SE_ACT_TANH: conv_tc_store<EPI_COLS, SE_ACT_TANH>(p, sum, orow, n0, cout); break;
        default: conv_tc_store<EPI_COLS, SE_ACT_NONE>(p, sum, orow, n0, cout); break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    __syncwarp();
    cluster_sync_all();   // the peer's TMEM / barriers stay alive until the leader's last MMA and commit have landed
  }
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static inline int make_act_map(CUtensorMap* map, const void* ptr, int B, int T, int F, int C, int Fout, int sf, int Tbox, bool f16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SE_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t es = f16 ? 2 : 4;
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)F * C * es, (cuuint64_t)T * F * C * es};
  cuuint32_t box[4] = {(cuuint32_t)(f16 ? CT_BK16 : CT_BK), (cuuint32_t)((Fout - 1) * sf + 1), (cuuint32_t)Tbox, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)sf, 1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed with CUresult %d (B=%d T=%d F=%d C=%d Fout=%d sf=%d Tbox=%d)", (int)r, B,
              T, F, C, Fout, sf, Tbox);
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

static inline int make_w_map(CUtensorMap* map, const void* ptr, int rows, int K, int box_rows, bool f16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SE_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(f16 ? CT_BK16 : CT_BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

template <int BN, int PAIR, bool F16>
static int launch_conv_tc(const CUtensorMap* m, const ConvTcParams& p, int sms, cudaStream_t s) {
  constexpr int SMEM = CtCfg<BN, PAIR>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(conv_tf32x3_kernel<BN, PAIR, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  if (e != cudaSuccess) {
    set_error("se_conv_tf32x3: smem attribute: %s", cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  const int mtiles = p.B * ceil_div(p.T, p.Tbox);
  if constexpr (PAIR) {
    const int tiles = ceil_div(mtiles, 2) * ceil_div(p.Cout, BN);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * min(sms / 2, tiles)));
    cfg.blockDim = dim3(CT_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, conv_tf32x3_kernel<BN, PAIR, F16>, m[0], m[1], m[2], m[3], m[4], m[5], p);
    if (e != cudaSuccess) {
      set_error("se_conv_tf32x3 (CTA pairs): cluster launch: %s", cudaGetErrorString(e));
      return SE_ERR_CUDA;
    }
  } else {
    const int tiles = mtiles * ceil_div(p.Cout, BN);
    conv_tf32x3_kernel<BN, PAIR, F16><<<min(sms * CtCfg<BN, PAIR>::CTAS_PER_SM, tiles), CT_THREADS, SMEM, s>>>(m[0], m[1], m[2], m[3], m[4], m[5], p);
  }
  return SE_OK;
}


// Arguments of one launch, common to the TF32-pair (se_conv_tf32x3) and fp16-pair (se_conv_f16x3) entry points.
struct ConvTcArgs {
  const void *src0_hi, *src0_lo, *src1_hi, *src1_lo;
  int C0, C1, B, T, Fin, Fout, ntaps;
  const int *dt, *df;
  int sf;
  const void *w_hi, *w_lo;
  const float* bias;
  int Cout, act;
  float act_param;
  float *out, *out_hi, *out_lo;
  unsigned short *out16_hi, *out16_lo;
  float out_scale, out16_scale;
  int dstF, dst_f0, dst_fstep, glu;
  const float *glu_scale, *glu_shift;
  int ncls, fout1;                 // ncls = 2: two parity classes of Cout / 2 channels (see ConvTcParams::cls_cols)
};

template <bool F16>
static int conv_tc_run(const ConvTcArgs* d, const char* what, cudaStream_t s) {
  constexpr int BK = F16 ? CT_BK16 : CT_BK;
  SE_REQUIRE(d->src0_hi && d->src0_lo && d->w_hi && d->w_lo, "%s: null pointer", what);
  SE_REQUIRE(d->out || d->out_hi || d->out16_hi, "%s: no output", what);
  SE_REQUIRE((d->out_hi == nullptr) == (d->out_lo == nullptr) && (d->out16_hi == nullptr) == (d->out16_lo == nullptr),
             "%s: hi / lo outputs go together", what);
  if constexpr (F16) {
    SE_REQUIRE(d->C0 > 0 && d->C0 % 8 == 0 && d->C1 >= 0 && d->C1 % 8 == 0, "%s: C0=%d C1=%d (%%8)", what, d->C0, d->C1);
  } else {
    SE_REQUIRE(d->C0 > 0 && d->C0 % CT_BK == 0 && d->C1 >= 0 && d->C1 % CT_BK == 0, "%s: C0=%d C1=%d (%%32)", what, d->C0, d->C1);
  }
  SE_REQUIRE(d->C1 == 0 || (d->src1_hi && d->src1_lo), "%s: second source missing", what);
  SE_REQUIRE(d->ntaps >= 1 && d->ntaps <= SE_MAX_TAPS, "%s: ntaps=%d", what, d->ntaps);
  SE_REQUIRE(d->Fout >= 1 && d->Fout <= CT_BM && d->sf >= 1 && (d->Fout - 1) * d->sf + 1 <= 256,
             "%s: Fout=%d sf=%d unsupported", what, d->Fout, d->sf);
  SE_REQUIRE(d->Cout >= 4 && d->Cout % 4 == 0, "%s: Cout=%d (%%4)", what, d->Cout);
  ConvTcParams p{};
  p.B = d->B;
  p.T = d->T;
  p.Fout = d->Fout;
  p.Tbox = min(CT_BM / d->Fout, d->T);
  if (p.Tbox > 256) p.Tbox = 256;
  p.ntaps = d->ntaps;
  for (int i = 0; i < d->ntaps; ++i) {
    p.dt[i] = d->dt[i];
    p.df[i] = d->df[i];
  }
  p.kb0 = ceil_div(d->C0, BK);
  p.kb1 = ceil_div(d->C1, BK);
  p.Cout = d->Cout;
  p.bias = d->bias;
  p.act = d->act;
  p.act_param = d->act_param;
  p.out = d->out;
  p.out_hi = d->out_hi;
  p.out_lo = d->out_lo;
  p.out16_hi = d->out16_hi;
  p.out16_lo = d->out16_lo;
  p.out_scale = d->out_scale;
  p.out16_scale = d->out16_scale;
  p.dstF = d->dstF;
  p.dst_f0 = d->dst_f0;
  p.dst_fstep = d->dst_fstep;
  p.a_bytes = p.Tbox * p.Fout * 128;       // one A box: Tbox * Fout rows of 128 bytes (32 fp32 or 64 fp16 channels)
  p.glu = d->glu;
  p.glu_scale = d->glu_scale;
  p.glu_shift = d->glu_shift;
  SE_REQUIRE(!d->glu || (d->Cout % 8 == 0 && d->Cout >= 32 && (d->act == SE_ACT_ELU || d->act == SE_ACT_NONE)),
             "%s (gated): Cout=%d must be a multiple of 8 and >= 32, act ELU or none", what, d->Cout);
  const int K = d->ntaps * (p.kb0 + p.kb1) * BK;      // columns of the packed weights (per-source padding with F16)
  // engine 1 (se_set_gemm_engine): CTA pairs where at least two activation tiles exist and the tile is >= 64 wide
  const bool pair = gemm_engine_is_pair() && d->Cout > 32 && (long long)d->B * ceil_div(d->T, p.Tbox) >= 2;
  const int BN = (pair && d->Cout > 128) ? 256 : (d->Cout > 64 ? 128 : (d->Cout > 32 ? 64 : (d->Cout > 16 ? 32 : 16)));
  if (d->ncls == 2) {
    SE_REQUIRE(!d->glu && d->Cout % 8 == 0 && (d->Cout / 2) % (BN / 2) == 0 && d->fout1 >= 0 && d->fout1 <= d->Fout,
               "%s (two parity classes): Cout=%d must be 2 x a multiple of %d, fout1=%d <= Fout=%d", what, d->Cout, BN / 2,
               d->fout1, d->Fout);
    p.cls_cols = d->Cout / 2;
    p.fout1 = d->fout1;
  } else {
    SE_REQUIRE(d->ncls == 0 || d->ncls == 1, "%s: ncls=%d", what, d->ncls);
  }
  CUtensorMap m[6];
  int rc;
  if ((rc = make_act_map(&m[0], d->src0_hi, d->B, d->T, d->Fin, d->C0, d->Fout, d->sf, p.Tbox, F16))) return rc;
  if ((rc = make_act_map(&m[1], d->src0_lo, d->B, d->T, d->Fin, d->C0, d->Fout, d->sf, p.Tbox, F16))) return rc;
  if (d->C1 > 0) {
    if ((rc = make_act_map(&m[2], d->src1_hi, d->B, d->T, d->Fin, d->C1, d->Fout, d->sf, p.Tbox, F16))) return rc;
    if ((rc = make_act_map(&m[3], d->src1_lo, d->B, d->T, d->Fin, d->C1, d->Fout, d->sf, p.Tbox, F16))) return rc;
  } else {
    m[2] = m[0];
    m[3] = m[1];
  }
  if ((rc = make_w_map(&m[4], d->w_hi, d->Cout, K, pair ? BN / 2 : BN, F16))) return rc;
  if ((rc = make_w_map(&m[5], d->w_lo, d->Cout, K, pair ? BN / 2 : BN, F16))) return rc;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  switch (BN) {
    case 256: rc = launch_conv_tc<256, 1, F16>(m, p, sms, s); break;
    case 128: rc = pair ? launch_conv_tc<128, 1, F16>(m, p, sms, s) : launch_conv_tc<128, 0, F16>(m, p, sms, s); break;
    case 64: rc = pair ? launch_conv_tc<64, 1, F16>(m, p, sms, s) : launch_conv_tc<64, 0, F16>(m, p, sms, s); break;
    case 32: rc = launch_conv_tc<32, 0, F16>(m, p, sms, s); break;
    default: rc = launch_conv_tc<16, 0, F16>(m, p, sms, s); break;
  }
  if (rc) return rc;
  return check_launch(what);
}

}  // namespace se
