// Persistent LSTM recurrence (nn.LSTM, gate order i,f,g,o; CRN/CRN.py:20,29, LSTM/LSTM.py:17-18).
//
// The input projection for all T steps is hoisted into one GEMM (se_conv_gemm); this kernel
// runs the T dependent steps  g_t = xproj_t + h_{t-1} W_hh^T  in ONE launch:
//   * one CTA per slice of 8 hidden units (= 32 gate columns); H = 1024 -> 128 CTAs,
//     co-resident on the 148 SMs (cooperative launch);
//   * the CTA's W_hh slice [H][32] fp32 (128 KB at H = 1024) stays in shared memory for
//     the whole sequence -- HBM sees W_hh once per launch instead of once per step;
//   * h_{t-1} is exchanged through a k-major scratch hT[2][H][64] in L2 (double buffered
//     by step parity) and streamed into shared memory with cp.async.cg in 16-row chunks,
//     one private double-buffered pipeline per warp (8 warps split K);
//   * per-warp 64x32 partial tiles (8x8 per lane, packed FFMA2) are reduced through
//     shared memory, gates/cell update fused, h_t written to hseq and hT;
//   * steps are separated by a device-wide counter barrier (release/acquire).
// FP32 throughout: see gemm.cu for why tensor cores are not used on this path yet.
#include <stdlib.h>

#include "common.cuh"

namespace se {

constexpr int kLstmThreads = 256;
constexpr int kLstmWarps = 8;
constexpr int kHU = 8;       // hidden units per CTA
constexpr int kNC = 4 * kHU; // gate columns per CTA
constexpr int kBT = 64;      // batch tile (rows of the per-step GEMM)
constexpr int kKC = 16;      // k rows per pipeline stage per warp
constexpr int LT_H_PUBLIC = 1024;   // hidden size the tcgen05 engine is built for
constexpr int kRep = 1;      // replicas of the exchanged state h_t (CTA s reads copy s % kRep).  Measured on B200: 8 copies
                             // do NOT help (13.7 -> 14.4 us/step): L2 same-line contention is not what bounds a step

struct LstmParams {
  const float* xproj;  // [B, T, xp_stride >= 4H] slice-ordered columns
  long long xp_stride;
  const float* whh;    // [H/8][H][32]
  int B, T, H;
  float* hseq;
  long long hs_sb, hs_st;
  float* hT;           // [ngroups][2][H][64]
  unsigned* sync;      // [ngroups] step counters
  // several independent LSTMs of the same shape in one launch (DCCRN's four real passes per complex
  // LSTM layer): group g reads xproj columns [g*xp_goff, +4H), weights whh + g*whh_gstride, writes
  // hseq columns [g*hs_goff, +H).  CTA = (group, slice).
  int ngroups;
  long long xp_goff, whh_gstride, hs_goff;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kLstmThreads, 1) lstm_seq_kernel(const LstmParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* Ws = reinterpret_cast<float*>(smem_raw);                 // [H][32]
  float* stage = Ws + (size_t)p.H * kNC;                           // [8 warps][2][kKC][64]  (aliased by red)
  float* cst = stage + kLstmWarps * 2 * kKC * kBT;                 // [64][8] cell state
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H;
  const int G = H / kHU;                 // CTAs per group
  const int group = blockIdx.x / G;
  const int slice = blockIdx.x - group * G;
  const float* xproj = p.xproj + group * p.xp_goff;
  const float* whh = p.whh + group * p.whh_gstride;
  float* hseq = p.hseq + group * p.hs_goff;
  float* hT = p.hT + (size_t)group * kRep * 2 * H * kBT;
  unsigned* sync = p.sync + 2 * group;
  const float* hT_rd = hT + (size_t)(slice % kRep) * 2 * H * kBT;

  // resident weights
  {
    const float4* src = reinterpret_cast<const float4*>(whh + (size_t)slice * H * kNC);
    float4* dst = reinterpret_cast<float4*>(Ws);
    for (int i = tid; i < H * kNC / 4; i += kLstmThreads) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < kBT * kHU; i += kLstmThreads) cst[i] = 0.f;
  __syncthreads();

  const int bg = lane >> 2, cg = lane & 3;  // 8 batch groups x 4 column groups per warp
  const int kper = H / kLstmWarps;          // k range of this warp
  const int kbase = warp * kper;
  const int nchunk = kper / kKC;
  float* my_stage = stage + warp * (2 * kKC * kBT);

  for (int t = 0; t < p.T; ++t) {
    // -- prefetch this step's input projection (independent of the recurrence) -------------
    float xg[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * kLstmThreads;
      const int b = idx >> 3, j = idx & 7;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        xg[r][g] = 0.f;
        if (b < p.B)
          xg[r][g] = __ldg(xproj + ((size_t)b * p.T + t) * (size_t)p.xp_stride + slice * kNC + g * kHU + j);
      }
    }

    if (t > 0) {
      // -- wait until every CTA has published h_{t-1} ------------------------------------------
      if (tid == 0) {
        const unsigned target = (unsigned)t * (unsigned)G;
        while (ld_acquire_u32(sync) < target) {
        }
      }
      __syncthreads();

      const float* hprev = hT_rd + (size_t)((t - 1) & 1) * H * kBT;
      float2 acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

      auto issue = [&](int c, int buf) {
        // 16 rows x 256 B = 256 x 16-byte pieces, 8 per lane
        const float* g = hprev + (size_t)(kbase + c * kKC) * kBT;
        float* s = my_stage + buf * (kKC * kBT);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int piece = lane + q * 32;
          cp_async16(s + piece * 4, g + piece * 4);
        }
        cp_async_commit();
      };

      issue(0, 0);
      for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) {
          issue(c + 1, buf ^ 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        const float* hs = my_stage + buf * (kKC * kBT);
        const float* wk = Ws + (size_t)(kbase + c * kKC) * kNC;
#pragma unroll
        for (int k = 0; k < kKC; ++k) {
          const float4 h0 = *reinterpret_cast<const float4*>(hs + k * kBT + bg * 8);
          const float4 h1 = *reinterpret_cast<const float4*>(hs + k * kBT + bg * 8 + 4);
          const float4 w0 = *reinterpret_cast<const float4*>(wk + k * kNC + cg * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wk + k * kNC + cg * 8 + 4);
          const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                make_float2(w1.z, w1.w)};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 hh = make_float2(hv[i], hv[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = ffma2(hh, wv[j], acc[i][j]);
          }
        }
        __syncwarp();  // all lanes done with this buffer before it is refilled
      }

      // -- per-warp partial tile -> shared (aliases this warp's own stage buffers) --------------
      // red[warp][b][col'], col' = ((col>>3) ^ (b&3))*8 + (col&7)  (bank swizzle for the gate phase)
      float* red = my_stage;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int b = bg * 8 + i;
        const int slot = cg ^ (b & 3);
        float* dst = red + b * kNC + slot * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y);
      }
      __syncthreads();
    }

    // -- gates, cell and hidden update for this CTA's 8 units -------------------------------------
    float* hcur = hT + (size_t)(t & 1) * H * kBT;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * kLstmThreads;
      const int b = idx >> 3, j = idx & 7;
      float g4[4] = {xg[r][0], xg[r][1], xg[r][2], xg[r][3]};
      if (t > 0) {
#pragma unroll
        for (int w = 0; w < kLstmWarps; ++w) {
          const float* red = stage + w * (2 * kKC * kBT) + b * kNC;
#pragma unroll
          for (int g = 0; g < 4; ++g) g4[g] += red[((g ^ (b & 3)) << 3) + j];
        }
      }
      const float ig = sigmoid_f(g4[0]);
      const float fg = sigmoid_f(g4[1]);
      const float gg = tanhf(g4[2]);
      const float og = sigmoid_f(g4[3]);
      const float c = fg * cst[idx] + ig * gg;
      const float h = og * tanhf(c);
      cst[idx] = c;
      const int u = slice * kHU + j;
#pragma unroll
      for (int rep = 0; rep < kRep; ++rep) hcur[(size_t)rep * 2 * H * kBT + (size_t)u * kBT + b] = (b < p.B) ? h : 0.f;
      if (b < p.B) hseq[(size_t)b * p.hs_sb + (size_t)t * p.hs_st + u] = h;
    }
    __syncthreads();
    if (tid == 0 && t + 1 < p.T) {
      __threadfence();
      red_release_add(sync, 1u);
    }
  }
}


// ================================================================================================
// Sequence-parallel recurrence for SMALL hidden sizes (H = 128: DPCRN's inter-chunk LSTM, DCCRN's real / imaginary
// LSTMs).  W_hh is only 4H x H x 4 B = 256 KB there, so ONE CTA can hold all of it -- half in registers (64 per
// thread: thread j owns gate column j, k = 0..63), half in shared memory ([64][512] floats, conflict-free) -- and run
// NS whole sequences through all T steps on its own: no device-wide barrier, no h exchange through L2.  The slice
// kernel above spends ~9 us per step on those (barrier + L2 round trip) for 33 MFLOP of work; here a step is the
// matvec (128 FFMA per sequence and thread) plus two __syncthreads.
// Same contract as lstm_seq_kernel (packed whh / xproj column order, groups); grid = groups x ceil(B / NS).
// ================================================================================================
constexpr int kSmallH = 128;
constexpr int kSmallThreads = 4 * kSmallH;          // one thread per gate column
constexpr int kSmallKReg = kSmallH / 2;             // k rows of W_hh kept in registers

template <int NS>
__global__ void __launch_bounds__(kSmallThreads, 1) lstm_seq_small_kernel(const LstmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = kSmallH, NCOL = 4 * kSmallH;
  float* Ws = reinterpret_cast<float*>(smem_raw);              // [H - kSmallKReg][NCOL]   k = 64..127
  float* hs = Ws + (size_t)(H - kSmallKReg) * NCOL;            // [NS][H]     h_{t-1}
  float* gs = hs + NS * H;                                     // [NS][NCOL]  gate pre-activations
  const int j = threadIdx.x;                                   // packed gate column: slice * 32 + gate * 8 + unit
  const int per_group = (p.B + NS - 1) / NS;
  const int group = blockIdx.x / per_group;
  const int b0 = (blockIdx.x - group * per_group) * NS;        // first sequence of this CTA
  const float* xproj = p.xproj + group * p.xp_goff;
  const float* whh = p.whh + group * p.whh_gstride;
  float* hseq = p.hseq + group * p.hs_goff;

  // resident weights: whh[(slice * H + k) * 32 + (j % 32)], slice = j / 32
  const float* wcol = whh + (size_t)(j >> 5) * H * kNC + (j & 31);
  float wreg[kSmallKReg];
#pragma unroll
  for (int k = 0; k < kSmallKReg; ++k) wreg[k] = __ldg(wcol + (size_t)k * kNC);
  for (int k = kSmallKReg; k < H; ++k) Ws[(size_t)(k - kSmallKReg) * NCOL + j] = __ldg(wcol + (size_t)k * kNC);

  // gate phase: thread (s, u) = (j / H, j % H) for j < NS * H owns cell state c of unit u of sequence s
  const int gs_s = j / H, gs_u = j - gs_s * H;
  const bool gate_thread = j < NS * H && b0 + gs_s < p.B;
  const int gcol = (gs_u >> 3) * kNC + (gs_u & 7);             // column of gate 0 of unit u; gate g at + 8 g
  float c_state = 0.f;

  float xg[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s)
    xg[s] = (b0 + s < p.B) ? __ldg(xproj + ((size_t)(b0 + s) * p.T) * (size_t)p.xp_stride + j) : 0.f;
  __syncthreads();

  for (int t = 0; t < p.T; ++t) {
    float acc[NS], acc2[NS];      // two partial sums per sequence: halves the dependent-FMA chain
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      acc[s] = xg[s];
      acc2[s] = 0.f;
    }
    if (t + 1 < p.T) {   // next step's input projection: independent of the recurrence, in flight during the matvec
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (b0 + s < p.B) xg[s] = __ldg(xproj + ((size_t)(b0 + s) * p.T + t + 1) * (size_t)p.xp_stride + j);
    }
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < kSmallKReg; k += 4) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float4 h4 = *reinterpret_cast<const float4*>(hs + s * H + k);     // broadcast
          acc[s] = fmaf(wreg[k], h4.x, acc[s]);
          acc2[s] = fmaf(wreg[k + 1], h4.y, acc2[s]);
          acc[s] = fmaf(wreg[k + 2], h4.z, acc[s]);
          acc2[s] = fmaf(wreg[k + 3], h4.w, acc2[s]);
        }
      }
#pragma unroll 4
      for (int k = kSmallKReg; k < H; k += 4) {
        const float w0 = Ws[(size_t)(k - kSmallKReg) * NCOL + j], w1 = Ws[(size_t)(k + 1 - kSmallKReg) * NCOL + j],
                    w2 = Ws[(size_t)(k + 2 - kSmallKReg) * NCOL + j], w3 = Ws[(size_t)(k + 3 - kSmallKReg) * NCOL + j];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float4 h4 = *reinterpret_cast<const float4*>(hs + s * H + k);
          acc[s] = fmaf(w0, h4.x, acc[s]);
          acc2[s] = fmaf(w1, h4.y, acc2[s]);
          acc[s] = fmaf(w2, h4.z, acc[s]);
          acc2[s] = fmaf(w3, h4.w, acc2[s]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) gs[s * NCOL + j] = acc[s] + acc2[s];
    __syncthreads();              // all pre-activations written; every thread is done reading h_{t-1}
    if (gate_thread) {
      // ex2.approx gates (|err| ~ 2e-7, as in the tcgen05 recurrence): this phase is a dependent chain on a quarter of
      // the warps while the rest wait at the barrier
      const float* g = gs + gs_s * NCOL + gcol;
      const float ig = fast_sigmoid(g[0]);
      const float fg = fast_sigmoid(g[kHU]);
      const float gg = fast_tanh(g[2 * kHU]);
      const float og = fast_sigmoid(g[3 * kHU]);
      c_state = fg * c_state + ig * gg;
      const float h = og * fast_tanh(c_state);
      hs[gs_s * H + gs_u] = h;
      hseq[(size_t)(b0 + gs_s) * p.hs_sb + (size_t)t * p.hs_st + gs_u] = h;
    }
    __syncthreads();              // h_t visible before the next matvec
  }
}

template <int NS>
static cudaError_t launch_lstm_small(const LstmParams& p, cudaStream_t s) {
  const size_t smem = ((size_t)(kSmallH - kSmallKReg) * 4 * kSmallH + (size_t)NS * kSmallH + (size_t)NS * 4 * kSmallH) *
                      sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(lstm_seq_small_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = p.ngroups * ((p.B + NS - 1) / NS);
  lstm_seq_small_kernel<NS><<<grid, kSmallThreads, smem, s>>>(p);
  return cudaGetLastError();
}

// ================================================================================================
// Tensor-core variant of the same recurrence: legacy mma.sync m16n8k8 TF32 with the 3xTF32 split.
//   * W_hh_hi (TF32) stays in REGISTERS for the whole sequence (A fragments: 2 m-tiles x KT k-tiles
//     x 4 = 128 registers per thread at H = 1024), W_hh_lo in shared memory in fragment order as
//     bf16 (|lo| <= 2^-11 |w|, so its 8-bit mantissa costs 2^-20 relative: still fp32 class), which
//     leaves room for a 4-deep cp.async ring on h (the step is L2-latency bound, not MMA bound);
//   * h_{t-1} is split hi/lo on the fly while loading the B fragments from the staged chunk;
//   * per step and warp: 3 x 2 x 8 x KT MMAs; accumulate in fp32 registers (only 3*KT accumulation
//     steps per warp, then an exact fp32 cross-warp reduction) -> fp32-class accuracy.
// tcgen05 is not used here: its M >= 64 tiles need a W_hh slice (hi + lo) of >= 512 KB per CTA, so it
// only pays with W_hh split across a cluster and h multicast (planned); mma.sync measured 278 TFLOP/s
// on this part (tools/microbench.cu), 5x the fp32 FMA rate this recurrence ran at.
// ================================================================================================
constexpr int kHS = 72;  // padded row stride of a staged h chunk (conflict-free B-fragment loads)
constexpr int kNST = 4;  // cp.async ring depth of the mma kernel

__device__ __forceinline__ unsigned tf32_rna(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kHB = 32;   // batch rows per half (the two halves of the 64-row tile ping-pong)
constexpr int kHS2 = 40;  // padded row stride of a staged half-batch h chunk (conflict-free B fragments)

// The 64-row batch tile is processed as two independent halves A and B that alternate inside a step:
// while the device-wide barrier of half A (step t) propagates, the CTA computes half B, so the barrier
// and first-fetch latency (~5 us of a 13.7 us step in the single-phase kernels) is hidden.
template <int KT>  // k-tiles (of 8) per warp: H = 64 * KT
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_seq_mma_kernel(const LstmParams p) {
  constexpr int H = 64 * KT;
  constexpr int NCH = (KT + 1) / 2;        // chunks of up to 16 k rows
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint2* wlo = reinterpret_cast<uint2*>(smem_raw);                           // [8 warps][2][KT][32 lanes] 4 x bf16
  float* stage = reinterpret_cast<float*>(wlo + kLstmWarps * 2 * KT * 32);   // [8 warps][kNST][16][kHS2] (aliased by red)
  float* cst = stage + kLstmWarps * kNST * kKC * kHS2;                       // [2 halves][32][8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int G = H / kHU;
  const int group = blockIdx.x / G;
  const int slice = blockIdx.x - group * G;
  const float* xproj = p.xproj + group * p.xp_goff;
  const float* whh = p.whh + group * p.whh_gstride + (size_t)slice * H * kNC;   // [H][32]
  float* hseq = p.hseq + group * p.hs_goff;
  float* hT = p.hT + (size_t)group * kRep * 2 * H * kBT;    // viewed as [half][parity][H][32]
  unsigned* sync = p.sync + 2 * group;                      // one counter per half
  const int kbase = warp * (H / kLstmWarps);
  const int nhalves = p.B > kHB ? 2 : 1;

  unsigned ahi[2][KT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      float lo[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = kbase + kt * 8 + tig + ((r >> 1) << 2);
        const int m = mt * 16 + gid + ((r & 1) << 3);
        const float w = __ldg(whh + (size_t)k * kNC + m);
        ahi[mt][kt][r] = tf32_rna(w);
        lo[r] = w - __uint_as_float(ahi[mt][kt][r]);
      }
      auto bf = [](float v) { unsigned u = __float_as_uint(v); u += 0x7FFFu + ((u >> 16) & 1u); return u >> 16; };
      wlo[((warp * 2 + mt) * KT + kt) * 32 + lane] = make_uint2(bf(lo[0]) | (bf(lo[1]) << 16), bf(lo[2]) | (bf(lo[3]) << 16));
    }
  for (int i = tid; i < 2 * kHB * kHU; i += kLstmThreads) cst[i] = 0.f;
  __syncthreads();

  float* my_stage = stage + warp * (kNST * kKC * kHS2);
  const int gb = tid >> 3, gj = tid & 7;   // gate phase: one (batch row, unit) pair per thread

  for (int t = 0; t < p.T; ++t) {
    for (int half = 0; half < nhalves; ++half) {
      const int bglob = half * kHB + gb;
      float xg[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        xg[g] = 0.f;
        if (bglob < p.B)
          xg[g] = __ldg(xproj + ((size_t)bglob * p.T + t) * (size_t)p.xp_stride + slice * kNC + g * kHU + gj);
      }
      float* hbase = hT + (size_t)half * 2 * H * kHB;      // [parity][H][32]

      if (t > 0) {
        if (tid == 0) {
          const unsigned target = (unsigned)t * (unsigned)G;
          while (ld_acquire_u32(sync + half) < target) {
          }
        }
        __syncthreads();
        const float* hprev = hbase + (size_t)((t - 1) & 1) * H * kHB;
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;

        auto issue = [&](int c) {   // always commits a group (possibly empty) so wait_group counts stay uniform
          if (c < NCH) {
            const int rows = min(kKC, KT * 8 - c * kKC);
            const float* g = hprev + (size_t)(kbase + c * kKC) * kHB;
            float* s = my_stage + (c % kNST) * (kKC * kHS2);
            for (int piece = lane; piece < rows * 8; piece += 32) {   // 16-byte pieces: 8 per 32-float row
              const int r = piece >> 3, q = piece & 7;
              cp_async16(s + r * kHS2 + q * 4, g + r * kHB + q * 4);
            }
          }
          cp_async_commit();
        };

#pragma unroll
        for (int c = 0; c < kNST - 1; ++c) issue(c);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          issue(c + kNST - 1);
          cp_async_wait<kNST - 1>();
          __syncwarp();
          const float* hs = my_stage + (c % kNST) * (kKC * kHS2);
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int kt = c * 2 + k2;
            if (kt < KT) {
              unsigned alo[2][4];
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                const uint2 v = wlo[((warp * 2 + mt) * KT + kt) * 32 + lane];
                alo[mt][0] = v.x << 16; alo[mt][1] = v.x & 0xFFFF0000u;
                alo[mt][2] = v.y << 16; alo[mt][3] = v.y & 0xFFFF0000u;
              }
#pragma unroll
              for (int np = 0; np < 2; ++np) {   // two batch tiles at a time: 12 MMAs on 4 independent accumulators
                unsigned bh[2][2], bl[2][2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const int nt = np * 2 + q;
                  const float h0 = hs[(k2 * 8 + tig) * kHS2 + nt * 8 + gid];
                  const float h1 = hs[(k2 * 8 + tig + 4) * kHS2 + nt * 8 + gid];
                  // hi = truncation to TF32 (bit mask, full-rate ALU); lo = exact remainder, the MMA ignores
                  // its low 13 bits: together 21 mantissa bits
                  bh[q][0] = __float_as_uint(h0) & 0xFFFFE000u; bh[q][1] = __float_as_uint(h1) & 0xFFFFE000u;
                  bl[q][0] = __float_as_uint(h0 - __uint_as_float(bh[q][0]));
                  bl[q][1] = __float_as_uint(h1 - __uint_as_float(bh[q][1]));
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][np * 2 + q], alo[mt], bh[q][0], bh[q][1]);
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][np * 2 + q], ahi[mt][kt], bl[q][0], bl[q][1]);
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][np * 2 + q], ahi[mt][kt], bh[q][0], bh[q][1]);
              }
            }
          }
          __syncwarp();
        }
        cp_async_wait<0>();

        // partial tile -> red[warp][b][col'] (b < 32; swizzle as in the FMA kernel; aliases this warp's stage)
        float* red = my_stage;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int col = mt * 16 + gid + ((r >> 1) << 3);
              const int b = nt * 8 + 2 * tig + (r & 1);
              red[b * kNC + ((((col >> 3) ^ (b & 3)) << 3) | (col & 7))] = acc[mt][nt][r];
            }
        __syncthreads();
      }

      float g4[4] = {xg[0], xg[1], xg[2], xg[3]};
      if (t > 0) {
#pragma unroll
        for (int w = 0; w < kLstmWarps; ++w) {
          const float* red = stage + w * (kNST * kKC * kHS2) + gb * kNC;
#pragma unroll
          for (int g = 0; g < 4; ++g) g4[g] += red[((g ^ (gb & 3)) << 3) + gj];
        }
      }
      const float ig = sigmoid_f(g4[0]);
      const float fg = sigmoid_f(g4[1]);
      const float gg = tanhf(g4[2]);
      const float og = sigmoid_f(g4[3]);
      const float c = fg * cst[half * kHB * kHU + tid] + ig * gg;
      const float h = og * tanhf(c);
      cst[half * kHB * kHU + tid] = c;
      const int u = slice * kHU + gj;
      hbase[(size_t)(t & 1) * H * kHB + (size_t)u * kHB + gb] = (bglob < p.B) ? h : 0.f;
      if (bglob < p.B) hseq[(size_t)bglob * p.hs_sb + (size_t)t * p.hs_st + u] = h;
      __syncthreads();
      if (tid == 0 && t + 1 < p.T) {
        __threadfence();
        red_release_add(sync + half, 1u);
      }
    }
  }
}

// 0: fp32 FMA kernel, 1: mma.sync 3xTF32 kernel, 2: tcgen05 cluster kernel (lstm_tc.cu) where it applies (H = 1024, one
// group, clusters fit the device), the FMA kernel elsewhere; 3: as 2, plus the sequence-parallel kernel
// (lstm_seq_small_kernel) for H = 128; 4 (default): as 3 with the fp16-pair / tagged-state tcgen05 kernel (lstm_f16.cu) at
// H = 1024; 5: as 4 with two independent half-batch chains per CTA (lstm_seq_f16_kernel<true>).
// -1 = not chosen yet: SE_LSTM_ENGINE in the environment (A/B runs), else the default.
static int g_lstm_engine = -1;
constexpr int kDefaultLstmEngine = 4;
constexpr int kMaxLstmEngine = 5;
static int lstm_engine() {
  if (g_lstm_engine < 0) {
    g_lstm_engine = kDefaultLstmEngine;
    if (const char* e = getenv("SE_LSTM_ENGINE")) {
      const int v = atoi(e);
      if (v >= 0 && v <= kMaxLstmEngine) g_lstm_engine = v;
    }
  }
  return g_lstm_engine;
}

int lstm_tc_supported();
void lstm_tc_set_profile(long long* dev_buf, int first_step, int nsteps);
int lstm_seq_tc_launch(const float* xproj, long long xp_stride, const float* whh, int B, int T, float* hseq,
                       long long hs_sb, long long hs_st, float* work, unsigned* sync, cudaStream_t s);
int lstm_f16_supported();
void lstm_f16_set_profile(long long* dev_buf, int first_step, int nsteps);
int lstm_seq_f16_launch(const float* xproj, long long xp_stride, const float* whh, int B, int T, float* hseq,
                        long long hs_sb, long long hs_st, float* work, int pingpong, cudaStream_t s);

template <int KT>
static cudaError_t launch_lstm_mma(const LstmParams& p, int G, cudaStream_t s) {
  const size_t smem = (size_t)kLstmWarps * 2 * KT * 32 * sizeof(uint2) +
                      ((size_t)kLstmWarps * kNST * kKC * kHS2 + kBT * kHU) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(lstm_seq_mma_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((const void*)lstm_seq_mma_kernel<KT>, dim3(G), dim3(kLstmThreads), args, smem, s);
}

}  // namespace se

using namespace se;

extern "C" long long se_lstm_seq_work_bytes(int B, int H) {
  (void)B;
  return 16ll * H * kBT * (long long)sizeof(float);   // tcgen05 engine: {h_hi, h_lo} x 2 parities x up to 4 replicas
}

extern "C" int se_lstm_seq_multi(const float* xproj, long long xproj_stride, long long xproj_group_off,
                                 const float* whh, long long whh_group_stride, int ngroups, int B, int T, int H,
                                 float* hseq, long long hseq_sb, long long hseq_st, long long hseq_group_off,
                                 float* work, unsigned* sync, se_stream_t stream) {
  SE_REQUIRE(xproj && whh && hseq && work && sync, "se_lstm_seq: null pointer");
  SE_REQUIRE(ngroups >= 1 && ngroups <= 8, "se_lstm_seq: ngroups=%d (1..8)", ngroups);
  SE_REQUIRE(B > 0 && B <= kBT, "se_lstm_seq: B=%d (1..%d per call)", B, kBT);
  SE_REQUIRE(T > 0 && H > 0 && H % (kLstmWarps * kKC) == 0, "se_lstm_seq: H=%d must be a multiple of %d", H,
             kLstmWarps * kKC);
  SE_REQUIRE(xproj_stride >= 4ll * H, "se_lstm_seq: xproj_stride=%lld < 4H", xproj_stride);
  SE_REQUIRE((((uintptr_t)whh) & 15) == 0 && (((uintptr_t)work) & 15) == 0 && (whh_group_stride & 3) == 0,
             "se_lstm_seq: unaligned buffers");
  const int G = ngroups * (H / kHU);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  SE_REQUIRE(G <= sms, "se_lstm_seq: H=%d x %d groups needs %d co-resident CTAs but the device has %d SMs", H, ngroups,
             G, sms);
  const size_t smem = ((size_t)H * kNC + kLstmWarps * 2 * kKC * kBT + kBT * kHU) * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(lstm_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("se_lstm_seq: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  e = cudaMemsetAsync(sync, 0, 16 * sizeof(unsigned), s);
  if (e != cudaSuccess) {
    set_error("se_lstm_seq: memset: %s", cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  LstmParams p{xproj, xproj_stride, whh, B, T, H, hseq, hseq_sb, hseq_st, work, sync, ngroups, xproj_group_off,
               whh_group_stride, hseq_group_off};
  void* args[] = {(void*)&p};
  const int engine = lstm_engine();
  if (engine >= 3 && H == kSmallH) {
    // sequences per CTA: as few as keep the grid within one wave (fewer sequences = shorter steps)
    const int ns = ngroups * B <= sms ? 1 : (ngroups * ((B + 1) / 2) <= sms ? 2 : 4);
    e = ns == 1 ? launch_lstm_small<1>(p, s) : (ns == 2 ? launch_lstm_small<2>(p, s) : launch_lstm_small<4>(p, s));
    if (e != cudaSuccess) {
      set_error("se_lstm_seq (sequence-parallel, H = 128): launch: %s", cudaGetErrorString(e));
      return SE_ERR_CUDA;
    }
    return SE_OK;
  }
  if (engine >= 4 && H == LT_H_PUBLIC && ngroups == 1 && lstm_f16_supported())
    return lstm_seq_f16_launch(xproj, xproj_stride, whh, B, T, hseq, hseq_sb, hseq_st, work, engine >= 5 ? 1 : 0, s);
  if (engine >= 2 && H == LT_H_PUBLIC && ngroups == 1 && lstm_tc_supported())
    return lstm_seq_tc_launch(xproj, xproj_stride, whh, B, T, hseq, hseq_sb, hseq_st, work, sync, s);
  if (engine == 1 && (H == 1024 || H == 512 || H == 128)) {
    e = H == 1024 ? launch_lstm_mma<16>(p, G, s) : (H == 512 ? launch_lstm_mma<8>(p, G, s) : launch_lstm_mma<2>(p, G, s));
  } else {
    e = cudaLaunchCooperativeKernel((const void*)lstm_seq_kernel, dim3(G), dim3(kLstmThreads), args, smem, s);
  }
  if (e != cudaSuccess) {
    set_error("se_lstm_seq: cooperative launch: %s", cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  return check_launch("se_lstm_seq");
}

extern "C" int se_lstm_seq(const float* xproj, long long xproj_stride, const float* whh, int B, int T, int H,
                           float* hseq, long long hseq_sb, long long hseq_st, float* work, unsigned* sync,
                           se_stream_t stream) {
  return se_lstm_seq_multi(xproj, xproj_stride, 0, whh, 0, 1, B, T, H, hseq, hseq_sb, hseq_st, 0, work, sync, stream);
}

extern "C" int se_debug_lstm_tc_profile(long long* dev_buf, int first_step, int nsteps) {
  SE_REQUIRE(dev_buf == nullptr || (first_step >= 1 && nsteps >= 1), "se_debug_lstm_tc_profile: bad step range");
  lstm_tc_set_profile(dev_buf, first_step, nsteps);
  lstm_f16_set_profile(dev_buf, first_step, nsteps);
  return SE_OK;
}

extern "C" int se_set_lstm_engine(int engine) {
  SE_REQUIRE(engine >= 0 && engine <= kMaxLstmEngine,
             "se_set_lstm_engine: 0 (fp32 FMA), 1 (mma.sync 3xTF32), 2 (tcgen05 3xTF32), 3 (2 + sequence-parallel H = 128), 4 (3 with "
             "the fp16-pair tcgen05 kernel) or 5 (4 with two half-batch chains per CTA)");
  g_lstm_engine = engine;
  return SE_OK;
}
