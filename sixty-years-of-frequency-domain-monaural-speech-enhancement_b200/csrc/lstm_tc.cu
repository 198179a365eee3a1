// LSTM recurrence on the 5th-generation tensor cores (tcgen05 + TMEM + TMA + clusters), H = 1024.
//
// Same contract as lstm_seq_kernel (lstm.cu): T dependent steps  g_t = xproj_t + h_{t-1} W_hh^T  in
// ONE launch, W_hh on chip for the whole sequence, h exchanged through L2, one device-wide counter
// barrier per step.  What changes is where the 2*64*4096*1024 flop of a step run and how much of
// h_{t-1} every SM has to pull out of L2:
//
//   * 128 CTAs = 32 clusters of 4.  Cluster c owns hidden units [32c, 32c+32) = 128 gate columns
//     (the M of the MMA); CTA rank q of the cluster owns the K slice [256q, 256q+256) of them.
//     So a CTA reads only a QUARTER of h_{t-1} per step (64 KB x {hi, lo} instead of 256 KB).
//   * the CTA's W_hh block [128 x 256] is split once into TF32 hi + lo: hi lives in TENSOR MEMORY as
//     the A operand (128 lanes x 256 columns, tcgen05.st at start-up), lo in shared memory as a
//     K-major SWIZZLE_128B operand (128 KB).  Nothing of W_hh moves after start-up.
//   * h_{t-1} is published already split (hi = rna_tf32(h), lo = rna_tf32(h - hi)) in batch-major
//     [64][1024] so a 2-D TMA box [64 batch x 32 k] is directly the K-major B operand (N = 64).
//   * 3xTF32:  D0 = W_hi h_hi            (32 accumulation steps -> truncation error of the tensor
//              D1 = W_hi h_lo + W_lo h_hi  core's fp32 accumulate stays ~1e-6, see gemm_tc.cu)
//     two TMEM accumulators of 64 columns; the small terms never meet the large one before fp32.
//   * K-split reduction over distributed shared memory: each CTA drains its [128 x 64] partial with
//     tcgen05.ld (lane = gate column) and stores the 32 columns of owner rank o into o's `red`
//     buffer (st.shared::cluster, 128-byte coalesced); after barrier.cluster the owner adds the 4
//     partials + xproj, runs the cell update for its 8 units x 64 batch rows (c in registers) and
//     publishes h_t (hseq, h_hi, h_lo), then arrives on the device-wide step counter.
//
// Warp roles (192 threads): 0-3 epilogue/gates (TMEM lane quarter = warp), 4 TMA producer + step
// barrier poller, 5 single-thread MMA issuer.  Every spin is bounded (trap after 4 s) so a protocol
// bug ends in a launch error instead of a hung device.
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

namespace se {

constexpr int LT_H = 1024;
constexpr int LT_CL = 4;                    // K split (CTAs that reduce over DSMEM); cluster = LT_CL * multicast width
constexpr int LT_KS = LT_H / LT_CL;         // 256 k per CTA
constexpr int LT_NB = 64;                   // batch rows (MMA N)
constexpr int LT_BK = 32;                   // k per TMA stage (128-byte rows)
constexpr int LT_KB = LT_KS / LT_BK;        // 8 stages' worth per step
constexpr int LT_STAGES = 4;
constexpr int LT_CTAS = LT_H / 8;           // 128
constexpr int LT_THREADS = 320;                 // 8 epilogue warps + TMA warp + MMA warp
constexpr int LT_TMA_WARP = 8, LT_MMA_WARP = 9;
constexpr int LT_WLO_BYTES = 128 * LT_KS * 4;          // 131072
constexpr int LT_BTILE = LT_NB * LT_BK * 4;            // 8192
constexpr int LT_STAGE_BYTES = 2 * LT_BTILE;           // h_hi + h_lo
constexpr int LT_RED_BYTES = LT_CL * LT_NB * 32 * 4;   // 32768
constexpr int LT_SMEM_BYTES = LT_WLO_BYTES + LT_STAGES * LT_STAGE_BYTES + LT_RED_BYTES + 1024 + 256;
constexpr int LT_REPLICAS = 1, LT_MAX_REPLICAS = 4;   // copies of the published state (SE_LSTM_TC_REPLICAS overrides)
constexpr unsigned LT_TMEM_COLS = 512;      // A: 0..255, D0: 256..319, D1: 320..383
constexpr unsigned long long LT_SPIN_NS = 4000000000ull;

struct LtParams {
  const float* xproj;
  long long xp_stride;
  const float* whh;  // [128 slices][1024 k][32]
  int B, T;
  float* hseq;
  long long hs_sb, hs_st;
  float* h_hi;       // [R replicas][2 parities][64][1024]
  float* h_lo;
  int writer_proxy_fence;   // 1: fence.proxy.async on the publishing side too (the reader always fences)
  int R;             // replicas of the published state: cluster c reads copy c % R (spreads the 32 readers of every
                     // line over R lines / L2 slices)
  unsigned* sync;
  // optional phase timestamps (se_debug_lstm_tc_profile): prof[(cta * prof_n + (t - prof_t0)) * LT_NEV + event]
  long long* prof;
  int prof_t0, prof_n;
};
constexpr int LT_NEV = 12;
// events: 0 producer starts polling, 1 step barrier observed, 2 last TMA issued, 3 first stage landed (MMA warp),
//         4 last stage landed, 5 accumulators complete (epilogue), 6 DSMEM partials sent, 7 cluster barrier passed,
//         8 cell update + stores issued, 10 proxy fence done, 11 CTA barrier passed, 9 release-arrival on the step counter issued
#define LT_STAMP(ev)                                                                                  \
  do {                                                                                                \
    if (p.prof && t >= p.prof_t0 && t < p.prof_t0 + p.prof_n)                                         \
      p.prof[((long long)blockIdx.x * p.prof_n + (t - p.prof_t0)) * LT_NEV + (ev)] = clock64();       \
  } while (0)

__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity, unsigned long long t0) {
  unsigned it = 0;
  while (!mbar_try(bar, parity)) {
    if (((++it) & 0xFFu) == 0 && gtimer_ns() - t0 > LT_SPIN_NS) __trap();
  }
}
__device__ __forceinline__ unsigned lt_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lt_red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(unsigned addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, uint64_t bdesc, unsigned idesc,
                                             unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(unsigned taddr, const unsigned (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// MC = multicast width: the cluster has 4*MC CTAs, rank = ug*4 + q.  The MC CTAs with the same K slice q (different
// unit groups ug) each fetch 1/MC of every h tile and multicast it to all of them, so L2 serves every byte of
// h_{t-1} 32/MC times per step instead of 32 (the v1 kernel was bound by exactly that: 7200 of 20000 cycles).
template <int MC>
__global__ void __launch_bounds__(LT_THREADS, 1)
    lstm_seq_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                       const LtParams p) {
  constexpr int RS = LT_NB / MC;                 // rows of a tile this CTA fetches
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* wlo = base;                                         // 8 k-blocks x [128 rows x 128 B], SW128
  unsigned char* stages = base + LT_WLO_BYTES;                       // LT_STAGES x {h_hi, h_lo} tiles [64 x 128 B]
  float* red = reinterpret_cast<float*>(stages + LT_STAGES * LT_STAGE_BYTES);   // [4 src][64 b][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(red) + LT_RED_BYTES);
  uint64_t* full = bars;                 // [LT_STAGES]
  uint64_t* empty = bars + LT_STAGES;    // [LT_STAGES]
  uint64_t* accfull = bars + 2 * LT_STAGES;
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(accfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned rank = cluster_rank();
  const unsigned q = rank & 3u;                    // K slice of this CTA / owner index of its 8 hidden units in the group
  const unsigned ug = rank >> 2;                   // unit group inside the cluster
  const int slice = blockIdx.x;                    // hidden units [8*slice, +8): what this CTA's gate phase owns
  const int grp = blockIdx.x >> 2;                 // unit group: hidden units [32*grp, +32) = the M of this CTA's MMAs
  unsigned short mc_mask = 0;
#pragma unroll
  for (int u = 0; u < MC; ++u) mc_mask |= (unsigned short)(1u << (u * 4 + q));
  const unsigned long long t0 = gtimer_ns();

  if (tid == 0) {
    for (int s = 0; s < LT_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MC);
    }
    mbar_init(accfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_hi);
    tma_prefetch_desc(&map_lo);
  }
  if (warp == LT_MMA_WARP) tmem_alloc(tmem_slot, LT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  const unsigned tmem_a = tmem_base, tmem_d0 = tmem_base + 256, tmem_d1 = tmem_base + 320;

  // ---- resident weights: row m = 32*o + l  <-  W_hh slice (4*grp + o), gate column l; k in this CTA's slice ----
  if (warp < 4) {
    const int m = warp * 32 + lane;
    const float* wsrc = p.whh + ((size_t)(grp * 4 + warp) * LT_H + (size_t)q * LT_KS) * 32 + lane;
    for (int k0 = 0; k0 < LT_KS; k0 += 32) {
      unsigned hi[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float w = __ldg(wsrc + (size_t)(k0 + i) * 32);
        float h, l;
        split_tf32_dev(w, h, l);
        hi[i] = __float_as_uint(h);
        // k-block k0/32, row m, 16-byte chunk (i/4) swizzled by (m & 7)
        unsigned char* dst = wlo + (k0 / 32) * (128 * 128) + m * 128 + ((((i >> 2) ^ (m & 7))) << 4) + ((i & 3) << 2);
        *reinterpret_cast<float*>(dst) = l;
      }
      tmem_st_32x32(tmem_a + ((unsigned)(warp * 32) << 16) + (unsigned)k0, hi);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  fence_proxy_async();          // generic-proxy smem writes (W_lo) -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();             // barriers initialised and every CTA running before anyone touches remote smem
  cluster_wait();

  if (warp == LT_TMA_WARP) {
    // ===================== TMA producer + step-barrier poller =====================
    int stage = 0;
    unsigned phase = 0;
    for (int t = 1; t < p.T; ++t) {
      if (elect_one()) {
        const unsigned target = (unsigned)t * (unsigned)LT_CTAS;
        unsigned it = 0;
        LT_STAMP(0);
        while (lt_ld_acquire(p.sync) < target) {
          if (((++it) & 0xFFu) == 0 && gtimer_ns() - t0 > LT_SPIN_NS) __trap();
        }
        LT_STAMP(1);
        fence_proxy_async();    // h_{t-1} was written with generic stores by other SMs; TMA reads it
        const int row0 = (((grp % p.R) * 2 + ((t - 1) & 1)) * LT_NB) + (int)ug * RS;
        for (int kb = 0; kb < LT_KB; ++kb) {
          mbar_wait_bounded(&empty[stage], phase ^ 1, t0);     // all MC consumers of this slot have drained it
          unsigned char* st = stages + stage * LT_STAGE_BYTES + (int)ug * RS * 128;
          mbar_expect_tx(&full[stage], LT_STAGE_BYTES);        // my share + the MC-1 shares multicast by my peers
          const int kc = (int)q * LT_KS + kb * LT_BK;
          if (MC == 1) {
            tma_load_2d(&map_hi, &full[stage], st, kc, row0);
            tma_load_2d(&map_lo, &full[stage], st + LT_BTILE, kc, row0);
          } else {
            tma_load_2d_mc(&map_hi, &full[stage], st, kc, row0, mc_mask);
            tma_load_2d_mc(&map_lo, &full[stage], st + LT_BTILE, kc, row0, mc_mask);
          }
          if (++stage == LT_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        LT_STAMP(2);
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    }
  } else if (warp == LT_MMA_WARP) {
    // ===================== MMA issuer =====================
    // h_hi and h_lo tiles of a stage are adjacent 64-row K-major tiles = ONE 128-row B operand: a single N = 128
    // MMA computes [W_hi h_hi | W_hi h_lo] (D0 | D1 are adjacent TMEM columns); W_lo h_hi accumulates into D1.
    // tcgen05.mma issue costs ~45-50 cycles whatever N is (tools/umma_bench.cu), so 2 MMAs per k-step beat 3.
    constexpr unsigned idesc_wide = make_idesc_tf32(128, 2 * LT_NB);
    constexpr unsigned idesc = make_idesc_tf32(128, LT_NB);
    int stage = 0;
    unsigned phase = 0;
    for (int t = 1; t < p.T; ++t) {
      if (elect_one()) {
        for (int kb = 0; kb < LT_KB; ++kb) {
          mbar_wait_bounded(&full[stage], phase, t0);
          if (kb == 0) LT_STAMP(3);
          if (kb == LT_KB - 1) LT_STAMP(4);
          tc_fence_after();
          unsigned char* st = stages + stage * LT_STAGE_BYTES;
          const uint64_t d_b = make_smem_desc(st);              // rows 0..63 h_hi, rows 64..127 h_lo
          const uint64_t d_alo = make_smem_desc(wlo + kb * (128 * 128));
#pragma unroll
          for (int k = 0; k < LT_BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
            const unsigned a_t = tmem_a + (unsigned)(kb * LT_BK + k * 8);
            umma_tf32_ts(tmem_d0, a_t, d_b + adv, idesc_wide, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_d1, d_alo + adv, d_b + adv, idesc, 1u);
          }
          if (MC == 1)
            umma_commit(&empty[stage]);
          else
            umma_commit_mc(&empty[stage], mc_mask);   // frees the slot in every CTA that multicasts into mine
          if (++stage == LT_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(accfull);
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    }
  } else {
    // ===================== epilogue: K-split reduction over DSMEM, gates, publish =====================
    const int quarter = warp & 3;        // TMEM lanes 32*quarter .. +31 = gate columns owned by rank ug*4 + quarter
    const int half = warp >> 2;          // batch rows [32*half, +32)
    float cstate[2] = {0.f, 0.f};
    const unsigned red_remote = map_to_rank(smem_u32(red), ug * 4u + (unsigned)quarter);
    for (int t = 0; t < p.T; ++t) {
      float xg[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pr = tid + 256 * i, b = pr >> 3, j = pr & 7;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          xg[i][g] = 0.f;
          if (b < p.B) xg[i][g] = __ldg(p.xproj + ((size_t)b * p.T + t) * (size_t)p.xp_stride + slice * 32 + g * 8 + j);
        }
      }
      if (t > 0) {
        mbar_wait_bounded(accfull, (unsigned)((t - 1) & 1), t0);
        if (tid == 0) LT_STAMP(5);
        tc_fence_after();
        float d0[32], d1[32];
        tmem_ld_32x32(tmem_d0 + ((unsigned)(quarter * 32) << 16) + (unsigned)(half * 32), d0);
        tmem_ld_32x32(tmem_d1 + ((unsigned)(quarter * 32) << 16) + (unsigned)(half * 32), d1);
        // lane = gate column (g = lane >> 3, unit j = lane & 7) of the owner; register i = batch row 32*half + i
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int b = half * 32 + i;
          const unsigned off = (q * 2048u + (unsigned)b * 32u + (unsigned)((((lane >> 3) ^ (b & 3)) << 3) | (lane & 7))) * 4u;
          st_cluster_f32(red_remote + off, d0[i] + d1[i]);
        }
        tc_fence_before();
        if (tid == 0) LT_STAMP(6);
        cluster_arrive();
        cluster_wait();
        if (tid == 0) LT_STAMP(7);
      }
      const int par = t & 1;
      float hv[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pr = tid + 256 * i, b = pr >> 3, j = pr & 7;
        float g4[4] = {xg[i][0], xg[i][1], xg[i][2], xg[i][3]};
        if (t > 0) {
#pragma unroll
          for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int g = 0; g < 4; ++g) g4[g] += red[s * 2048 + b * 32 + (((g ^ (b & 3)) << 3) | j)];
        }
        const float ig = fast_sigmoid(g4[0]);
        const float fg = fast_sigmoid(g4[1]);
        const float gg = fast_tanh(g4[2]);
        const float og = fast_sigmoid(g4[3]);
        const float c = fg * cstate[i] + ig * gg;
        float h = og * fast_tanh(c);
        cstate[i] = c;
        const int u = slice * 8 + j;
        if (b >= p.B) h = 0.f;
        float hh, hl;
        split_tf32_dev(h, hh, hl);
        const size_t o = ((size_t)par * LT_NB + b) * LT_H + u;
        for (int r = 0; r < p.R; ++r) {
          p.h_hi[o + (size_t)r * (2 * LT_NB * LT_H)] = hh;
          p.h_lo[o + (size_t)r * (2 * LT_NB * LT_H)] = hl;
        }
        hv[i] = h;
      }
      if (tid == 0) LT_STAMP(8);
      if (t + 1 < p.T) {
        if (p.writer_proxy_fence) fence_proxy_async();   // generic-proxy stores -> async-proxy (TMA) readers
        if (tid == 0) LT_STAMP(10);
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        if (tid == 0) LT_STAMP(11);
        if (tid == 0) lt_red_release(p.sync, 1u);        // release is cumulative over the CTA's stores (bar.sync above)
      }
      if (tid == 0) LT_STAMP(9);
      // the sequence output is not on the step's critical path: store it after the arrival
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pr = tid + 256 * i, b = pr >> 3, j = pr & 7;
        if (b < p.B) p.hseq[(size_t)b * p.hs_sb + (size_t)t * p.hs_st + slice * 8 + j] = hv[i];
      }
    }
  }
  tc_fence_before();
  cluster_arrive();             // no CTA leaves while peers may still signal its barriers / write its smem
  cluster_wait();
  __syncthreads();
  if (warp == LT_MMA_WARP) tmem_dealloc(tmem_base, LT_TMEM_COLS);
}

static int make_h_map(CUtensorMap* map, const float* ptr, int box_rows, int replicas) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SE_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)LT_H, (cuuint64_t)(2 * LT_NB * replicas)};
  cuuint64_t strides[1] = {(cuuint64_t)LT_H * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)LT_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("se_lstm_seq (tcgen05): cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

typedef void (*LtKernel)(const CUtensorMap, const CUtensorMap, const LtParams);
static LtKernel lt_kernel(int mc) {
  return mc == 4 ? lstm_seq_tc_kernel<4> : (mc == 2 ? lstm_seq_tc_kernel<2> : lstm_seq_tc_kernel<1>);
}

// can LT_CTAS / (4*mc) clusters of 4*mc CTAs of this kernel be co-resident on the current device?
static bool lt_fits(int mc) {
  LtKernel k = lt_kernel(mc);
  const bool verbose = getenv("SE_LSTM_TC_VERBOSE") != nullptr;
  cudaError_t e1 = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM_BYTES);
  cudaError_t e2 = 4 * mc > 8 ? cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)
                              : cudaSuccess;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    if (verbose) fprintf(stderr, "lstm_tc mc=%d: attributes: %s / %s\n", mc, cudaGetErrorString(e1), cudaGetErrorString(e2));
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LT_CTAS);
  cfg.blockDim = dim3(LT_THREADS);
  cfg.dynamicSmemBytes = LT_SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4 * mc;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nclusters = 0;
  cudaError_t e3 = cudaOccupancyMaxActiveClusters(&nclusters, (const void*)k, &cfg);
  if (verbose)
    fprintf(stderr, "lstm_tc mc=%d: cudaOccupancyMaxActiveClusters -> %d clusters of %d (need %d): %s\n", mc, nclusters,
            4 * mc, LT_CTAS / (4 * mc), cudaGetErrorString(e3));
  if (e3 != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return nclusters >= LT_CTAS / (4 * mc);
}

// multicast width to run with: the widest of {4, 2, 1} whose clusters fit the device (0 = engine unavailable);
// SE_LSTM_TC_MC=1|2|4 in the environment pins it (A/B measurements).
static int lt_pick_mc() {
  static int cached = -1;
  if (cached >= 0) return cached;
  cached = 0;
  int want = 0;
  if (const char* e = getenv("SE_LSTM_TC_MC")) want = atoi(e);
  const int order[3] = {4, 2, 1};
  for (int i = 0; i < 3; ++i) {
    if (want && order[i] != want) continue;
    if (lt_fits(order[i])) {
      cached = order[i];
      break;
    }
  }
  return cached;
}

int lstm_tc_supported() { return lt_pick_mc() > 0; }

static long long* g_prof = nullptr;
static int g_prof_t0 = 0, g_prof_n = 0;
void lstm_tc_set_profile(long long* dev_buf, int first_step, int nsteps) {
  g_prof = dev_buf;
  g_prof_t0 = first_step;
  g_prof_n = nsteps;
}

// work: [2 arrays][R <= 4 replicas][2 parities][64][1024] fp32; sync[0] zeroed by the caller on `s`
int lstm_seq_tc_launch(const float* xproj, long long xp_stride, const float* whh, int B, int T, float* hseq,
                       long long hs_sb, long long hs_st, float* work, unsigned* sync, cudaStream_t s) {
  const int mc = lt_pick_mc();
  if (mc <= 0) {
    set_error("se_lstm_seq (tcgen05): clusters of this kernel do not fit the device");
    return SE_ERR_CUDA;
  }
  static int reps = -1;
  if (reps < 0) {
    reps = LT_REPLICAS;
    if (const char* e = getenv("SE_LSTM_TC_REPLICAS")) reps = atoi(e);
    if (reps < 1 || reps > LT_MAX_REPLICAS) reps = LT_REPLICAS;
  }
  CUtensorMap map_hi, map_lo;
  float* h_hi = work;
  float* h_lo = work + (size_t)reps * 2 * LT_NB * LT_H;
  int rc = make_h_map(&map_hi, h_hi, LT_NB / mc, reps);
  if (rc != SE_OK) return rc;
  rc = make_h_map(&map_lo, h_lo, LT_NB / mc, reps);
  if (rc != SE_OK) return rc;
  // No proxy fence on the publishing side by default: the reader's fence.proxy.async, executed after its acquire of
  // the step counter, already sits between the generic-proxy stores and its own async-proxy (TMA) reads in causality
  // order, and TMA reads go to L2, where the released stores are.  Measured 7.54 -> 6.92 us/step
  // (profiles/lstm_tc_phases_nowriterfence_r01.json); SE_LSTM_TC_WRITER_FENCE=1 restores the second fence.
  static int wfence = -1;
  if (wfence < 0) {
    wfence = 0;
    if (const char* e = getenv("SE_LSTM_TC_WRITER_FENCE")) wfence = atoi(e) ? 1 : 0;
  }
  LtParams p{xproj, xp_stride, whh, B, T, hseq, hs_sb, hs_st, h_hi, h_lo, wfence, reps, sync, g_prof, g_prof_t0, g_prof_n};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LT_CTAS);
  cfg.blockDim = dim3(LT_THREADS);
  cfg.dynamicSmemBytes = LT_SMEM_BYTES;
  cfg.stream = s;
  // SE_LSTM_TC_COOP=0 drops the cooperative attribute: Nsight Compute cannot replay a cooperative CLUSTER launch
  // (LaunchFailed under ncu), so profiling runs rely on the idle device + the bounded spins instead.
  static int coop = -1;
  if (coop < 0) {
    coop = 1;
    if (profiler_attached()) coop = 0;
    if (const char* e = getenv("SE_LSTM_TC_COOP")) coop = atoi(e) ? 1 : 0;
  }
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4 * mc;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;   // all 128 CTAs co-resident or the launch fails (never a deadlock)
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = coop ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lt_kernel(mc), map_hi, map_lo, p);
  if (e != cudaSuccess && coop) {
    // tools that replay launches (Nsight Compute) reject a cooperative cluster launch: retry without the attribute;
    // the bounded spins turn a co-residency failure into a launch error instead of a hang
    cudaGetLastError();
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, lt_kernel(mc), map_hi, map_lo, p);
    if (e == cudaSuccess) coop = 0;
  }
  if (e != cudaSuccess) {
    set_error("se_lstm_seq (tcgen05, multicast %d): cooperative cluster launch: %s", mc, cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  return check_launch("se_lstm_seq (tcgen05)");
}

}  // namespace se
