// Utterance-level normalisations of the TCM family (CTSNet, TaylorSENet, G2Net) + the CTSNet two-stage glue.
//
//   nn.InstanceNorm{1,2}d(affine=True)  (CTSNet/Step1_network.py:48,164): statistics per (clip, channel) over the
//       whole utterance (T or T x F) -- also in eval(), the reference never tracks running statistics -- so a tile
//       cannot be normalised before the clip has been seen: two passes.
//         se_chan_stats   pass 1: mean / rstd per (b, c) of pre(x)        (fp64 combination, deterministic)
//         se_chan_norm    pass 2: post( (pre(x) - mean) * rstd * gamma + beta )
//   CumulativeLayerNorm{1,2}d  (CTSNet_new/Step1_network.py:213-286): statistics over (C[,F]) of all frames <= t.
//         se_cum_stats    per-frame sums -> prefix scan over T -> mean / rstd per (b, t); se_chan_norm applies them.
//   pre  = what the reference computes between the convolution and the norm: the Gate_Conv product
//          a * sigmoid(b) (Step1_network.py:150-151), a per-channel PReLU (:163,170,178), or both (the TCM output
//          path: left * sigmoid(right) -> PReLU -> norm, :188-189);
//   post = per-channel PReLU (encoder / decoder blocks, :49) or ShareSepConv (:195-209): ONE causal FIR shared by
//          all channels of a branch, applied to the normalised signal with zero history.
// All tensors are channels-last [B, rows, C]; rows = T * F (F = 1 inside the TCMs).  HBM-bound.
#include "tc_common.cuh"

namespace se {

constexpr int NT = 256;

__device__ __forceinline__ float prelu_f(float x, float a) { return x >= 0.f ? x : a * x; }

// value of pre(x) at (row r, channel c).  xr = x + r * Cin
__device__ __forceinline__ float pre_value(const float* __restrict__ xr, int c, int Cin, int C, int pre,
                                           const float* __restrict__ slope) {
  switch (pre) {
    case SE_NORM_PRE_GLU: return __ldg(xr + c) * sigmoid_f(__ldg(xr + C + c));
    case SE_NORM_PRE_PRELU: return prelu_f(__ldg(xr + (c % Cin)), __ldg(slope + c));
    case SE_NORM_PRE_GLU_PRELU: return prelu_f(__ldg(xr + c) * sigmoid_f(__ldg(xr + C + c)), __ldg(slope + c));
    default: return __ldg(xr + c);
  }
}

// ---- pass 1, instance statistics -----------------------------------------------------------------------------
// grid (chunks, B).  Thread tid owns channel c = tid % CL (CL = C rounded so that NT % CL == 0) and row lane tid / CL.
// ws: unsigned ticket[1024] (zero on entry, left zero on exit), then double partial[B][chunks][C][2].
__global__ void __launch_bounds__(NT) chan_stats_kernel(const float* __restrict__ x, long long rows, int Cin, int C,
                                                       int pre, const float* __restrict__ slope, float eps,
                                                       float* __restrict__ mean, float* __restrict__ rstd,
                                                       double* __restrict__ partial, unsigned* __restrict__ ticket) {
  __shared__ double sh[NT][2];
  __shared__ bool last;
  const int tid = threadIdx.x, b = blockIdx.y, chunks = gridDim.x;
  const int lanes = NT / C, c = tid % C, lane = tid / C;
  const long long per = (rows + chunks - 1) / chunks;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(rows, r0 + per);
  const float* xb = x + (long long)b * rows * Cin;
  double s = 0.0, ss = 0.0;
  if (lane < lanes) {
    float fs = 0.f, fss = 0.f;
    int n = 0;
    for (long long r = r0 + lane; r < r1; r += lanes) {
      const float v = pre_value(xb + r * Cin, c, Cin, C, pre, slope);
      fs += v;
      fss = fmaf(v, v, fss);
      if (++n == 64) {          // short fp32 runs, combined in fp64
        s += fs;
        ss += fss;
        fs = fss = 0.f;
        n = 0;
      }
    }
    s += fs;
    ss += fss;
  }
  sh[tid][0] = s;
  sh[tid][1] = ss;
  __syncthreads();
  if (tid < C) {
    for (int l = 1; l < lanes; ++l) {
      s += sh[tid + l * C][0];
      ss += sh[tid + l * C][1];
    }
    double* p = partial + (((long long)b * chunks + blockIdx.x) * C + tid) * 2;
    p[0] = s;
    p[1] = ss;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(ticket + b, 1u);
    last = prev == (unsigned)chunks - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (tid < C) {
    double ts = 0.0, tss = 0.0;
    for (int k = 0; k < chunks; ++k) {
      const double* p = partial + (((long long)b * chunks + k) * C + tid) * 2;
      ts += p[0];
      tss += p[1];
    }
    const double m = ts / (double)rows;
    double var = tss / (double)rows - m * m;
    if (var < 0.0) var = 0.0;
    mean[(long long)b * C + tid] = (float)m;
    rstd[(long long)b * C + tid] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (tid == 0) ticket[b] = 0u;
}


// four consecutive channels of pre(x) at once (C, Cin multiples of 4; 16-byte loads)
__device__ __forceinline__ void pre_value4(const float* __restrict__ xr, int c, int Cin, int C, int pre,
                                           const float* __restrict__ slope, float (&v)[4]) {
  float4 a, g = make_float4(0.f, 0.f, 0.f, 0.f), sl = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool glu = pre == SE_NORM_PRE_GLU || pre == SE_NORM_PRE_GLU_PRELU;
  const bool pr = pre == SE_NORM_PRE_PRELU || pre == SE_NORM_PRE_GLU_PRELU;
  a = __ldg(reinterpret_cast<const float4*>(xr + (glu ? c : c % Cin)));
  if (glu) g = __ldg(reinterpret_cast<const float4*>(xr + C + c));
  if (pr) sl = __ldg(reinterpret_cast<const float4*>(slope + c));
  const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w}, sv[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float t = av[e];
    if (glu) t *= sigmoid_f(gv[e]);
    if (pr) t = prelu_f(t, sv[e]);
    v[e] = t;
  }
}

// pass 1, instance statistics, 4 channels per thread: thread tid owns channels 4*(tid % C4) .. +3 and row lane tid / C4
__global__ void __launch_bounds__(NT) chan_stats_vec_kernel(const float* __restrict__ x, long long rows, int Cin, int C,
                                                           int pre, const float* __restrict__ slope, float eps,
                                                           float* __restrict__ mean, float* __restrict__ rstd,
                                                           double* __restrict__ partial, unsigned* __restrict__ ticket) {
  __shared__ double sh[NT][8];
  __shared__ bool last;
  const int tid = threadIdx.x, b = blockIdx.y, chunks = gridDim.x;
  const int C4 = C >> 2, lanes = NT / C4, cq = tid % C4, lane = tid / C4;
  const long long per = (rows + chunks - 1) / chunks;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(rows, r0 + per);
  const float* xb = x + (long long)b * rows * Cin;
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  {
    float fs[4] = {0, 0, 0, 0}, fss[4] = {0, 0, 0, 0};
    int n = 0;
    for (long long r = r0 + lane; r < r1; r += lanes) {
      float v[4];
      pre_value4(xb + r * Cin, 4 * cq, Cin, C, pre, slope, v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        fs[e] += v[e];
        fss[e] = fmaf(v[e], v[e], fss[e]);
      }
      if (++n == 64) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[e] += fs[e];
          ss[e] += fss[e];
          fs[e] = fss[e] = 0.f;
        }
        n = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[e] += fs[e];
      ss[e] += fss[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sh[tid][e] = s[e];
    sh[tid][4 + e] = ss[e];
  }
  __syncthreads();
  if (tid < C) {                         // channel tid: quad tid / 4, component tid % 4, summed over the row lanes
    const int q = tid >> 2, e = tid & 3;
    double ts = 0.0, tss = 0.0;
    for (int l = 0; l < lanes; ++l) {
      ts += sh[q + l * C4][e];
      tss += sh[q + l * C4][4 + e];
    }
    double* p = partial + (((long long)b * chunks + blockIdx.x) * C + tid) * 2;
    p[0] = ts;
    p[1] = tss;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(ticket + b, 1u);
    last = prev == (unsigned)chunks - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (tid < C) {
    double ts = 0.0, tss = 0.0;
    for (int k = 0; k < chunks; ++k) {
      const double* p = partial + (((long long)b * chunks + k) * C + tid) * 2;
      ts += p[0];
      tss += p[1];
    }
    const double m = ts / (double)rows;
    double var = tss / (double)rows - m * m;
    if (var < 0.0) var = 0.0;
    mean[(long long)b * C + tid] = (float)m;
    rstd[(long long)b * C + tid] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (tid == 0) ticket[b] = 0u;
}

// ---- pass 1, cumulative statistics -----------------------------------------------------------------------------
// step sums: sum and sum of squares of pre(x) over the F x C elements of every frame.
// one WARP per (frame, channel group): group g = channels [g*C/G, (g+1)*C/G) (branches that share a tensor keep their
// own statistics).  Inside the TCMs a frame is only 64 values, so a block per frame would be 7/8 idle.
__global__ void __launch_bounds__(NT) cum_step_kernel(const float* __restrict__ x, long long frames, int F, int Cin, int C,
                                                     int G, int pre, const float* __restrict__ slope,
                                                     double* __restrict__ step) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (w >= frames * G) return;
  const long long bt = w / G;
  const int g = (int)(w - bt * G), cg = C / G;
  const float* xb = x + bt * F * Cin;
  const int n = F * cg;
  float fs = 0.f, fss = 0.f;
  double s = 0.0, ss = 0.0;
  if (((cg | Cin) & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)slope)) & 15) == 0) {
    int cnt = 0;
    for (int i = 4 * lane; i < n; i += 128) {
      const int f = i / cg, c = g * cg + (i - f * cg);
      float v[4];
      pre_value4(xb + (long long)f * Cin, c, Cin, C, pre, slope, v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        fs += v[e];
        fss = fmaf(v[e], v[e], fss);
      }
      if (++cnt == 16) {
        s += fs, ss += fss, fs = fss = 0.f, cnt = 0;
      }
    }
  } else {
    int cnt = 0;
    for (int i = lane; i < n; i += 32) {
      const int f = i / cg, c = g * cg + (i - f * cg);
      const float v = pre_value(xb + (long long)f * Cin, c, Cin, C, pre, slope);
      fs += v;
      fss = fmaf(v, v, fss);
      if (++cnt == 64) {
        s += fs, ss += fss, fs = fss = 0.f, cnt = 0;
      }
    }
  }
  s += fs;
  ss += fss;
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) {
    step[w * 2] = s;
    step[w * 2 + 1] = ss;
  }
}
// prefix over T (one warp per clip; T is a few hundred): mean_t, rstd_t of everything up to and including frame t
__global__ void __launch_bounds__(32) cum_scan_kernel(const double* __restrict__ step, int T, int G, int per_frame,
                                                     float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  const int b = blockIdx.x, g = blockIdx.y, lane = threadIdx.x;
  double cs = 0.0, css = 0.0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    double s = t < T ? step[(((long long)b * T + t) * G + g) * 2] : 0.0;
    double ss = t < T ? step[(((long long)b * T + t) * G + g) * 2 + 1] : 0.0;
    for (int o = 1; o < 32; o <<= 1) {
      const double a = __shfl_up_sync(0xffffffffu, s, o), a2 = __shfl_up_sync(0xffffffffu, ss, o);
      if (lane >= o) {
        s += a;
        ss += a2;
      }
    }
    s += cs;
    ss += css;
    if (t < T) {
      const double cnt = (double)per_frame * (double)(t + 1);
      const double m = s / cnt;
      double var = ss / cnt - m * m;      // == (cum_pow - 2 mean cum_sum) / cnt + mean^2  (Step1_network.py:250)
      if (var < 0.0) var = 0.0;
      mean[((long long)b * T + t) * G + g] = (float)m;
      rstd[((long long)b * T + t) * G + g] = (float)(1.0 / sqrt(var + (double)eps));
    }
    cs = __shfl_sync(0xffffffffu, s, 31);
    css = __shfl_sync(0xffffffffu, ss, 31);
  }
}

// ---- pass 2 -----------------------------------------------------------------------------------------------------
struct NormParams {
  const float* x;
  int B;
  long long rows;
  int Cin, C, pre;
  const float* pre_slope;
  const float *mean, *rstd;
  int stat_mode, rows_per_t, stat_groups;
  const float *gamma, *beta;
  int post;
  const float* post_slope;
  const float* fir_w;
  int fir_k, fir_groups;
  float *out, *out_hi, *out_lo;
};

__global__ void __launch_bounds__(NT) chan_norm_kernel(const NormParams p) {
  const long long n = (long long)p.B * p.rows * p.C;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % p.C);
    const long long br = i / p.C;
    const int b = (int)(br / p.rows);
    const long long r = br - (long long)b * p.rows;
    const float g = __ldg(p.gamma + c), be = __ldg(p.beta + c);
    float y;
    if (p.post != SE_NORM_POST_FIR) {
      const long long si = p.stat_mode == SE_NORM_STAT_INSTANCE
                               ? (long long)b * p.C + c
                               : ((long long)b * (p.rows / p.rows_per_t) + r / p.rows_per_t) * p.stat_groups +
                                     c / (p.C / p.stat_groups);
      const float v = pre_value(p.x + br * p.Cin, c, p.Cin, p.C, p.pre, p.pre_slope);
      y = (v - __ldg(p.mean + si)) * __ldg(p.rstd + si) * g + be;
      if (p.post == SE_NORM_POST_PRELU) y = prelu_f(y, __ldg(p.post_slope + c));
    } else {
      // ShareSepConv on the normalised signal: y[t] = sum_k w[k] z[t - (K-1) + k], z = 0 before the clip starts
      const float* w = p.fir_w + (long long)(c / (p.C / p.fir_groups)) * p.fir_k;
      float acc = 0.f;
      for (int k = 0; k < p.fir_k; ++k) {
        const long long rr = r - (p.fir_k - 1) + k;
        if (rr < 0) continue;
        const long long si = p.stat_mode == SE_NORM_STAT_INSTANCE
                                 ? (long long)b * p.C + c
                                 : ((long long)b * p.rows + rr) * p.stat_groups + c / (p.C / p.stat_groups);
        const float v = pre_value(p.x + ((long long)b * p.rows + rr) * p.Cin, c, p.Cin, p.C, p.pre, p.pre_slope);
        const float z = (v - __ldg(p.mean + si)) * __ldg(p.rstd + si) * g + be;
        acc = fmaf(__ldg(w + k), z, acc);
      }
      y = acc;
    }
    if (p.out) p.out[i] = y;
    if (p.out_hi) split_tf32_dev(y, p.out_hi[i], p.out_lo[i]);
  }
}


// pass 2 with the ShareSepConv FIR (CTSNet TCM branches, Step1_network.py:127-157): y[t] = sum_k w[k] z[t - (K-1) + k] on the
// NORMALISED signal z.  The scalar kernel above re-normalises every tap of every output from global memory (K up to 63:
// 225 us per call on a 13 MB slab, 17 % of CTSNet).  Here a CTA owns FIR_TT output rows of one clip: the normalised rows
// [t0 - (K-1), t0 + FIR_TT) are computed ONCE into shared memory (4 channels per thread, 16-byte loads), then every thread
// accumulates FIR_TT * C / 256 outputs of one channel from shared memory.
constexpr int FIR_TT = 32;
__global__ void __launch_bounds__(NT) chan_norm_fir_kernel(const NormParams p) {
  extern __shared__ __align__(16) float fir_smem[];
  const int C = p.C, K = p.fir_k, C4 = C >> 2;
  float* z = fir_smem;                         // [FIR_TT + K - 1][C]
  float* wsm = z + (size_t)(FIR_TT + K - 1) * C;   // [fir_groups][K]
  const int b = blockIdx.y, tid = threadIdx.x;
  const long long t0 = (long long)blockIdx.x * FIR_TT;
  const int nz = FIR_TT + K - 1;
  for (int i = tid; i < p.fir_groups * K; i += NT) wsm[i] = __ldg(p.fir_w + i);
  for (int i = tid; i < nz * C4; i += NT) {
    const int j = i / C4, c = 4 * (i - j * C4);
    const long long r = t0 - (K - 1) + j;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0 && r < p.rows) {
      float v[4];
      pre_value4(p.x + ((long long)b * p.rows + r) * p.Cin, c, p.Cin, C, p.pre, p.pre_slope, v);
      float m[4], rs[4];
      if (p.stat_mode == SE_NORM_STAT_INSTANCE) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mean + (long long)b * C + c));
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.rstd + (long long)b * C + c));
        m[0] = m4.x, m[1] = m4.y, m[2] = m4.z, m[3] = m4.w;
        rs[0] = r4.x, rs[1] = r4.y, rs[2] = r4.z, rs[3] = r4.w;
      } else {                                 // cumulative statistics, one row per frame (rows_per_t == 1)
        const long long si = ((long long)b * p.rows + r) * p.stat_groups + c / (C / p.stat_groups);
        const float mm = __ldg(p.mean + si), rr = __ldg(p.rstd + si);
#pragma unroll
        for (int e = 0; e < 4; ++e) m[e] = mm, rs[e] = rr;
      }
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + c));
      o.x = (v[0] - m[0]) * rs[0] * g4.x + b4.x;
      o.y = (v[1] - m[1]) * rs[1] * g4.y + b4.y;
      o.z = (v[2] - m[2]) * rs[2] * g4.z + b4.z;
      o.w = (v[3] - m[3]) * rs[3] * g4.w + b4.w;
    }
    *reinterpret_cast<float4*>(z + (size_t)j * C + c) = o;
  }
  __syncthreads();
  // thread = (channel c, row lane): rows lane, lane + L, ... of the tile (L = NT / C row lanes)
  const int L = NT / C, c = tid % C, lane = tid / C;
  const float* w = wsm + (c / (C / p.fir_groups)) * K;
  constexpr int MAXO = 16;                     // FIR_TT / L outputs per thread (C >= 64: L <= 4 -> <= 8 ... C = 128: 16)
  float acc[MAXO];
#pragma unroll
  for (int o = 0; o < MAXO; ++o) acc[o] = 0.f;
  const int no = FIR_TT / L;                   // <= MAXO (host checks)
  for (int k = 0; k < K; ++k) {
    const float wk = w[k];
#pragma unroll
    for (int o = 0; o < MAXO; ++o)
      if (o < no) acc[o] = fmaf(wk, z[(size_t)(lane * no + o + k) * C + c], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < MAXO; ++o) {
    if (o >= no) break;
    const long long r = t0 + lane * no + o;
    if (r >= p.rows) break;
    const long long i = ((long long)b * p.rows + r) * C + c;
    if (p.out) p.out[i] = acc[o];
    if (p.out_hi) split_tf32_dev(acc[o], p.out_hi[i], p.out_lo[i]);
  }
}

// pass 2 without FIR, 4 channels per thread (C, Cin, C / stat_groups multiples of 4)
__global__ void __launch_bounds__(NT) chan_norm_vec_kernel(const NormParams p) {
  const int C4 = p.C >> 2;
  const long long n = (long long)p.B * p.rows * C4;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int c = 4 * (int)(i % C4);
    const long long br = i / C4;
    const int b = (int)(br / p.rows);
    const long long r = br - (long long)b * p.rows;
    float v[4];
    pre_value4(p.x + br * p.Cin, c, p.Cin, p.C, p.pre, p.pre_slope, v);
    float m[4], rs[4];
    if (p.stat_mode == SE_NORM_STAT_INSTANCE) {
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mean + (long long)b * p.C + c));
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.rstd + (long long)b * p.C + c));
      m[0] = m4.x, m[1] = m4.y, m[2] = m4.z, m[3] = m4.w;
      rs[0] = r4.x, rs[1] = r4.y, rs[2] = r4.z, rs[3] = r4.w;
    } else {
      const long long si = ((long long)b * (p.rows / p.rows_per_t) + r / p.rows_per_t) * p.stat_groups +
                           c / (p.C / p.stat_groups);
      const float mm = __ldg(p.mean + si), rr = __ldg(p.rstd + si);
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = mm, rs[e] = rr;
    }
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + c));
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float sv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.post == SE_NORM_POST_PRELU) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.post_slope + c));
      sv[0] = s4.x, sv[1] = s4.y, sv[2] = s4.z, sv[3] = s4.w;
    }
    float y[4], hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      y[e] = (v[e] - m[e]) * rs[e] * gv[e] + bv[e];
      if (p.post == SE_NORM_POST_PRELU) y[e] = prelu_f(y[e], sv[e]);
      if (p.out_hi) split_tf32_dev(y[e], hi[e], lo[e]);
    }
    const long long o = br * p.C + c;
    if (p.out) *reinterpret_cast<float4*>(p.out + o) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.out_hi) {
      *reinterpret_cast<float4*>(p.out_hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(p.out_lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ---- small elementwise helpers -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float ca,
                                                  float cb, long long n, float* __restrict__ out,
                                                  float* __restrict__ out_hi, float* __restrict__ out_lo) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    // ca == cb == 1 is the plain sum (no multiply, so it rounds exactly like torch's a + b)
    const float y = (ca == 1.f && cb == 1.f) ? __ldg(a + i) + __ldg(b + i) : ca * __ldg(a + i) + cb * __ldg(b + i);
    if (out) out[i] = y;
    if (out_hi) split_tf32_dev(y, out_hi[i], out_lo[i]);
  }
}

// CTSNet stage glue (CTSNet/two_stage_com_decode_vb.py:79-84).
//   stage 1 -> 2: s1 = est_mag * (cos, sin)(phase of the noisy spectrum); s2_in = cat(noisy RI, s1 RI) [.., 4]
__global__ void __launch_bounds__(NT) cts_glue1_kernel(const float2* __restrict__ x, const float* __restrict__ est,
                                                      long long n, float4* __restrict__ s2in) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float2 v = __ldg(x + i);
    const float ph = atan2f(v.y, v.x);      // :74 phase_x (0 for a zero bin, as torch.atan2)
    float s, c;
    sincosf(ph, &s, &c);
    const float e = __ldg(est + i);
    s2in[i] = make_float4(v.x, v.y, e * c, e * s);
  }
}
//   stage 2 output: model2(s2_in) + s1  (:84), interleaved (re, im) for the iSTFT prologue
__global__ void __launch_bounds__(NT) cts_glue2_kernel(const float* __restrict__ out_r, const float* __restrict__ out_i,
                                                      const float4* __restrict__ s2in, long long n,
                                                      float2* __restrict__ est) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float4 s = __ldg(s2in + i);
    est[i] = make_float2(__ldg(out_r + i) + s.z, __ldg(out_i + i) + s.w);
  }
}

// TaylorSENet zeroth-order term (TaylorSENet/TaylorSENet.py:73-76): gain * |X| * (cos, sin)(angle X) written as one
// "RI row" per frame: [re(F) | im(F) | zero pad] of width ld (the layout the high-order GEMMs read and write).
__global__ void __launch_bounds__(NT) taylor_zero_kernel(const float2* __restrict__ x, const float* __restrict__ gain,
                                                        long long rows, int F, int ld, float* __restrict__ out,
                                                        float* __restrict__ out_hi, float* __restrict__ out_lo) {
  const long long n = rows * ld;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long r = i / ld;
    const int c = (int)(i - r * ld);
    float y = 0.f;
    if (c < 2 * F) {
      const int f = c < F ? c : c - F;
      const float2 v = __ldg(x + r * F + f);
      const float mag = sqrtf(v.x * v.x + v.y * v.y);          // torch.norm(inputs, dim=1)
      const float ph = atan2f(v.y, v.x);
      const float zm = __ldg(gain + r * F + f) * mag;
      y = c < F ? zm * cosf(ph) : zm * sinf(ph);
    }
    out[i] = y;
    if (out_hi) split_tf32_dev(y, out_hi[i], out_lo[i]);
  }
}

// G2Net stage update (G2Net_new/gaf_net_320.py:104-115): x' = gain * |pre| * (cos, sin)(angle pre) + com_resi, with
// every complex tensor held as "RI rows" [rows, ld]: re at column 0, im at column im_off, zeros elsewhere.  pre is read
// through (pointer, row stride, element stride) so that the very first stage can read the channels-last network input
// [rows, F, 2] directly.  gain == NULL: plain relayout of pre into RI rows (feature for the first in_conv GEMM).
__global__ void __launch_bounds__(NT) gaf_update_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im,
                                                       long long xs_r, int xs_f, const float* __restrict__ gain,
                                                       const float* __restrict__ resi, long long rows, int F, int ld,
                                                       int im_off, float* __restrict__ out, float* __restrict__ out_hi,
                                                       float* __restrict__ out_lo) {
  const long long n = rows * ld;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long r = i / ld;
    const int c = (int)(i - r * ld);
    float y = 0.f;
    const bool is_re = c < F, is_im = c >= im_off && c < im_off + F;
    if (is_re || is_im) {
      const int f = is_re ? c : c - im_off;
      const float re = __ldg(x_re + r * xs_r + (long long)f * xs_f), im = __ldg(x_im + r * xs_r + (long long)f * xs_f);
      if (gain) {
        const float mag = sqrtf(re * re + im * im);             // torch.norm(pre_x, dim=1)
        const float ph = atan2f(im, re);
        const float xm = mag * __ldg(gain + r * F + f);
        y = (is_re ? xm * cosf(ph) : xm * sinf(ph)) + __ldg(resi + i);
      } else {
        y = is_re ? re : im;
      }
    }
    if (out) out[i] = y;
    if (out_hi) split_tf32_dev(y, out_hi[i], out_lo[i]);
  }
}

static int grid_for(long long n) {
  long long g = (n + NT - 1) / NT;
  return (int)(g < 148 * 16 ? (g < 1 ? 1 : g) : 148 * 16);
}

}  // namespace se

using namespace se;

static bool norm_c_ok(int C) { return C >= 1 && C <= NT && NT % C == 0; }

extern "C" long long se_chan_stats_ws_bytes(int B, long long rows, int C) {
  (void)rows;
  // upper bound on chunks: 1184 / B + 1
  const long long chunks = 148 * 8 / (B > 0 ? B : 1) + 1;
  return 4096 + (long long)B * chunks * C * 2 * 8;
}

extern "C" int se_chan_stats(const float* x, int B, long long rows, int Cin, int C, int pre, const float* pre_slope,
                             float eps, float* mean, float* rstd, void* ws, se_stream_t stream) {
  SE_REQUIRE(x && mean && rstd && ws && B > 0 && B <= 1024 && rows > 0, "se_chan_stats: bad arguments (B <= 1024)");
  SE_REQUIRE(norm_c_ok(C), "se_chan_stats: C=%d must divide %d", C, NT);
  SE_REQUIRE(pre >= SE_NORM_PRE_NONE && pre <= SE_NORM_PRE_GLU_PRELU, "se_chan_stats: pre=%d", pre);
  SE_REQUIRE((pre == SE_NORM_PRE_GLU || pre == SE_NORM_PRE_GLU_PRELU) ? Cin == 2 * C
                                                                      : (Cin >= 1 && C % Cin == 0 && (pre != SE_NORM_PRE_NONE || Cin == C)),
             "se_chan_stats: Cin=%d C=%d pre=%d", Cin, C, pre);
  SE_REQUIRE(!(pre == SE_NORM_PRE_PRELU || pre == SE_NORM_PRE_GLU_PRELU) || pre_slope, "se_chan_stats: slopes missing");
  const int lanes = NT / C;
  long long chunks = 148 * 8 / B;
  const long long maxc = (rows + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  // tickets FIRST, at a fixed place: a workspace reused across shapes must find them zero wherever the partials were
  unsigned* ticket = reinterpret_cast<unsigned*>(ws);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + 4096);
  const bool vec = (C & 3) == 0 && (Cin & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)pre_slope)) & 15) == 0;
  if (vec)
    chan_stats_vec_kernel<<<dim3((unsigned)chunks, (unsigned)B), NT, 0, (cudaStream_t)stream>>>(
        x, rows, Cin, C, pre, pre_slope, eps, mean, rstd, partial, ticket);
  else
    chan_stats_kernel<<<dim3((unsigned)chunks, (unsigned)B), NT, 0, (cudaStream_t)stream>>>(
        x, rows, Cin, C, pre, pre_slope, eps, mean, rstd, partial, ticket);
  return check_launch("se_chan_stats");
}

extern "C" int se_cum_stats(const float* x, int B, int T, int F, int Cin, int C, int groups, int pre,
                            const float* pre_slope, float eps, float* mean, float* rstd, void* ws, se_stream_t stream) {
  SE_REQUIRE(x && mean && rstd && ws && B > 0 && T > 0 && F > 0 && C > 0, "se_cum_stats: bad arguments");
  SE_REQUIRE(groups >= 1 && C % groups == 0, "se_cum_stats: groups=%d C=%d", groups, C);
  SE_REQUIRE(pre >= SE_NORM_PRE_NONE && pre <= SE_NORM_PRE_GLU_PRELU, "se_cum_stats: pre=%d", pre);
  SE_REQUIRE((pre == SE_NORM_PRE_GLU || pre == SE_NORM_PRE_GLU_PRELU) ? Cin == 2 * C
                                                                      : (Cin >= 1 && C % Cin == 0 && (pre != SE_NORM_PRE_NONE || Cin == C)),
             "se_cum_stats: Cin=%d C=%d pre=%d", Cin, C, pre);
  SE_REQUIRE(!(pre == SE_NORM_PRE_PRELU || pre == SE_NORM_PRE_GLU_PRELU) || pre_slope, "se_cum_stats: slopes missing");
  double* step = reinterpret_cast<double*>(ws);        // [B*T][G][2]
  const long long frames = (long long)B * T;
  cum_step_kernel<<<(unsigned)((frames * groups + NT / 32 - 1) / (NT / 32)), NT, 0, (cudaStream_t)stream>>>(
      x, frames, F, Cin, C, groups, pre, pre_slope, step);
  int rc = check_launch("se_cum_stats (step sums)");
  if (rc) return rc;
  cum_scan_kernel<<<dim3((unsigned)B, (unsigned)groups), 32, 0, (cudaStream_t)stream>>>(step, T, groups, F * (C / groups),
                                                                                          eps, mean, rstd);
  return check_launch("se_cum_stats (scan)");
}

extern "C" int se_chan_norm(const float* x, int B, long long rows, int Cin, int C, int pre, const float* pre_slope,
                            const float* mean, const float* rstd, int stat_mode, int rows_per_t, int stat_groups,
                            const float* gamma, const float* beta, int post, const float* post_slope, const float* fir_w, int fir_k,
                            int fir_groups, float* out, float* out_hi, float* out_lo, se_stream_t stream) {
  SE_REQUIRE(x && mean && rstd && gamma && beta && (out || out_hi) && B > 0 && rows > 0 && C > 0,
             "se_chan_norm: bad arguments");
  SE_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "se_chan_norm: out_hi/out_lo go together");
  SE_REQUIRE(pre >= SE_NORM_PRE_NONE && pre <= SE_NORM_PRE_GLU_PRELU, "se_chan_norm: pre=%d", pre);
  SE_REQUIRE((pre == SE_NORM_PRE_GLU || pre == SE_NORM_PRE_GLU_PRELU) ? Cin == 2 * C
                                                                      : (Cin >= 1 && C % Cin == 0 && (pre != SE_NORM_PRE_NONE || Cin == C)),
             "se_chan_norm: Cin=%d C=%d pre=%d", Cin, C, pre);
  SE_REQUIRE(!(pre == SE_NORM_PRE_PRELU || pre == SE_NORM_PRE_GLU_PRELU) || pre_slope, "se_chan_norm: slopes missing");
  SE_REQUIRE(stat_mode == SE_NORM_STAT_INSTANCE || (stat_mode == SE_NORM_STAT_CUMULATIVE && rows_per_t >= 1 &&
                                                     rows % rows_per_t == 0 && stat_groups >= 1 && C % stat_groups == 0),
             "se_chan_norm: stat_mode=%d rows_per_t=%d stat_groups=%d", stat_mode, rows_per_t, stat_groups);
  SE_REQUIRE(post >= SE_NORM_POST_NONE && post <= SE_NORM_POST_FIR, "se_chan_norm: post=%d", post);
  SE_REQUIRE(post != SE_NORM_POST_PRELU || post_slope, "se_chan_norm: post slopes missing");
  SE_REQUIRE(post != SE_NORM_POST_FIR || (fir_w && fir_k >= 1 && fir_groups >= 1 && C % fir_groups == 0 &&
                                          (stat_mode == SE_NORM_STAT_INSTANCE || rows_per_t == 1)),
             "se_chan_norm: FIR arguments");
  NormParams p{x, B, rows, Cin, C, pre, pre_slope, mean, rstd, stat_mode, rows_per_t, stat_groups, gamma, beta, post, post_slope,
               fir_w, fir_k, fir_groups, out, out_hi, out_lo};
  auto al16 = [](const void* q) { return (((uintptr_t)q) & 15) == 0; };
  const bool vec = post != SE_NORM_POST_FIR && (C & 3) == 0 && (Cin & 3) == 0 &&
                   (stat_mode == SE_NORM_STAT_INSTANCE || ((C / stat_groups) & 3) == 0) && al16(x) && al16(pre_slope) &&
                   al16(mean) && al16(rstd) && al16(gamma) && al16(beta) && al16(post_slope) && al16(out) && al16(out_hi) &&
                   al16(out_lo);
  // tiled FIR kernel: channels divide the block, FIR_TT / (NT / C) outputs per thread fit its accumulators
  const bool fir_tiled = post == SE_NORM_POST_FIR && (C & 3) == 0 && (Cin & 3) == 0 && NT % C == 0 && FIR_TT % (NT / C) == 0 &&
                         FIR_TT / (NT / C) <= 16 && (stat_mode == SE_NORM_STAT_INSTANCE || ((C / stat_groups) & 3) == 0) &&
                         al16(x) && al16(pre_slope) && al16(mean) && al16(rstd) && al16(gamma) && al16(beta);
  const size_t fir_smem_bytes = ((size_t)(FIR_TT + fir_k - 1) * C + (size_t)fir_groups * fir_k) * sizeof(float);
  if (fir_tiled && fir_smem_bytes <= 200 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(chan_norm_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fir_smem_bytes);
    if (e != cudaSuccess) {
      set_error("se_chan_norm: %zu bytes of shared memory: %s", fir_smem_bytes, cudaGetErrorString(e));
      return SE_ERR_CUDA;
    }
    chan_norm_fir_kernel<<<dim3((unsigned)((rows + FIR_TT - 1) / FIR_TT), (unsigned)B), NT, fir_smem_bytes, (cudaStream_t)stream>>>(p);
  } else if (vec)
    chan_norm_vec_kernel<<<grid_for((long long)B * rows * (C / 4)), NT, 0, (cudaStream_t)stream>>>(p);
  else
    chan_norm_kernel<<<grid_for((long long)B * rows * C), NT, 0, (cudaStream_t)stream>>>(p);
  return check_launch("se_chan_norm");
}

extern "C" int se_axpby(const float* a, const float* b, float ca, float cb, long long n, float* out, float* out_hi,
                        float* out_lo, se_stream_t stream) {
  SE_REQUIRE(a && b && n > 0 && (out || out_hi), "se_axpby: bad arguments");
  SE_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "se_axpby: out_hi/out_lo go together");
  axpby_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(a, b, ca, cb, n, out, out_hi, out_lo);
  return check_launch("se_axpby");
}

extern "C" int se_add(const float* a, const float* b, long long n, float* out, float* out_hi, float* out_lo,
                      se_stream_t stream) {
  return se_axpby(a, b, 1.f, 1.f, n, out, out_hi, out_lo, stream);
}

extern "C" int se_taylor_zero(const float* x_ri, const float* gain, long long rows, int F, int ld, float* out,
                              float* out_hi, float* out_lo, se_stream_t stream) {
  SE_REQUIRE(x_ri && gain && out && rows > 0 && F > 0 && ld >= 2 * F, "se_taylor_zero: bad arguments");
  SE_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "se_taylor_zero: out_hi/out_lo go together");
  SE_REQUIRE((((uintptr_t)x_ri) & 7) == 0, "se_taylor_zero: unaligned");
  taylor_zero_kernel<<<grid_for(rows * ld), NT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x_ri), gain,
                                                                           rows, F, ld, out, out_hi, out_lo);
  return check_launch("se_taylor_zero");
}

extern "C" int se_cts_glue1(const float* x_ri, const float* est_mag, long long n, float* s2_in, se_stream_t stream) {
  SE_REQUIRE(x_ri && est_mag && s2_in && n > 0, "se_cts_glue1: bad arguments");
  SE_REQUIRE(((((uintptr_t)x_ri) & 7) | (((uintptr_t)s2_in) & 15)) == 0, "se_cts_glue1: unaligned");
  cts_glue1_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x_ri), est_mag, n,
                                                                 reinterpret_cast<float4*>(s2_in));
  return check_launch("se_cts_glue1");
}

extern "C" int se_cts_glue2(const float* out_r, const float* out_i, const float* s2_in, long long n, float* est,
                            se_stream_t stream) {
  SE_REQUIRE(out_r && out_i && s2_in && est && n > 0, "se_cts_glue2: bad arguments");
  SE_REQUIRE(((((uintptr_t)est) & 7) | (((uintptr_t)s2_in) & 15)) == 0, "se_cts_glue2: unaligned");
  cts_glue2_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(out_r, out_i, reinterpret_cast<const float4*>(s2_in), n,
                                                                 reinterpret_cast<float2*>(est));
  return check_launch("se_cts_glue2");
}

extern "C" int se_gaf_update(const float* x_re, const float* x_im, long long xs_r, int xs_f, const float* gain,
                             const float* resi, long long rows, int F, int ld, int im_off, float* out, float* out_hi,
                             float* out_lo, se_stream_t stream) {
  SE_REQUIRE(x_re && x_im && (out || out_hi) && rows > 0 && F > 0 && im_off >= F && ld >= im_off + F,
             "se_gaf_update: bad arguments");
  SE_REQUIRE((gain == nullptr) == (resi == nullptr), "se_gaf_update: gain and resi go together");
  SE_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "se_gaf_update: out_hi/out_lo go together");
  gaf_update_kernel<<<grid_for(rows * ld), NT, 0, (cudaStream_t)stream>>>(x_re, x_im, xs_r, xs_f, gain, resi, rows, F, ld,
                                                                          im_off, out, out_hi, out_lo);
  return check_launch("se_gaf_update");
}
