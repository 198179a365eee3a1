// Implicit-GEMM causal Conv2d / ConvTranspose2d / Linear on channels-last fp32 activations.
//
// FP32 FMA (CUDA-core) path.  SURVEY.md Appendix B: rounding only the weights to TF32
// already costs 1.1e-4 RMS on CRN -- above the 1e-4 parity gate -- so single-pass
// tensor-core math is out; this kernel is the exact-fp32 engine (packed FFMA2 on
// sm_100a), the 3xTF32 tcgen05 engine replaces it for the large contractions.
//
//   out[m, n] = act(bias[n] + sum_k A[m, k] * W[k, n]),   m = (b, t, fo),  k = (tap, ci)
//   A[m, k]   = in[b, t + dt[tap], fo*sf + df[tap], ci]   (zero outside the tensor)
//
// CTA tile BM x BN, K step 16, 256 threads, register tile TM x TN, double-buffered
// shared memory with register-staged global loads.
#include "common.cuh"

namespace se {

constexpr int BK = 16;
constexpr int kGemmThreads = 256;

struct ConvParams {
  se_conv_desc d;
  int M, K, Ctot;
};

template <int BM, int BN, int TM, int TN, bool ALIGNED>
__global__ void __launch_bounds__(kGemmThreads) conv_gemm_kernel(const ConvParams P) {
  static_assert((BM / TM) * (BN / TN) == kGemmThreads, "tile/thread mismatch");
  static_assert(TM == 4 || TM == 8, "TM");
  static_assert(TN == 4 || TN == 8, "TN");
  constexpr int NA = BM / 64;                    // float4 A loads per thread per K step
  constexpr int NBQ = (BK * BN / 4);             // float4 B loads per CTA per K step
  constexpr int NB = (NBQ + kGemmThreads - 1) / kGemmThreads;
  constexpr int TNX = BN / TN;                   // threads along n

  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const se_conv_desc& d = P.d;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tn = tid % TNX, tm = tid / TNX;

  // ---- per-thread A row bookkeeping (rows are fixed over the K loop) ---------------------
  int r_t[NA], r_fo[NA];
  long long r_bt[NA];
  bool r_ok[NA];
  const int kq = tid & 3;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int row = (tid >> 2) + i * 64;
    const int m = m0 + row;
    r_ok[i] = m < P.M;
    const int mm = r_ok[i] ? m : 0;
    const int fo = mm % d.Fout;
    const int bt = mm / d.Fout;
    r_fo[i] = fo;
    r_t[i] = bt % d.T;
    r_bt[i] = (long long)(bt - r_t[i]);  // b*T
  }

  float4 ra[NA];
  float4 rb[NB];

  auto load_tiles = [&](int k0) {
    // ---- A ----
    if (ALIGNED) {
      const int tap = k0 / P.Ctot;
      const int c = k0 - tap * P.Ctot;
      const float* src;
      int cs, Cs;
      if (c < d.C0) {
        src = d.src0; cs = c; Cs = d.C0;
      } else {
        src = d.src1; cs = c - d.C0; Cs = d.C1;
      }
      const int dt = d.dt[tap], df = d.df[tap];
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int ti = r_t[i] + dt;
        const int fi = r_fo[i] * d.sf + df;
        const bool ok = r_ok[i] && ti >= 0 && ti < d.T && fi >= 0 && fi < d.Fin;
        if (ok) {
          const long long pos = (r_bt[i] + ti) * d.Fin + fi;
          ra[i] = __ldg(reinterpret_cast<const float4*>(src + pos * Cs + cs + kq * 4));
        } else {
          ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k0 + kq * 4 + e;
          float x = 0.f;
          if (k < P.K && r_ok[i]) {
            const int tap = k / P.Ctot;
            const int c = k - tap * P.Ctot;
            const int ti = r_t[i] + d.dt[tap];
            const int fi = r_fo[i] * d.sf + d.df[tap];
            if (ti >= 0 && ti < d.T && fi >= 0 && fi < d.Fin) {
              const long long pos = (r_bt[i] + ti) * d.Fin + fi;
              x = (c < d.C0) ? __ldg(d.src0 + pos * d.C0 + c) : __ldg(d.src1 + pos * d.C1 + (c - d.C0));
            }
          }
          v[e] = x;
        }
        ra[i] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    // ---- B ----
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int f = tid + i * kGemmThreads;
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < NBQ) {
        const int kr = f / (BN / 4);
        const int c4 = f - kr * (BN / 4);
        const int k = k0 + kr;
        const int n = n0 + c4 * 4;
        if (k < P.K && n < d.ldw) rb[i] = __ldg(reinterpret_cast<const float4*>(d.W + (long long)k * d.ldw + n));
      }
    }
  };

  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int row = (tid >> 2) + i * 64;
      const int slot = (row >> 2) ^ (2 * kq);       // float4-slot swizzle keyed by k>>2
      const int mcol = slot * 4 + (row & 3);
      As[buf][kq * 4 + 0][mcol] = ra[i].x;
      As[buf][kq * 4 + 1][mcol] = ra[i].y;
      As[buf][kq * 4 + 2][mcol] = ra[i].z;
      As[buf][kq * 4 + 3][mcol] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int f = tid + i * kGemmThreads;
      if (f < NBQ) {
        const int kr = f / (BN / 4);
        const int c4 = f - kr * (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kr][c4 * 4]) = rb[i];
      }
    }
  };

  float2 acc[TM][TN / 2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  const int nk = (P.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const int q2 = 2 * ((k >> 2) & 3);
      float a[TM];
      float2 b[TN / 2];
      if constexpr (TM == 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][((2 * tm) ^ q2) * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][((2 * tm + 1) ^ q2) * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][(tm ^ q2) * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      }
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
        b[0] = make_float2(b0.x, b0.y);
        b[1] = make_float2(b0.z, b0.w);
        if constexpr (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tn * 4]);
          b[2] = make_float2(b1.x, b1.y);
          b[3] = make_float2(b1.z, b1.w);
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc[i][j] = ffma2(aa, b[j], acc[i][j]);
      }
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias + activation + channels-last store ---------------------------------------
  const bool vec_ok = ((d.Cout & 3) == 0) && ((((uintptr_t)d.dst) & 15) == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm * TM + i;
    if (m >= P.M) continue;
    const int fo = m % d.Fout;
    const int bt = m / d.Fout;
    float* orow = d.dst + ((long long)bt * d.dstF + d.dst_f0 + (long long)fo * d.dst_fstep) * d.Cout;
#pragma unroll
    for (int h = 0; h < TN / 4; ++h) {
      const int n = n0 + (h == 0 ? tn * 4 : BN / 2 + tn * 4);
      float v[4] = {acc[i][2 * h].x, acc[i][2 * h].y, acc[i][2 * h + 1].x, acc[i][2 * h + 1].y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n + e < d.Cout) {
          const float bb = d.bias ? __ldg(d.bias + n + e) : 0.f;
          v[e] = apply_act(v[e] + bb, d.act, d.act_param);
        }
      }
      if (vec_ok && n + 3 < d.Cout) {
        *reinterpret_cast<float4*>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < d.Cout) orow[n + e] = v[e];
      }
    }
  }
}

// left-pad column that still goes through BN + activation (CRN de4)
__global__ void fill_col_kernel(float* dst, long long rows, int dstF, int Cout, int fill_f, const float* fill,
                                int act, float act_param) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Cout) return;
  const long long r = idx / Cout;
  const int co = (int)(idx - r * Cout);
  dst[(r * dstF + fill_f) * Cout + co] = apply_act(__ldg(fill + co), act, act_param);
}

// ---------------------------------------------------------------------------------------------
// First encoder layer: Conv2d(1 -> Cout, k(2,3), s(1,2)), causal in T.  HBM-bound (writes
// Fout*Cout floats per frame from Fin inputs).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_in1_kernel(const float* __restrict__ src, int B, int T, int Fin,
                                                      const float* __restrict__ W, const float* __restrict__ bias,
                                                      int Cout, int act, float* __restrict__ dst, int Fout) {
  __shared__ float ws[6 * 64 + 64];
  for (int i = threadIdx.x; i < 6 * Cout; i += blockDim.x) ws[i] = W[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[6 * 64 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cq = Cout >> 2;
  const long long total = (long long)B * T * Fout * cq;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % cq);
    const long long pos = idx / cq;
    const int fo = (int)(pos % Fout);
    const long long bt = pos / Fout;
    const int t = (int)(bt % T);
    const float* r1 = src + bt * Fin + 2 * fo;  // frame t   (kt = 1)
    const float* r0 = r1 - Fin;                  // frame t-1 (kt = 0), zero for t == 0
    float x[6];
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      x[kf] = t > 0 ? __ldg(r0 + kf) : 0.f;
      x[3 + kf] = __ldg(r1 + kf);
    }
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = q * 4 + e;
      float a = ws[6 * 64 + co];
#pragma unroll
      for (int tp = 0; tp < 6; ++tp) a = fmaf(x[tp], ws[tp * Cout + co], a);
      o[e] = apply_act(a, act);
    }
    *reinterpret_cast<float4*>(dst + pos * Cout + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Last decoder layer: ConvTranspose2d(C0+C1 -> 1, k(2,3), s(1,2)), last frame dropped.
// out[b,t,f'] = act(bias + sum_{kt,kf: f'=2f+kf} sum_ci in[b,t-kt,f,ci] * W[kt*3+kf][ci])
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deconv_out1_kernel(const float* __restrict__ src0,
                                                         const float* __restrict__ src1, int C0, int C1, int B,
                                                         int T, int Fin, const float* __restrict__ W, float bias,
                                                         int act, float* __restrict__ dst) {
  extern __shared__ float wsm[];  // [6][C0+C1]
  const int Ct = C0 + C1;
  for (int i = threadIdx.x; i < 6 * Ct; i += blockDim.x) wsm[i] = W[i];
  __syncthreads();
  const int Fo = 2 * Fin + 1;
  const long long total = (long long)B * T * Fo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int fp = (int)(idx % Fo);
    const long long bt = idx / Fo;
    const int t = (int)(bt % T);
    float acc = bias;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      if (t - kt < 0) continue;
      const long long rowpos = (bt - kt) * Fin;
      for (int kf = (fp & 1); kf < 3; kf += 2) {
        const int f = (fp - kf) >> 1;
        if (f < 0 || f >= Fin) continue;
        const float* w = wsm + (kt * 3 + kf) * Ct;
        const float4* p0 = reinterpret_cast<const float4*>(src0 + (rowpos + f) * C0);
        for (int c = 0; c < C0 / 4; ++c) {
          const float4 v = __ldg(p0 + c);
          acc = fmaf(v.x, w[4 * c], acc);
          acc = fmaf(v.y, w[4 * c + 1], acc);
          acc = fmaf(v.z, w[4 * c + 2], acc);
          acc = fmaf(v.w, w[4 * c + 3], acc);
        }
        if (C1 > 0) {
          const float4* p1 = reinterpret_cast<const float4*>(src1 + (rowpos + f) * C1);
          for (int c = 0; c < C1 / 4; ++c) {
            const float4 v = __ldg(p1 + c);
            acc = fmaf(v.x, w[C0 + 4 * c], acc);
            acc = fmaf(v.y, w[C0 + 4 * c + 1], acc);
            acc = fmaf(v.z, w[C0 + 4 * c + 2], acc);
            acc = fmaf(v.w, w[C0 + 4 * c + 3], acc);
          }
        }
      }
    }
    dst[idx] = apply_act(acc, act);
  }
}

// The same layer, one thread per INPUT position (round 2): thread (row r, column f) of a CTA of R x Fin threads computes
//   a_kf = <in[t, f, :], W[kf]> + <in[t-1, f, :], W[3 + kf]>,   kf = 0, 1, 2
// from two contiguous channel vectors (float4 loads; a warp reads one contiguous span per source and row), hands a_2 to
// its right-hand neighbour through shared memory and stores out[t, 2f] = a_0(f) + a_2(f-1), out[t, 2f+1] = a_1(f)
// (+ out[t, 2 Fin] = a_2(Fin-1)): every input is read once per time offset instead of once per output that touches it,
// with the same FMA count.  The per-output kernel above ran at 17 % of the HBM roofline (264 MB in 237 us).
template <int CT>
__global__ void __launch_bounds__(256) deconv_out1_rows_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                                              int C0, int C1, int B, int T, int Fin, int R,
                                                              const float* __restrict__ W, float bias, int act,
                                                              float* __restrict__ dst) {
  __shared__ __align__(16) float wsm[6 * CT];
  __shared__ float a2s[256];
  for (int i = threadIdx.x; i < 6 * CT; i += blockDim.x) wsm[i] = W[i];
  const int r = threadIdx.x / Fin, f = threadIdx.x - r * Fin;
  const int tiles_t = (T + R - 1) / R;
  const int b = blockIdx.x / tiles_t;
  const int t = (blockIdx.x - b * tiles_t) * R + r;
  const bool live = r < R && t < T;
  const bool wide = (C0 & 7) == 0 && (C1 & 7) == 0 && ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(src1)) & 31) == 0;
  __syncthreads();
  float a[3] = {0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      if (t - kt < 0) continue;
      const long long pos = ((long long)b * T + (t - kt)) * Fin + f;
      const float4* w4 = reinterpret_cast<const float4*>(wsm + kt * 3 * CT);
      // a thread reads its own contiguous channel vector: 32-byte loads (whole sectors) where the vectors are 32-byte multiples
      auto dot8 = [&](const float* src, int cbase) {
        float v[8];
        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(src));
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
          const float4 w0 = w4[kf * (CT / 4) + cbase], w1 = w4[kf * (CT / 4) + cbase + 1];
          a[kf] = fmaf(v[0], w0.x, fmaf(v[1], w0.y, fmaf(v[2], w0.z, fmaf(v[3], w0.w, a[kf]))));
          a[kf] = fmaf(v[4], w1.x, fmaf(v[5], w1.y, fmaf(v[6], w1.z, fmaf(v[7], w1.w, a[kf]))));
        }
      };
      auto dot4 = [&](const float* src, int cbase) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
          const float4 w = w4[kf * (CT / 4) + cbase];
          a[kf] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, a[kf]))));
        }
      };
      const float* p0 = src0 + pos * C0;
      if (wide) {
#pragma unroll 2
        for (int c = 0; c < C0 / 8; ++c) dot8(p0 + 8 * c, 2 * c);
      } else {
#pragma unroll 4
        for (int c = 0; c < C0 / 4; ++c) dot4(p0 + 4 * c, c);
      }
      if (C1 > 0) {
        const float* p1 = src1 + pos * C1;
        if (wide) {
#pragma unroll 2
          for (int c = 0; c < C1 / 8; ++c) dot8(p1 + 8 * c, C0 / 4 + 2 * c);
        } else {
#pragma unroll 4
          for (int c = 0; c < C1 / 4; ++c) dot4(p1 + 4 * c, C0 / 4 + c);
        }
      }
    }
  }
  a2s[threadIdx.x] = a[2];
  __syncthreads();
  if (!live) return;
  float* o = dst + ((long long)b * T + t) * (2 * Fin + 1) + 2 * f;
  o[0] = apply_act(bias + a[0] + (f > 0 ? a2s[threadIdx.x - 1] : 0.f), act);
  o[1] = apply_act(bias + a[1], act);
  if (f == Fin - 1) o[2] = apply_act(bias + a[2], act);
}

// ---------------------------------------------------------------------------------------------
// Narrow-output variant of the same contract (Cout <= 4: the last decoder layers that emit the 2-channel RI spectrum
// or a 1-channel mask -- CTSNet/Step2_network.py de5, DCCRN decoder.5, DPCRN's CRM head, Uformer's mask heads).  The
// tiled kernel above computes a 16- or 32-wide column tile whatever Cout is, i.e. 8-16x the useful FMAs, and re-reads
// every activation once per tap from L2.  Here one CTA owns one output frame (b, t):
//   * the <= 4 input frames it needs (distinct dt of the taps) are staged ONCE in shared memory as [Fin][C0 + C1] rows
//     (float4, coalesced; frames outside the tensor are zero), the K x NCO weights next to them;
//   * a warp computes 4 neighbouring output columns at a time: lanes stride over the channels (float4 from shared
//     memory, conflict-free), each weight vector is loaded once for the 4 columns, 4 x NCO partial sums per lane,
//     butterfly reduction, lanes 0 .. 4 NCO - 1 store.
// A first version with one warp per output and global loads (no staging) was 1.4-2x SLOWER than the tiled kernel: the
// im2col-expanded L2 -> SM traffic, not the FMAs, is what these layers cost (profiles/models_r01h.jsonl).
// ---------------------------------------------------------------------------------------------
constexpr int kRowsMaxDt = 4;
struct ConvRowsParams {
  ConvParams c;
  int ndt;
  int dtv[kRowsMaxDt];          // distinct time offsets of the taps
  int tap_dti[SE_MAX_TAPS];     // tap -> index into dtv
  // blockIdx.y = chunk of fc output columns: only the input columns [fo*sf + dfmin, fo*sf + dfmax] of the chunk are staged,
  // so a CTA needs tens of KB instead of a whole frame per time offset and several CTAs share an SM (one stages while
  // another computes; with whole frames -- 164 KB at Fin = 80, C = 256 -- the two phases of the only resident CTA were serial)
  int fc, dfmin, dfmax, nin_max;
};

template <int NCO>
__global__ void __launch_bounds__(256) conv_rows_kernel(const ConvRowsParams R) {
  extern __shared__ __align__(16) float smem_rows[];
  const ConvParams& P = R.c;
  const se_conv_desc& d = P.d;
  const int Ct = P.Ctot;
  const int f_lo = blockIdx.y * R.fc, f_hi = min(d.Fout, f_lo + R.fc);          // output columns of this CTA
  const int in_lo = max(0, f_lo * d.sf + R.dfmin), in_hi = min(d.Fin - 1, (f_hi - 1) * d.sf + R.dfmax);
  const int nin = max(0, in_hi - in_lo + 1);                  // staged input columns [in_lo, in_hi]
  float* rows = smem_rows;                                   // [ndt][nin_max][Ct]
  float* wsm = rows + (size_t)R.ndt * R.nin_max * Ct;        // [K][NCO]
  const int bt = blockIdx.x;                                 // b * T + t
  const int t = bt % d.T;
  for (int i = threadIdx.x; i < P.K * NCO; i += blockDim.x) {
    const int k = i / NCO, co = i - k * NCO;
    wsm[i] = co < d.Cout ? __ldg(d.W + (size_t)k * d.ldw + co) : 0.f;
  }
  const int c4 = Ct >> 2, c04 = d.C0 >> 2;
  for (int di = 0; di < R.ndt; ++di) {
    const int ti = t + R.dtv[di];
    const bool ok = ti >= 0 && ti < d.T;
    const long long pos0 = (long long)(bt - t + ti) * d.Fin + in_lo;
    float4* dst = reinterpret_cast<float4*>(rows + (size_t)di * R.nin_max * Ct);
    for (int i = threadIdx.x; i < nin * c4; i += blockDim.x) {
      const int f = i / c4, q = i - f * c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok)
        v = q < c04 ? __ldg(reinterpret_cast<const float4*>(d.src0 + (pos0 + f) * d.C0) + q)
                    : __ldg(reinterpret_cast<const float4*>(d.src1 + (pos0 + f) * d.C1) + (q - c04));
      dst[i] = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int fo0 = f_lo + warp * 4; fo0 < f_hi; fo0 += nwarp * 4) {
    float acc[4][NCO];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int co = 0; co < NCO; ++co) acc[r][co] = 0.f;
    for (int tap = 0; tap < d.ntaps; ++tap) {
      const float* xr = rows + (size_t)R.tap_dti[tap] * R.nin_max * Ct;
      const float* wt = wsm + (size_t)tap * Ct * NCO;
      int fi[4];
      bool fok[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        fi[r] = (fo0 + r) * d.sf + d.df[tap];
        fok[r] = fo0 + r < f_hi && fi[r] >= 0 && fi[r] < d.Fin;
        fi[r] -= in_lo;
      }
      for (int c = lane * 4; c < Ct; c += 128) {
        float wv[4 * NCO];
#pragma unroll
        for (int i = 0; i < NCO; ++i) {
          const float4 w4 = *reinterpret_cast<const float4*>(wt + c * NCO + 4 * i);
          wv[4 * i] = w4.x;
          wv[4 * i + 1] = w4.y;
          wv[4 * i + 2] = w4.z;
          wv[4 * i + 3] = w4.w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (!fok[r]) continue;                                       // warp-uniform
          const float4 v = *reinterpret_cast<const float4*>(xr + (size_t)fi[r] * Ct + c);
          const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int co = 0; co < NCO; ++co) acc[r][co] = fmaf(xv[e], wv[e * NCO + co], acc[r][co]);
        }
      }
    }
    float mine = 0.f;                                                  // lane r * NCO + co keeps sum (r, co)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int co = 0; co < NCO; ++co) {
        float v = acc[r][co];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == r * NCO + co) mine = v;
      }
    const int r = lane / NCO, co = lane - r * NCO;
    if (lane < 4 * NCO && fo0 + r < f_hi && co < d.Cout) {
      const float v = apply_act(mine + (d.bias ? __ldg(d.bias + co) : 0.f), d.act, d.act_param);
      d.dst[(((long long)bt * d.dstF) + d.dst_f0 + (long long)(fo0 + r) * d.dst_fstep) * d.Cout + co] = v;
    }
  }
}

// returns 1 when the launch was made, 0 when the shape does not fit (caller falls back to the tiled kernel)
template <int NCO>
static int launch_conv_rows(const ConvParams& P, cudaStream_t s, int* rc) {
  ConvRowsParams R;
  R.c = P;
  R.ndt = 0;
  const se_conv_desc& d = P.d;
  for (int tap = 0; tap < d.ntaps; ++tap) {
    int di = 0;
    while (di < R.ndt && R.dtv[di] != d.dt[tap]) ++di;
    if (di == R.ndt) {
      if (R.ndt == kRowsMaxDt) return 0;
      R.dtv[R.ndt++] = d.dt[tap];
    }
    R.tap_dti[tap] = di;
  }
  R.dfmin = R.dfmax = d.df[0];
  for (int tap = 1; tap < d.ntaps; ++tap) {
    R.dfmin = d.df[tap] < R.dfmin ? d.df[tap] : R.dfmin;
    R.dfmax = d.df[tap] > R.dfmax ? d.df[tap] : R.dfmax;
  }
  // output columns per CTA: a multiple of 32 (8 warps x 4 columns) whose staged input fits ~64 KB, i.e. 3 CTAs per SM
  R.fc = ((d.Fout + 31) / 32) * 32;
  auto nin_of = [&](int fc) { return (fc - 1) * d.sf + (R.dfmax - R.dfmin) + 1; };
  while (R.fc > 32 && (size_t)R.ndt * nin_of(R.fc) * P.Ctot * sizeof(float) > 64 * 1024) R.fc -= 32;
  R.nin_max = nin_of(R.fc) < d.Fin ? nin_of(R.fc) : d.Fin;
  const size_t smem = ((size_t)R.ndt * R.nin_max * P.Ctot + (size_t)P.K * NCO) * sizeof(float);
  if (smem > 200 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(conv_rows_kernel<NCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("se_conv_gemm (rows): %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    *rc = SE_ERR_CUDA;
    return 1;
  }
  conv_rows_kernel<NCO><<<dim3((unsigned)(d.B * d.T), (unsigned)((d.Fout + R.fc - 1) / R.fc)), 256, smem, s>>>(R);
  *rc = SE_OK;
  return 1;
}

template <int BM, int BN, int TM, int TN>
static void launch_conv(const ConvParams& P, bool aligned, cudaStream_t s) {
  dim3 grid(ceil_div(P.M, BM), ceil_div(P.d.Cout, BN));
  if (aligned)
    conv_gemm_kernel<BM, BN, TM, TN, true><<<grid, kGemmThreads, 0, s>>>(P);
  else
    conv_gemm_kernel<BM, BN, TM, TN, false><<<grid, kGemmThreads, 0, s>>>(P);
}

}  // namespace se

using namespace se;

extern "C" int se_conv_gemm(const se_conv_desc* desc, se_stream_t stream) {
  SE_REQUIRE(desc != nullptr, "se_conv_gemm: null descriptor");
  ConvParams P;
  P.d = *desc;
  const se_conv_desc& d = P.d;
  SE_REQUIRE(d.src0 && d.W && d.dst, "se_conv_gemm: null pointer");
  SE_REQUIRE(d.ntaps >= 1 && d.ntaps <= SE_MAX_TAPS, "se_conv_gemm: ntaps=%d", d.ntaps);
  SE_REQUIRE(d.C0 > 0 && d.C1 >= 0 && (d.C1 == 0 || d.src1), "se_conv_gemm: bad channel split %d+%d", d.C0, d.C1);
  SE_REQUIRE(d.B > 0 && d.T > 0 && d.Fin > 0 && d.Fout > 0 && d.Cout > 0, "se_conv_gemm: bad shape");
  SE_REQUIRE(d.ldw >= d.Cout && (d.ldw & 3) == 0 && ((((uintptr_t)d.W) & 15) == 0),
             "se_conv_gemm: W must be 16-byte aligned with ldw %% 4 == 0 (ldw=%d)", d.ldw);
  P.Ctot = d.C0 + d.C1;
  P.K = d.ntaps * P.Ctot;
  const long long M = (long long)d.B * d.T * d.Fout;
  SE_REQUIRE(M < (1ll << 31), "se_conv_gemm: M too large");
  P.M = (int)M;
  const bool aligned = (d.C0 % BK == 0) && (d.C1 % BK == 0) && ((((uintptr_t)d.src0) & 15) == 0) &&
                       (d.C1 == 0 || (((uintptr_t)d.src1) & 15) == 0);
  cudaStream_t s = (cudaStream_t)stream;
  // Cout <= 4 with float4-addressable channels: frame-per-CTA kernel with the input frames staged in shared memory
  const bool narrow = d.Cout <= 4 && (d.C0 & 3) == 0 && (d.C1 & 3) == 0 && ((((uintptr_t)d.src0) & 15) == 0) &&
                      (d.C1 == 0 || (((uintptr_t)d.src1) & 15) == 0) && d.Fout >= 16;
  int rcn = SE_OK;
  if (narrow && (d.Cout <= 2 ? launch_conv_rows<2>(P, s, &rcn) : launch_conv_rows<4>(P, s, &rcn))) {
    if (rcn) return rcn;
  } else if (d.Cout > 64)
    launch_conv<128, 128, 8, 8>(P, aligned, s);
  else if (d.Cout > 32)
    launch_conv<128, 64, 8, 4>(P, aligned, s);
  else if (d.Cout > 16)
    launch_conv<256, 32, 8, 4>(P, aligned, s);
  else
    launch_conv<256, 16, 4, 4>(P, aligned, s);
  int rc = check_launch("se_conv_gemm");
  if (rc) return rc;
  if (d.fill_f >= 0) {
    SE_REQUIRE(d.fill && d.fill_f < d.dstF, "se_conv_gemm: bad fill column");
    const long long rows = (long long)d.B * d.T;
    const long long n = rows * d.Cout;
    fill_col_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, s>>>(d.dst, rows, d.dstF, d.Cout, d.fill_f, d.fill,
                                                                  d.act, d.act_param);
    rc = check_launch("se_conv_gemm(fill)");
  }
  return rc;
}

extern "C" int se_fill_column(float* dst, long long rows, int dstF, int Cout, int fill_f, const float* fill, int act,
                              float act_param, se_stream_t stream) {
  SE_REQUIRE(dst && fill && rows > 0 && fill_f >= 0 && fill_f < dstF && Cout > 0, "se_fill_column: bad arguments");
  const long long n = rows * Cout;
  fill_col_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, rows, dstF, Cout, fill_f, fill,
                                                                                   act, act_param);
  return check_launch("se_fill_column");
}

extern "C" int se_conv_in1(const float* src, int B, int T, int Fin, const float* W, const float* bias, int Cout,
                           int act, float* dst, int Fout, se_stream_t stream) {
  SE_REQUIRE(src && W && dst, "se_conv_in1: null pointer");
  SE_REQUIRE(Cout > 0 && Cout <= 64 && (Cout & 3) == 0, "se_conv_in1: Cout=%d (<=64, %%4)", Cout);
  SE_REQUIRE(Fout == (Fin - 3) / 2 + 1 && Fout > 0, "se_conv_in1: Fout=%d for Fin=%d", Fout, Fin);
  const long long total = (long long)B * T * Fout * (Cout / 4);
  const int blocks = (int)min((long long)148 * 16, ceil_div_ll(total, 256));
  conv_in1_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, B, T, Fin, W, bias, Cout, act, dst, Fout);
  return check_launch("se_conv_in1");
}

extern "C" int se_deconv_out1(const float* src0, const float* src1, int C0, int C1, int B, int T, int Fin,
                              const float* W, float bias, int act, float* dst, se_stream_t stream) {
  SE_REQUIRE(src0 && W && dst, "se_deconv_out1: null pointer");
  SE_REQUIRE(C0 > 0 && (C0 & 3) == 0 && C1 >= 0 && (C1 & 3) == 0 && (C1 == 0 || src1), "se_deconv_out1: channels");
  static const bool legacy = []() { const char* e = getenv("SE_DECONV_OUT1_LEGACY"); return e && e[0] == '1'; }();
  if (!legacy && Fin <= 256 && (C0 + C1 == 32 || C0 + C1 == 64) && (long long)B * T < (1ll << 30)) {
    const int R = 256 / Fin;                    // frames per CTA: R x Fin threads (240 at Fin = 80)
    const int grid = B * ceil_div(T, R);
    if (C0 + C1 == 32)
      deconv_out1_rows_kernel<32><<<grid, R * Fin, 0, (cudaStream_t)stream>>>(src0, src1, C0, C1, B, T, Fin, R, W, bias, act, dst);
    else
      deconv_out1_rows_kernel<64><<<grid, R * Fin, 0, (cudaStream_t)stream>>>(src0, src1, C0, C1, B, T, Fin, R, W, bias, act, dst);
    return check_launch("se_deconv_out1");
  }
  const long long total = (long long)B * T * (2 * Fin + 1);
  const int blocks = (int)min((long long)148 * 16, ceil_div_ll(total, 256));
  const int smem = 6 * (C0 + C1) * 4;
  deconv_out1_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(src0, src1, C0, C1, B, T, Fin, W, bias, act, dst);
  return check_launch("se_deconv_out1");
}
