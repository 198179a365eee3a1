// DSP front/back end of the decode loop on sm_100a:
//   se_rms_scale  (a1)      c = sqrt(N / sum x^2)
//   se_stft       (a3+a4)   reflect-pad + framing + Hann + rFFT + |X|^p / RI split
//   se_istft      (a7..a9)  recombination prologue + irFFT + window + OLA + envelope + 1/c
// Both transforms are HBM-bound: every audio sample and spectrum bin is touched once
// (halo frames of neighbouring CTAs hit L2).  See DESIGN.md for the byte accounting.
#include <mutex>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "fft.cuh"
#include "fft_thread.cuh"

namespace se {

constexpr int kFramesPerCta = 32;   // STFT: frames per CTA (8 warps x 4)
constexpr int kHopBlocksPerCta = 32;  // iSTFT: hop-sized output blocks per CTA
constexpr int kDspThreads = 256;

// -------------------------------------------------------------------------------------------
// a1: RMS scale
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) rms_scale_kernel(const float* __restrict__ wav, long long stride, int N,
                                                       const int* __restrict__ lengths, int reciprocal,
                                                       float* __restrict__ c, float* __restrict__ inv_c) {
  const float* x = wav + (long long)blockIdx.x * stride;
  if (lengths) N = max(1, min(N, __ldg(lengths + blockIdx.x)));   // tail-padded batch: this clip's own sample count
  double acc = 0.0;
  const bool vec = ((((uintptr_t)x) & 15) == 0);
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int n4 = N >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 v = __ldg(x4 + i);
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < N; i += blockDim.x) acc += (double)x[i] * x[i];
  } else {
    for (int i = threadIdx.x; i < N; i += blockDim.x) acc += (double)x[i] * x[i];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[16];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    if (!reciprocal) {
      const double cc = sqrt((double)N / s);  // crn_decode.py:39: x * c ... y / c
      c[blockIdx.x] = (float)cc;
      inv_c[blockIdx.x] = (float)(1.0 / cc);
    } else {
      const double cc = sqrt(s / (double)N);  // G2Net_new/com_decode.py:43-44: x / c ... y * c
      c[blockIdx.x] = (float)(1.0 / cc);
      inv_c[blockIdx.x] = (float)cc;
    }
  }
}

// -------------------------------------------------------------------------------------------
// a3+a4: STFT
// -------------------------------------------------------------------------------------------
struct StftParams {
  const float* wav;
  long long wav_stride;
  int B, N;
  const float* scale;
  int win, hop, T;
  float *mag, *re, *im;
  long long msb, mst, msf;
  long long sb, st, sf;
  float p_mag, p_ri;
  const float2* twiddles;   // [32][2R + 5], see WarpFFT<R>::fill_table
  const int* lengths;       // optional [B]: samples of every clip of a tail-padded batch (N = row length = maximum)
};

__device__ __forceinline__ float pow_pos(float m, float p) {
  // m >= 0.  p = 1, 0.5, 2 are the exponents the decode scripts use.
  if (p == 1.0f) return m;
  if (p == 0.5f) return sqrtf(m);
  if (p == 2.0f) return m * m;
  return m > 0.0f ? powf(m, p) : 0.0f;
}
__device__ __forceinline__ float pow_scale(float m, float e) {
  // m^e with the convention 0^e = 0 (the reference multiplies a zero magnitude by cos/sin)
  if (e == 0.0f) return 1.0f;
  if (m <= 0.0f) return 0.0f;
  if (e == 1.0f) return m;
  if (e == -0.5f) return rsqrtf(m);
  return powf(m, e);
}

template <int R>
__global__ void __launch_bounds__(kDspThreads) stft_kernel(StftParams p) {
  constexpr int NFFT = 64 * R, N2 = 32 * R, F = N2 + 1;
  constexpr int FT = kFramesPerCta;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wtab = reinterpret_cast<float*>(smem_raw);                    // [NFFT]
  float2* spec = reinterpret_cast<float2*>(wtab + NFFT);               // [FT][R][33]
  float* nyq_s = reinterpret_cast<float*>(spec + FT * R * 33);         // [FT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(nyq_s + FT);             // 8-byte aligned (FT even)
  float* tile = reinterpret_cast<float*>(bar + 2);                     // 16-byte aligned

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FT;
  const int nf = min(FT, p.T - t0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = p.wav + (long long)b * p.wav_stride;
  const int start = t0 * p.hop - NFFT / 2;
  const int tile_len = (nf - 1) * p.hop + NFFT;
  // per-clip length of a tail-padded batch: the clip ends (and is reflected) at ITS last sample and has 1 + Nb / hop
  // frames, exactly as when it is decoded alone (crn_decode_vb.py:31-37 loops one file at a time); later frames are 0
  const int Nb = p.lengths ? max(NFFT, min(p.N, __ldg(p.lengths + b))) : p.N;
  const int Tb = 1 + Nb / p.hop;

  // --- stage the audio span of this frame tile --------------------------------------------
  const bool interior = (start >= 0) && (start + tile_len <= Nb) && ((((uintptr_t)(x + start)) & 15) == 0) &&
                        ((tile_len & 3) == 0);
  if (interior) {
    // one bulk (TMA-engine) copy, completion on an mbarrier
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, (unsigned)tile_len * 4u);
      bulk_g2s(tile, x + start, (unsigned)tile_len * 4u, bar);
    }
  } else {
    // clip edges: reflect padding cannot be described to the copy engine
    for (int i = tid; i < tile_len; i += kDspThreads) {
      int j = start + i;
      if (j < 0) j = -j;
      if (j >= Nb) j = 2 * (Nb - 1) - j;
      j = max(0, min(j, Nb - 1));
      tile[i] = __ldg(x + j);
    }
  }
  // window (periodic Hann of length win, centred in NFFT) times the clip's RMS scale
  {
    const float sc = p.scale ? __ldg(p.scale + b) : 1.0f;
    const int left = (NFFT - p.win) / 2;
    for (int i = tid; i < NFFT; i += kDspThreads) {
      const int n = i - left;
      float w = 0.0f;
      if (n >= 0 && n < p.win) {
        const float sn = sinpif((float)n / (float)p.win);  // hann = sin^2: no cancellation near the ends
        w = sn * sn;
      }
      wtab[i] = w * sc;
    }
  }
  WarpFFT<R> fft;
  fft.init(lane, p.twiddles);
  if (interior) mbar_wait(bar, 0);
  __syncthreads();

  // --- one warp per frame -------------------------------------------------------------------
  for (int i = warp; i < nf; i += kDspThreads / 32) {
    const float2* fr = reinterpret_cast<const float2*>(tile + i * p.hop);
    const float2* w2 = reinterpret_cast<const float2*>(wtab);
    float2 z[R];
#pragma unroll
    for (int m1 = 0; m1 < R; ++m1) {
      const int m = 32 * m1 + lane;
      const float2 v = fr[m], w = w2[m];
      z[m1] = make_float2(v.x * w.x, v.y * w.y);
    }
    const float nyq = fft.forward(z);
    const int k2 = bitrev5(lane);
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) spec[(i * R + k1) * 33 + k2] = z[k1];
    if (lane == 0) nyq_s[i] = nyq;
  }
  __syncthreads();

  // --- epilogue: feature split + layout-aware store ------------------------------------------
  const bool time_major = p.re ? (p.sf <= p.st) : (p.msf <= p.mst);
  const int total = nf * F;
  for (int idx = tid; idx < total; idx += kDspThreads) {
    int i, k;
    if (time_major) {
      i = idx / F;
      k = idx - i * F;
    } else {
      k = idx / nf;
      i = idx - k * nf;
    }
    float2 X;
    if (k == N2) {
      X = make_float2(nyq_s[i], 0.0f);
    } else {
      const int k2 = k / R, k1 = k - k2 * R;
      X = spec[(i * R + k1) * 33 + k2];
    }
    if (t0 + i >= Tb) X = make_float2(0.0f, 0.0f);     // frame past this clip's end (tail-padded batch)
    const float m = sqrtf(X.x * X.x + X.y * X.y);
    if (p.mag)
      p.mag[(long long)b * p.msb + (long long)(t0 + i) * p.mst + (long long)k * p.msf] = pow_pos(m, p.p_mag);
    if (p.re) {
      const long long off = (long long)b * p.sb + (long long)(t0 + i) * p.st + (long long)k * p.sf;
      const float s = pow_scale(m, p.p_ri - 1.0f);
      p.re[off] = X.x * s;
      p.im[off] = X.y * s;
    }
  }
}

// -------------------------------------------------------------------------------------------
// a7+a8+a9: iSTFT
// -------------------------------------------------------------------------------------------
struct IstftParams {
  int mode;
  const float *a_re, *a_im;
  long long a_sb, a_st, a_sf;
  const float *b_re, *b_im;
  long long b_sb, b_st, b_sf;
  float inv_p, p_x;
  int B, T, win, hop;
  const float* out_scale;
  float* out;
  long long out_stride;
  int L;
  int nf_max;
  const float2* twiddles;   // [32][2R + 5], see WarpFFT<R>::fill_table
  const int* lengths;       // optional [B]: clip b has 1 + lengths[b] / hop frames and min(L, lengths[b]) output samples
  int ob;                   // istft2_kernel: hop-sized output blocks per CTA (chosen so that nf_max is a multiple of 16)
};

template <int R>
__global__ void __launch_bounds__(kDspThreads) istft_kernel(IstftParams p) {
  constexpr int NFFT = 64 * R, N2 = 32 * R, F = N2 + 1;
  constexpr int FRS = R * 66;  // floats per frame buffer: spectrum [R][33] float2 aliases NFFT samples
  constexpr int OB = kHopBlocksPerCta;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wtab = reinterpret_cast<float*>(smem_raw);  // [NFFT]
  float* nyq_s = wtab + NFFT;                        // [nf_max]
  float* buf = nyq_s + ((p.nf_max + 3) & ~3);        // [nf_max][FRS]

  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * OB * p.hop;  // first output sample of this CTA
  const int s0 = n0 + NFFT / 2;            // same, in overlap-add coordinates
  const int s1 = s0 + OB * p.hop;
  const int q = s0 - NFFT;
  const int t_lo = q < 0 ? 0 : q / p.hop + 1;
  // tail-padded batch: only this clip's own frames overlap-add (and enter the window envelope), output past its end is 0
  const int Nb = p.lengths ? max(NFFT, __ldg(p.lengths + b)) : 0x7fffffff;
  const int Tb = p.lengths ? min(p.T, 1 + Nb / p.hop) : p.T;
  const int Lb = min(p.L, Nb);
  const int t_hi = min(Tb - 1, (s1 - 1) / p.hop);
  const int nf = t_hi - t_lo + 1;

  {
    const int left = (NFFT - p.win) / 2;
    for (int i = tid; i < NFFT; i += kDspThreads) {
      const int n = i - left;
      float w = 0.0f;
      if (n >= 0 && n < p.win) {
        const float sn = sinpif((float)n / (float)p.win);  // hann = sin^2: no cancellation near the ends
        w = sn * sn;
      }
      wtab[i] = w;
    }
  }

  // --- stage 1: recombination prologue, spectrum of every contributing frame -> smem --------
  if (nf > 0) {
    const bool time_major = (p.a_sf <= p.a_st);
    const int total = nf * F;
#pragma unroll 4   // independent global loads of 4 bins in flight per thread (the loop was long-scoreboard bound)
    for (int idx = tid; idx < total; idx += kDspThreads) {
      int i, k;
      if (time_major) {
        i = idx / F;
        k = idx - i * F;
      } else {
        k = idx / nf;
        i = idx - k * nf;
      }
      const int t = t_lo + i;
      const long long oa = (long long)b * p.a_sb + (long long)t * p.a_st + (long long)k * p.a_sf;
      float2 Y;
      if (p.mode == SE_ISTFT_SPEC) {
        Y = make_float2(__ldg(p.a_re + oa), __ldg(p.a_im + oa));
      } else if (p.mode == SE_ISTFT_RI_DECOMP) {
        const float2 A = make_float2(__ldg(p.a_re + oa), __ldg(p.a_im + oa));
        const float s = pow_scale(sqrtf(A.x * A.x + A.y * A.y), p.inv_p - 1.0f);
        Y = make_float2(A.x * s, A.y * s);
      } else {
        const long long ob = (long long)b * p.b_sb + (long long)t * p.b_st + (long long)k * p.b_sf;
        const float2 X = make_float2(__ldg(p.b_re + ob), __ldg(p.b_im + ob));
        const float m = sqrtf(X.x * X.x + X.y * X.y);
        if (p.mode == SE_ISTFT_MAG_PHASE) {
          // est^(1/p) * exp(j angle(X));  angle(0) = 0
          const float g = pow_pos(__ldg(p.a_re + oa), p.inv_p);
          const float2 ph = m > 0.0f ? make_float2(X.x / m, X.y / m) : make_float2(1.0f, 0.0f);
          Y = make_float2(g * ph.x, g * ph.y);
        } else {  // SE_ISTFT_CMASK
          const float sx = pow_scale(m, p.p_x - 1.0f);
          const float2 Xc = make_float2(X.x * sx, X.y * sx);
          const float2 A = make_float2(__ldg(p.a_re + oa), __ldg(p.a_im + oa));
          const float2 C = make_float2(A.x * Xc.x - A.y * Xc.y, A.y * Xc.x + A.x * Xc.y);
          const float s = pow_scale(sqrtf(C.x * C.x + C.y * C.y), p.inv_p - 1.0f);
          Y = make_float2(C.x * s, C.y * s);
        }
      }
      if (k == N2) {
        nyq_s[i] = Y.x;
      } else {
        const int k2 = k / R, k1 = k - k2 * R;
        reinterpret_cast<float2*>(buf + (size_t)i * FRS)[k1 * 33 + k2] = Y;
      }
    }
  }
  WarpFFT<R> fft;
  fft.init(lane, p.twiddles);
  __syncthreads();

  // --- stage 2: one warp per frame: irFFT, window, time-domain frame back into its buffer ---
  for (int i = warp; i < nf; i += kDspThreads / 32) {
    float* fb = buf + (size_t)i * FRS;
    const float2* sp = reinterpret_cast<const float2*>(fb);
    float2 x[R];
    const int k2 = bitrev5(lane);
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) x[k1] = sp[k1 * 33 + k2];
    const float nyq = nyq_s[i];
    __syncwarp();
    fft.inverse(x, nyq);
    __syncwarp();
    const float2* w2 = reinterpret_cast<const float2*>(wtab);
    float2* fo = reinterpret_cast<float2*>(fb);
#pragma unroll
    for (int m1 = 0; m1 < R; ++m1) {
      const int m = 32 * m1 + lane;
      const float2 w = w2[m];
      fo[m] = make_float2(x[m1].x * w.x, x[m1].y * w.y);
    }
  }
  __syncthreads();

  // --- stage 3: gather overlap-add, envelope, trim, scale ------------------------------------
  const float osc = p.out_scale ? __ldg(p.out_scale + b) : 1.0f;
  float* out = p.out + (long long)b * p.out_stride;
  const int span = OB * p.hop;
  for (int j = tid; j < span; j += kDspThreads) {
    const int n = n0 + j;
    if (n >= p.L) break;
    if (n >= Lb) {
      out[n] = 0.0f;
      continue;
    }
    const int s = s0 + j;
    const int first = s - NFFT + 1;
    int ta = first <= 0 ? 0 : (first + p.hop - 1) / p.hop;
    ta = max(ta, t_lo);
    const int tb = min(t_hi, s / p.hop);
    float acc = 0.0f, env = 0.0f;
    for (int t = ta; t <= tb; ++t) {
      const int off = s - t * p.hop;
      acc += buf[(size_t)(t - t_lo) * FRS + off];
      const float w = wtab[off];
      env += w * w;
    }
    if (env > FLT_MIN) acc /= env;
    out[n] = acc * osc;
  }
}

// -------------------------------------------------------------------------------------------
// Second-generation STFT / iSTFT: per-thread radix-16 x radix-N2 FFT through shared memory (fft_thread.cuh) instead of
// the five-stage warp-shuffle network.  16 threads per frame (two frames per warp), M = NFFT / 2 = 16 * N2 packed complex
// points: pass 1 = 16-point DFTs in registers by N2 of the threads, twiddle, exchange through the frame's own buffer,
// pass 2 = N2-point DFTs by all 16.  Same staging, feature split, recombination and overlap-add as the kernels above
// (same StftParams / IstftParams), so every parity test of tests/test_gpu_dsp.py applies unchanged.
// -------------------------------------------------------------------------------------------
// Frames (STFT) / hop blocks (iSTFT) per CTA are parameters: smaller tiles = less shared memory = more CTAs per SM.
// Where the kernels stand (profiles/ncu_dsp2_r02_summary.txt): stft2_kernel issues in 65 % of the cycles at 33 % occupancy,
// ~970 warp instructions per 320-point frame of which the FFT is now ~250 -- the layout-generic feature split (index
// division, three 64-bit multiplies per store) and the per-CTA table generation are what is left.

template <int R>
__device__ __forceinline__ void dsp2_tables(float2* tw1, float2* tw2, int tid) {
  constexpr int M = 32 * R, N2 = 2 * R;
  for (int i = tid; i < 16 * N2; i += kDspThreads) {
    const int k1 = i / N2, n2 = i - k1 * N2;
    float sn, cs;
    sincospif(2.0f * (float)((n2 * k1) % M) / (float)M, &sn, &cs);     // W_M^(n2 k1): (cos, sin) of the positive angle
    tw1[i] = make_float2(cs, sn);
  }
  for (int k = tid; k <= M; k += kDspThreads) {
    float sn, cs;
    sincospif((float)k / (float)M, &sn, &cs);                          // W_2M^k
    tw2[k] = make_float2(cs, sn);
  }
}

template <int R, int FT>
__global__ void __launch_bounds__(kDspThreads, 3) stft2_kernel(StftParams p) {
  constexpr int NFFT = 64 * R, M = 32 * R, N2 = 2 * R, F = M + 1;
  constexpr int FS = M + 16;          // FS: float2 per frame buffer (exchange rows padded to N2 + 1)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wtab = reinterpret_cast<float*>(smem_raw);                    // [NFFT]
  float2* tw1 = reinterpret_cast<float2*>(wtab + NFFT);                // [16][N2]
  float2* tw2 = tw1 + 16 * N2;                                         // [M + 1]  (+1 pad keeps 16-byte alignment below)
  float2* zb = tw2 + M + 2;                                            // [FT][FS]
  uint64_t* bar = reinterpret_cast<uint64_t*>(zb + FT * FS);
  float* tile = reinterpret_cast<float*>(bar + 2);

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FT;
  const int nf = min(FT, p.T - t0);
  const int tid = threadIdx.x;
  const float* x = p.wav + (long long)b * p.wav_stride;
  const int start = t0 * p.hop - NFFT / 2;
  const int tile_len = (nf - 1) * p.hop + NFFT;
  const int Nb = p.lengths ? max(NFFT, min(p.N, __ldg(p.lengths + b))) : p.N;
  const int Tb = 1 + Nb / p.hop;

  const bool interior = (start >= 0) && (start + tile_len <= Nb) && ((((uintptr_t)(x + start)) & 15) == 0) &&
                        ((tile_len & 3) == 0);
  if (interior) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, (unsigned)tile_len * 4u);
      bulk_g2s(tile, x + start, (unsigned)tile_len * 4u, bar);
    }
  } else {
    for (int i = tid; i < tile_len; i += kDspThreads) {
      int j = start + i;
      if (j < 0) j = -j;
      if (j >= Nb) j = 2 * (Nb - 1) - j;
      j = max(0, min(j, Nb - 1));
      tile[i] = __ldg(x + j);
    }
  }
  {
    const float sc = p.scale ? __ldg(p.scale + b) : 1.0f;
    const int left = (NFFT - p.win) / 2;
    for (int i = tid; i < NFFT; i += kDspThreads) {
      const int n = i - left;
      float w = 0.0f;
      if (n >= 0 && n < p.win) {
        const float sn = sinpif((float)n / (float)p.win);
        w = sn * sn;
      }
      wtab[i] = w * sc;
    }
  }
  dsp2_tables<R>(tw1, tw2, tid);
  if (interior) mbar_wait(bar, 0);
  __syncthreads();

  // --- 16 threads per frame ---------------------------------------------------------------------------
  const int grp = tid >> 4, l16 = tid & 15;
  const float2* w2 = reinterpret_cast<const float2*>(wtab);
#pragma unroll 1
  for (int pass = 0; pass < FT / 16; ++pass) {
    const int i = pass * 16 + grp;
    const bool active = i < nf;
    float2* fb = zb + i * FS;
    if (active && l16 < N2) {                       // pass 1: thread n2 = l16
      const float2* fr = reinterpret_cast<const float2*>(tile + i * p.hop);
      float2 a[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int n = N2 * n1 + l16;
        const float2 v = fr[n], w = w2[n];
        a[n1] = make_float2(v.x * w.x, v.y * w.y);
      }
      ft::Dft<16, false>::run(a);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const float2 t = tw1[k1 * N2 + l16];
        fb[k1 * (N2 + 1) + l16] = ft::twmul<false>(a[k1], t.x, t.y);
      }
    }
    __syncwarp();
    float2 bq[N2];
    if (active) {                                    // pass 2: thread k1 = l16
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) bq[n2] = fb[l16 * (N2 + 1) + n2];
    }
    __syncwarp();
    if (active) {
      ft::Dft<N2, false>::run(bq);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) fb[l16 + 16 * k2] = bq[k2];
    }
  }
  __syncthreads();

  // --- epilogue: real-FFT split + feature split + layout-aware store ------------------------------------
  const bool time_major = p.re ? (p.sf <= p.st) : (p.msf <= p.mst);
  const bool ri_pair = p.re && p.im == p.re + 1 && p.sf == 2 && (p.st & 1) == 0 && (p.sb & 1) == 0 && ((uintptr_t)p.re & 7) == 0;
  const int total = nf * F;
#pragma unroll 4
  for (int idx = tid; idx < total; idx += kDspThreads) {
    int i, k;
    if (time_major) {
      i = idx / F;
      k = idx - i * F;
    } else {
      k = idx / nf;
      i = idx - k * nf;
    }
    const float2* fb = zb + i * FS;
    const float2 tw = tw2[k];
    float2 X = ft::rfft_split(fb[k == M ? 0 : k], fb[k == 0 ? 0 : M - k], tw.x, tw.y);
    if (k == 0 || k == M) X.y = 0.0f;
    if (t0 + i >= Tb) X = make_float2(0.0f, 0.0f);
    const float m = sqrtf(X.x * X.x + X.y * X.y);
    if (p.mag)
      p.mag[(long long)b * p.msb + (long long)(t0 + i) * p.mst + (long long)k * p.msf] = pow_pos(m, p.p_mag);
    if (p.re) {
      const long long off = (long long)b * p.sb + (long long)(t0 + i) * p.st + (long long)k * p.sf;
      const float s = pow_scale(m, p.p_ri - 1.0f);
      if (ri_pair) {
        *reinterpret_cast<float2*>(p.re + off) = make_float2(X.x * s, X.y * s);
      } else {
        p.re[off] = X.x * s;
        p.im[off] = X.y * s;
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(kDspThreads, 3) istft2_kernel(IstftParams p) {
  constexpr int NFFT = 64 * R, M = 32 * R, N2 = 2 * R, F = M + 1;
  const int OB = p.ob;
  constexpr int FS = M + 16;       // float2 per frame buffer: spectrum [M + 1] -> exchange [16][N2 + 1] -> NFFT samples
  constexpr int FRS = 2 * FS;      // the same in floats
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wtab = reinterpret_cast<float*>(smem_raw);                    // [NFFT]   window / NFFT
  float2* tw1 = reinterpret_cast<float2*>(wtab + NFFT);                // [16][N2]
  float2* tw2 = tw1 + 16 * N2;                                         // [M + 1] (+1 pad)
  float* envtab = reinterpret_cast<float*>(tw2 + M + 2);               // [512]: 1 / sum_m w^2[r + m hop], r < hop <= NFFT
  float* buf = envtab + 512;                                           // [nf_max][FRS]

  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * OB * p.hop;
  const int s0 = n0 + NFFT / 2;
  const int s1 = s0 + OB * p.hop;
  const int q = s0 - NFFT;
  const int t_lo = q < 0 ? 0 : q / p.hop + 1;
  const int Nb = p.lengths ? max(NFFT, __ldg(p.lengths + b)) : 0x7fffffff;
  const int Tb = p.lengths ? min(p.T, 1 + Nb / p.hop) : p.T;
  const int Lb = min(p.L, Nb);
  const int t_hi = min(Tb - 1, (s1 - 1) / p.hop);
  const int nf = t_hi - t_lo + 1;

  {
    const int left = (NFFT - p.win) / 2;
    for (int i = tid; i < NFFT; i += kDspThreads) {
      const int n = i - left;
      float w = 0.0f;
      if (n >= 0 && n < p.win) {
        const float sn = sinpif((float)n / (float)p.win);
        w = sn * sn;
      }
      wtab[i] = w;
    }
  }
  dsp2_tables<R>(tw1, tw2, tid);

  // --- stage 1: recombination prologue, bins 0..M of every contributing frame -> smem ---------------------
  // (re, im) pairs that are interleaved in memory (complex64 tensors: im = re + 1, bin stride 2) are read with one
  // 8-byte load; 8 bins per thread are in flight (the stage is bound by memory-level parallelism, not by issue)
  if (nf > 0) {
    const bool time_major = (p.a_sf <= p.a_st);
    const bool a_pair = p.a_im == p.a_re + 1 && p.a_sf == 2 && (p.a_st & 1) == 0 && (p.a_sb & 1) == 0 &&
                        ((uintptr_t)p.a_re & 7) == 0;
    const bool b_pair = p.b_re && p.b_im == p.b_re + 1 && p.b_sf == 2 && (p.b_st & 1) == 0 && (p.b_sb & 1) == 0 &&
                        ((uintptr_t)p.b_re & 7) == 0;
    const int total = nf * F;
#pragma unroll 8
    for (int idx = tid; idx < total; idx += kDspThreads) {
      int i, k;
      if (time_major) {
        i = idx / F;
        k = idx - i * F;
      } else {
        k = idx / nf;
        i = idx - k * nf;
      }
      const int t = t_lo + i;
      const long long oa = (long long)b * p.a_sb + (long long)t * p.a_st + (long long)k * p.a_sf;
      float2 A = make_float2(0.f, 0.f), X = make_float2(0.f, 0.f);
      if (p.mode == SE_ISTFT_MAG_PHASE)
        A.x = __ldg(p.a_re + oa);
      else if (a_pair)
        A = __ldg(reinterpret_cast<const float2*>(p.a_re + oa));
      else
        A = make_float2(__ldg(p.a_re + oa), __ldg(p.a_im + oa));
      if (p.mode >= SE_ISTFT_MAG_PHASE) {
        const long long ob = (long long)b * p.b_sb + (long long)t * p.b_st + (long long)k * p.b_sf;
        X = b_pair ? __ldg(reinterpret_cast<const float2*>(p.b_re + ob)) : make_float2(__ldg(p.b_re + ob), __ldg(p.b_im + ob));
      }
      float2 Y;
      if (p.mode == SE_ISTFT_SPEC) {
        Y = A;
      } else if (p.mode == SE_ISTFT_RI_DECOMP) {
        const float s = pow_scale(sqrtf(A.x * A.x + A.y * A.y), p.inv_p - 1.0f);
        Y = make_float2(A.x * s, A.y * s);
      } else {
        const float m = sqrtf(X.x * X.x + X.y * X.y);
        if (p.mode == SE_ISTFT_MAG_PHASE) {
          const float g = pow_pos(A.x, p.inv_p);
          const float2 ph = m > 0.0f ? make_float2(X.x / m, X.y / m) : make_float2(1.0f, 0.0f);
          Y = make_float2(g * ph.x, g * ph.y);
        } else {  // SE_ISTFT_CMASK
          const float sx = pow_scale(m, p.p_x - 1.0f);
          const float2 Xc = make_float2(X.x * sx, X.y * sx);
          const float2 C = make_float2(A.x * Xc.x - A.y * Xc.y, A.y * Xc.x + A.x * Xc.y);
          const float s = pow_scale(sqrtf(C.x * C.x + C.y * C.y), p.inv_p - 1.0f);
          Y = make_float2(C.x * s, C.y * s);
        }
      }
      if (k == 0 || k == M) Y.y = 0.0f;          // irfft ignores the imaginary part of the DC and Nyquist bins
      reinterpret_cast<float2*>(buf + (size_t)i * FRS)[k] = Y;
    }
  }
  __syncthreads();

  for (int rr = tid; rr < p.hop; rr += kDspThreads) {   // wtab is complete (barrier above); consumed after the next barrier
    float e = 0.0f;
    for (int o = rr; o < NFFT; o += p.hop) e += wtab[o] * wtab[o];
    envtab[rr] = e > FLT_MIN ? 1.0f / e : 1.0f;
  }
  // --- stage 2: 16 threads per frame: merge, inverse FFT, window, time-domain frame back into its buffer ---
  const int grp = tid >> 4, l16 = tid & 15;
  const float2* w2 = reinterpret_cast<const float2*>(wtab);
  constexpr float kNorm = 1.0f / (float)NFFT;
#pragma unroll 1
  for (int i0 = 0; i0 < nf; i0 += 16) {
    const int i = i0 + grp;
    const bool active = i < nf;
    float2* fb = reinterpret_cast<float2*>(buf + (size_t)(active ? i : 0) * FRS);
    float2 a[16];
    if (active && l16 < N2) {
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int n = N2 * n1 + l16;
        const float2 t = tw2[n];
        a[n1] = ft::irfft_merge(fb[n], fb[M - n], t.x, t.y);
      }
    }
    __syncwarp();                                   // every spectrum read of the group precedes the exchange writes
    if (active && l16 < N2) {
      ft::Dft<16, true>::run(a);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const float2 t = tw1[k1 * N2 + l16];
        fb[k1 * (N2 + 1) + l16] = ft::twmul<true>(a[k1], t.x, t.y);
      }
    }
    __syncwarp();
    float2 bq[N2];
    if (active) {
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) bq[n2] = fb[l16 * (N2 + 1) + n2];
    }
    __syncwarp();
    if (active) {
      ft::Dft<N2, true>::run(bq);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) {
        const int m = l16 + 16 * k2;
        const float2 w = w2[m];
        fb[m] = make_float2(bq[k2].x * w.x * kNorm, bq[k2].y * w.y * kNorm);
      }
    }
  }
  __syncthreads();

  // --- stage 3: gather overlap-add, envelope, trim, scale ------------------------------------
  // Division-free indexing (j -> hop block hb, offset r advance by constants) and, for samples whose covering frames all
  // exist, the window-sum-square envelope from a per-CTA table envtab[s mod hop]: this stage was the heaviest of the
  // kernel (two integer divisions, a w^2 sum and an IEEE division per sample).
  const float osc = p.out_scale ? __ldg(p.out_scale + b) : 1.0f;
  float* out = p.out + (long long)b * p.out_stride;
  const int hop = p.hop;
  const int span = OB * hop;
  const int c0 = (NFFT / 2) % hop;                 // s0 mod hop (n0 is a multiple of hop)
  const int tq0 = s0 / hop;
  const int q1 = (NFFT - 1) / hop, r1 = (NFFT - 1) % hop;
  const int dq = kDspThreads / hop, dr = kDspThreads % hop;
  int hb = tid / hop, r = tid % hop;
  for (int j = tid; j < span; j += kDspThreads) {
    const int n = n0 + j;
    if (n >= p.L) break;
    int rr = r + c0, th = tq0 + hb;                // s = s0 + j:  rr = s mod hop, th = s / hop
    if (rr >= hop) {
      rr -= hop;
      ++th;
    }
    r += dr;
    hb += dq;
    if (r >= hop) {
      r -= hop;
      ++hb;
    }
    if (n >= Lb) {
      out[n] = 0.0f;
      continue;
    }
    const int mmax = q1 - (rr > r1 ? 1 : 0);       // frames th, th-1, ..., th-mmax cover sample s (offsets rr + m hop < NFFT)
    float acc = 0.0f;
    if (th - mmax >= 0 && th <= t_hi) {            // interior: every covering frame exists
      const float* fp = buf + (size_t)(th - t_lo) * FRS + rr;
      for (int m = 0; m <= mmax; ++m) acc += fp[m * hop - (long long)m * FRS];
      acc *= envtab[rr];
    } else {
      float env = 0.0f;
      for (int m = 0; m <= mmax; ++m) {
        const int t = th - m;
        if (t < 0 || t < t_lo || t > t_hi) continue;
        const int off = rr + m * hop;
        acc += buf[(size_t)(t - t_lo) * FRS + off];
        const float w = wtab[off];
        env += w * w;
      }
      if (env > FLT_MIN) acc /= env;
    }
    out[n] = acc * osc;
  }
}

static int dsp_engine() {
  static int e = -1;
  if (e < 0) {
    e = 2;
    if (const char* v = getenv("SE_DSP_ENGINE")) e = atoi(v) == 1 ? 1 : 2;    // 1 = round-1 warp-shuffle kernels (A/B)
  }
  return e;
}
// measured (profiles/dsp_only_r02.jsonl): 16 frames per CTA win at the 512-point geometries (less shared memory, more
// CTAs per SM), 32 at 320 points; SE_DSP_TILE=16|32 overrides (A/B)
static int dsp_tile(int R) {
  static int forced = -1;
  if (forced < 0) {
    forced = 0;
    if (const char* v = getenv("SE_DSP_TILE")) forced = atoi(v) == 32 ? 32 : (atoi(v) == 16 ? 16 : 0);
  }
  return forced ? forced : (R == 5 ? 32 : 16);
}
static int dsp2_smem_stft(int R, int hop, int ft) {
  const int nfft = 64 * R, M = 32 * R, N2 = 2 * R;
  const int tile = (ft - 1) * hop + nfft;
  return nfft * 4 + (16 * N2 + M + 2) * 8 + ft * (M + 16) * 8 + 16 + tile * 4 + 16;
}
template <int R, int FT>
static cudaError_t launch_stft2(const StftParams& p, int smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(stft2_kernel<R, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) stft2_kernel<R, FT><<<dim3(ceil_div(p.T, FT), p.B), kDspThreads, smem, s>>>(p);
  return e;
}
template <int R>
static cudaError_t launch_istft2(const IstftParams& p, int smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(istft2_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) istft2_kernel<R><<<dim3(ceil_div(p.L, p.ob * p.hop), p.B), kDspThreads, smem, s>>>(p);
  return e;
}
static int dsp2_smem_istft(int R, int nf_max) {
  const int nfft = 64 * R, M = 32 * R, N2 = 2 * R;
  return nfft * 4 + (16 * N2 + M + 2) * 8 + 512 * 4 + nf_max * 2 * (M + 16) * 4;
}

// -------------------------------------------------------------------------------------------
// front step: band-limited resampling (resampy.interpn.resample_f restated; one output sample per thread)
// -------------------------------------------------------------------------------------------
// HBM-bound in principle (reads 4/ratio bytes, writes 4 bytes per output sample); the ~2 * num_zeros / min(1, ratio)
// taps per sample (385 at 48k -> 16k) re-read the input and the 128 KB table from L1/L2.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long x_stride, int n_in,
                                                      float* __restrict__ y, long long y_stride, int n_out, int n_valid,
                                                      double ratio, const double* __restrict__ time_reg_tab,
                                                      const float* __restrict__ win,
                                                      const float* __restrict__ delta, int nwin, int num_table) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const float* xb = x + (long long)blockIdx.y * x_stride;
  float* yb = y + (long long)blockIdx.y * y_stride;
  if (t >= n_valid) {
    yb[t] = 0.f;
    return;
  }
  const double scale = ratio < 1.0 ? ratio : 1.0;
  const int step = (int)(scale * (double)num_table);
  // resample_f accumulates time_register += 1/ratio in float64.  Where the exact time is an integer (t = 160 k at
  // 44.1 -> 16 kHz) the rounding of that sum decides floor(time), and with it which table entries are used (the table
  // stride is truncated to an integer, so the two choices differ by ~1e-3): the caller passes the sequentially
  // accumulated register; without it t * (1/ratio) is used (identical for exactly representable increments).
  const double time_reg = time_reg_tab ? __ldg(time_reg_tab + t) : (double)t * (1.0 / ratio);
  const int n = (int)time_reg;
  double frac = scale * (time_reg - (double)n);
  double index_frac = frac * (double)num_table;
  int offset = (int)index_frac;
  float eta = (float)(index_frac - (double)offset);
  float acc = 0.f;
  const int i_max = min(n + 1, (nwin - offset) / step);
  for (int i = 0; i < i_max; ++i) {
    const int idx = offset + i * step;
    acc = fmaf(__ldg(win + idx) + eta * __ldg(delta + idx), __ldg(xb + n - i), acc);
  }
  frac = scale - frac;
  index_frac = frac * (double)num_table;
  offset = (int)index_frac;
  eta = (float)(index_frac - (double)offset);
  const int k_max = min(n_in - n - 1, (nwin - offset) / step);
  for (int k = 0; k < k_max; ++k) {
    const int idx = offset + k * step;
    acc = fmaf(__ldg(win + idx) + eta * __ldg(delta + idx), __ldg(xb + n + k + 1), acc);
  }
  yb[t] = acc;
}

// Twiddle tables of the two FFT sizes in static device memory (no allocation behind the caller's back), filled on the
// first call per device: 32 lanes x (2 R + 5) float2.
__device__ float2 g_twiddles5[32 * WarpFFT<5>::kTwPerLane];
__device__ float2 g_twiddles8[32 * WarpFFT<8>::kTwPerLane];

static const float2* dsp_twiddles(int R, cudaStream_t s) {
  constexpr int kMaxDev = 64;
  static bool ready[kMaxDev][2] = {};
  static std::mutex mu;                       // first calls from several host threads / streams of one device
  std::lock_guard<std::mutex> guard(mu);
  (void)s;
  int dev = 0;
  cudaGetDevice(&dev);
  const int which = R == 5 ? 0 : 1;
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, R == 5 ? (const void*)g_twiddles5 : (const void*)g_twiddles8) != cudaSuccess) return nullptr;
  if (dev < 0 || dev >= kMaxDev || !ready[dev][which]) {
    float2 host[32 * WarpFFT<8>::kTwPerLane];
    size_t bytes;
    if (R == 5) {
      WarpFFT<5>::fill_table(host);
      bytes = sizeof(float2) * 32 * WarpFFT<5>::kTwPerLane;
    } else {
      WarpFFT<8>::fill_table(host);
      bytes = sizeof(float2) * 32 * WarpFFT<8>::kTwPerLane;
    }
    // SYNCHRONOUS copy: complete when it returns, so a launch on ANY stream that follows sees the table (an async copy on
    // the caller's stream let a first STFT on another stream read zeros).  Not legal inside a stream capture: the first
    // call per device must be made outside one (decode.GraphedEnhance / se_enhance_crn warm up before they capture).
    if (cudaMemcpy(sym, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    if (dev >= 0 && dev < kMaxDev) ready[dev][which] = true;
  }
  return reinterpret_cast<const float2*>(sym);
}

static int dsp_smem_stft(int R, int hop) {
  const int nfft = 64 * R;
  const int tile = (kFramesPerCta - 1) * hop + nfft;
  return nfft * 4 + kFramesPerCta * R * 33 * 8 + kFramesPerCta * 4 + 16 + tile * 4 + 16;
}

}  // namespace se

using namespace se;

extern "C" int se_rms_scale_len(const float* wav, long long wav_stride, int B, int N, const int* lengths, int reciprocal,
                                float* c, float* inv_c, se_stream_t stream) {
  SE_REQUIRE(wav && c && inv_c && B > 0 && N > 0, "se_rms_scale: bad arguments (B=%d N=%d)", B, N);
  rms_scale_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(wav, wav_stride, N, lengths, reciprocal, c, inv_c);
  return check_launch("se_rms_scale");
}

extern "C" int se_rms_scale(const float* wav, long long wav_stride, int B, int N, int reciprocal, float* c,
                            float* inv_c, se_stream_t stream) {
  return se_rms_scale_len(wav, wav_stride, B, N, nullptr, reciprocal, c, inv_c, stream);
}

static int geom_ok(const char* who, int n_fft, int win, int hop) {
  if (!(n_fft == 320 || n_fft == 512)) {
    set_error("%s: n_fft=%d unsupported (320 or 512)", who, n_fft);
    return 0;
  }
  if (win < 2 || win > n_fft || ((n_fft - win) & 1) || hop < 2 || (hop & 1) || hop > n_fft) {
    set_error("%s: unsupported win=%d hop=%d for n_fft=%d", who, win, hop, n_fft);
    return 0;
  }
  return 1;
}

extern "C" int se_stft(const float* wav, long long wav_stride, int B, int N, const float* scale, int n_fft, int win,
                       int hop, int T, float* mag, long long msb, long long mst, long long msf, float* re,
                       float* im, long long sb, long long st, long long sf, float p_mag, float p_ri,
                       se_stream_t stream) {
  return se_stft_len(wav, wav_stride, B, N, nullptr, scale, n_fft, win, hop, T, mag, msb, mst, msf, re, im, sb, st, sf, p_mag,
                     p_ri, stream);
}

extern "C" int se_stft_len(const float* wav, long long wav_stride, int B, int N, const int* lengths, const float* scale,
                           int n_fft, int win, int hop, int T, float* mag, long long msb, long long mst, long long msf,
                           float* re, float* im, long long sb, long long st, long long sf, float p_mag, float p_ri,
                           se_stream_t stream) {
  if (!geom_ok("se_stft", n_fft, win, hop)) return SE_ERR_SHAPE;
  SE_REQUIRE(wav && B > 0 && N >= n_fft, "se_stft: need N >= n_fft (N=%d)", N);
  SE_REQUIRE(T == 1 + N / hop, "se_stft: T=%d but 1+N/hop=%d", T, 1 + N / hop);
  SE_REQUIRE((re == nullptr) == (im == nullptr), "se_stft: re and im must both be given or both NULL");
  SE_REQUIRE(mag || re, "se_stft: no output plane");
  StftParams p{wav, wav_stride, B, N, scale, win, hop, T, mag, re, im, msb, mst, msf, sb, st, sf, p_mag, p_ri, nullptr,
               lengths};
  p.twiddles = dsp_twiddles(n_fft / 64, (cudaStream_t)stream);
  SE_REQUIRE(p.twiddles != nullptr, "se_stft: twiddle table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  const int R = n_fft / 64;
  if (dsp_engine() == 2) {
    const int ft = dsp_tile(R);
    const int smem2 = dsp2_smem_stft(R, hop, ft);
    cudaStream_t cs = (cudaStream_t)stream;
    const cudaError_t e2 = R == 5 ? (ft == 16 ? launch_stft2<5, 16>(p, smem2, cs) : launch_stft2<5, 32>(p, smem2, cs))
                                  : (ft == 16 ? launch_stft2<8, 16>(p, smem2, cs) : launch_stft2<8, 32>(p, smem2, cs));
    if (e2 != cudaSuccess) {
      set_error("se_stft: cudaFuncSetAttribute(%d bytes): %s", smem2, cudaGetErrorString(e2));
      return SE_ERR_CUDA;
    }
    return check_launch("se_stft");
  }
  dim3 grid(ceil_div(T, kFramesPerCta), B);
  const int smem = dsp_smem_stft(R, hop);
  cudaError_t e;
  if (R == 5) {
    e = cudaFuncSetAttribute(stft_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) stft_kernel<5><<<grid, kDspThreads, smem, (cudaStream_t)stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(stft_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) stft_kernel<8><<<grid, kDspThreads, smem, (cudaStream_t)stream>>>(p);
  }
  if (e != cudaSuccess) {
    set_error("se_stft: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  return check_launch("se_stft");
}

extern "C" int se_istft(int mode, const float* a_re, const float* a_im, long long a_sb, long long a_st,
                        long long a_sf, const float* b_re, const float* b_im, long long b_sb, long long b_st,
                        long long b_sf, float inv_p, float p_x, int B, int T, int n_fft, int win, int hop,
                        const float* out_scale, float* out, long long out_stride, int L, se_stream_t stream) {
  return se_istft_len(mode, a_re, a_im, a_sb, a_st, a_sf, b_re, b_im, b_sb, b_st, b_sf, inv_p, p_x, B, T, n_fft, win, hop,
                      out_scale, out, out_stride, L, nullptr, stream);
}

extern "C" int se_istft_len(int mode, const float* a_re, const float* a_im, long long a_sb, long long a_st,
                            long long a_sf, const float* b_re, const float* b_im, long long b_sb, long long b_st,
                            long long b_sf, float inv_p, float p_x, int B, int T, int n_fft, int win, int hop,
                            const float* out_scale, float* out, long long out_stride, int L, const int* lengths,
                            se_stream_t stream) {
  if (!geom_ok("se_istft", n_fft, win, hop)) return SE_ERR_SHAPE;
  SE_REQUIRE(mode >= SE_ISTFT_SPEC && mode <= SE_ISTFT_CMASK, "se_istft: bad mode %d", mode);
  SE_REQUIRE(a_re && out && B > 0 && T > 0 && L > 0, "se_istft: bad arguments");
  SE_REQUIRE(mode == SE_ISTFT_MAG_PHASE || a_im, "se_istft: a_im required for mode %d", mode);
  SE_REQUIRE(mode < SE_ISTFT_MAG_PHASE || (b_re && b_im), "se_istft: noisy spectrum (b_re,b_im) required");
  IstftParams p{mode, a_re, a_im, a_sb, a_st, a_sf, b_re, b_im, b_sb, b_st, b_sf, inv_p, p_x, B, T, win, hop,
                out_scale, out, out_stride, L, 0, nullptr, lengths, kHopBlocksPerCta};
  p.twiddles = dsp_twiddles(n_fft / 64, (cudaStream_t)stream);
  SE_REQUIRE(p.twiddles != nullptr, "se_istft: twiddle table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  const int R = n_fft / 64;
  if (dsp_engine() == 2) {
    // hop blocks per CTA: the frames that overlap them (at most ob + ceil(n_fft / hop)) fill whole rounds of 16 frame groups
    p.nf_max = dsp_tile(R);
    p.ob = p.nf_max - ceil_div(n_fft, hop);
    const int smem2 = dsp2_smem_istft(R, p.nf_max);
    cudaStream_t cs = (cudaStream_t)stream;
    const cudaError_t e2 = R == 5 ? launch_istft2<5>(p, smem2, cs) : launch_istft2<8>(p, smem2, cs);
    if (e2 != cudaSuccess) {
      set_error("se_istft: cudaFuncSetAttribute(%d bytes): %s", smem2, cudaGetErrorString(e2));
      return SE_ERR_CUDA;
    }
    return check_launch("se_istft");
  }
  p.nf_max = kHopBlocksPerCta + ceil_div(n_fft, hop) + 1;
  dim3 grid(ceil_div(L, kHopBlocksPerCta * hop), B);
  const int smem = n_fft * 4 + ((p.nf_max + 3) & ~3) * 4 + p.nf_max * R * 66 * 4;
  cudaError_t e;
  if (R == 5) {
    e = cudaFuncSetAttribute(istft_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) istft_kernel<5><<<grid, kDspThreads, smem, (cudaStream_t)stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(istft_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) istft_kernel<8><<<grid, kDspThreads, smem, (cudaStream_t)stream>>>(p);
  }
  if (e != cudaSuccess) {
    set_error("se_istft: cudaFuncSetAttribute(%d bytes): %s", smem, cudaGetErrorString(e));
    return SE_ERR_CUDA;
  }
  return check_launch("se_istft");
}

extern "C" int se_resample(const float* x, long long x_stride, int B, int n_in, float* y, long long y_stride, int n_out,
                           int n_valid, double ratio, const double* time_reg, const float* win, const float* delta,
                           int nwin, int num_table, se_stream_t stream) {
  SE_REQUIRE(x && y && win && delta && B > 0 && n_in > 0 && n_out > 0, "se_resample: bad arguments");
  SE_REQUIRE(ratio > 0.0 && num_table > 0 && nwin > num_table, "se_resample: ratio=%g num_table=%d nwin=%d", ratio,
             num_table, nwin);
  SE_REQUIRE((int)((ratio < 1.0 ? ratio : 1.0) * num_table) >= 1, "se_resample: ratio %g too small for the table", ratio);
  SE_REQUIRE(n_valid >= 0 && n_valid <= n_out && (double)(n_valid - 1) / ratio < (double)n_in,
             "se_resample: n_valid=%d reads past the input (n_in=%d ratio=%g)", n_valid, n_in, ratio);
  dim3 grid(ceil_div(n_out, 256), B);
  resample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_stride, n_in, y, y_stride, n_out, n_valid, ratio,
                                                         time_reg, win, delta, nwin, num_table);
  return check_launch("se_resample");
}
