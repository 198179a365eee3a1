// Per-thread small DFTs for the second-generation STFT / iSTFT kernels (dsp.cu: stft2_kernel / istft2_kernel).
//
// A real frame of NFFT = 2M samples (M = 160 or 256) is packed as M complex points z[n] = x[2n] + i x[2n+1]; the length-M
// complex FFT is split  M = 16 * N2  (N2 = 10 or 16, Cooley-Tukey, n = N2 n1 + n2, k = k1 + 16 k2):
//     pass 1   thread n2 (N2 of the 16 threads of a frame): 16-point DFT over n1 in REGISTERS, times W_M^(n2 k1)
//     exchange through shared memory  S[k1][n2]
//     pass 2   thread k1 (all 16 threads): N2-point DFT over n2 in registers  ->  Z[k1 + 16 k2]
// instead of the five __shfl_xor butterfly stages of fft.cuh (whose ~470 warp instructions per frame made the round-1
// kernels issue-bound at 16-27 % of the HBM roofline).  The real-FFT split / merge
//     X[k] = (Z[k] + conj Z[M-k]) / 2  -  i W_2M^k (Z[k] - conj Z[M-k]) / 2
// is applied where the bins are read for the epilogue (forward) or written by the prologue (inverse).
//
// Plain C++ (no CUDA intrinsics) so the index algebra is unit-tested on the CPU: tests/test_host_logic.py compiles
// tools/fft_thread_host_test.cpp against numpy.fft.
#pragma once

#if defined(__CUDACC__)
#define SE_FT_HD __host__ __device__ __forceinline__
#else
#define SE_FT_HD inline
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}
#endif

namespace se {
namespace ft {

SE_FT_HD float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SE_FT_HD float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SE_FT_HD float2 scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
SE_FT_HD float2 conj(float2 a) { return make_float2(a.x, -a.y); }
// a * (c - i s) forward, a * (c + i s) inverse   (c, s = cos, sin of a positive angle)
template <bool INV>
SE_FT_HD float2 twmul(float2 a, float c, float s) {
  return INV ? make_float2(a.x * c - a.y * s, a.y * c + a.x * s) : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
SE_FT_HD float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
SE_FT_HD void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
  const float2 s0 = add(c0, c2), s1 = sub(c0, c2), s2 = add(c1, c3), s3 = mul_mi<INV>(sub(c1, c3));
  c0 = add(s0, s2);
  c1 = add(s1, s3);
  c2 = sub(s0, s2);
  c3 = sub(s1, s3);
}

template <bool INV>
SE_FT_HD void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;    // cos(2 pi / 5), cos(4 pi / 5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;     // sin(2 pi / 5), sin(4 pi / 5)
  const float2 t1 = add(x1, x4), t2 = add(x2, x3), t3 = sub(x1, x4), t4 = sub(x2, x3);
  const float2 m1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
  const float2 u1 = mul_mi<INV>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
  const float2 u2 = mul_mi<INV>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
  x0 = add(x0, add(t1, t2));
  x1 = add(m1, u1);
  x4 = sub(m1, u1);
  x2 = add(m2, u2);
  x3 = sub(m2, u2);
}

// 16-point DFT, natural order in and out:  n = q + 4 p, k = r + 4 s
template <bool INV>
SE_FT_HD void dft16(float2 (&a)[16]) {
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;     // cos, sin(2 pi / 16)
  const float c2 = 0.70710678118654752440f;                                   // cos = sin(4 pi / 16)
#pragma unroll
  for (int q = 0; q < 4; ++q) dft4<INV>(a[q], a[q + 4], a[q + 8], a[q + 12]);   // a[q + 4 r] <- sum_p a[q + 4 p] W4^(p r)
  // twiddles W16^(q r)
  a[1 + 4] = twmul<INV>(a[1 + 4], c1, s1);       // q = 1, r = 1: W16^1
  a[1 + 8] = twmul<INV>(a[1 + 8], c2, c2);       // q = 1, r = 2: W16^2
  a[1 + 12] = twmul<INV>(a[1 + 12], s1, c1);     // q = 1, r = 3: W16^3
  a[2 + 4] = twmul<INV>(a[2 + 4], c2, c2);       // q = 2, r = 1: W16^2
  a[2 + 8] = mul_mi<INV>(a[2 + 8]);              // q = 2, r = 2: W16^4 = -i
  a[2 + 12] = twmul<INV>(a[2 + 12], -c2, c2);    // q = 2, r = 3: W16^6
  a[3 + 4] = twmul<INV>(a[3 + 4], s1, c1);       // q = 3, r = 1: W16^3
  a[3 + 8] = twmul<INV>(a[3 + 8], -c2, c2);      // q = 3, r = 2: W16^6
  a[3 + 12] = twmul<INV>(a[3 + 12], -c1, -s1);   // q = 3, r = 3: W16^9
#pragma unroll
  for (int r = 0; r < 4; ++r) dft4<INV>(a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3]);   // a[4 r + s] = X[r + 4 s]
  // transpose to natural order: X[r + 4 s] sits in a[4 r + s]
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = r + 1; s < 4; ++s) {
      const float2 t = a[4 * r + s];
      a[4 * r + s] = a[4 * s + r];
      a[4 * s + r] = t;
    }
}

// 10-point DFT, natural order in and out:  n = q + 2 p, k = r + 5 s
template <bool INV>
SE_FT_HD void dft10(float2 (&a)[10]) {
  // W10^r, r = 1..4: (cos, sin)(2 pi r / 10)
  const float c1 = 0.80901699437494742410f, s1 = 0.58778525229247312917f;
  const float c2 = 0.30901699437494742410f, s2 = 0.95105651629515357212f;
  dft5<INV>(a[0], a[2], a[4], a[6], a[8]);       // a[2 r]     <- sum_p a[2 p] W5^(p r)
  dft5<INV>(a[1], a[3], a[5], a[7], a[9]);       // a[2 r + 1] <- sum_p a[2 p + 1] W5^(p r)
  a[3] = twmul<INV>(a[3], c1, s1);
  a[5] = twmul<INV>(a[5], c2, s2);
  a[7] = twmul<INV>(a[7], -c2, s2);
  a[9] = twmul<INV>(a[9], -c1, s1);
  float2 o[10];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    o[r] = add(a[2 * r], a[2 * r + 1]);
    o[r + 5] = sub(a[2 * r], a[2 * r + 1]);
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) a[k] = o[k];
}

template <int N, bool INV>
struct Dft;
template <bool INV>
struct Dft<16, INV> {
  static SE_FT_HD void run(float2 (&a)[16]) { dft16<INV>(a); }
};
template <bool INV>
struct Dft<10, INV> {
  static SE_FT_HD void run(float2 (&a)[10]) { dft10<INV>(a); }
};

// real-FFT split of bin k from the packed transform:  X[k] = (Z[k] + conj Z[M-k]) / 2 - i W_2M^k (Z[k] - conj Z[M-k]) / 2
// zk = Z[k mod M], zm = Z[(M - k) mod M], (c, s) = (cos, sin)(pi k / M)
SE_FT_HD float2 rfft_split(float2 zk, float2 zm, float c, float s) {
  const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));      // (Z[k] + conj Z[M-k]) / 2
  const float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));      // (Z[k] - conj Z[M-k]) / 2
  const float2 w = twmul<false>(o, c, s);                                          // W_2M^k * o
  return make_float2(e.x + w.y, e.y - w.x);                                        // e - i w
}
// inverse: Z[k] = (X[k] + conj X[M-k]) + i conj(W_2M^k) (X[k] - conj X[M-k])   (so that IDFT_M(Z)[n] = x[2n] + i x[2n+1]
// times 2M when the IDFT is unnormalised and the 1/(2M) of irfft is applied afterwards)
SE_FT_HD float2 irfft_merge(float2 xk, float2 xm, float c, float s) {
  const float2 e = make_float2(xk.x + xm.x, xk.y - xm.y);
  const float2 o = make_float2(xk.x - xm.x, xk.y + xm.y);
  const float2 w = twmul<true>(o, c, s);                                           // conj(W_2M^k) * o
  return make_float2(e.x - w.y, e.y + w.x);                                        // e + i w
}

}  // namespace ft
}  // namespace se
