// Warp-level real FFT of length N = 64*R (R = 5 -> 320 points, R = 8 -> 512 points).
//
// One warp transforms one frame.  The real frame x[0..N) is packed as N/2 complex
// points z[m] = x[2m] + i x[2m+1]; the length-N/2 complex FFT is split R x 32
// (Cooley-Tukey): a radix-R butterfly in registers (lane m2 holds z[32*m1 + m2],
// m1 < R), a twiddle, and a 32-point FFT ACROSS LANES done with five radix-2
// decimation-in-frequency stages of __shfl_xor butterflies.  After the stages lane l
// holds bins k = k1 + R*bitrev5(l); the real-FFT split X[k] = E[k] + W_N^k O[k] needs
// Z[N/2-k], which lives in lane l^31 (register R-k1) -- one more shuffle.  The inverse
// runs the same network backwards (DIT, conjugate twiddles).  Index algebra verified
// against numpy.fft in a lane-level simulation before this was written.
#pragma once
#include "common.cuh"

namespace se {

__device__ __forceinline__ int bitrev5(int l) { return (int)(__brev((unsigned)l) >> 27); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
  float2 s0 = cadd(c0, c2), s1 = csub(c0, c2), s2 = cadd(c1, c3), s3 = mul_mi<INV>(csub(c1, c3));
  c0 = cadd(s0, s2);
  c1 = cadd(s1, s3);
  c2 = csub(s0, s2);
  c3 = csub(s1, s3);
}

template <int R, bool INV>
struct SmallDft;

template <bool INV>
struct SmallDft<8, INV> {
  __device__ static __forceinline__ void run(float2 (&x)[8]) {
    const float s = 0.70710678118654752440f;
    float2 a0 = cadd(x[0], x[4]), a1 = cadd(x[1], x[5]), a2 = cadd(x[2], x[6]), a3 = cadd(x[3], x[7]);
    float2 b0 = csub(x[0], x[4]), b1 = csub(x[1], x[5]), b2 = csub(x[2], x[6]), b3 = csub(x[3], x[7]);
    // b_j *= W8^j   (forward W8 = exp(-i pi/4))
    b1 = INV ? make_float2(s * (b1.x - b1.y), s * (b1.x + b1.y)) : make_float2(s * (b1.x + b1.y), s * (b1.y - b1.x));
    b2 = mul_mi<INV>(b2);
    b3 = INV ? make_float2(-s * (b3.x + b3.y), s * (b3.x - b3.y)) : make_float2(s * (b3.y - b3.x), -s * (b3.x + b3.y));
    dft4<INV>(a0, a1, a2, a3);  // X[0],X[2],X[4],X[6]
    dft4<INV>(b0, b1, b2, b3);  // X[1],X[3],X[5],X[7]
    x[0] = a0; x[2] = a1; x[4] = a2; x[6] = a3;
    x[1] = b0; x[3] = b1; x[5] = b2; x[7] = b3;
  }
};

template <bool INV>
struct SmallDft<5, INV> {
  __device__ static __forceinline__ void run(float2 (&x)[5]) {
    const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
    const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
    const float s1 = 0.95105651629515357212f;   // sin(2pi/5)
    const float s2 = 0.58778525229247312917f;   // sin(4pi/5)
    float2 t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
    float2 m1 = make_float2(x[0].x + c1 * t1.x + c2 * t2.x, x[0].y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(x[0].x + c2 * t1.x + c1 * t2.x, x[0].y + c2 * t1.y + c1 * t2.y);
    float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    x[0] = make_float2(x[0].x + t1.x + t2.x, x[0].y + t1.y + t2.y);
    float2 in1 = mul_mi<INV>(n1), in2 = mul_mi<INV>(n2);  // -/+ i * n
    x[1] = cadd(m1, in1);
    x[4] = csub(m1, in1);
    x[2] = cadd(m2, in2);
    x[3] = csub(m2, in2);
  }
};

template <int R>
struct WarpFFT {
  static constexpr int N2 = 32 * R;
  static constexpr int N = 64 * R;
  float2 tw2[R];  // exp(-2 pi i lane*k1 / N2)
  float2 twp[R];  // exp(-2 pi i k / N),  k = k1 + R*bitrev5(lane)
  float2 tws[5];  // stage h = 16,8,4,2,1: exp(-2 pi i (lane&(h-1)) / (2h))
  int lane, src0;

  // Twiddles of one lane: kTwPerLane float2 = tw2[0..R) | twp[0..R) | tws[0..5), rounded from float64.  They are computed
  // ONCE per device on the host (fill_table) and read back with 2 R + 5 cached 8-byte loads: evaluating them in the
  // kernel cost 15 double-precision sincospi per thread and CTA, 43 % of all instructions of the STFT (ncu, round 1).
  static constexpr int kTwPerLane = 2 * R + 5;

  static void fill_table(float2* tab /* [32][kTwPerLane] */) {
    const double pi = 3.14159265358979323846;
    for (int l = 0; l < 32; ++l) {
      int br = 0;
      for (int bit = 0; bit < 5; ++bit) br |= ((l >> bit) & 1) << (4 - bit);
      float2* t = tab + l * kTwPerLane;
      for (int k1 = 0; k1 < R; ++k1) {
        const double a = 2.0 * pi * (double)((l * k1) % N2) / (double)N2;
        const double b = 2.0 * pi * (double)(k1 + R * br) / (double)N;
        t[k1] = make_float2((float)cos(a), (float)-sin(a));
        t[R + k1] = make_float2((float)cos(b), (float)-sin(b));
      }
      for (int si = 0; si < 5; ++si) {
        const int h = 16 >> si;
        const double a = pi * (double)(l & (h - 1)) / (double)h;
        t[2 * R + si] = make_float2((float)cos(a), (float)-sin(a));
      }
    }
  }

  __device__ __forceinline__ void init(int lane_, const float2* __restrict__ tab) {
    lane = lane_;
    const int br = bitrev5(lane);
    const float2* t = tab + lane * kTwPerLane;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      tw2[k1] = __ldg(t + k1);
      twp[k1] = __ldg(t + R + k1);
    }
#pragma unroll
    for (int si = 0; si < 5; ++si) tws[si] = __ldg(t + 2 * R + si);
    src0 = bitrev5((32 - br) & 31);
  }

  // in : z[m1] = (x[2m], x[2m+1]), m = 32*m1 + lane
  // out: z[k1] = X[k1 + R*bitrev5(lane)];  returns Re X[N/2] (valid in lane 0)
  __device__ __forceinline__ float forward(float2 (&z)[R]) const {
    SmallDft<R, false>::run(z);
#pragma unroll
    for (int k1 = 1; k1 < R; ++k1) z[k1] = cmul(z[k1], tw2[k1]);
#pragma unroll
    for (int si = 0; si < 5; ++si) {
      const int h = 16 >> si;
      const bool up = (lane & h) == 0;
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, z[k1].x, h);
        o.y = __shfl_xor_sync(0xffffffffu, z[k1].y, h);
        if (up) {
          z[k1] = cadd(z[k1], o);
        } else {
          float2 d = csub(o, z[k1]);
          z[k1] = (h == 1) ? d : cmul(d, tws[si]);
        }
      }
    }
    const float nyq = z[0].x - z[0].y;
    float2 out[R];
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      float2 part;
      if (k1 == 0) {
        part.x = __shfl_sync(0xffffffffu, z[0].x, src0);
        part.y = __shfl_sync(0xffffffffu, z[0].y, src0);
      } else {
        part.x = __shfl_xor_sync(0xffffffffu, z[R - k1].x, 31);
        part.y = __shfl_xor_sync(0xffffffffu, z[R - k1].y, 31);
      }
      const float2 zp = cconj(part);
      const float2 e = make_float2(0.5f * (z[k1].x + zp.x), 0.5f * (z[k1].y + zp.y));
      const float2 d = csub(z[k1], zp);
      const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // -i/2 * d
      out[k1] = cadd(e, cmul(twp[k1], o));
    }
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) z[k1] = out[k1];
    return nyq;
  }

  // in : x[k1] = X[k1 + R*bitrev5(lane)], nyq_re = Re X[N/2] (lane 0's value is used)
  // out: x[m1] = (y[2m], y[2m+1]), m = 32*m1 + lane, y = irfft(X) (1/N normalised)
  __device__ __forceinline__ void inverse(float2 (&x)[R], float nyq_re) const {
    float2 zz[R];
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      float2 part;
      if (k1 == 0) {
        part.x = __shfl_sync(0xffffffffu, x[0].x, src0);
        part.y = __shfl_sync(0xffffffffu, x[0].y, src0);
      } else {
        part.x = __shfl_xor_sync(0xffffffffu, x[R - k1].x, 31);
        part.y = __shfl_xor_sync(0xffffffffu, x[R - k1].y, 31);
      }
      const float2 xp = cconj(part);
      const float2 e = make_float2(0.5f * (x[k1].x + xp.x), 0.5f * (x[k1].y + xp.y));
      const float2 wo = make_float2(0.5f * (x[k1].x - xp.x), 0.5f * (x[k1].y - xp.y));
      const float2 o = cmulc(wo, twp[k1]);
      zz[k1] = make_float2(e.x - o.y, e.y + o.x);  // E + i O
    }
    if (lane == 0) {  // DC / Nyquist: imaginary parts are ignored (C2R semantics)
      zz[0] = make_float2(0.5f * (x[0].x + nyq_re), 0.5f * (x[0].x - nyq_re));
    }
#pragma unroll
    for (int si = 4; si >= 0; --si) {
      const int h = 16 >> si;
      const bool up = (lane & h) == 0;
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        float2 mine = (up || h == 1) ? zz[k1] : cmulc(zz[k1], tws[si]);
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, mine.x, h);
        o.y = __shfl_xor_sync(0xffffffffu, mine.y, h);
        zz[k1] = up ? cadd(mine, o) : csub(o, mine);
      }
    }
#pragma unroll
    for (int k1 = 1; k1 < R; ++k1) zz[k1] = cmulc(zz[k1], tw2[k1]);
    SmallDft<R, true>::run(zz);
    const float inv = 1.0f / (float)N2;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) x[k1] = make_float2(zz[k1].x * inv, zz[k1].y * inv);
  }
};

}  // namespace se
