// CPU emulation of the two-pass per-thread FFT of csrc/fft_thread.cuh exactly as stft2_kernel / istft2_kernel drive it
// (thread n2 / k1 loops become plain loops; the shared-memory exchange is an array).  Prints, for M = argv[1] in {160, 256}:
// line 1: the 2M real input samples; line 2: the M+1 complex bins (re im ...) of the forward transform; line 3: the 2M
// samples reconstructed by the inverse path.  tests/test_host_logic.py compares with numpy.fft.rfft / irfft.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../sixty-years-of-frequency-domain-monaural-speech-enhancement_b200/csrc/fft_thread.cuh"

using namespace se::ft;

template <int N2>
static void run(int M) {
  const double PI = 3.14159265358979323846;
  std::vector<float> x(2 * M);
  unsigned s = 12345u;
  for (auto& v : x) {
    s = s * 1664525u + 1013904223u;
    v = (float)((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
  }
  std::vector<float2> tw1(16 * N2), tw2(M + 1);      // W_M^(n2 k1), W_2M^k as (cos, sin) of the positive angle
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < N2; ++n2) tw1[k1 * N2 + n2] = make_float2((float)cos(2 * PI * n2 * k1 / M), (float)sin(2 * PI * n2 * k1 / M));
  for (int k = 0; k <= M; ++k) tw2[k] = make_float2((float)cos(PI * k / M), (float)sin(PI * k / M));
  // ---- forward
  std::vector<float2> S(16 * N2), Z(M), X(M + 1);
  for (int n2 = 0; n2 < N2; ++n2) {                   // pass 1: thread n2
    float2 a[16];
    for (int n1 = 0; n1 < 16; ++n1) a[n1] = make_float2(x[2 * (N2 * n1 + n2)], x[2 * (N2 * n1 + n2) + 1]);
    Dft<16, false>::run(a);
    for (int k1 = 0; k1 < 16; ++k1) S[k1 * N2 + n2] = twmul<false>(a[k1], tw1[k1 * N2 + n2].x, tw1[k1 * N2 + n2].y);
  }
  for (int k1 = 0; k1 < 16; ++k1) {                   // pass 2: thread k1
    float2 b[N2];
    for (int n2 = 0; n2 < N2; ++n2) b[n2] = S[k1 * N2 + n2];
    Dft<N2, false>::run(b);
    for (int k2 = 0; k2 < N2; ++k2) Z[k1 + 16 * k2] = b[k2];
  }
  for (int k = 0; k <= M; ++k) X[k] = rfft_split(Z[k % M], Z[(M - k) % M], tw2[k].x, tw2[k].y);
  // ---- inverse: merge, the same two passes with conjugate twiddles, 1 / (2M)
  std::vector<float2> Zi(M), Si(16 * N2), z(M);
  for (int k = 0; k < M; ++k) Zi[k] = irfft_merge(X[k], X[M - k], tw2[k].x, tw2[k].y);
  for (int n2 = 0; n2 < N2; ++n2) {
    float2 a[16];
    for (int n1 = 0; n1 < 16; ++n1) a[n1] = Zi[N2 * n1 + n2];
    Dft<16, true>::run(a);
    for (int k1 = 0; k1 < 16; ++k1) Si[k1 * N2 + n2] = twmul<true>(a[k1], tw1[k1 * N2 + n2].x, tw1[k1 * N2 + n2].y);
  }
  for (int k1 = 0; k1 < 16; ++k1) {
    float2 b[N2];
    for (int n2 = 0; n2 < N2; ++n2) b[n2] = Si[k1 * N2 + n2];
    Dft<N2, true>::run(b);
    for (int k2 = 0; k2 < N2; ++k2) z[k1 + 16 * k2] = scale(b[k2], 1.0f / (2.0f * M));
  }
  for (int i = 0; i < 2 * M; ++i) printf("%.9g ", x[i]);
  printf("\n");
  for (int k = 0; k <= M; ++k) printf("%.9g %.9g ", X[k].x, X[k].y);
  printf("\n");
  for (int n = 0; n < M; ++n) printf("%.9g %.9g ", z[n].x, z[n].y);
  printf("\n");
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 256;
  if (M == 256) run<16>(M);
  else if (M == 160) run<10>(M);
  else return 2;
  return 0;
}
