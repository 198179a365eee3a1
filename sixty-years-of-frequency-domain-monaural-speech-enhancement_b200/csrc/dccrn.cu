// DCCRN-specific glue: the mask application of DCCRN.forward (DCCRN/DCCRN_cprs.py:201-224): masking_mode 'E' (polar, the
// mode every shipped checkpoint uses), 'C' (complex product) and 'R' (real mask per component).
// Everything else in DCCRN runs on the shared engines: complex convolutions are real
// convolutions on stacked (real | imag) channels (gemm.cu), the complex LSTM is four real
// LSTMs sharing one projection GEMM (gemm_tc.cu + lstm.cu).  HBM-bound elementwise kernel.
#include "common.cuh"

namespace se {

__global__ void __launch_bounds__(256) dccrn_mask_kernel(const float2* __restrict__ m, const float* __restrict__ x_re,
                                                        const float* __restrict__ x_im, long long xsb, long long xst,
                                                        long long xsf, int B, int T, int F, int mode, float* __restrict__ e_re,
                                                        float* __restrict__ e_im, long long esb, long long est,
                                                        long long esf) {
  const long long total = (long long)B * T * F;
  const bool f_fast = esf <= est;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int b, t, f;
    if (f_fast) {
      f = (int)(idx % F);
      const long long bt = idx / F;
      t = (int)(bt % T);
      b = (int)(bt / T);
    } else {
      t = (int)(idx % T);
      const long long bf = idx / T;
      f = (int)(bf % F);
      b = (int)(bf / F);
    }
    float2 out = make_float2(0.f, 0.f);
    if (f > 0) {  // the DC bin of the mask is zero padding (:203-204): tanh(0) * |X| = 0
      const float2 mk = __ldg(m + ((long long)b * T + t) * (F - 1) + (f - 1));
      const float xr = __ldg(x_re + b * xsb + t * xst + f * xsf);
      const float xi = __ldg(x_im + b * xsb + t * xst + f * xsf);
      if (mode == SE_DCCRN_MASK_E) {
        const float mm = sqrtf(mk.x * mk.x + mk.y * mk.y);
        if (mm > 0.f) {
          // est = tanh(|M|) |X| e^{j(angle X + angle M)} = tanh(|M|) * X * M/|M|
          const float g = tanhf(mm) / mm;
          out = make_float2(g * (xr * mk.x - xi * mk.y), g * (xr * mk.y + xi * mk.x));
        }
      } else if (mode == SE_DCCRN_MASK_C) {   // :221-222  X * M
        out = make_float2(xr * mk.x - xi * mk.y, xr * mk.y + xi * mk.x);
      } else {                                // :223-224  (X_r M_r, X_i M_i)
        out = make_float2(xr * mk.x, xi * mk.y);
      }
    }
    e_re[b * esb + t * est + f * esf] = out.x;
    e_im[b * esb + t * est + f * esf] = out.y;
  }
}

}  // namespace se

using namespace se;

extern "C" int se_dccrn_mask_ex(const float* m, const float* x_re, const float* x_im, long long x_sb, long long x_st,
                                long long x_sf, int B, int T, int F, int mode, float* e_re, float* e_im, long long e_sb,
                                long long e_st, long long e_sf, se_stream_t stream) {
  SE_REQUIRE(m && x_re && x_im && e_re && e_im && B > 0 && T > 0 && F > 1, "se_dccrn_mask: bad arguments");
  SE_REQUIRE(mode >= SE_DCCRN_MASK_E && mode <= SE_DCCRN_MASK_R, "se_dccrn_mask: mode %d (0 = E, 1 = C, 2 = R)", mode);
  SE_REQUIRE((((uintptr_t)m) & 7) == 0, "se_dccrn_mask: mask must be 8-byte aligned");
  const long long total = (long long)B * T * F;
  const int blocks = (int)min((long long)148 * 16, ceil_div_ll(total, 256));
  dccrn_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(m), x_re, x_im, x_sb,
                                                              x_st, x_sf, B, T, F, mode, e_re, e_im, e_sb, e_st, e_sf);
  return check_launch("se_dccrn_mask");
}

extern "C" int se_dccrn_mask(const float* m, const float* x_re, const float* x_im, long long x_sb, long long x_st,
                             long long x_sf, int B, int T, int F, float* e_re, float* e_im, long long e_sb,
                             long long e_st, long long e_sf, se_stream_t stream) {
  return se_dccrn_mask_ex(m, x_re, x_im, x_sb, x_st, x_sf, B, T, F, SE_DCCRN_MASK_E, e_re, e_im, e_sb, e_st, e_sf, stream);
}
