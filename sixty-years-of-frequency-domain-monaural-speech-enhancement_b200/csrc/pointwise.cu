// Small elementwise kernels shared by the CRN-family models (HBM-bound).
//   se_glu_affine_act  GluConv2d / GluConvTranspose2d tail (GCRN/GCRN_noncprs.py:42-83,138-157):
//                      y = act( (a * sigmoid(b)) * scale[c] + shift[c] ),  x rows = [a (C) | b (C)]
//   se_unary           y = act(x)   (the second ELU the GCRN decoder applies to its skip inputs, :149-152)
// Both can also emit the TF32 split of y for a following tensor-core layer.
#include "tc_common.cuh"

namespace se {

__global__ void __launch_bounds__(256) glu_affine_act_kernel(const float* __restrict__ x, long long rows, int C,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int act, float act_param,
                                                            float* __restrict__ out, float* __restrict__ out_hi,
                                                            float* __restrict__ out_lo) {
  const long long n = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float a = __ldg(x + r * 2 * C + c), b = __ldg(x + r * 2 * C + C + c);
    float y = a * sigmoid_f(b);
    y = y * (scale ? __ldg(scale + c) : 1.f) + (shift ? __ldg(shift + c) : 0.f);
    y = apply_act(y, act, act_param);
    if (out) out[i] = y;
    if (out_hi) split_tf32_dev(y, out_hi[i], out_lo[i]);
  }
}

__global__ void __launch_bounds__(256) unary_kernel(const float* __restrict__ x, long long n, int act, float act_param,
                                                   float* __restrict__ out, float* __restrict__ out_hi,
                                                   float* __restrict__ out_lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float y = apply_act(__ldg(x + i), act, act_param);
    if (out) out[i] = y;
    if (out_hi) split_tf32_dev(y, out_hi[i], out_lo[i]);
  }
}

}  // namespace se

using namespace se;

extern "C" int se_glu_affine_act(const float* x, long long rows, int C, const float* scale, const float* shift, int act,
                                 float act_param, float* out, float* out_hi, float* out_lo, se_stream_t stream) {
  SE_REQUIRE(x && rows > 0 && C > 0 && (out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)),
             "se_glu_affine_act: bad arguments");
  const long long n = rows * C;
  const int blocks = (int)min((long long)148 * 16, ceil_div_ll(n, 256));
  glu_affine_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, C, scale, shift, act, act_param, out, out_hi,
                                                                  out_lo);
  return check_launch("se_glu_affine_act");
}

extern "C" int se_unary(const float* x, long long n, int act, float act_param, float* out, float* out_hi, float* out_lo,
                        se_stream_t stream) {
  SE_REQUIRE(x && n > 0 && (out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)), "se_unary: bad arguments");
  const int blocks = (int)min((long long)148 * 16, ceil_div_ll(n, 256));
  unary_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, act, act_param, out, out_hi, out_lo);
  return check_launch("se_unary");
}
