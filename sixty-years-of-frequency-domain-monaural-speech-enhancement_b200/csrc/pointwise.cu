// Small elementwise kernels shared by the CRN-family models (HBM-bound).
//   se_glu_affine_act  GluConv2d / GluConvTranspose2d tail (GCRN/GCRN_noncprs.py:42-83,138-157):
//                      y = act( (a * sigmoid(b)) * scale[c] + shift[c] ),  x rows = [a (C) | b (C)]
//   se_unary           y = act(x)   (the second ELU the GCRN decoder applies to its skip inputs, :149-152)
//   se_cmul            complex ratio mask applied inside DPCRN's forward (DPCRN/DPCRN.py:33-42), interleaved (re, im)
// The first two can also emit the TF32 split of y for a following tensor-core layer.
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

// C % 4 == 0: four channels per thread, 16-byte accesses
__global__ void __launch_bounds__(256) glu_affine_act_vec_kernel(const float4* __restrict__ x, long long rows, int C4,
                                                                const float4* __restrict__ scale,
                                                                const float4* __restrict__ shift, int act,
                                                                float act_param, float4* __restrict__ out,
                                                                float4* __restrict__ out_hi, float4* __restrict__ out_lo) {
  const long long n = rows * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C4;
    const int c = (int)(i - r * C4);
    const float4 a = __ldg(x + r * 2 * C4 + c), b = __ldg(x + r * 2 * C4 + C4 + c);
    const float4 sc = scale ? __ldg(scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sh = shift ? __ldg(shift + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
    float y[4], hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      y[e] = apply_act(av[e] * sigmoid_f(bv[e]) * scv[e] + shv[e], act, act_param);
      if (out_hi) split_tf32_dev(y[e], hi[e], lo[e]);
    }
    if (out) out[i] = make_float4(y[0], y[1], y[2], y[3]);
    if (out_hi) {
      out_hi[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
      out_lo[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
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

__global__ void __launch_bounds__(256) cmul_kernel(const float2* __restrict__ x, const float2* __restrict__ m, long long n,
                                                  float2* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 a = __ldg(x + i), b = __ldg(m + i);
    out[i] = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
  }
}

}  // namespace se

using namespace se;

extern "C" int se_cmul(const float* x, const float* m, long long n, float* out, se_stream_t stream) {
  SE_REQUIRE(x && m && out && n > 0, "se_cmul: bad arguments");
  SE_REQUIRE(((((uintptr_t)x) | ((uintptr_t)m) | ((uintptr_t)out)) & 7) == 0, "se_cmul: pointers must be 8-byte aligned");
  const int blocks = (int)min((long long)148 * 16, ceil_div_ll(n, 256));
  cmul_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x),
                                                        reinterpret_cast<const float2*>(m), n,
                                                        reinterpret_cast<float2*>(out));
  return check_launch("se_cmul");
}

extern "C" int se_glu_affine_act(const float* x, long long rows, int C, const float* scale, const float* shift, int act,
                                 float act_param, float* out, float* out_hi, float* out_lo, se_stream_t stream) {
  SE_REQUIRE(x && rows > 0 && C > 0 && (out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)),
             "se_glu_affine_act: bad arguments");
  const long long n = rows * C;
  auto al16 = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
  if ((C & 3) == 0 && al16(x) && al16(scale) && al16(shift) && al16(out) && al16(out_hi) && al16(out_lo)) {
    const int blocks = (int)min((long long)148 * 16, ceil_div_ll(n / 4, 256));
    glu_affine_act_vec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), rows, C / 4, reinterpret_cast<const float4*>(scale),
        reinterpret_cast<const float4*>(shift), act, act_param, reinterpret_cast<float4*>(out),
        reinterpret_cast<float4*>(out_hi), reinterpret_cast<float4*>(out_lo));
    return check_launch("se_glu_affine_act");
  }
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
