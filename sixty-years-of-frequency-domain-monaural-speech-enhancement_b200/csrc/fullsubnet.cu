// FullSubNet glue kernels (FullSubNet/fullsubnet_net_sa/model.py:68-118): everything between the
// two stacked-LSTM sequence models that the reference does with pad / mean / F.unfold / cat /
// reshape.  All HBM-bound elementwise / reduction work; the FLOPs live in lstm.cu and gemm_tc.cu.
//
//   se_fsn_clip_inv_mean   offline_laplace_norm statistics (base_model.py:196-209)
//   se_fsn_fb_input        look-ahead pad + norm + [B,1,F,T] -> time-major [B,T+la,F]  (model.py:79,84)
//   se_fsn_sb_assemble     reflect-unfold(15) of the noisy magnitude ++ full-band output, norm,
//                          TF32 split, laid out [T+la][B*F][32] for the per-step cell GEMM (:88-110)
//   se_fsn_sb_fc           Linear(384,2) on the step's hidden state -> mask[t][b*F+f][2]   (:113-114)
#include "tc_common.cuh"

namespace se {

// out_inv[b] = 1 / ( (sum_{t,f} w[f]*x[b,t,f] + sum extra[b,:]) / denom + 1e-5 )
__global__ void __launch_bounds__(1024) fsn_clip_inv_mean_kernel(const float* __restrict__ x, long long sb,
                                                                long long st, long long sf, int T, int F,
                                                                const float* __restrict__ wgt,
                                                                const float* __restrict__ extra, long long extra_sb,
                                                                long long n_extra, double denom,
                                                                float* __restrict__ out_inv) {
  const int b = blockIdx.x;
  const float* xb = x + (long long)b * sb;
  double acc = 0.0;
  const long long n = (long long)T * F;
  const bool f_fast = sf <= st;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    int t, f;
    if (f_fast) {
      t = (int)(i / F);
      f = (int)(i - (long long)t * F);
    } else {
      f = (int)(i / T);
      t = (int)(i - (long long)f * T);
    }
    const float v = __ldg(xb + (long long)t * st + (long long)f * sf);
    acc += (double)(wgt ? __ldg(wgt + f) * v : v);
  }
  if (extra) {
    const float* eb = extra + (long long)b * extra_sb;
    for (long long i = threadIdx.x; i < n_extra; i += blockDim.x) acc += (double)__ldg(eb + i);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    out_inv[b] = (float)(1.0 / (s / denom + 1e-5));
  }
}

// mag [B,T,F] (any strides) -> mag_tm [B,Tp,F] (zero look-ahead frames) and xn [B,Tp,F] = mag_tm*inv[b]
__global__ void __launch_bounds__(256) fsn_fb_input_kernel(const float* __restrict__ x, long long sb, long long st,
                                                          long long sf, int B, int T, int Tp, int F,
                                                          const float* __restrict__ inv, float* __restrict__ mag_tm,
                                                          float* __restrict__ xn) {
  // 32x32 tiles over (t, f) so that both the strided read and the time-major write coalesce
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool f_fast = sf <= st;
  const float* xb = x + (long long)b * sb;
  for (int r = ty; r < 32; r += 8) {
    // read: fastest index along the input's contiguous dimension
    const int t = f_fast ? t0 + r : t0 + tx;
    const int f = f_fast ? f0 + tx : f0 + r;
    float v = 0.f;
    if (t < T && f < F) v = __ldg(xb + (long long)t * st + (long long)f * sf);
    if (f_fast)
      tile[r][tx] = v;  // [t][f]
    else
      tile[tx][r] = v;
  }
  __syncthreads();
  const float s = __ldg(inv + b);
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, f = f0 + tx;
    if (t < Tp && f < F) {
      const float v = tile[r][tx];
      const long long o = ((long long)b * Tp + t) * F + f;
      mag_tm[o] = v;
      xn[o] = v * s;
    }
  }
}

// out[t][b*F+f][j] = inv[b] * (j < 2n+1 ? mag_tm[b,t,reflect(f+j-n)] : fb[b,t,f]),  split hi/lo.
// One warp per (t, b, f-group of 4): each thread produces a float4 of one row.
template <bool F16>
__global__ void __launch_bounds__(256) fsn_sb_assemble_kernel(const float* __restrict__ mag_tm,
                                                             const float* __restrict__ fb, int B, int Tp, int F,
                                                             int nn, const float* __restrict__ inv,
                                                             float* __restrict__ out_hi, float* __restrict__ out_lo,
                                                             float scale16) {
  const int W = 2 * nn + 2;  // 32 features per row
  const long long rows = (long long)Tp * B * F;
  const long long total = rows * (W / 4);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % (W / 4));
    const long long row = idx / (W / 4);
    const int f = (int)(row % F);
    const long long tb = row / F;
    const int b = (int)(tb % B);
    const int t = (int)(tb / B);
    const float* m = mag_tm + ((long long)b * Tp + t) * F;
    const float s = __ldg(inv + b);
    float hi[4], lo[4];
    unsigned short h16[4], l16[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = q * 4 + e;
      float v;
      if (j < 2 * nn + 1) {
        int ff = f + j - nn;
        if (ff < 0) ff = -ff;
        if (ff >= F) ff = 2 * (F - 1) - ff;
        v = __ldg(m + ff);
      } else {
        v = __ldg(fb + ((long long)b * Tp + t) * F + f);
      }
      if constexpr (F16) split_f16_dev(v * s, scale16, h16[e], l16[e]);
      else split_tf32_dev(v * s, hi[e], lo[e]);
    }
    if constexpr (F16) {
      unsigned short* o_hi = reinterpret_cast<unsigned short*>(out_hi);
      unsigned short* o_lo = reinterpret_cast<unsigned short*>(out_lo);
      *reinterpret_cast<uint2*>(o_hi + row * W + q * 4) = make_uint2(h16[0] | ((unsigned)h16[1] << 16), h16[2] | ((unsigned)h16[3] << 16));
      *reinterpret_cast<uint2*>(o_lo + row * W + q * 4) = make_uint2(l16[0] | ((unsigned)l16[1] << 16), l16[2] | ((unsigned)l16[3] << 16));
    } else {
      *reinterpret_cast<float4*>(out_hi + row * W + q * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(out_lo + row * W + q * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// mask[row][c] = bias[c] + sum_k h[row][k] * W[c][k],  c in {0,1};  one warp per row.
__global__ void __launch_bounds__(256) fsn_sb_fc_kernel(const float* __restrict__ h, int M, int H,
                                                       const float* __restrict__ W, const float* __restrict__ bias,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp; row < M; row += nwarps) {
    const float* hr = h + row * H;
    float a0 = 0.f, a1 = 0.f;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(hr + k);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + k));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + H + k));
      a0 += v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
      a1 += v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
    }
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == 0) {
      out[row * 2] = a0 + __ldg(bias);
      out[row * 2 + 1] = a1 + __ldg(bias + 1);
    }
  }
}

}  // namespace se

using namespace se;

extern "C" int se_fsn_clip_inv_mean(const float* x, long long sb, long long st, long long sf, int B, int T, int F,
                                    const float* wgt, const float* extra, long long extra_sb, long long n_extra,
                                    double denom, float* out_inv, se_stream_t stream) {
  SE_REQUIRE(x && out_inv && B > 0 && T > 0 && F > 0 && denom > 0, "se_fsn_clip_inv_mean: bad arguments");
  fsn_clip_inv_mean_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(x, sb, st, sf, T, F, wgt, extra, extra_sb, n_extra,
                                                                 denom, out_inv);
  return check_launch("se_fsn_clip_inv_mean");
}

extern "C" int se_fsn_fb_input(const float* x, long long sb, long long st, long long sf, int B, int T, int Tp, int F,
                               const float* inv, float* mag_tm, float* xn, se_stream_t stream) {
  SE_REQUIRE(x && inv && mag_tm && xn && B > 0 && T > 0 && Tp >= T && F > 0, "se_fsn_fb_input: bad arguments");
  dim3 grid(ceil_div(Tp, 32), ceil_div(F, 32), B);
  fsn_fb_input_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, sb, st, sf, B, T, Tp, F, inv, mag_tm, xn);
  return check_launch("se_fsn_fb_input");
}

extern "C" int se_fsn_sb_assemble(const float* mag_tm, const float* fb, int B, int Tp, int F, int num_neighbors,
                                  const float* inv, float* out_hi, float* out_lo, se_stream_t stream) {
  SE_REQUIRE(mag_tm && fb && inv && out_hi && out_lo, "se_fsn_sb_assemble: null pointer");
  SE_REQUIRE((2 * num_neighbors + 2) % 4 == 0 && num_neighbors < F, "se_fsn_sb_assemble: neighbours=%d", num_neighbors);
  const long long total = (long long)Tp * B * F * ((2 * num_neighbors + 2) / 4);
  const int blocks = (int)min((long long)148 * 32, ceil_div_ll(total, 256));
  fsn_sb_assemble_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(mag_tm, fb, B, Tp, F, num_neighbors, inv, out_hi,
                                                                          out_lo, 1.0f);
  return check_launch("se_fsn_sb_assemble");
}

extern "C" int se_fsn_sb_assemble_f16(const float* mag_tm, const float* fb, int B, int Tp, int F, int num_neighbors,
                                      const float* inv, int scale_log2, unsigned short* out_hi, unsigned short* out_lo,
                                      se_stream_t stream) {
  SE_REQUIRE(mag_tm && fb && inv && out_hi && out_lo, "se_fsn_sb_assemble_f16: null pointer");
  SE_REQUIRE((2 * num_neighbors + 2) % 8 == 0 && num_neighbors < F, "se_fsn_sb_assemble_f16: neighbours=%d", num_neighbors);
  const long long total = (long long)Tp * B * F * ((2 * num_neighbors + 2) / 4);
  const int blocks = (int)min((long long)148 * 32, ceil_div_ll(total, 256));
  fsn_sb_assemble_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(
      mag_tm, fb, B, Tp, F, num_neighbors, inv, reinterpret_cast<float*>(out_hi), reinterpret_cast<float*>(out_lo),
      ldexpf(1.0f, scale_log2));
  return check_launch("se_fsn_sb_assemble_f16");
}

extern "C" int se_fsn_sb_fc(const float* h, int M, int H, const float* W, const float* bias, float* out,
                            se_stream_t stream) {
  SE_REQUIRE(h && W && bias && out && M > 0 && H > 0 && H % 4 == 0, "se_fsn_sb_fc: bad arguments");
  const int blocks = (int)min((long long)148 * 8, ceil_div_ll((long long)M * 32, 256));
  fsn_sb_fc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(h, M, H, W, bias, out);
  return check_launch("se_fsn_sb_fc");
}
