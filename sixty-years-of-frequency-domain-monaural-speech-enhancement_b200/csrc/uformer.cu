// Uformer-specific kernels (Uformer/uformer.py, dilated_dualpath_conformer.py and friends).
// Convolutions and all linear layers run on the shared conv / GEMM engines; what is left is
// elementwise / normalisation / small-d attention work, all HBM- or latency-bound:
//   se_uf_prep         magnitude / phase split of the noisy spectrum with the reference's eps rules
//   se_uf_fusion       cross-branch fusion (fusion.py:13-19)
//   se_group_layernorm nn.LayerNorm over the channel axis of a channels-last tensor, per (re|im)
//                      group, with optional gate pre-op, PReLU / swish post-op, residual, TF32 split
//   se_attention       softmax(QK^T/sqrt(d))V for d = 16 heads with signed head combination
//                      (t_att_cplx.py:58-67: real = A-B-C-D, imag = E+F+G-H), over T or over F
//   se_uf_mask         sigmoid magnitude branch + polar mask branch, averaged (uformer.py:236-262)
#include <float.h>

#include "tc_common.cuh"

namespace se {

#define UF_EPS 1.1920928955078125e-07f  // torch.finfo(torch.float32).eps

// ------------------------------------------------------------------------------------------------
// x [B,T,F,2] (re,im interleaved) -> mag [B,T,F], phase [B,T,F], and the network inputs for bins 1..F-1:
// cplx_in [B,T,F-1,2] = (mag cos(phase), mag sin(phase)), mag_in [B,T,F-1]      (uformer.py:197-210)
__global__ void __launch_bounds__(256) uf_prep_kernel(const float2* __restrict__ x, long long n, int F,
                                                     float* __restrict__ mag, float* __restrict__ phase,
                                                     float2* __restrict__ cplx_in, float* __restrict__ mag_in) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = __ldg(x + i);
    const float m = sqrtf(fmaxf(v.x * v.x + v.y * v.y, UF_EPS));
    const float ph = atan2f(v.y + UF_EPS, v.x);
    mag[i] = m;
    phase[i] = ph;
    const int f = (int)(i % F);
    if (f > 0) {
      const long long o = (i / F) * (F - 1) + (f - 1);
      float s, c;
      sincosf(ph, &s, &c);
      cplx_in[o] = make_float2(m * c, m * s);
      mag_in[o] = m;
    }
  }
}

// c [rows, 2C] = (re C | im C), m [rows, C]:  m' = m + sigmoid(|c|);  c' = c + sigmoid(m)   (fusion.py:13-19)
// V channels per thread (V = 4: float4 traffic, C % 4 == 0).  Each result goes out as fp32 and / or as the TF32 (hi, lo)
// pair the tensor-core convs read, so the U-Net levels need no separate split pass after the fusion.
struct FusionParams {
  const float *c, *m;
  long long rows;
  int C;
  float *c_out, *c_hi, *c_lo, *m_out, *m_hi, *m_lo;
};

template <int V>
__device__ __forceinline__ void fusion_store(float* out, float* hi, float* lo, long long off, const float (&v)[V]) {
  if constexpr (V == 4) {
    if (out) *reinterpret_cast<float4*>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
    if (hi) {
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32_dev(v[e], h[e], l[e]);
      *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
  } else {
    if (out) out[off] = v[0];
    if (hi) split_tf32_dev(v[0], hi[off], lo[off]);
  }
}

template <int V>
__global__ void __launch_bounds__(256) uf_fusion_kernel(const FusionParams p) {
  const int C = p.C, CV = C / V;
  const long long n = p.rows * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / CV;
    const int ch = (int)(i - r * CV) * V;
    float cr[V], ci[V], mv[V];
    if constexpr (V == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.c + r * 2 * C + ch));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.c + r * 2 * C + C + ch));
      const float4 d = __ldg(reinterpret_cast<const float4*>(p.m + r * C + ch));
      cr[0] = a.x; cr[1] = a.y; cr[2] = a.z; cr[3] = a.w;
      ci[0] = b.x; ci[1] = b.y; ci[2] = b.z; ci[3] = b.w;
      mv[0] = d.x; mv[1] = d.y; mv[2] = d.z; mv[3] = d.w;
    } else {
      cr[0] = __ldg(p.c + r * 2 * C + ch);
      ci[0] = __ldg(p.c + r * 2 * C + C + ch);
      mv[0] = __ldg(p.m + r * C + ch);
    }
    float mo[V], cro[V], cio[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float cm = sqrtf(fmaxf(cr[e] * cr[e] + ci[e] * ci[e], UF_EPS));
      const float sg = sigmoid_f(mv[e]);
      mo[e] = mv[e] + sigmoid_f(cm);
      cro[e] = cr[e] + sg;
      cio[e] = ci[e] + sg;
    }
    fusion_store<V>(p.m_out, p.m_hi, p.m_lo, r * C + ch, mo);
    fusion_store<V>(p.c_out, p.c_hi, p.c_lo, r * 2 * C + ch, cro);
    fusion_store<V>(p.c_out, p.c_hi, p.c_lo, r * 2 * C + C + ch, cio);
  }
}

// One warp per (row, group): LayerNorm over C contiguous channels, gamma/beta [C] shared by the groups.
//   v = x (* sigmoid(gate) if gate);  y = (v - mean)/sqrt(var + eps) * gamma + beta
//   post: 0 none, 1 PReLU(slope), 2 swish (y * sigmoid(y));  out = post(y) (+ res)
struct LnParams {
  const float *x, *gate, *gamma, *beta, *res;
  long long rows;
  int G, C;
  float eps, slope;
  int post;
  float *out, *out_hi, *out_lo;
  const int* out_index;
};

template <int PER_MAX>   // channels per lane: C <= 32 * PER_MAX
__global__ void __launch_bounds__(256) group_layernorm_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = p.rows * p.G;
  for (long long w = warp; w < total; w += nwarps) {
    const long long base = w * p.C;
    float v[PER_MAX];
    float s = 0.f;
    const int per = (p.C + 31) / 32;
#pragma unroll
    for (int i = 0; i < PER_MAX; ++i) {
      if (i >= per) break;
      const int ch = lane + 32 * i;
      float t = 0.f;
      if (ch < p.C) {
        t = __ldg(p.x + base + ch);
        if (p.gate) t *= sigmoid_f(__ldg(p.gate + base + ch));
      }
      v[i] = t;
      s += t;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)p.C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER_MAX; ++i) {
      if (i >= per) break;
      const int ch = lane + 32 * i;
      const float d = ch < p.C ? v[i] - mean : 0.f;
      q += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)p.C + p.eps);
#pragma unroll
    for (int i = 0; i < PER_MAX; ++i) {
      if (i >= per) break;
      const int ch = lane + 32 * i;
      if (ch >= p.C) continue;
      float y = (v[i] - mean) * rstd * __ldg(p.gamma + ch) + __ldg(p.beta + ch);
      if (p.post == 1)
        y = y >= 0.f ? y : p.slope * y;
      else if (p.post == 2)
        y = y * sigmoid_f(y);
      if (p.res) y += __ldg(p.res + base + ch);
      const long long o = base + (p.out_index ? __ldg(p.out_index + ch) : ch);
      if (p.out) p.out[o] = y;
      if (p.out_hi) {
        float hi, lo;
        split_tf32_dev(y, hi, lo);
        p.out_hi[o] = hi;
        p.out_lo[o] = lo;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention with head dim 16.  qkv [R, nheads*48] rows = (q16 | k16 | v16) per head.
struct AttnParams {
  const float* qkv;
  int ld, nheads, nout;
  int head_out[8];
  float head_sign[8];
  int L;
  long long lstride;          // rows between consecutive positions of a sequence
  int n_inner;
  long long outer_stride, inner_stride;   // first row of sequence (o, i) = o*outer_stride + i*inner_stride
  float scale;
  float* out;
  int ldo;
};

// long sequences (attention over T): one CTA = one sequence x 128 queries; K/V of one head in smem
__global__ void __launch_bounds__(128) attention_long_kernel(const AttnParams p) {
  extern __shared__ float kv[];  // K [L][16], V [L][16]
  float* Ks = kv;
  float* Vs = kv + (size_t)p.L * 16;
  const int seq = blockIdx.y;
  const int o = seq / p.n_inner, i = seq - o * p.n_inner;
  const long long row0 = (long long)o * p.outer_stride + (long long)i * p.inner_stride;
  const int l = blockIdx.x * 128 + threadIdx.x;
  const bool active = l < p.L;
  float acc_out[2][16];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int d = 0; d < 16; ++d) acc_out[a][d] = 0.f;
  for (int h = 0; h < p.nheads; ++h) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < p.L * 8; idx += 128) {  // float4 pieces: 4 for K + 4 for V per position
      const int pos = idx >> 3, part = idx & 7;
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.qkv + (row0 + (long long)pos * p.lstride) * p.ld + h * 48 +
                                                             16 + part * 4));
      float* dst = (part < 4) ? (Ks + pos * 16 + part * 4) : (Vs + pos * 16 + (part - 4) * 4);
      *reinterpret_cast<float4*>(dst) = v;
    }
    __syncthreads();
    if (active) {
      float q[16];
      const float* qp = p.qkv + (row0 + (long long)l * p.lstride) * p.ld + h * 48;
#pragma unroll
      for (int d = 0; d < 16; d += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(qp + d));
        q[d] = v.x * p.scale; q[d + 1] = v.y * p.scale; q[d + 2] = v.z * p.scale; q[d + 3] = v.w * p.scale;
      }
      float mx = -FLT_MAX, den = 0.f, acc[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[d] = 0.f;
      for (int j = 0; j < p.L; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + j * 16);
        float s = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < 4; ++d4) {
          const float4 k = kr[d4];
          s += q[4 * d4] * k.x + q[4 * d4 + 1] * k.y + q[4 * d4 + 2] * k.z + q[4 * d4 + 3] * k.w;
        }
        if (s > mx) {
          const float corr = expf(mx - s);
          den *= corr;
#pragma unroll
          for (int d = 0; d < 16; ++d) acc[d] *= corr;
          mx = s;
        }
        const float pj = expf(s - mx);
        den += pj;
        const float4* vr = reinterpret_cast<const float4*>(Vs + j * 16);
#pragma unroll
        for (int d4 = 0; d4 < 4; ++d4) {
          const float4 v = vr[d4];
          acc[4 * d4] += pj * v.x; acc[4 * d4 + 1] += pj * v.y; acc[4 * d4 + 2] += pj * v.z; acc[4 * d4 + 3] += pj * v.w;
        }
      }
      const float inv = p.head_sign[h] / den;
      const int oo = p.head_out[h];
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        if (oo == 0)
          acc_out[0][d] += acc[d] * inv;
        else
          acc_out[1][d] += acc[d] * inv;
      }
    }
  }
  if (active) {
    float* op = p.out + (row0 + (long long)l * p.lstride) * p.ldo;
    for (int a = 0; a < p.nout; ++a)
#pragma unroll
      for (int d = 0; d < 16; d += 4)
        *reinterpret_cast<float4*>(op + a * 16 + d) =
            make_float4(acc_out[a][d], acc_out[a][d + 1], acc_out[a][d + 2], acc_out[a][d + 3]);
  }
}

// short sequences (attention over F = 4): one thread per (sequence, query)
__global__ void __launch_bounds__(128) attention_short_kernel(const AttnParams p, long long nseq) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nseq * p.L) return;
  const long long seq = idx / p.L;
  const int l = (int)(idx - seq * p.L);
  const long long o = seq / p.n_inner, i = seq - o * p.n_inner;
  const long long row0 = o * p.outer_stride + i * p.inner_stride;
  float acc_out[2][16];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int d = 0; d < 16; ++d) acc_out[a][d] = 0.f;
  for (int h = 0; h < p.nheads; ++h) {
    float q[16];
    const float* qp = p.qkv + (row0 + (long long)l * p.lstride) * p.ld + h * 48;
#pragma unroll
    for (int d = 0; d < 16; ++d) q[d] = __ldg(qp + d) * p.scale;
    float s[8];
    float mx = -FLT_MAX;
    for (int j = 0; j < p.L; ++j) {
      const float* kp = p.qkv + (row0 + (long long)j * p.lstride) * p.ld + h * 48 + 16;
      float t = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) t += q[d] * __ldg(kp + d);
      s[j] = t;
      mx = fmaxf(mx, t);
    }
    float den = 0.f;
    for (int j = 0; j < p.L; ++j) {
      s[j] = expf(s[j] - mx);
      den += s[j];
    }
    const float inv = p.head_sign[h] / den;
    const int oo = p.head_out[h];
    for (int j = 0; j < p.L; ++j) {
      const float* vp = p.qkv + (row0 + (long long)j * p.lstride) * p.ld + h * 48 + 32;
      const float w = s[j] * inv;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float v = w * __ldg(vp + d);
        if (oo == 0)
          acc_out[0][d] += v;
        else
          acc_out[1][d] += v;
      }
    }
  }
  float* op = p.out + (row0 + (long long)l * p.lstride) * p.ldo;
  for (int a = 0; a < p.nout; ++a)
#pragma unroll
    for (int d = 0; d < 16; ++d) op[a * 16 + d] = acc_out[a][d];
}

// ------------------------------------------------------------------------------------------------
// final recombination (uformer.py:236-262).  cmask [B,T,F-1,2], mdec [B,T,F-1], mag/phase [B,T,F] -> est [B,T,F,2]
__global__ void __launch_bounds__(256) uf_mask_kernel(const float2* __restrict__ cmask, const float* __restrict__ mdec,
                                                     const float* __restrict__ mag, const float* __restrict__ phase,
                                                     long long n, int F, float2* __restrict__ est) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const float m_in = __ldg(mag + i), ph = __ldg(phase + i);
    float mmag = 0.f, mph = 0.f, msig = 0.f;  // zero padding at the DC bin (F.pad after the nonlinearities)
    if (f > 0) {
      const long long o = (i / F) * (F - 1) + (f - 1);
      const float2 mk = __ldg(cmask + o);
      const float mm = sqrtf(fmaxf(mk.x * mk.x + mk.y * mk.y, UF_EPS));
      const float rp = mk.x / (mm + UF_EPS), ip = mk.y / (mm + UF_EPS);
      mmag = tanhf(mm + UF_EPS);
      mph = atan2f(ip + UF_EPS, rp);
      msig = sigmoid_f(__ldg(mdec + o));
    }
    const float est_mag = 0.5f * (mmag * m_in + msig * m_in);
    float s, c;
    sincosf(ph + mph, &s, &c);
    est[i] = make_float2(est_mag * c, est_mag * s);
  }
}

static int grid_for(long long n, int threads) { return (int)min((long long)148 * 16, ceil_div_ll(n, threads)); }

}  // namespace se

using namespace se;

extern "C" int se_uf_prep(const float* x, int B, int T, int F, float* mag, float* phase, float* cplx_in, float* mag_in,
                          se_stream_t stream) {
  SE_REQUIRE(x && mag && phase && cplx_in && mag_in && B > 0 && T > 0 && F > 1, "se_uf_prep: bad arguments");
  const long long n = (long long)B * T * F;
  uf_prep_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x), n, F, mag, phase,
                                                                     reinterpret_cast<float2*>(cplx_in), mag_in);
  return check_launch("se_uf_prep");
}

extern "C" int se_uf_fusion_ex(const float* c, const float* m, long long rows, int C, float* c_out, float* c_hi,
                               float* c_lo, float* m_out, float* m_hi, float* m_lo, se_stream_t stream) {
  SE_REQUIRE(c && m && rows > 0 && C > 0, "se_uf_fusion: bad arguments");
  SE_REQUIRE((c_out || c_hi) && (m_out || m_hi), "se_uf_fusion: every result needs an fp32 or a split output");
  SE_REQUIRE((c_hi == nullptr) == (c_lo == nullptr) && (m_hi == nullptr) == (m_lo == nullptr),
             "se_uf_fusion: hi / lo outputs go together");
  const FusionParams p{c, m, rows, C, c_out, c_hi, c_lo, m_out, m_hi, m_lo};
  const uintptr_t all = (uintptr_t)c | (uintptr_t)m | (uintptr_t)c_out | (uintptr_t)c_hi | (uintptr_t)c_lo |
                        (uintptr_t)m_out | (uintptr_t)m_hi | (uintptr_t)m_lo;
  if ((C & 3) == 0 && (all & 15) == 0)
    uf_fusion_kernel<4><<<grid_for(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(p);
  else
    uf_fusion_kernel<1><<<grid_for(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("se_uf_fusion");
}

extern "C" int se_uf_fusion(const float* c, const float* m, long long rows, int C, float* c_out, float* m_out,
                            se_stream_t stream) {
  return se_uf_fusion_ex(c, m, rows, C, c_out, nullptr, nullptr, m_out, nullptr, nullptr, stream);
}

extern "C" int se_group_layernorm(const float* x, const float* gate, long long rows, int G, int C, const float* gamma,
                                  const float* beta, float eps, int post, float slope, const float* res,
                                  const int* out_index, float* out, float* out_hi, float* out_lo, se_stream_t stream) {
  SE_REQUIRE(x && gamma && beta && rows > 0 && G > 0 && C > 0 && C <= 1024, "se_group_layernorm: bad arguments (C=%d)", C);
  SE_REQUIRE(out || out_hi, "se_group_layernorm: no output");
  SE_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "se_group_layernorm: out_hi/out_lo go together");
  LnParams p{x, gate, gamma, beta, res, rows, G, C, eps, slope, post, out, out_hi, out_lo, out_index};
  if (C <= 256)
    group_layernorm_kernel<8><<<grid_for(rows * G * 32, 256), 256, 0, (cudaStream_t)stream>>>(p);
  else
    group_layernorm_kernel<32><<<grid_for(rows * G * 32, 256), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("se_group_layernorm");
}

extern "C" int se_attention(const float* qkv, int ld, int nheads, const int* head_out, const float* head_sign, int nout,
                            int L, long long lstride, int n_outer, long long outer_stride, int n_inner,
                            long long inner_stride, float scale, float* out, int ldo, se_stream_t stream) {
  SE_REQUIRE(qkv && out && head_out && head_sign, "se_attention: null pointer");
  SE_REQUIRE(nheads >= 1 && nheads <= 8 && nout >= 1 && nout <= 2 && ld >= nheads * 48 && (ld & 3) == 0 && (ldo & 3) == 0,
             "se_attention: nheads=%d nout=%d ld=%d", nheads, nout, ld);
  SE_REQUIRE(L >= 1 && n_outer >= 1 && n_inner >= 1, "se_attention: bad sequence geometry");
  AttnParams p{};
  p.qkv = qkv; p.ld = ld; p.nheads = nheads; p.nout = nout;
  for (int h = 0; h < nheads; ++h) {
    SE_REQUIRE(head_out[h] >= 0 && head_out[h] < nout, "se_attention: head_out[%d]=%d", h, head_out[h]);
    p.head_out[h] = head_out[h];
    p.head_sign[h] = head_sign[h];
  }
  p.L = L; p.lstride = lstride; p.n_inner = n_inner; p.outer_stride = outer_stride; p.inner_stride = inner_stride;
  p.scale = scale; p.out = out; p.ldo = ldo;
  const long long nseq = (long long)n_outer * n_inner;
  cudaStream_t s = (cudaStream_t)stream;
  if (L <= 8) {
    attention_short_kernel<<<(unsigned)ceil_div_ll(nseq * L, 128), 128, 0, s>>>(p, nseq);
  } else {
    const size_t smem = (size_t)L * 32 * sizeof(float);
    SE_REQUIRE(smem <= 200 * 1024 && nseq < 65536, "se_attention: L=%d / %lld sequences too large", L, nseq);
    cudaError_t e = cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("se_attention: smem attribute: %s", cudaGetErrorString(e));
      return SE_ERR_CUDA;
    }
    attention_long_kernel<<<dim3(ceil_div(L, 128), (unsigned)nseq), 128, smem, s>>>(p);
  }
  return check_launch("se_attention");
}

extern "C" int se_uf_mask(const float* cmask, const float* mdec, const float* mag, const float* phase, int B, int T, int F,
                          float* est, se_stream_t stream) {
  SE_REQUIRE(cmask && mdec && mag && phase && est && B > 0 && T > 0 && F > 1, "se_uf_mask: bad arguments");
  const long long n = (long long)B * T * F;
  uf_mask_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(cmask), mdec, mag,
                                                                     phase, n, F, reinterpret_cast<float2*>(est));
  return check_launch("se_uf_mask");
}
