// se_conv_f16x3: the tensor-core implicit-GEMM convolution on fp16 operand pairs (tcgen05 kind::f16; kernel:
// conv_tc_impl.cuh with F16 = true).  A separate translation unit so that the two instantiation sets compile in parallel.
#include "conv_tc_impl.cuh"

using namespace se;

extern "C" int se_conv_f16x3(const se_conv_f16_desc* d, se_stream_t stream) {
  SE_REQUIRE(d != nullptr, "se_conv_f16x3: null descriptor");
  ConvTcArgs a{};
  a.src0_hi = d->src0_hi, a.src0_lo = d->src0_lo, a.src1_hi = d->src1_hi, a.src1_lo = d->src1_lo;
  a.C0 = d->C0, a.C1 = d->C1, a.B = d->B, a.T = d->T, a.Fin = d->Fin, a.Fout = d->Fout, a.ntaps = d->ntaps;
  a.dt = d->dt, a.df = d->df, a.sf = d->sf;
  a.w_hi = d->w_hi, a.w_lo = d->w_lo, a.bias = d->bias, a.Cout = d->Cout, a.act = d->act, a.act_param = d->act_param;
  a.out = d->out, a.out_hi = d->out_hi, a.out_lo = d->out_lo, a.out16_hi = d->out16_hi, a.out16_lo = d->out16_lo;
  a.out_scale = ldexpf(1.0f, -(d->scale_log2_a + d->scale_log2_w));
  a.out16_scale = ldexpf(1.0f, d->out16_scale_log2);
  a.dstF = d->dstF, a.dst_f0 = d->dst_f0, a.dst_fstep = d->dst_fstep;
  a.glu = d->glu, a.glu_scale = d->glu_scale, a.glu_shift = d->glu_shift;
  a.ncls = d->ncls, a.fout1 = d->fout1;
  return conv_tc_run<true>(&a, "se_conv_f16x3", (cudaStream_t)stream);
}
