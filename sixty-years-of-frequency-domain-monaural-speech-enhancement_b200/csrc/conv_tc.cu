// se_conv_tf32x3: the tensor-core implicit-GEMM convolution on TF32 operand pairs (kernel: conv_tc_impl.cuh).
#include "conv_tc_impl.cuh"

using namespace se;

extern "C" int se_conv_tf32x3(const se_conv_tc_desc* d, se_stream_t stream) {
  SE_REQUIRE(d != nullptr, "se_conv_tf32x3: null descriptor");
  ConvTcArgs a{};
  a.src0_hi = d->src0_hi, a.src0_lo = d->src0_lo, a.src1_hi = d->src1_hi, a.src1_lo = d->src1_lo;
  a.C0 = d->C0, a.C1 = d->C1, a.B = d->B, a.T = d->T, a.Fin = d->Fin, a.Fout = d->Fout, a.ntaps = d->ntaps;
  a.dt = d->dt, a.df = d->df, a.sf = d->sf;
  a.w_hi = d->w_hi, a.w_lo = d->w_lo, a.bias = d->bias, a.Cout = d->Cout, a.act = d->act, a.act_param = d->act_param;
  a.out = d->out, a.out_hi = d->out_hi, a.out_lo = d->out_lo;
  a.out_scale = 1.0f, a.out16_scale = 1.0f;
  a.dstF = d->dstF, a.dst_f0 = d->dst_f0, a.dst_fstep = d->dst_fstep;
  a.glu = d->glu, a.glu_scale = d->glu_scale, a.glu_shift = d->glu_shift;
  return conv_tc_run<false>(&a, "se_conv_tf32x3", (cudaStream_t)stream);
}
