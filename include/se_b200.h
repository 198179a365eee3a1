/*
 * se_b200.h -- C ABI of libse_b200.so: the B200 (sm_100a) decode hot path
 *              STFT -> mask/mapping network -> recombine -> overlap-add iSTFT.
 *
 * The reference (cszheng-ioa/Sixty-years-of-frequency-domain-monaural-speech-enhancement)
 * is pure Python and has NO FFI of its own; the boundary its
 * decode scripts would bind is therefore defined here, one entry point per step of
 * the shared decode loop, each citing the reference lines it replaces.  All pointers
 * are DEVICE pointers owned by the caller unless the name says "host"; nothing in
 * this library allocates device memory except se_*_plan objects (none yet); every
 * launch takes the caller's cudaStream_t (passed as void*).  No C++ types and no
 * exceptions cross this boundary.  Every function returns 0 (SE_OK) or a negative
 * se_status; se_last_error() gives the message for the calling thread.
 */
#ifndef SE_B200_H_
#define SE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE_B200_ABI_VERSION 1

typedef enum se_status {
  SE_OK = 0,
  SE_ERR_SHAPE = -1, /* bad argument / unsupported geometry */
  SE_ERR_ARCH = -2,  /* current device is not sm_100 */
  SE_ERR_CUDA = -3   /* CUDA runtime error, see se_last_error() */
} se_status;

typedef enum se_act {
  SE_ACT_NONE = 0,
  SE_ACT_ELU = 1,      /* nn.ELU(alpha=1)              CRN/CRN.py:43 */
  SE_ACT_SOFTPLUS = 2, /* nn.Softplus(beta=1,thr=20)   CRN/CRN.py:102, LSTM/LSTM.py:22 */
  SE_ACT_RELU = 3,
  SE_ACT_SIGMOID = 4,
  SE_ACT_TANH = 5,
  SE_ACT_PRELU = 6     /* nn.PReLU() with one shared slope = act_param   DCCRN/DCCRN_cprs.py:76 */
} se_act;

typedef void* se_stream_t; /* cudaStream_t */

int se_abi_version(void);
const char* se_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, else SE_ERR_ARCH. */
int se_device_check(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long se_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * a1  RMS normalisation constant.  Replaces `c = np.sqrt(len(x)/np.sum(x**2)); x = x*c`
 *     (CRN/crn_decode.py:39-40, LSTM/lstm_decode_vb.py:35-36, DCCRN/dccrn_decode.py:31-32,
 *     FullSubNet/fullsubnet_sa_decode.py:46-47, Uformer/uformer_decode_vb.py:35-36).
 *     wav [B, N] fp32 (row stride wav_stride floats).  Writes c[b] and inv_c[b] = 1/c[b]
 *     (sum of squares accumulated in fp64).  The scaled waveform is never materialised:
 *     se_stft() takes c as `scale`, se_istft() takes inv_c as `out_scale`.
 *     reciprocal != 0 selects the G2Net convention c = sqrt(sum x^2 / N)
 *     (G2Net_new/com_decode.py:43-44): then c[b] holds 1/that and inv_c[b] that, so the
 *     same two downstream arguments apply.
 * ------------------------------------------------------------------------------------- */
int se_rms_scale(const float* wav, long long wav_stride, int B, int N, int reciprocal, float* c, float* inv_c,
                 se_stream_t stream);
/* Tail-padded batches of clips of DIFFERENT lengths (the scripts decode any directory one file
 * at a time, CRN/crn_decode_vb.py:31-33): `lengths` (device int32 [B], or NULL = every clip has N samples) gives the
 * sample count of each row; the statistics of clip b run over its own lengths[b] samples. */
int se_rms_scale_len(const float* wav, long long wav_stride, int B, int N, const int* lengths, int reciprocal, float* c,
                     float* inv_c, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a3+a4  Fused reflect-pad + framing + periodic-Hann + one-sided FFT + feature split.
 *     Replaces librosa.stft(x, n_fft, hop, window='hanning') / torch.stft(x, n_fft, hop, win,
 *     hann_window(win)) with center=True (CRN/crn_decode.py:41, DCCRN/dccrn_decode.py:41,
 *     FullSubNet/fullsubnet_sa_decode.py:53, Uformer/uformer.py:178) AND the feature split
 *     that follows it: |X|**p (crn_decode.py:44), |X|**p * cos/sin(angle)
 *     (gcrn_decode.py:45-49, dccrn_decode.py:44-46).
 *     wav [B,N]; scale [B] or NULL (multiplies the waveform, i.e. the `* c` of a1).
 *     Supported (n_fft, win, hop): n_fft in {320, 512}, win <= n_fft (window centred,
 *     zero padded), hop even, hop <= n_fft.  T must equal 1 + N/hop.
 *     Up to three output planes, any may be NULL (re/im together), addressed as
 *        mag[b*msb + t*mst + f*msf]  and  re|im[b*sb + t*st + f*sf]
 *     (strides in floats; F = n_fft/2+1 bins):
 *        mag = |X|^p_mag ; re,im = X * |X|^(p_ri-1)  (p_ri = 1: the plain spectrum).
 *     An interleaved complex64 [B,T,F] tensor is (re = base, im = base+1, sb=2TF, st=2F, sf=2).
 * ------------------------------------------------------------------------------------- */
int se_stft(const float* wav, long long wav_stride, int B, int N, const float* scale, int n_fft, int win, int hop,
            int T, float* mag, long long msb, long long mst, long long msf, float* re, float* im, long long sb,
            long long st, long long sf, float p_mag, float p_ri, se_stream_t stream);
/* se_stft on a tail-padded batch: clip b is reflect-padded around ITS last sample lengths[b] - 1 and owns the frames
 * t < 1 + lengths[b] / hop, bit-identical to the frames of that clip transformed alone; its later frames (up to
 * T = 1 + N / hop of the padded row) are written as zeros.  lengths[b] >= n_fft.  NULL = se_stft. */
int se_stft_len(const float* wav, long long wav_stride, int B, int N, const int* lengths, const float* scale, int n_fft,
                int win, int hop, int T, float* mag, long long msb, long long mst, long long msf, float* re, float* im,
                long long sb, long long st, long long sf, float p_mag, float p_ri, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a7+a8+a9  Fused recombination prologue + irFFT + window + overlap-add + envelope
 *     normalisation + trim + 1/c.  Replaces the numpy/torch glue and librosa.istft /
 *     torch.istft of every decode script (CRN/crn_decode.py:51-57, gcrn_decode.py:52-63,
 *     dccrn_decode.py:49-60, fullsubnet_sa_decode.py:64-78, Uformer/uformer.py:276).
 *     mode:
 *       SE_ISTFT_SPEC       Y = (a_re, a_im)                              (already a spectrum)
 *       SE_ISTFT_RI_DECOMP  Y = A * |A|^(inv_p-1), A=(a_re,a_im)         (backend rule (ii))
 *       SE_ISTFT_MAG_PHASE  Y = a_re^inv_p * X/|X|, X=(b_re,b_im)        (rule (i); a_im unused)
 *       SE_ISTFT_CMASK      C = A * (X*|X|^(p_x-1)); Y = C*|C|^(inv_p-1) (rule (iii); A = mask)
 *     a_* / b_* planes are addressed plane[b*sb + t*st + f*sf] with their own stride sets.
 *     out[b*out_stride + n], n in [0, L): sample n + n_fft/2 of the overlap-add, divided by
 *     the window-sum-square envelope where it exceeds FLT_MIN, times out_scale[b] (or 1).
 *     Samples beyond the overlap-add extent are written as 0 (librosa fix_length).
 * ------------------------------------------------------------------------------------- */
typedef enum se_istft_mode {
  SE_ISTFT_SPEC = 0,
  SE_ISTFT_RI_DECOMP = 1,
  SE_ISTFT_MAG_PHASE = 2,
  SE_ISTFT_CMASK = 3
} se_istft_mode;

int se_istft(int mode, const float* a_re, const float* a_im, long long a_sb, long long a_st, long long a_sf,
             const float* b_re, const float* b_im, long long b_sb, long long b_st, long long b_sf, float inv_p,
             float p_x, int B, int T, int n_fft, int win, int hop, const float* out_scale, float* out,
             long long out_stride, int L, se_stream_t stream);
/* se_istft on a tail-padded batch: only the frames t < 1 + lengths[b] / hop of clip b are overlap-added (and enter its
 * window envelope), samples n >= lengths[b] are written as 0 -- librosa.istft(..., length=len(x)) per file
 * (CRN/crn_decode_vb.py:50-51).  NULL = se_istft. */
int se_istft_len(int mode, const float* a_re, const float* a_im, long long a_sb, long long a_st, long long a_sf,
                 const float* b_re, const float* b_im, long long b_sb, long long b_st, long long b_sf, float inv_p,
                 float p_x, int B, int T, int n_fft, int win, int hop, const float* out_scale, float* out,
                 long long out_stride, int L, const int* lengths, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a6 building block: causal Conv2d / ConvTranspose2d / Linear as one implicit GEMM.
 *     Activations are channels-last [B, T, F, C] fp32.  For an output position (b, t, fo)
 *     and output channel co:
 *        out[b, t, dst_f0 + fo*dst_fstep, co] =
 *           act( bias[co] + sum_{tap < ntaps} sum_{ci < C0+C1}
 *                  in[b, t + dt[tap], fo*sf + df[tap], ci] * W[(tap*(C0+C1) + ci) * ldw + co] )
 *     where `in` is the channel concatenation of src0 (C0 channels) and src1 (C1, may be
 *     0): the decoder's torch.cat((x, skip), dim=1) (CRN/CRN.py:107) without the copy.
 *     Out-of-range (t+dt < 0, f outside [0,Fin)) taps read zero: the causal top pad of
 *     CRN/CRN.py:38 and the transposed-convolution borders.  BatchNorm (eval) is folded
 *     into W / bias by the caller.  Covers nn.Conv2d k(2,3) s(1,2) (CRN.py:40), the two
 *     output-column parities of nn.ConvTranspose2d k(2,3) s(1,2) (CRN.py:77), nn.Linear and
 *     the hoisted LSTM input projection (ntaps=1, F=1).
 *     fill_f >= 0 additionally writes act(fill[co]) into output column fill_f (the left F-pad
 *     of de4, CRN.py:92-97, which passes through BN+ELU).
 * ------------------------------------------------------------------------------------- */
#define SE_MAX_TAPS 16
typedef struct se_conv_desc {
  const float* src0;
  const float* src1;
  int C0, C1;
  int B, T, Fin;
  int Fout;       /* output columns computed by this launch */
  int ntaps;
  int dt[SE_MAX_TAPS];
  int df[SE_MAX_TAPS];
  int sf;
  const float* W; /* [ntaps*(C0+C1)][ldw] */
  int ldw;        /* >= Cout, multiple of 4 */
  const float* bias; /* [Cout] or NULL */
  int Cout;
  int act;        /* se_act */
  float act_param; /* PReLU slope */
  float* dst;     /* [B, T, dstF, Cout] */
  int dstF, dst_f0, dst_fstep;
  int fill_f;     /* -1: none */
  const float* fill; /* [Cout] */
} se_conv_desc;

int se_conv_gemm(const se_conv_desc* desc, se_stream_t stream);

/* dst[b, t, fill_f, co] = act(fill[co]) for every (b,t): a zero-padded output column that still passes
 * through BatchNorm + activation (the left F-pad of CRN de4, CRN/CRN.py:92-97). */
int se_fill_column(float* dst, long long rows, int dstF, int Cout, int fill_f, const float* fill, int act,
                   float act_param, se_stream_t stream);

/* First encoder layer (C_in = 1): Conv2d(1,Cout,k(2,3),s(1,2)) + folded BN + act on a
 * [B,T,Fin] plane -> channels-last [B,T,Fout,Cout].  CRN/CRN.py:37-43.  W [6][Cout] (tap
 * major: kt*3+kf), Cout <= 64. */
int se_conv_in1(const float* src, int B, int T, int Fin, const float* W, const float* bias, int Cout, int act,
                float* dst, int Fout, se_stream_t stream);

/* Last decoder layer (C_out = 1): ConvTranspose2d(C0+C1,1,k(2,3),s(1,2)) + drop last frame
 * + folded BN + act on channels-last inputs -> [B,T,2*Fin+1] plane.  CRN/CRN.py:98-102.
 * W [6][C0+C1] (tap major kt*3+kf). */
int se_deconv_out1(const float* src0, const float* src1, int C0, int C1, int B, int T, int Fin, const float* W,
                   float bias, int act, float* dst, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a6 building block: the LSTM recurrence (nn.LSTM, batch_first, zero initial state, gate
 *     order i,f,g,o; CRN/CRN.py:20,29; LSTM/LSTM.py:17-18).  The input projection
 *     x W_ih^T + b_ih + b_hh for all T is computed beforehand by se_conv_gemm into
 *     xproj [B, T, 4H] whose columns are in "slice order": column  s*4*HU + gate*HU + j
 *     is gate `gate` of hidden unit u = s*HU + j  (HU = H / nslices).  whh is packed the
 *     same way: whh[s][k][gate*HU + j] = W_hh[gate*H + s*HU + j][k]  (size nslices*H*4*HU).
 *     xproj_stride = floats between consecutive (b,t) rows (>= 4H; lets several LSTMs share one
 *     projection GEMM output).  One persistent CTA per slice keeps its W_hh slice resident in shared memory for all
 *     T steps; steps are separated by a device-wide barrier on `sync` (>= 16 unsigned,
 *     zeroed by the caller before each call is NOT required: the kernel is given a base
 *     epoch).  hseq [B, T, H] receives h_t (natural unit order).  work: >= 2*H*Bpad floats,
 *     Bpad = 8*ceil(B/8), scratch for the transposed state.
 *     Requires H % nslices == 0, 4*HU == 32 (HU = 8), H % 128 == 0, B <= 64.
 * ------------------------------------------------------------------------------------- */
int se_lstm_seq(const float* xproj, long long xproj_stride, const float* whh, int B, int T, int H, float* hseq,
                long long hseq_sb, long long hseq_st, float* work, unsigned* sync, se_stream_t stream);
/* `ngroups` (<= 8) independent LSTMs of identical shape in ONE launch (DCCRN's four real passes per
 * NavieComplexLSTM, complexnn): group g reads xproj columns [g*xproj_group_off, +4H), weights
 * whh + g*whh_group_stride, writes hseq columns [g*hseq_group_off, +H).  work: ngroups x the single
 * size; sync: >= 16 unsigned. */
int se_lstm_seq_multi(const float* xproj, long long xproj_stride, long long xproj_group_off, const float* whh,
                      long long whh_group_stride, int ngroups, int B, int T, int H, float* hseq, long long hseq_sb,
                      long long hseq_st, long long hseq_group_off, float* work, unsigned* sync, se_stream_t stream);
/* Recurrence engine (process-global; for A/B measurements and tests):
 *   4 (default) = as 3, with the second-generation tcgen05 kernel at H = 1024 (csrc/lstm_f16.cu: FP16 hi/lo operand
 *       pairs with power-of-two scales instead of TF32 pairs -- same three-term product, twice the MMA rate, half the
 *       operand bytes -- and a barrier-free state exchange: h_t is published as self-validating tagged 32-bit words that
 *       the consumers poll with plain L2 loads, no device-wide counter / proxy fence / TMA in the step's chain);
 *   3 = as 2, plus the sequence-parallel kernel for H = 128 (csrc/lstm.cu: lstm_seq_small_kernel -- all of
 *       W_hh resident in one CTA's registers + shared memory, 1 / 2 / 4 whole sequences per CTA, no device-wide
 *       barrier): DPCRN's inter-chunk LSTM, DCCRN's real / imaginary LSTMs;
 *   2 = tcgen05 cluster kernel (csrc/lstm_tc.cu: W_hh hi part in tensor memory, K split over a cluster of
 *       4 CTAs with a DSMEM reduction) where it applies -- H = 1024, one group, 32 clusters of 4 co-resident -- and
 *       the FMA kernel elsewhere;
 *       (SE_LSTM_ENGINE=0..4 in the environment picks the start-up value for A/B runs);
 *   1 = legacy mma.sync TF32 path with the 3xTF32 split (H in {128, 512, 1024});
 *   0 = fp32 FMA kernel (any H %% 128 == 0).
 * Measured on B200 (H = 1024, B = 64): 7.9 / 14.4 / 13.8 us per step, see DESIGN.md. */
int se_set_lstm_engine(int engine);
/* tcgen05 GEMM engine (process-global; se_gemm_tf32x3*, se_lstm_cell_tf32x3*, se_conv_tf32x3):
 *   0 = one CTA per 128 x 128 output tile (csrc/gemm_tc.cu: gemm_tf32x3_kernel) for every shape;
 *   1 = CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles, each SM stages half of the operands) where M >= 256 and
 *       N >= 256 (GEMM) / two activation tiles exist and Cout > 32 (conv), the one-CTA kernel elsewhere;
 *   2 / 3 / 4 = the one-CTA MMA in clusters of 2 x 2 / 1 x 2 / 2 x 1 CTAs that read the operand tiles they have in
 *       common from L2 once (TMA multicast): GEMM only, the conv kernel stays on engine 0;
 *   5 (default) = CTA pairs for GEMMs with K >= 384, M >= 256 and N >= 256, engine 0 for everything else.
 * SE_GEMM_ENGINE in the environment picks the start-up value. */
int se_set_gemm_engine(int engine);
/* Measurement aid for the tcgen05 engine: dev_buf (device, 128 * nsteps * 12 int64, or NULL to switch off)
 * receives clock64() stamps of 12 phase boundaries per CTA for steps [first_step, first_step + nsteps); see
 * csrc/lstm_tc.cu for the event list and tools/lstm_tc_phases.py for the reader. */
int se_debug_lstm_tc_profile(long long* dev_buf, int first_step, int nsteps);
/* Bytes of `work` se_lstm_seq needs (per group). */
long long se_lstm_seq_work_bytes(int B, int H);

/* ---------------------------------------------------------------------------------------
 * a6 building block on the tensor cores: fp32-accurate GEMM as 3xTF32 (tcgen05 + TMEM + TMA).
 *     se_split_tf32:  x -> hi = rna_tf32(x), lo = rna_tf32(x - hi)   (n %% 4 == 0, 16-byte aligned)
 *     se_gemm_tf32x3: C[M,N] = act(bias + (A_hi+A_lo)[M,K] * (B_hi+B_lo)[N,K]^T), dropping only
 *                     the A_lo*B_lo term; both operands K-major (row strides lda / ldb floats),
 *                     K %% 32 == 0.  Same contract as nn.Linear(K, N) with weight [N, K]
 *                     (LSTM/LSTM.py:19-22) and the hoisted nn.LSTM input projection
 *                     weight_ih_l* [4H, K] (CRN/CRN.py:20).
 * ------------------------------------------------------------------------------------- */
int se_split_tf32(const float* x, float* hi, float* lo, long long n, se_stream_t stream);
/* The same for rows of K floats that are zero-padded to Kpad >= K (Kpad %% 4 == 0) on the way: x [rows, K] ->
 * hi / lo [rows, Kpad].  Lets a K that is not a multiple of 32 (the 161 bins entering LSTM/LSTM.py:17) use
 * se_gemm_tf32x3 against weights padded the same way. */
int se_pad_split_tf32(const float* x, long long rows, int K, int Kpad, float* hi, float* lo, se_stream_t stream);
int se_gemm_tf32x3(const float* a_hi, const float* a_lo, long long lda, const float* b_hi, const float* b_lo,
                   long long ldb, int M, int N, int K, const float* bias, int act, float* C, long long ldc,
                   se_stream_t stream);
/* Extended epilogue: C = alpha * act(A B^T + bias, act_param) + res (res [M, ldc] or NULL), optionally also the
 * TF32 split of C (c_hi / c_lo, row stride ldc) for a following tensor-core layer.  C may be NULL when only
 * the split is wanted.  Covers y*0.5 + x of the conformer feed-forward blocks (Uformer/ff_cplx.py:26-32) and the
 * x + y residual of DSConv2d (dsconv2d_cplx.py:58-59). */
int se_gemm_tf32x3_ex(const float* a_hi, const float* a_lo, long long lda, const float* b_hi, const float* b_lo,
                      long long ldb, int M, int N, int K, const float* bias, int act, float act_param, float alpha,
                      const float* res, float* C, float* c_hi, float* c_lo, long long ldc, se_stream_t stream);

/* One LSTM time step for MANY independent sequences, fused on the tensor cores:
 *     gates[M, 4H] = [x_t | h_{t-1}] * W^T + bias ;  c, h updated in the GEMM epilogue.
 *     FullSubNet's sub-band model runs B*257 sequences of width 384 (model.py:106-113): the
 *     per-step contraction is a real GEMM (M = 8224 per GPU), so the recurrence is T launches of
 *     this kernel instead of a persistent weight-stationary kernel.
 *     x_hi/x_lo [M, Kx], h_hi/h_lo [M, H] (state in, TF32 split), W hi/lo [4H, Kx+H] K-major with
 *     rows in TILE order: row j*128 + g*32 + u  <-  gate g (i,f,g,o) of unit 32j+u; bias likewise.
 *     c_state [M, H] in/out; h_hi_out/h_lo_out [M, H] (must differ from the inputs: other CTAs
 *     still read h_{t-1}); h_out [M, H] plain fp32 or NULL.  Kx %% 32 == 0, H %% 32 == 0. */
int se_lstm_cell_tf32x3(const float* x_hi, const float* x_lo, long long ldx, int Kx, const float* h_hi,
                        const float* h_lo, long long ldh, int H, const float* w_hi, const float* w_lo, long long ldw,
                        const float* bias, int M, float* c_state, float* h_hi_out, float* h_lo_out, float* h_out,
                        se_stream_t stream);
/* Same step with (a) a row stride ld_hout for the three h outputs, so a step can write straight into a
 * [M, L, D*H] sequence buffer at (position, direction) -- the output of step l is the state input of step
 * l+1 and the layer output at once (DPCRN's intra Bi-LSTM over F = 4, DPCRN/DPCRN.py:52,70); and (b)
 * first_step = 1: h_{-1} = c_{-1} = 0, the recurrent part of K is skipped, h_hi / h_lo may be NULL and
 * c_state is write-only. */
int se_lstm_cell_tf32x3_ex(const float* x_hi, const float* x_lo, long long ldx, int Kx, const float* h_hi,
                           const float* h_lo, long long ldh, int H, const float* w_hi, const float* w_lo, long long ldw,
                           const float* bias, int M, float* c_state, float* h_hi_out, float* h_lo_out, float* h_out,
                           long long ld_hout, int first_step, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * FullSubNet glue (FullSubNet/fullsubnet_net_sa/model.py:68-118); see csrc/fullsubnet.cu.
 *   se_fsn_clip_inv_mean: out_inv[b] = 1/((sum_{t<T,f<F} wgt[f]*x[b,t,f] + sum extra[b,:n_extra])/denom + 1e-5)
 *                         -- offline_laplace_norm (base_model.py:196-209); wgt / extra may be NULL.
 *   se_fsn_fb_input:      x [B,T,F] (strides) -> mag_tm [B,Tp,F] (zero look-ahead frames, model.py:79)
 *                         and xn = mag_tm * inv[b] (the full-band model input, model.py:84).
 *   se_fsn_sb_assemble:   rows [t][b*F+f] of 2n+2 features: reflect-unfolded noisy magnitude
 *                         (base_model.py:12-42) ++ full-band output, times inv[b], TF32 hi/lo split.
 *   se_fsn_sb_fc:         out[row][c] = bias[c] + h[row,:] . W[c,:], c in {0,1}  (Linear(384,2)).
 * ------------------------------------------------------------------------------------- */
int se_fsn_clip_inv_mean(const float* x, long long sb, long long st, long long sf, int B, int T, int F,
                         const float* wgt, const float* extra, long long extra_sb, long long n_extra, double denom,
                         float* out_inv, se_stream_t stream);
int se_fsn_fb_input(const float* x, long long sb, long long st, long long sf, int B, int T, int Tp, int F,
                    const float* inv, float* mag_tm, float* xn, se_stream_t stream);
int se_fsn_sb_assemble(const float* mag_tm, const float* fb, int B, int Tp, int F, int num_neighbors,
                       const float* inv, float* out_hi, float* out_lo, se_stream_t stream);
int se_fsn_sb_fc(const float* h, int M, int H, const float* W, const float* bias, float* out, se_stream_t stream);
/* se_fsn_sb_assemble for the fp16-pair cell GEMM (se_lstm_cell_f16x3): rows scaled by 2^scale_log2. */
int se_fsn_sb_assemble_f16(const float* mag_tm, const float* fb, int B, int Tp, int F, int num_neighbors,
                           const float* inv, int scale_log2, unsigned short* out_hi, unsigned short* out_lo,
                           se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * The same fp32-class products on fp16 operand PAIRS (tcgen05 kind::f16: twice the MMA rate of kind::tf32 and half
 * the operand bytes).  An operand x travels as two IEEE fp16 arrays with  x * 2^s = hi + lo  (s = its scale_log2):
 * 22 significand bits wherever |x * 2^s| >= 2^-3, an absolute floor of 2^-25-s below, saturation (not inf) beyond
 * +-65504 * 2^-s.  Weights are split once at load time with a per-tensor s that puts max|w| in [2^13, 2^14);
 * activations use a fixed s (4 by default: |x| < 4094, floor 2e-9).  Same three-term product as se_gemm_tf32x3
 * (A_lo*B_lo dropped); the epilogue multiplies the fp32 sums by 2^-(s_A + s_B).
 *     se_split_f16:   x [rows, K] (row stride ldx floats) -> hi / lo [rows, Kpad] fp16, columns K.. zero, Kpad %% 8 == 0
 *     se_gemm_f16x3:  C = alpha * act(A B^T + bias, act_param) + res, A [M,K] and B [N,K] fp16 pairs (lda / ldb in
 *                     elements, %% 8), K %% 8 == 0 (k-blocks of 64, zero-filled past K); scale_log2_ab = s_A + s_B;
 *                     optional outputs: fp32 C, the TF32 pair (c_hi / c_lo) and / or the fp16 pair (c16_hi / c16_lo,
 *                     scaled by 2^c16_scale_log2) of C for the layer that follows; all with row stride ldc.
 *                     Replaces the same reference ops as se_gemm_tf32x3 (nn.Linear, hoisted nn.LSTM input projections:
 *                     CRN/CRN.py:20, LSTM/LSTM.py:17-22).
 *     se_lstm_cell_f16x3: se_lstm_cell_tf32x3_ex on fp16 pairs: gates = [x | h] [W_ih | W_hh]^T with W laid out as
 *                     [4H][Kx rounded up to 64 | H] (tile row order as for the TF32 cell), x and h scaled by 2^scale_log2_a,
 *                     W by 2^scale_log2_w; h_hi_out / h_lo_out leave as fp16 pairs scaled by 2^scale_log2_a.
 *                     (FullSubNet sub-band LSTM steps, FullSubNet/fullsubnet_net_sa/model.py:106-112.)
 * ------------------------------------------------------------------------------------- */
int se_split_f16(const float* x, long long rows, int K, long long ldx, int Kpad, int scale_log2, unsigned short* hi,
                 unsigned short* lo, se_stream_t stream);
int se_gemm_f16x3(const unsigned short* a_hi, const unsigned short* a_lo, long long lda, const unsigned short* b_hi,
                  const unsigned short* b_lo, long long ldb, int M, int N, int K, int scale_log2_ab, const float* bias,
                  int act, float act_param, float alpha, const float* res, float* C, float* c_hi, float* c_lo,
                  unsigned short* c16_hi, unsigned short* c16_lo, int c16_scale_log2, long long ldc, se_stream_t stream);
int se_lstm_cell_f16x3(const unsigned short* x_hi, const unsigned short* x_lo, long long ldx, int Kx,
                       const unsigned short* h_hi, const unsigned short* h_lo, long long ldh, int H,
                       const unsigned short* w_hi, const unsigned short* w_lo, long long ldw, int scale_log2_a,
                       int scale_log2_w, const float* bias, int M, float* c_state, unsigned short* h_hi_out,
                       unsigned short* h_lo_out, float* h_out, long long ld_hout, int first_step, se_stream_t stream);

/* Tensor-core twin of se_conv_gemm (tcgen05 3xTF32, 4-D TMA im2col-free A tiles; csrc/conv_tc.cu)
 * for layers with C0, C1 multiples of 32 and Fout <= 128.  Activations and weights are TF32 hi/lo
 * pairs (se_split_tf32); weights are K-major [Cout][ntaps*(C0+C1)] with the same (tap, channel)
 * K order as se_conv_gemm.  Writes fp32 `out` and/or the hi/lo pair for the next layer. */
typedef struct se_conv_tc_desc {
  const float *src0_hi, *src0_lo, *src1_hi, *src1_lo;
  int C0, C1;
  int B, T, Fin, Fout;
  int ntaps;
  int dt[SE_MAX_TAPS];
  int df[SE_MAX_TAPS];
  int sf;
  const float *w_hi, *w_lo;
  const float* bias;
  int Cout;
  int act;
  float act_param;
  float *out, *out_hi, *out_lo;
  int dstF, dst_f0, dst_fstep;
  /* gated conv (GluConv2d / GluConvTranspose2d, GCRN/GCRN_noncprs.py:42-83): when glu != 0 the Cout GEMM columns are
   * (conv1, conv2) pairs -- column 2j = conv1 of channel j, 2j+1 = conv2 -- and the outputs have Cout / 2 channels:
   * act((conv1 * sigmoid(conv2)) * glu_scale[j] + glu_shift[j])  (gate, eval BatchNorm, ELU in the epilogue;
   * scale / shift may be NULL).  Cout %% 8 == 0, act ELU or none. */
  int glu;
  const float *glu_scale, *glu_shift;
} se_conv_tc_desc;
int se_conv_tf32x3(const se_conv_tc_desc* desc, se_stream_t stream);

/* se_conv_tf32x3 on fp16 operand pairs (see "fp16 operand PAIRS" above; csrc/conv_f16.cu): activations [B,T,F,C] and
 * weights are (hi, lo) fp16 pairs scaled by 2^scale_log2_a / 2^scale_log2_w.  A k-block is a 64-channel slice: C0 / C1
 * only need to be multiples of 8 (the TMA unit zero-fills the rest of a slice), the packed weights are
 * [Cout][ntaps * (pad64(C0) + pad64(C1))] with zero columns in the padding.  Outputs: fp32 `out`, the TF32 pair and / or
 * the fp16 pair (scaled by 2^out16_scale_log2) for the layer that follows.  Replaces the same reference layers as
 * se_conv_tf32x3 / se_conv_gemm (nn.Conv2d / nn.ConvTranspose2d of CRN/CRN.py:35-109, DCCRN/DCCRN_cprs.py:60-132, ...). */
typedef struct se_conv_f16_desc {
  const unsigned short *src0_hi, *src0_lo, *src1_hi, *src1_lo;
  int C0, C1;
  int B, T, Fin, Fout;
  int ntaps;
  int dt[SE_MAX_TAPS];
  int df[SE_MAX_TAPS];
  int sf;
  const unsigned short *w_hi, *w_lo;
  int scale_log2_a, scale_log2_w;
  const float* bias;
  int Cout;
  int act;
  float act_param;
  float *out, *out_hi, *out_lo;
  unsigned short *out16_hi, *out16_lo;
  int out16_scale_log2;
  int dstF, dst_f0, dst_fstep;
  int glu;
  const float *glu_scale, *glu_shift;
  /* ncls == 2: both output-column parity classes of a stride-2 ConvTranspose2d in ONE launch (the activation tiles are
   * read once instead of twice).  The Cout GEMM columns are two classes of Cout / 2 channels: columns [0, Cout/2) are
   * written at output column dst_f0 + fo * dst_fstep, columns [Cout/2, Cout) at dst_f0 + 1 + fo * dst_fstep for
   * fo < fout1; the tap list is the union of the two classes' taps (zero weights where a class does not use a tap),
   * bias has Cout / 2 entries, the outputs have Cout / 2 channels.  0 / 1: one class (the fields above as they are). */
  int ncls, fout1;
} se_conv_f16_desc;
int se_conv_f16x3(const se_conv_f16_desc* d, se_stream_t stream);

/* GCRN gated-conv tail and skip re-activation (csrc/pointwise.cu):
 *   se_glu_affine_act: x [rows, 2C] = [conv1 | conv2] -> act((conv1 * sigmoid(conv2)) * scale[c] + shift[c])
 *                      (GluConv2d + eval BatchNorm2d + ELU, GCRN/GCRN_noncprs.py:55-57,138); scale/shift may be NULL.
 *   se_unary:          y = act(x)  (the ELU applied to cat(BN(deconv), skip), :149-152, re-activates the skip).
 *   Both write fp32 `out` and/or its TF32 split. */
int se_glu_affine_act(const float* x, long long rows, int C, const float* scale, const float* shift, int act,
                      float act_param, float* out, float* out_hi, float* out_lo, se_stream_t stream);
int se_unary(const float* x, long long n, int act, float act_param, float* out, float* out_hi, float* out_lo,
             se_stream_t stream);
/* out[i] = x[i] * m[i] for n interleaved (re, im) complex numbers: the complex ratio mask DPCRN applies inside
 * forward (DPCRN/DPCRN.py:33-42). */
int se_cmul(const float* x, const float* m, long long n, float* out, se_stream_t stream);

/* DCCRN polar mask, masking_mode 'E' (DCCRN/DCCRN_cprs.py:201-220):
 *     est = tanh(|M|) * |X| * exp(j(angle X + angle M)),  M = 0 at the DC bin (:203-204).
 *     m [B,T,F-1,2] channels-last mask for bins 1..F-1; x / est: planes (re, im) addressed
 *     plane[b*sb + t*st + f*sf] with their own strides. */
int se_dccrn_mask(const float* m, const float* x_re, const float* x_im, long long x_sb, long long x_st, long long x_sf,
                  int B, int T, int F, float* e_re, float* e_im, long long e_sb, long long e_st, long long e_sf,
                  se_stream_t stream);
/* The same with the other two mask rules of DCCRN.forward (DCCRN/DCCRN_cprs.py:221-224; no shipped checkpoint uses
 * them): 'C' est = X * M (complex product), 'R' est = (X_r M_r, X_i M_i). */
enum { SE_DCCRN_MASK_E = 0, SE_DCCRN_MASK_C = 1, SE_DCCRN_MASK_R = 2 };
int se_dccrn_mask_ex(const float* m, const float* x_re, const float* x_im, long long x_sb, long long x_st,
                     long long x_sf, int B, int T, int F, int mode, float* e_re, float* e_im, long long e_sb,
                     long long e_st, long long e_sf, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Front step of the *_decode_vb.py scripts: librosa.resample(x, orig_fs, 16000, fix=True, scale=False)
 * (LSTM/lstm_decode_vb.py:33-34, DCCRN/dccrn_decode_vb.py:25-26) = resampy 'kaiser_best' band-limited
 * interpolation (un-vendored dependency; published algorithm restated, see oracle/resample.py):
 *     y[b, t] = sum_i (win[o_l + i*step] + eta_l * delta[o_l + i*step]) * x[b, n - i]            (left wing)
 *             + sum_k (win[o_r + k*step] + eta_r * delta[o_r + k*step]) * x[b, n + k + 1]        (right wing)
 *     time = t / ratio, n = floor(time), step = int(min(1, ratio) * num_table), offsets / eta from the fractional
 *     part as resampy.interpn.resample_f computes them (float64 index arithmetic).
 * x [B, n_in] (row stride x_stride), y [B, n_out] (row stride y_stride); samples t >= n_valid = int(n_in * ratio)
 * are zero (librosa's fix_length to ceil(n_in * ratio)).  win / delta: device tables of nwin floats (the half
 * window, already multiplied by ratio when decimating, and its forward difference).  time_reg: device table of
 * n_valid doubles holding resample_f's time register (0, then += 1/ratio sequentially in float64) or NULL to use
 * t * (1/ratio) -- identical when 1/ratio is exactly representable (48 k -> 16 k); at 44.1 k -> 16 k the exact
 * time is an integer every 160 samples and the rounding of the running sum decides floor(time). */
int se_resample(const float* x, long long x_stride, int B, int n_in, float* y, long long y_stride, int n_out,
                int n_valid, double ratio, const double* time_reg, const float* win, const float* delta, int nwin,
                int num_table, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Uformer glue (Uformer/uformer.py:172-287, dilated_dualpath_conformer.py, fusion.py, t_att_*.py,
 * f_att_*.py); see csrc/uformer.cu.  Tensors are channels-last; complex tensors carry (re C | im C).
 *   se_uf_prep:   x [B,T,F,2] -> mag, phase [B,T,F] (sqrt(clamp(.,eps)), atan2(im+eps, re); uformer.py:197)
 *                 and the network inputs for bins 1..F-1 (uformer.py:204-210).
 *   se_uf_fusion: m' = m + sigmoid(|c|), c' = c + sigmoid(m)                         (fusion.py:13-19)
 *   se_group_layernorm: nn.LayerNorm(C) over each of the G channel groups of every row (G = 2: the
 *                 real and imaginary halves share gamma/beta, as x.transpose(1,4) does); optional
 *                 gate (x * sigmoid(gate), dsconv2d_cplx.py:54), post op (1 PReLU, 2 swish), residual add,
 *                 fp32 and/or TF32-split output.  C <= 1024.  out_index (NULL = identity, else C ints): channel
 *                 ch of a group is stored at position out_index[ch] (GCRN's (c,f) -> (f,c) view+transpose after
 *                 the grouped LSTM, GCRN_noncprs.py:145, without a separate permute pass).
 *   se_attention: single-head-dim-16 attention for nheads (<= 8) heads whose outputs are combined with
 *                 signs into nout (<= 2) groups (t_att_cplx.py:58-67).  qkv [R, ld] rows hold
 *                 (q16|k16|v16) per head; sequence (o, i) starts at row o*outer_stride + i*inner_stride and
 *                 advances lstride rows per position; out [R, ldo].
 *   se_uf_mask:   sigmoid magnitude branch and tanh/phase mask branch averaged, DC bin zero padded
 *                 (uformer.py:236-262) -> est [B,T,F,2].
 * ------------------------------------------------------------------------------------- */
int se_uf_prep(const float* x, int B, int T, int F, float* mag, float* phase, float* cplx_in, float* mag_in,
               se_stream_t stream);
int se_uf_fusion(const float* c, const float* m, long long rows, int C, float* c_out, float* m_out, se_stream_t stream);
/* The same with each result as fp32 (c_out / m_out) and / or as the TF32 (hi, lo) pair the tensor-core convs read
 * (c_hi, c_lo / m_hi, m_lo); unused outputs NULL.  Saves the split pass between a fusion and the next U-Net conv. */
int se_uf_fusion_ex(const float* c, const float* m, long long rows, int C, float* c_out, float* c_hi, float* c_lo,
                    float* m_out, float* m_hi, float* m_lo, se_stream_t stream);
int se_group_layernorm(const float* x, const float* gate, long long rows, int G, int C, const float* gamma,
                       const float* beta, float eps, int post, float slope, const float* res, const int* out_index,
                       float* out, float* out_hi, float* out_lo, se_stream_t stream);
int se_attention(const float* qkv, int ld, int nheads, const int* head_out, const float* head_sign, int nout, int L,
                 long long lstride, int n_outer, long long outer_stride, int n_inner, long long inner_stride, float scale,
                 float* out, int ldo, se_stream_t stream);
int se_uf_mask(const float* cmask, const float* mdec, const float* mag, const float* phase, int B, int T, int F,
               float* est, se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Utterance-level normalisations of the TCM family + CTSNet stage glue; see csrc/norm.cu.
 *   nn.InstanceNorm{1,2}d(affine=True), statistics per (clip, channel) over T (x F) even in eval()
 *       (CTSNet/Step1_network.py:48,164; Step2_network.py:44,131)           -> se_chan_stats + se_chan_norm
 *   CumulativeLayerNorm{1,2}d, statistics over (C[,F]) of all frames <= t
 *       (CTSNet_new/Step1_network.py:213-286)                              -> se_cum_stats  + se_chan_norm
 * x is channels-last [B, rows, Cin] (rows = T*F).  `pre` is applied on the fly before the statistics / the
 * normalisation (it is what the reference computes between the convolution and the norm):
 *     NONE       v = x[r, c]                                   (Cin == C)
 *     GLU        v = x[r, c] * sigmoid(x[r, C + c])            (Cin == 2C; Gate_Conv, Step1_network.py:150-151)
 *     PRELU      v = prelu(x[r, c % Cin], pre_slope[c])        (C a multiple of Cin: several branches, each with
 *                                                               its own PReLU, read the same tensor; :163,170)
 *     GLU_PRELU  v = prelu(GLU, pre_slope[c])                  (TCM output path, :188-189)
 * se_chan_norm: y = (v - mean) * rstd * gamma[c] + beta[c], then `post`:
 *     PRELU  per-channel slopes (:49);   FIR  ShareSepConv (:195-209): y[t] = sum_k fir_w[g][k] y[t-(K-1)+k] with zero
 *     history, one filter per channel group g (fir_groups groups of C/fir_groups channels); rows must be T.
 *   stat_mode INSTANCE: mean/rstd [B, C];  CUMULATIVE: mean/rstd [B, T, G], T = rows / rows_per_t, one set per channel
 *   group (G = `groups` / `stat_groups`: branches that were stacked on the channel axis keep their own statistics).
 *   Outputs: fp32 `out` and/or the TF32 split pair.  eps as the reference module's (1e-5).
 *   ws: se_chan_stats_ws_bytes(B, rows, C) bytes whose first 4096 (the per-clip tickets; B <= 1024) are ZERO-filled
 *       once by the caller -- the kernel leaves them zeroed, so one buffer serves every later call;
 *       se_cum_stats needs 16 * B * T * groups bytes (no initialisation).  C must divide 256 for se_chan_stats.
 * se_add: out = a + b (x_acc of Step1_network.py:28-33);  se_axpby: out = ca*a + cb*b (the Taylor recursion
 *   update = block(..) + k * pre_term, out += update / (k+1)!, TaylorSENet/TaylorSENet.py:84-93).
 * se_taylor_zero: TaylorSENet's zeroth-order term gain * |X| * (cos, sin)(angle X) (TaylorSENet.py:73-76) from
 *   x_ri [rows, F, 2] and gain [rows, F], written as "RI rows" [rows, ld] = [re(F) | im(F) | zero pad].
 * se_cts_glue1 / se_cts_glue2: CTSNet/two_stage_com_decode_vb.py:79-84 -- stage-1 magnitude x noisy phase,
 *   cat(noisy RI, stage-1 RI) -> s2_in [n, 4];  est [n, 2] = stage-2 output + stage-1 RI.
 * ------------------------------------------------------------------------------------- */
enum { SE_NORM_PRE_NONE = 0, SE_NORM_PRE_GLU = 1, SE_NORM_PRE_PRELU = 2, SE_NORM_PRE_GLU_PRELU = 3 };
enum { SE_NORM_POST_NONE = 0, SE_NORM_POST_PRELU = 1, SE_NORM_POST_FIR = 2 };
enum { SE_NORM_STAT_INSTANCE = 0, SE_NORM_STAT_CUMULATIVE = 1 };
long long se_chan_stats_ws_bytes(int B, long long rows, int C);
int se_chan_stats(const float* x, int B, long long rows, int Cin, int C, int pre, const float* pre_slope, float eps,
                  float* mean, float* rstd, void* ws, se_stream_t stream);
int se_cum_stats(const float* x, int B, int T, int F, int Cin, int C, int groups, int pre, const float* pre_slope,
                 float eps, float* mean, float* rstd, void* ws, se_stream_t stream);
int se_chan_norm(const float* x, int B, long long rows, int Cin, int C, int pre, const float* pre_slope,
                 const float* mean, const float* rstd, int stat_mode, int rows_per_t, int stat_groups,
                 const float* gamma, const float* beta, int post, const float* post_slope, const float* fir_w, int fir_k, int fir_groups,
                 float* out, float* out_hi, float* out_lo, se_stream_t stream);
int se_add(const float* a, const float* b, long long n, float* out, float* out_hi, float* out_lo, se_stream_t stream);
int se_axpby(const float* a, const float* b, float ca, float cb, long long n, float* out, float* out_hi, float* out_lo,
             se_stream_t stream);
int se_taylor_zero(const float* x_ri, const float* gain, long long rows, int F, int ld, float* out, float* out_hi,
                   float* out_lo, se_stream_t stream);
 /* se_gaf_update: G2Net stage update x' = gain * |pre| * (cos, sin)(angle pre) + com_resi
 *   (G2Net_new/gaf_net_320.py:104-115) on RI rows [rows, ld] (re at column 0, im at column im_off, zero elsewhere);
 *   pre is addressed as x_re/x_im[r * xs_r + f * xs_f]; gain == resi == NULL: relayout of pre into RI rows. */
int se_gaf_update(const float* x_re, const float* x_im, long long xs_r, int xs_f, const float* gain, const float* resi,
                  long long rows, int F, int ld, int im_off, float* out, float* out_hi, float* out_lo,
                  se_stream_t stream);
int se_cts_glue1(const float* x_ri, const float* est_mag, long long n, float* s2_in, se_stream_t stream);
int se_cts_glue2(const float* out_r, const float* out_i, const float* s2_in, long long n, float* est,
                 se_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Plan-level entry points: the whole CRN path behind four calls, for hosts that are not
 * Python.  csrc/plan_crn.cu restates crn.py + packing.py + decode.enhance_mag_mapping in C++ over the op-level entry
 * points above.
 *   se_crn_weights: HOST pointers to the fp32 tensors of the reference state-dict in torch's own layouts
 *       (CRN/CRN.py:35-109; the keys `crn_net().state_dict()` lists): en.en_module.i.1.{weight [Co,Ci,2,3], bias},
 *       en.en_module.i.2.{weight, bias, running_mean, running_var}, lstm.{weight_ih,weight_hh,bias_ih,bias_hh}_l{0,1},
 *       de.de_module.i.0.{weight [Ci,Co,2,3], bias}, de.de_module.i.{2, 3 for i = 3}.{weight, bias, running_mean,
 *       running_var}.
 *   se_plan_create_crn: packs the weights (eval BatchNorm folded, K-major conv matrices + TF32 pairs, output-parity
 *       classes of the transposed convs, LSTM gate rows in slice order) and allocates ONE set of device buffers for
 *       batches of up to B_max clips of up to N_max samples; nothing is allocated afterwards.
 *   se_query_workspace: bytes of device memory the plan holds.
 *   se_forward_crn: crn_net.forward (CRN/CRN.py:23-33), mag / est [B, T, 161] device fp32.
 *   se_enhance_crn: CRN/crn_decode.py:38-57 for B device waveforms of N samples (row strides in floats): RMS scale,
 *       STFT 320/320/160 with |X|^p, forward, est^(1/p) with the noisy phase, iSTFT(length = N), 1/c.  lengths:
 *       NULL or device int32 [B] per-clip sample counts of a tail-padded batch (se_stft_len).
 *   se_plan_set_graph(plan, 1): se_enhance_crn captures its launch sequence once per (B, N, ragged, p) in a CUDA graph
 *       and replays it (the batch is staged through plan-owned buffers because kernel arguments are baked in).
 * ------------------------------------------------------------------------------------- */
typedef struct se_plan se_plan_t;
typedef struct se_crn_weights {
  const float* en_w[5];
  const float* en_b[5];
  const float* en_bn[5][4];   /* weight, bias, running_mean, running_var */
  const float* lstm_w_ih[2];
  const float* lstm_w_hh[2];
  const float* lstm_b_ih[2];
  const float* lstm_b_hh[2];
  const float* de_w[5];
  const float* de_b[5];
  const float* de_bn[5][4];
} se_crn_weights;
int se_plan_create_crn(const se_crn_weights* w, int B_max, int N_max, se_plan_t** plan);
long long se_query_workspace(const se_plan_t* plan);
int se_plan_set_graph(se_plan_t* plan, int enabled);
int se_forward_crn(se_plan_t* plan, const float* mag, float* est, int B, int T, se_stream_t stream);
int se_enhance_crn(se_plan_t* plan, const float* wav, long long wav_stride, float* out, long long out_stride, int B, int N,
                   const int* lengths, float p, se_stream_t stream);
int se_plan_destroy(se_plan_t* plan);

#ifdef __cplusplus
}
#endif
#endif /* SE_B200_H_ */
