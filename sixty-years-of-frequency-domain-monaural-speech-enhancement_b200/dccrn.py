"""Drop-in ``DCCRN`` (reference: DCCRN/DCCRN_cprs.py:9-226; ``complexnn`` is not vendored by the
reference -- its published algorithm is restated in oracle/complexnn_restated.py and is the
de-facto specification here).

Same constructor, ``forward(inputs [B,2,257,T]) -> [B,2,257,T]`` and state-dict keys
(``encoder.0.0.real_conv.weight`` ... ``enhance.1.r_trans.weight``), so the shipped checkpoints
load unchanged (DCCRN/dccrn_decode.py:11-12).  Covers what the decode scripts instantiate:
masking_mode 'E', use_clstm=True, use_cbn=False.

How the complex network maps onto the real engines (channels-last [B,T,F,C], C = real | imag):
  * ComplexConv2d k(5,2) s(2,1): ONE real implicit GEMM with stacked channels,
        [out_r | out_i] = [x_r | x_i] * [[W_r, W_i], [-W_i, W_r]]      (DCCRN_cprs.py:66-72)
    eval BatchNorm2d folded, PReLU in the epilogue; causal left pad on T and symmetric pad 2 on F
    are tap offsets with zero fill.
  * ComplexConvTranspose2d k(5,2) s(2,1) p(2,0) op(1,0) + ``out[...,1:]`` (:108-115,199): the two
    output-bin parities (3 and 2 frequency taps) x 2 time taps (dt = +1, 0: the crop makes the
    decoder look one frame AHEAD); ``complex_cat`` skip (:197) is two source pointers.
  * NavieComplexLSTM x2 (:84-90,182): the four real LSTM passes per layer share one projection
    GEMM (block weights with the signs of  real = L_r(x_r) - L_i(x_i), imag = L_r(x_i) + L_i(x_r)
    folded into the NEXT projection), the recurrences run through the persistent kernel, and
    r_trans / i_trans write straight into the decoder's channels-last input.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import conv_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .param_tree import bn_rows, build_param_tree

BINS = 257


def _spec(kernel_num, rnn_units, kernel_size):
    kn = [2] + list(kernel_num)
    rows = []
    for i in range(len(kn) - 1):
        for part in ("real_conv", "imag_conv"):
            rows += [(f"encoder.{i}.0.{part}.weight", (kn[i + 1] // 2, kn[i] // 2, kernel_size, 2), "param"),
                     (f"encoder.{i}.0.{part}.bias", (kn[i + 1] // 2,), "param")]
        rows += bn_rows(f"encoder.{i}.1", kn[i + 1])
        rows += [(f"encoder.{i}.2.weight", (1,), "param")]
    di = 0
    for idx in range(len(kn) - 1, 0, -1):
        for part in ("real_conv", "imag_conv"):
            rows += [(f"decoder.{di}.0.{part}.weight", (kn[idx], kn[idx - 1] // 2, kernel_size, 2), "param"),
                     (f"decoder.{di}.0.{part}.bias", (kn[idx - 1] // 2,), "param")]
        if idx != 1:
            rows += bn_rows(f"decoder.{di}.1", kn[idx - 1])
            rows += [(f"decoder.{di}.2.weight", (1,), "param")]
        di += 1
    hidden_dim = 512 // (2 ** len(kn))
    in0 = hidden_dim * kn[-1] // 2
    h = rnn_units // 2
    for l in range(2):
        i = in0 if l == 0 else h
        for part in ("real_lstm", "imag_lstm"):
            rows += [(f"enhance.{l}.{part}.weight_ih_l0", (4 * h, i), "param"),
                     (f"enhance.{l}.{part}.weight_hh_l0", (4 * h, h), "param"),
                     (f"enhance.{l}.{part}.bias_ih_l0", (4 * h,), "param"),
                     (f"enhance.{l}.{part}.bias_hh_l0", (4 * h,), "param")]
    rows += [("enhance.1.r_trans.weight", (in0, h), "param"), ("enhance.1.r_trans.bias", (in0,), "param"),
             ("enhance.1.i_trans.weight", (in0, h), "param"), ("enhance.1.i_trans.bias", (in0,), "param")]
    # keep the reference's key ORDER (encoder, decoder, enhance): see torch.load listing
    return rows


ENC_TAPS = [(kt - 1, kf - 2) for kf in range(5) for kt in range(2)]          # (dt, df): in[t-1+kt, 2f+kf-2]
DEC_EVEN = [(1 - kt, 1 - kf // 2) for kf in (0, 2, 4) for kt in range(2)]    # f'=2m : f = m+1-kf/2 ; t+1-kt
DEC_ODD = [(1 - kt, (3 - kf) // 2) for kf in (1, 3) for kt in range(2)]      # f'=2m+1: f = m+(3-kf)/2


def _stack_complex(wr, wi, taps_kf, transpose):
    """Real weight of the stacked-channel conv for the listed (kf, kt) taps.
    conv:      w [Co/2, Ci/2, 5, 2]  -> rows (tap, [ci_r | ci_i]) cols [co_r | co_i]
    transpose: w [Ci/2, Co/2, 5, 2]  (ConvTranspose2d layout)."""
    blocks = []
    for kf, kt in taps_kf:
        a = wr[:, :, kf, kt]
        b = wi[:, :, kf, kt]
        if not transpose:
            a, b = a.t(), b.t()                       # -> [ci, co]
        top = torch.cat([a, b], dim=1)                # from x_r: (W_r -> out_r, W_i -> out_i)
        bot = torch.cat([-b, a], dim=1)               # from x_i: (-W_i -> out_r, W_r -> out_i)
        blocks.append(torch.cat([top, bot], dim=0))   # [2ci, 2co]
    return torch.cat(blocks, dim=0)


class DCCRN(nn.Module):
    def __init__(self, rnn_layers=2, rnn_units=128, win_len=512, win_inc=128, fft_len=512, win_type='hanning',
                 masking_mode='E', use_clstm=False, use_cbn=False, kernel_size=5,
                 kernel_num=[16, 32, 64, 128, 256, 256], crop_first=True):
        super().__init__()
        if masking_mode not in ('E', 'C', 'R') or not use_clstm or use_cbn or rnn_layers != 2 or kernel_size != 5 \
                or fft_len != 512:
            raise NotImplementedError("se_b200 DCCRN covers the decode scripts' configuration: masking_mode 'E' (or "
                                      "'C' / 'R'), use_clstm=True, use_cbn=False, 2 rnn layers, kernel 5, 512-point FFT")
        self.masking_mode = masking_mode    # DCCRN_cprs.py:206-224
        self.kernel_num = [2] + list(kernel_num)
        self.rnn_units = rnn_units
        self.crop_first = crop_first        # False reproduces DCCRN_SNR/DCCRN.py:159 (``[..., :-1]``)
        build_param_tree(self, _spec(kernel_num, rnn_units, kernel_size))
        self._packed = None
        self._packed_key = None

    def _state_key(self):
        p = next(self.parameters())
        return (p.device, tuple(int(t._version) for t in self.state_dict().values()))

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        dev = next(self.parameters()).device
        kn = self.kernel_num
        P = {}
        for i in range(6):
            pre = f"encoder.{i}"
            s, o = packing.bn_fold(*(sd[f"{pre}.1.{n}"] for n in ("weight", "bias", "running_mean", "running_var")))
            w = _stack_complex(sd[f"{pre}.0.real_conv.weight"], sd[f"{pre}.0.imag_conv.weight"],
                               [(kf, kt) for kf in range(5) for kt in range(2)], False)
            br, bi = sd[f"{pre}.0.real_conv.bias"], sd[f"{pre}.0.imag_conv.bias"]
            bias = torch.cat([br - bi, br + bi])
            P[f"enc{i}"] = (ConvWeights(w * s[None, :], kn[i + 1]), (bias * s + o).contiguous(),
                            float(sd[f"{pre}.2.weight"].item()))
        # ---- complex LSTM ----
        h = self.rnn_units // 2
        d = 512 // (2 ** len(kn))                # 4 frequency bins at the bottleneck
        c2 = kn[-1] // 2                         # 128 real (and 128 imag) channels
        q = torch.arange(d * kn[-1], device=dev)
        qd, qc = q // kn[-1], q % kn[-1]         # channels-last flatten index -> (bin, channel)
        is_real = qc < c2
        ref_feat = (qc % c2) * d + qd            # reference flatten index c*4 + d within a half
        rows = packing.slice_rows(h).to(dev)

        def lstm_block(part, l):
            pre = f"enhance.{l}.{part}"
            return (sd[f"{pre}.weight_ih_l0"][rows], sd[f"{pre}.weight_hh_l0"][rows],
                    (sd[f"{pre}.bias_ih_l0"] + sd[f"{pre}.bias_hh_l0"])[rows])

        # layer 0: one GEMM [M, 1024] x [1024, 4*512]; column blocks = (L_r on x_r, L_i on x_r, L_r on x_i, L_i on x_i)
        wr, whr, br_ = lstm_block("real_lstm", 0)
        wi, whi, bi_ = lstm_block("imag_lstm", 0)
        zeros = torch.zeros(d * kn[-1], 4 * h, device=dev)
        blocks, biases = [], []
        for wsel, bsel, use_real in ((wr, br_, True), (wi, bi_, True), (wr, br_, False), (wi, bi_, False)):
            blk = zeros.clone()
            mask = is_real if use_real else ~is_real
            blk[mask] = wsel[:, ref_feat[mask]].t()
            blocks.append(blk)
            biases.append(bsel)
        w0 = torch.cat(blocks, dim=1).t().contiguous()           # [2048, 1024]  (N, K) K-major
        P["l0_hi"], P["l0_lo"] = packing.split_tf32(w0)
        P["l0_16"] = packing.pack_linear_f16(w0)
        P["l0_kn"] = packing.pad_cols(w0.t().contiguous())
        P["l0_b"] = torch.cat(biases).contiguous()

        def whh_pack(wh):
            s_ = h // packing.HU
            return wh.reshape(s_, 4 * packing.HU, h).permute(0, 2, 1).contiguous()
        P["whh0"] = torch.stack([whh_pack(whr), whh_pack(whi), whh_pack(whr), whh_pack(whi)]).contiguous()
        # layer 1 input: the four hidden sequences [r2r | r2i | i2r | i2i]; real = r2r - i2i, imag = i2r + r2i
        wr1, whr1, br1 = lstm_block("real_lstm", 1)
        wi1, whi1, bi1 = lstm_block("imag_lstm", 1)
        z = torch.zeros(4 * h, h, device=dev)

        def from_parts(w, sign):       # rows of the [4h_in -> 4h gates] block for input (sign_r2r, r2i, i2r, i2i)
            return torch.cat([w * s if s != 0 else z for s in sign], dim=1)      # [4h, 4*h_in]
        real_in = (1, 0, 0, -1)
        imag_in = (0, 1, 1, 0)
        w1 = torch.cat([from_parts(wr1, real_in), from_parts(wi1, real_in), from_parts(wr1, imag_in),
                        from_parts(wi1, imag_in)], dim=0).contiguous()           # [2048, 512]
        P["l1_hi"], P["l1_lo"] = packing.split_tf32(w1)
        P["l1_16"] = packing.pack_linear_f16(w1)
        P["l1_kn"] = packing.pad_cols(w1.t().contiguous())
        P["l1_b"] = torch.cat([br1, bi1, br1, bi1]).contiguous()
        P["whh1"] = torch.stack([whh_pack(whr1), whh_pack(whi1), whh_pack(whr1), whh_pack(whi1)]).contiguous()
        # projection: real = r_trans(r2r - i2i), imag = i_trans(i2r + r2i) -> channels-last [4, 256] flatten
        rt, it = sd["enhance.1.r_trans.weight"], sd["enhance.1.i_trans.weight"]     # [512, 128]
        wp_r = torch.cat([rt, torch.zeros_like(rt), torch.zeros_like(rt), -rt], dim=1)   # [512, 512]
        wp_i = torch.cat([torch.zeros_like(it), it, it, torch.zeros_like(it)], dim=1)
        bp_r, bp_i = sd["enhance.1.r_trans.bias"], sd["enhance.1.i_trans.bias"]
        wp = torch.zeros(d * kn[-1], 4 * h, device=dev)
        bp = torch.zeros(d * kn[-1], device=dev)
        wp[is_real] = wp_r[ref_feat[is_real]]
        wp[~is_real] = wp_i[ref_feat[~is_real]]
        bp[is_real] = bp_r[ref_feat[is_real]]
        bp[~is_real] = bp_i[ref_feat[~is_real]]
        P["proj_hi"], P["proj_lo"] = packing.split_tf32(wp.contiguous())           # [1024, 512] (N, K)
        P["proj_16"] = packing.pack_linear_f16(wp.contiguous())
        P["proj_kn"] = packing.pad_cols(wp.t().contiguous())
        P["proj_b"] = bp.contiguous()
        # ---- decoder ----
        for i in range(6):
            pre = f"decoder.{i}"
            wr_, wi_ = sd[f"{pre}.0.real_conv.weight"], sd[f"{pre}.0.imag_conv.weight"]   # [Cin/2(cat), Co/2, 5, 2]
            half = wr_.shape[0] // 2      # complex_cat: first half of the real input = decoder path, second = skip
            co2 = wr_.shape[1]
            if i < 5:
                s, o = packing.bn_fold(*(sd[f"{pre}.1.{n}"] for n in ("weight", "bias", "running_mean", "running_var")))
                slope = float(sd[f"{pre}.2.weight"].item())
            else:
                s, o, slope = torch.ones(2 * co2, device=dev), torch.zeros(2 * co2, device=dev), 0.0
            br, bi = sd[f"{pre}.0.real_conv.bias"], sd[f"{pre}.0.imag_conv.bias"]
            bias = (torch.cat([br - bi, br + bi]) * s + o).contiguous()

            def parity(kfs):
                # K order per tap: src0 = decoder path [r | i], src1 = skip [r | i]
                out = []
                for kf in kfs:
                    for kt in range(2):
                        a0 = _stack_complex(wr_[:half], wi_[:half], [(kf, kt)], True)
                        a1 = _stack_complex(wr_[half:], wi_[half:], [(kf, kt)], True)
                        out += [a0, a1]
                return ConvWeights(torch.cat(out, dim=0) * s[None, :], 2 * co2)
            P[f"dec{i}"] = (parity((0, 2, 4)), parity((1, 3)), bias, slope)
            P[f"dec{i}_m"] = conv_engine.merge_parity(P[f"dec{i}"][0], DEC_EVEN, P[f"dec{i}"][1], DEC_ODD)
        self._packed = P

    def _ensure_packed(self):
        key = self._state_key()
        if self._packed is None or key != self._packed_key:
            self._pack()
            self._packed_key = key

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # -- public contract: [B,2,257,T] -> [B,2,257,T] -------------------------------------------------
    @torch.no_grad()
    def forward(self, inputs, lens=None):
        if not inputs.is_cuda:
            raise RuntimeError("DCCRN (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        x = inputs.float().permute(0, 3, 2, 1).contiguous()        # layout plumbing: [B,T,F,2]
        est = self._forward_nhwc(x)
        return est.permute(0, 3, 2, 1)                             # [B,2,F,T] view

    # -- fast path used by decode.enhance_dccrn: channels-last in and out ------------------------------
    def _forward_nhwc(self, x, taps=None):
        """x [B,T,257,2] compressed RI -> est [B,T,257,2]."""
        self._ensure_packed()
        P = self._packed
        b, t, f, _ = x.shape
        assert f == BINS
        dev = x.device
        kn = self.kernel_num
        use_tc = self._use_tc()
        from . import lstm_engine
        f16 = use_tc and lstm_engine.USE_F16_PAIRS      # tensor-core layers on fp16 operand pairs (SE_F16_PAIRS=0: TF32 pairs)
        # encoder: drop the DC bin (DCCRN_cprs.py:166) by pointing at bin 1 with row stride 257
        enc = []
        h = Act(x[:, :, 1:, :].contiguous())                       # [B,T,256,2] (small: 2 channels)
        fin = 256
        f32_of = lambda a: a.value()   # noqa: E731 (debug taps only)
        for i in range(6):
            w, bias, slope = P[f"enc{i}"]
            ci, co = kn[i], kn[i + 1]
            fo = fin // 2
            is_tc = conv_engine.tc_eligible(ci, 0, co, fo, 2, f16)
            out = conv_engine.new_act(b, t, fo, co, dev, want_f32=not is_tc, want_pair=is_tc, f16=f16)
            conv_engine.conv(h, None, b, t, fin, fo, ENC_TAPS, 2, w, bias, "prelu", out, fo, act_param=slope)
            h, fin = out, fo
            enc.append(h)
            if taps is not None:
                taps[f"enc{i}"] = f32_of(h)
        # complex LSTM
        m = b * t
        hid = self.rnn_units // 2
        seq = h.f32.view(m, fin * kn[-1]) if h.f32 is not None else None
        pair = (h.pair[0].view(m, -1), h.pair[1].view(m, -1)) if h.pair is not None else None
        hs = torch.empty(b, t, 4 * hid, device=dev, dtype=torch.float32)
        for l in range(2):
            xp = self._proj(seq, pair, P[f"l{l}_hi"], P[f"l{l}_lo"], P[f"l{l}_kn"], P[f"l{l}_b"], 16 * hid, use_tc,
                            P[f"l{l}_16"] if f16 else None)
            xp = xp.view(b, t, 16 * hid)
            ops.lstm_seq_multi(xp, P[f"whh{l}"], hid, 4, hs)       # the four real LSTM passes, one launch
            seq, pair = hs.view(m, 4 * hid), None
            if l == 0:
                hs = torch.empty(b, t, 4 * hid, device=dev, dtype=torch.float32)
        dec_in = self._proj(seq, None, P["proj_hi"], P["proj_lo"], P["proj_kn"], P["proj_b"], fin * kn[-1], use_tc,
                            P["proj_16"] if f16 else None)
        h = Act(dec_in.view(b, t, fin, kn[-1]))
        if taps is not None:
            taps["rnn_out"] = h.f32
        # decoder
        dt_shift = 0 if self.crop_first else -1
        for i in range(6):
            we, wo, bias, slope = P[f"dec{i}"]
            co = kn[5 - i]
            skip = enc[5 - i]
            fo = 2 * fin
            act = "prelu" if i < 5 else "none"
            c0, c1 = h.shape[-1], skip.shape[-1]
            is_tc = conv_engine.tc_eligible(c0, c1, co, fin, 1, f16)
            # the consumer of dec4 (Cout=2 last layer) and of dec5 (mask kernel) run on fp32
            nxt_tc = i < 4 and conv_engine.tc_eligible(co, kn[4 - i], kn[4 - i], fo, 1, f16)
            out = conv_engine.new_act(b, t, fo, co, dev, want_f32=(not is_tc) or (not nxt_tc),
                                      want_pair=is_tc and nxt_tc, f16=f16)
            ev = [(dt + dt_shift, df) for dt, df in DEC_EVEN]
            od = [(dt + dt_shift, df) for dt, df in DEC_ODD]
            wm = P[f"dec{i}_m"]
            if f16 and is_tc and conv_engine.parity2_eligible(h, skip, wm, fin):
                # both output-column parity classes in one launch (the odd class's taps are among the even class's)
                conv_engine.conv_parity2(h, skip, b, t, fin, fin, fin, ev, wm, bias, act, out, fo, act_param=slope)
            else:
                conv_engine.conv(h, skip, b, t, fin, fin, ev, 1, we, bias, act, out, fo, dst_f0=0, dst_fstep=2,
                                 act_param=slope)
                conv_engine.conv(h, skip, b, t, fin, fin, od, 1, wo, bias, act, out, fo, dst_f0=1, dst_fstep=2,
                                 act_param=slope)
            h, fin = out, fo
            if taps is not None:
                taps[f"dec{i}"] = f32_of(h)
        h = h.f32
        est = torch.empty(b, t, BINS, 2, device=dev, dtype=torch.float32)
        ops.dccrn_mask(h, x[..., 0], x[..., 1], est[..., 0], est[..., 1], mode=self.masking_mode)
        return est

    @staticmethod
    def _use_tc():
        from . import lstm_engine
        return lstm_engine.USE_TENSOR_CORES

    @staticmethod
    def _proj(seq, pair, w_hi, w_lo, w_kn, bias, n, use_tc, w16=None):
        k = (seq if seq is not None else pair[0]).shape[1]
        if w16 is not None and k % 8 == 0 and (pair is not None or seq.shape[0] >= 128) and \
                (pair is None or pair[0].dtype == torch.float16):
            return ops.gemm_f16x3(pair if pair is not None else ops.split_f16(seq), w16[:2], w16[2], bias, n)
        if use_tc and k % 32 == 0 and (pair is not None or seq.shape[0] >= 128):
            a_hi, a_lo = pair if pair is not None else ops.split_tf32(seq)
            return ops.gemm_tf32x3(a_hi, a_lo, w_hi, w_lo, bias, n)
        return ops.linear(seq, w_kn, bias, n)
