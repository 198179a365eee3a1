"""Drop-in GCRN ``Net`` (reference: GCRN/GCRN_noncprs.py:5-165).

forward(x [B,2,T,161] real/imag planes) -> [B,2,T,161]; same class name, constructor and state-dict keys
as the reference (``conv1.conv1.weight`` ... ``glstm.lstm_list1.0.weight_ih_l0`` ... ``fc2.bias``), so
``Net().load_state_dict(torch.load('BEST_MODEL/vb_gcrn_cprs_model.pth'))`` works unchanged
(GCRN/gcrn_decode_vb.py:16-21).  Inference only.

How the reference's modules map onto the kernels:
  * GluConv2d / GluConvTranspose2d, k(1,3) s(1,2) (:42-83): ``conv1(x) * sigmoid(conv2(x))`` -- the two convs are
    ONE implicit GEMM with 2*Cout output channels [a | b]; gate, eval BatchNorm (it sits AFTER the gate, so it
    cannot be folded into the weights) and ELU are the se_glu_affine_act pass, which also emits the TF32
    split the next tensor-core layer consumes.  ConvTranspose = even / odd output-column parity classes.
  * decoder ``elu(cat(bn(deconv), skip))`` (:149-152): the already-activated encoder output goes through ELU
    a second time; ELU(e_i) is computed once per level (se_unary) and shared by the two decoder branches;
    the concat itself is two source pointers of the next implicit GEMM.
  * GLSTM (:5-39), 2 groups x 2 stages of LSTM(512, 512): every stage is one projection GEMM with a
    block-structured weight [2*2048, 1024] + ONE two-group recurrence launch.  The stack+flatten interleave
    between the stages (:28-29) and the (c,f) <-> channels-last flatten orders (:23,:36) are permutations
    of LayerNorm parameters / projection columns / the LayerNorm store index -- no data movement.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import conv_engine, lstm_engine, ops, packing
from .conv_engine import Act, ConvWeights
from .param_tree import bn_rows, build_param_tree

_CH = [2, 16, 32, 64, 128, 256]               # GCRN_noncprs.py:90-94
_F = [161, 80, 39, 19, 9, 4]
_DEC = {5: (512, 128), 4: (256, 64), 3: (128, 32), 2: (64, 16), 1: (32, 1)}     # :98-108
_DEC_FOUT = {5: 9, 4: 19, 3: 39, 2: 80, 1: 161}                                   # lvl 2 has output_padding (0,1)

ENC_TAPS = [(0, 0), (0, 1), (0, 2)]           # in[t, 2f + kf]
DEC_EVEN = [(0, 0), (0, -1)]                  # f' = 2m  : kf = 0 <- f = m,  kf = 2 <- f = m - 1
DEC_ODD = [(0, 0)]                            # f' = 2m+1: kf = 1 <- f = m


def _spec():
    rows = []
    for i in range(1, 6):
        for c in ("conv1", "conv2"):
            rows += [(f"conv{i}.{c}.weight", (_CH[i], _CH[i - 1], 1, 3), "param"), (f"conv{i}.{c}.bias", (_CH[i],), "param")]
    for st in (1, 2):
        for g in range(2):
            pre = f"glstm.lstm_list{st}.{g}"
            rows += [(pre + ".weight_ih_l0", (2048, 512), "param"), (pre + ".weight_hh_l0", (2048, 512), "param"),
                     (pre + ".bias_ih_l0", (2048,), "param"), (pre + ".bias_hh_l0", (2048,), "param")]
    rows += [("glstm.ln1.weight", (1024,), "param"), ("glstm.ln1.bias", (1024,), "param"),
             ("glstm.ln2.weight", (1024,), "param"), ("glstm.ln2.bias", (1024,), "param")]
    for br in (1, 2):
        for lvl in (5, 4, 3, 2, 1):
            ci, co = _DEC[lvl]
            for c in ("conv1", "conv2"):
                rows += [(f"conv{lvl}_t_{br}.{c}.weight", (ci, co, 1, 3), "param"),
                         (f"conv{lvl}_t_{br}.{c}.bias", (co,), "param")]
    for i in range(1, 6):
        rows += bn_rows(f"bn{i}", _CH[i])
    for br in (1, 2):
        for lvl in (5, 4, 3, 2, 1):
            rows += bn_rows(f"bn{lvl}_t_{br}", _DEC[lvl][1])
    rows += [("fc1.weight", (161, 161), "param"), ("fc1.bias", (161,), "param"),
             ("fc2.weight", (161, 161), "param"), ("fc2.bias", (161,), "param")]
    return rows


class Net(nn.Module):
    N_BINS = 161

    def __init__(self):
        super().__init__()
        build_param_tree(self, _spec())
        self._packed = None
        self._packed_key = None

    # -- weight packing --------------------------------------------------------------------------
    def _state_key(self):
        p = next(self.parameters())
        return (p.device, tuple(int(t._version) for t in self.state_dict().values()))

    def _ensure_packed(self):
        key = self._state_key()
        if self._packed is None or key != self._packed_key:
            self._pack()
            self._packed_key = key

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items() if v.is_floating_point()}
        dev = next(self.parameters()).device
        P = {}

        def bn(pre):
            s, o = packing.bn_fold(sd[pre + ".weight"], sd[pre + ".bias"], sd[pre + ".running_mean"],
                                   sd[pre + ".running_var"])
            return s.contiguous(), o.contiguous()

        def interleaved(w_ab, bias_ab):
            """[a | b] columns -> (a_0, b_0, a_1, b_1, ...): the layout of the fused gate epilogue (se_conv_tc_desc.glu)."""
            c = w_ab.shape[1] // 2
            perm = torch.stack([torch.arange(c), torch.arange(c) + c], 1).reshape(-1).to(w_ab.device)
            return ConvWeights(w_ab[:, perm].contiguous(), 2 * c), bias_ab[perm].contiguous()

        for i in range(1, 6):
            w1, w2 = sd[f"conv{i}.conv1.weight"], sd[f"conv{i}.conv2.weight"]             # [Co, Ci, 1, 3]
            w = torch.cat([torch.cat([w1[:, :, 0, kf].t(), w2[:, :, 0, kf].t()], 1) for kf in range(3)], 0)
            bias = torch.cat([sd[f"conv{i}.conv1.bias"], sd[f"conv{i}.conv2.bias"]]).contiguous()
            P[f"enc{i}"] = (ConvWeights(w.contiguous(), 2 * _CH[i]), bias, *bn(f"bn{i}"))   # K = (kf, ci), N = [a | b]
            P[f"enc{i}_il"] = interleaved(w, bias)
        for br in (1, 2):
            for lvl in (5, 4, 3, 2, 1):
                w1, w2 = sd[f"conv{lvl}_t_{br}.conv1.weight"], sd[f"conv{lvl}_t_{br}.conv2.weight"]   # [Ci, Co, 1, 3]
                co = w1.shape[1]
                tap = lambda kf: torch.cat([w1[:, :, 0, kf], w2[:, :, 0, kf]], 1)      # noqa: E731  [Ci, 2Co]
                bias = torch.cat([sd[f"conv{lvl}_t_{br}.conv1.bias"], sd[f"conv{lvl}_t_{br}.conv2.bias"]]).contiguous()
                P[f"dec{lvl}_{br}"] = (ConvWeights(torch.cat([tap(0), tap(2)], 0).contiguous(), 2 * co),
                                       ConvWeights(tap(1).contiguous(), 2 * co), bias, *bn(f"bn{lvl}_t_{br}"))
                P[f"dec{lvl}_{br}_il"] = (interleaved(torch.cat([tap(0), tap(2)], 0), bias)[0],
                                          *interleaved(tap(1), bias))
            P[f"fc{br}"] = (packing.pad_cols(sd[f"fc{br}.weight"].t().contiguous()), sd[f"fc{br}.bias"].contiguous())

        # ---- GLSTM -----------------------------------------------------------------------------------
        # channels-last flatten of e5 [B,T,4,256]: column q = f*256 + c  <->  reference feature c*4 + f (:23)
        q = torch.arange(1024, device=dev)
        ref_of_q = (q % 256) * 4 + q // 256
        col_of_ref = torch.empty_like(q)
        col_of_ref[ref_of_q] = q

        def stage(st, in_cols):
            """in_cols[g][i]: which of MY 1024 input columns feeds reference input i of group g."""
            blocks, biases, whh = [], [], []
            for g in range(2):
                pre = f"glstm.lstm_list{st}.{g}"
                lay = packing.pack_lstm_layer(sd[pre + ".weight_ih_l0"], sd[pre + ".weight_hh_l0"],
                                              sd[pre + ".bias_ih_l0"], sd[pre + ".bias_hh_l0"])
                full = torch.zeros(2048, 1024, device=dev)
                full[:, in_cols[g]] = lay["wih_kn"][:, :2048].t()              # exact fp32 rows in slice order
                blocks.append(full)
                biases.append(lay["bias"])
                whh.append(lay["whh"])
            w = torch.cat(blocks, 0).contiguous()                          # [2 * 2048, 1024]  (N, K)
            hi, lo = packing.split_tf32(w)
            # the two H = 512 groups as ONE block-diagonal H = 1024 recurrence: slice 64 g + s of the big LSTM is slice s
            # of group g and only sees k in [512 g, 512 g + 512).  Twice the FLOPs, but the tcgen05 cluster recurrence
            # (H = 1024 only; 6.9 us/step, latency-bound) beats the fp32 slice kernel the grouped call runs on (8.9)
            whh_bd = torch.zeros(128, 1024, 32, device=dev)
            for g in range(2):
                whh_bd[64 * g:64 * g + 64, 512 * g:512 * g + 512] = whh[g]
            return {"wih_hi": hi, "wih_lo": lo, "wih_kn": packing.pad_cols(w.t().contiguous()),
                    "bias": torch.cat(biases).contiguous(), "whh": torch.stack(whh).contiguous(),
                    "whh_bd": whh_bd.contiguous(),
                    "hidden": 1024}      # 2 groups x 512: the projection has 4 * 1024 output columns

        P["st1"] = stage(1, [col_of_ref[512 * g:512 * g + 512] for g in range(2)])
        # stage-1 output column m = g*512 + j sits at reference position 2j + g after stack+flatten (:28-29)
        m = torch.arange(1024, device=dev)
        ref_of_m = (m % 512) * 2 + m // 512
        col_of_ref1 = torch.empty_like(m)
        col_of_ref1[ref_of_m] = m
        P["ln1"] = (sd["glstm.ln1.weight"][ref_of_m].contiguous(), sd["glstm.ln1.bias"][ref_of_m].contiguous())
        P["st2"] = stage(2, [col_of_ref1[512 * g:512 * g + 512] for g in range(2)])
        # stage-2 output is in reference order r = c*4 + f; LayerNorm stores channel r at f*256 + c (:36)
        P["ln2"] = (sd["glstm.ln2.weight"].contiguous(), sd["glstm.ln2.bias"].contiguous())
        r = torch.arange(1024, device=dev)
        P["ln2_index"] = ((r % 4) * 256 + r // 4).to(torch.int32).contiguous()
        self._packed = P

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("GCRN Net (se_b200) runs on CUDA sm_100a only; there is no CPU path")
        return self._forward_impl(x, taps)

    def _forward_impl(self, x, taps=None):
        assert x.dim() == 4 and x.shape[1] == 2 and x.shape[3] == self.N_BINS, tuple(x.shape)
        y = self.forward_nhwc(x.float().permute(0, 2, 3, 1).contiguous(), taps)      # [2][B,T,161]
        return torch.stack(y, dim=1)

    def forward_nhwc(self, x, taps=None):
        """x [B,T,161,2] channels-last -> (real [B,T,161], imag [B,T,161])."""
        self._ensure_packed()
        P = self._packed
        b, t = x.shape[0], x.shape[1]
        dev = x.device
        tc = conv_engine.tc_eligible

        def glu_layer(src, skip, fin, fout_classes, w_classes, bias, scale, shift, fout, taps_classes, sf, cons_tc,
                      want_f32, il=None):
            """One gated conv (all parity classes) + gate/BN/ELU.  cons_tc: the consumer reads the TF32 split.
            il = (interleaved weights per class, interleaved bias): on tensor-core layers gate, BatchNorm and ELU run in
            the conv epilogue and the [a | b] intermediate never exists."""
            co2 = w_classes[0].cout
            c0 = src.shape[-1]
            c1 = skip.shape[-1] if skip is not None else 0
            if il is not None and all(tc(c0, c1, co2, fo, sf) for fo in fout_classes):
                wil, bil = il
                out = conv_engine.new_act(b, t, fout, co2 // 2, dev, want_f32=want_f32 or not cons_tc, want_pair=cons_tc)
                for cls, (w, tp, fo) in enumerate(zip(wil, taps_classes, fout_classes)):
                    conv_engine.conv(src, skip, b, t, fin, fo, tp, sf, w, bil, "elu", out, fout, dst_f0=cls,
                                     dst_fstep=len(wil), glu=(scale, shift))
                return out
            tmp = Act(torch.empty(b, t, fout, co2, device=dev, dtype=torch.float32))
            for cls, (w, tp, fo) in enumerate(zip(w_classes, taps_classes, fout_classes)):
                step = len(w_classes)
                conv_engine.conv(src, skip, b, t, fin, fo, tp, sf, w, bias, "none", tmp, fout, dst_f0=cls, dst_fstep=step)
            f32, pair = ops.glu_affine_act(tmp.f32, scale, shift, "elu", want_f32=want_f32 or not cons_tc,
                                           want_pair=cons_tc)
            return Act(f32, pair)

        # ---- encoder (:138-142) ----
        enc = []
        h = Act(x)
        for i in range(1, 6):
            w, bias, s, o = P[f"enc{i}"]
            # consumers of e_i: conv_{i+1} (or the LSTM projection / first decoder layer for e5) and ELU(e_i)
            cons_tc = tc(_CH[i], 0, 2 * _CH[i + 1], _F[i + 1], 2) if i < 5 else lstm_engine.USE_TENSOR_CORES
            h = glu_layer(h, None, _F[i - 1], [_F[i]], [w], bias, s, o, _F[i], [ENC_TAPS], 2, cons_tc,
                          want_f32=(i < 5),          # ELU(e_i) (se_unary) reads the fp32 copy
                          il=([P[f"enc{i}_il"][0]], P[f"enc{i}_il"][1]))
            enc.append(h)
            if taps is not None:
                taps[f"e{i}"] = h.f32 if h.f32 is not None else h.pair[0] + h.pair[1]
        e5 = enc[4]

        # ---- grouped LSTM (:22-39) ----
        seq = e5.f32.view(b * t, 1024) if e5.f32 is not None else None
        pair = (e5.pair[0].view(b * t, 1024), e5.pair[1].view(b * t, 1024)) if e5.pair is not None else None
        tcl = lstm_engine.USE_TENSOR_CORES
        for st in (1, 2):
            lay = P[f"st{st}"]
            xp = lstm_engine.input_projection(seq, lay, pair)                      # [B*T, 2 * 2048]
            if tcl:      # block-diagonal H = 1024 recurrence on the tcgen05 cluster kernel
                hs = ops.lstm_seq(xp.view(b, t, 4096), lay["whh_bd"], 1024)
            else:
                hs = torch.empty(b, t, 1024, device=dev, dtype=torch.float32)
                ops.lstm_seq_multi(xp.view(b, t, 4096), lay["whh"], 512, 2, hs)
            g, be = P[f"ln{st}"]
            seq, pair = ops.group_layernorm(hs.view(b * t, 1024), 1, g, be, want_f32=not tcl, want_pair=tcl,
                                            out_index=P["ln2_index"] if st == 2 else None)
        d0 = Act(seq.view(b, t, 4, 256) if seq is not None else None,
                 (pair[0].view(b, t, 4, 256), pair[1].view(b, t, 4, 256)) if pair is not None else None)
        if taps is not None:
            taps["glstm_nhwc"] = d0.f32 if d0.f32 is not None else d0.pair[0] + d0.pair[1]

        # ---- ELU(e_i) for the decoder concats (:152) ----
        skips = {5: e5}
        for lvl in (4, 3, 2, 1):
            e = enc[lvl - 1]
            cons_tc = tc(_CH[lvl], _CH[lvl], 2 * _DEC[lvl][1], (_DEC_FOUT[lvl] + 1) // 2, 1)
            f32, pr = ops.unary(e.f32, "elu", want_f32=not cons_tc, want_pair=cons_tc)
            skips[lvl] = Act(f32, pr)

        # ---- two decoders (:149-162) ----
        outs = []
        for br in (1, 2):
            d = d0
            for lvl in (5, 4, 3, 2, 1):
                we, wo, bias, s, o = P[f"dec{lvl}_{br}"]
                fin, fout = _F[lvl], _DEC_FOUT[lvl]
                if lvl > 1:
                    nco = _DEC[lvl][1]
                    cons_tc = tc(nco, nco, 2 * _DEC[lvl - 1][1], (_DEC_FOUT[lvl - 1] + 1) // 2, 1)
                else:
                    cons_tc = False
                wie, wio, bil = P[f"dec{lvl}_{br}_il"]
                d = glu_layer(d, skips[lvl], fin, [(fout + 1) // 2, fout // 2], [we, wo], bias, s, o, fout,
                              [DEC_EVEN, DEC_ODD], 1, cons_tc, want_f32=False, il=([wie, wio], bil))
            if taps is not None:
                taps[f"d1_{br}"] = d.f32
            wfc, bfc = P[f"fc{br}"]
            outs.append(ops.linear(d.f32.view(b * t, self.N_BINS), wfc, bfc, self.N_BINS).view(b, t, self.N_BINS))
        return outs
