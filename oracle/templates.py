"""State-dict templates (key -> shape) of the reference models, so synthetic weights can be drawn
on a box that has neither the reference tree nor its checkpoints.  Shapes as listed by
``torch.load`` on CRN/BEST_MODEL/*.pth and LSTM/BEST_MODEL/*.pth.
TEST / BENCH INFRASTRUCTURE."""
from __future__ import annotations

import os as _os

# key -> shape tables recorded from the reference modules; ONE copy, kept with the product package (which needs them to build
# its parameter trees and cannot import oracle/)
_KEYS_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "sixty-years-of-frequency-domain-monaural-speech-enhancement_b200")


def _bn(prefix, c):
    return {f"{prefix}.weight": (c,), f"{prefix}.bias": (c,), f"{prefix}.running_mean": (c,),
            f"{prefix}.running_var": (c,), f"{prefix}.num_batches_tracked": ()}


def _lstm(prefix, i, h, layers):
    d = {}
    for l in range(layers):
        d[f"{prefix}.weight_ih_l{l}"] = (4 * h, i if l == 0 else h)
        d[f"{prefix}.weight_hh_l{l}"] = (4 * h, h)
        d[f"{prefix}.bias_ih_l{l}"] = (4 * h,)
        d[f"{prefix}.bias_hh_l{l}"] = (4 * h,)
    return d


def crn_template():
    d = {}
    ch = [1, 16, 32, 64, 128, 256]
    for i in range(5):
        d[f"en.en_module.{i}.1.weight"] = (ch[i + 1], ch[i], 2, 3)
        d[f"en.en_module.{i}.1.bias"] = (ch[i + 1],)
        d.update(_bn(f"en.en_module.{i}.2", ch[i + 1]))
    d.update(_lstm("lstm", 1024, 1024, 2))
    for i, (ci, co) in enumerate([(512, 128), (256, 64), (128, 32), (64, 16), (32, 1)]):
        d[f"de.de_module.{i}.0.weight"] = (ci, co, 2, 3)
        d[f"de.de_module.{i}.0.bias"] = (co,)
        d.update(_bn(f"de.de_module.{i}.{3 if i == 3 else 2}", co))
    return d


def lstm_template():
    d = _bn("bn", 161)
    d.update(_lstm("lstm1", 161, 1024, 1))
    d.update(_lstm("lstm2", 1024, 1024, 2))
    d["fc.0.weight"] = (161, 1024)
    d["fc.0.bias"] = (161,)
    return d


def fullsubnet_template():
    d = _lstm("fb_model.sequence_model", 257, 512, 2)
    d["fb_model.fc_output_layer.weight"] = (257, 512)
    d["fb_model.fc_output_layer.bias"] = (257,)
    d.update(_lstm("sb_model.sequence_model", 32, 384, 2))
    d["sb_model.fc_output_layer.weight"] = (2, 384)
    d["sb_model.fc_output_layer.bias"] = (2,)
    return d


def dccrn_template(kernel_num=(32, 64, 128, 256, 256, 256), rnn_units=256):
    """Key ORDER as torch.load lists the shipped DCCRN checkpoints (encoder, decoder, enhance)."""
    kn = [2] + list(kernel_num)
    d = {}
    for i in range(6):
        for part in ("real_conv", "imag_conv"):
            d[f"encoder.{i}.0.{part}.weight"] = (kn[i + 1] // 2, kn[i] // 2, 5, 2)
            d[f"encoder.{i}.0.{part}.bias"] = (kn[i + 1] // 2,)
        d.update(_bn(f"encoder.{i}.1", kn[i + 1]))
        d[f"encoder.{i}.2.weight"] = (1,)
    di = 0
    for idx in range(6, 0, -1):
        for part in ("real_conv", "imag_conv"):
            d[f"decoder.{di}.0.{part}.weight"] = (kn[idx], kn[idx - 1] // 2, 5, 2)
            d[f"decoder.{di}.0.{part}.bias"] = (kn[idx - 1] // 2,)
        if idx != 1:
            d.update(_bn(f"decoder.{di}.1", kn[idx - 1]))
            d[f"decoder.{di}.2.weight"] = (1,)
        di += 1
    h = rnn_units // 2
    in0 = 4 * kn[-1] // 2
    for l in range(2):
        for part in ("real_lstm", "imag_lstm"):
            d[f"enhance.{l}.{part}.weight_ih_l0"] = (4 * h, in0 if l == 0 else h)
            d[f"enhance.{l}.{part}.weight_hh_l0"] = (4 * h, h)
            d[f"enhance.{l}.{part}.bias_ih_l0"] = (4 * h,)
            d[f"enhance.{l}.{part}.bias_hh_l0"] = (4 * h,)
    d["enhance.1.r_trans.weight"] = (in0, h)
    d["enhance.1.r_trans.bias"] = (in0,)
    d["enhance.1.i_trans.weight"] = (in0, h)
    d["enhance.1.i_trans.bias"] = (in0,)
    return d


def gcrn_template():
    """GCRN/GCRN_noncprs.py:86-134 (``Net``): key order as ``Net().state_dict()`` lists it."""
    d = {}
    ch = [2, 16, 32, 64, 128, 256]
    for i in range(1, 6):
        for c in ("conv1", "conv2"):
            d[f"conv{i}.{c}.weight"] = (ch[i], ch[i - 1], 1, 3)
            d[f"conv{i}.{c}.bias"] = (ch[i],)
    for st in (1, 2):
        for g in range(2):
            d.update(_lstm(f"glstm.lstm_list{st}.{g}", 512, 512, 1))
    for n in ("ln1", "ln2"):
        d[f"glstm.{n}.weight"] = (1024,)
        d[f"glstm.{n}.bias"] = (1024,)
    dec = {5: (512, 128), 4: (256, 64), 3: (128, 32), 2: (64, 16), 1: (32, 1)}
    for br in (1, 2):
        for lvl in (5, 4, 3, 2, 1):
            for c in ("conv1", "conv2"):
                d[f"conv{lvl}_t_{br}.{c}.weight"] = (dec[lvl][0], dec[lvl][1], 1, 3)
                d[f"conv{lvl}_t_{br}.{c}.bias"] = (dec[lvl][1],)
    for i in range(1, 6):
        d.update(_bn(f"bn{i}", ch[i]))
    for br in (1, 2):
        for lvl in (5, 4, 3, 2, 1):
            d.update(_bn(f"bn{lvl}_t_{br}", dec[lvl][1]))
    for br in (1, 2):
        d[f"fc{br}.weight"] = (161, 161)
        d[f"fc{br}.bias"] = (161,)
    return d


def dpcrn_template():
    """DPCRN/DPCRN.py:16-179 (``dpcrn``): key order as ``dpcrn().state_dict()`` lists it."""
    d = {}
    ch = [2, 32, 32, 32, 64, 128]
    for i in range(5):
        d[f"en.en_module.{i}.1.weight"] = (ch[i + 1], ch[i], 2, 3)
        d[f"en.en_module.{i}.1.bias"] = (ch[i + 1],)
        d.update(_bn(f"en.en_module.{i}.2", ch[i + 1]))
        d[f"en.en_module.{i}.3.weight"] = (1,)
    for l in range(2):
        for sfx in ("", "_reverse"):
            d[f"dprnn.intra_rnn.weight_ih_l{l}{sfx}"] = (256, 128)
            d[f"dprnn.intra_rnn.weight_hh_l{l}{sfx}"] = (256, 64)
            d[f"dprnn.intra_rnn.bias_ih_l{l}{sfx}"] = (256,)
            d[f"dprnn.intra_rnn.bias_hh_l{l}{sfx}"] = (256,)
    d["dprnn.intra_fc.weight"] = (128, 128)
    d["dprnn.intra_fc.bias"] = (128,)
    d.update(_lstm("dprnn.inter_rnn", 128, 128, 2))
    d["dprnn.inter_fc.weight"] = (128, 128)
    d["dprnn.inter_fc.bias"] = (128,)
    for n in ("ln1", "ln2"):
        d[f"dprnn.{n}.weight"] = (4, 128)
        d[f"dprnn.{n}.bias"] = (4, 128)
    for i, (ci, co) in enumerate([(256, 64), (128, 32), (64, 32), (64, 32), (64, 2)]):
        d[f"de.de_module.{i}.0.weight"] = (ci, co, 2, 3)
        d[f"de.de_module.{i}.0.bias"] = (co,)
        if i < 4:
            bn = 3 if i == 3 else 2
            d.update(_bn(f"de.de_module.{i}.{bn}", co))
            d[f"de.de_module.{i}.{bn + 1}.weight"] = (1,)
    return d


def uformer_template():
    """The 668 state-dict entries of the shipped Uformer checkpoints (names, shapes, order), as listed by
    torch.load on Uformer/BEST_MODEL/*.pth and stored in <package>/uformer_keys.json."""
    import json
    import os
    keys = json.load(open(os.path.join(_KEYS_DIR, "uformer_keys.json")))
    return {k: tuple(shape) for k, shape, _ in keys}


def _cts_norm_rows(d, pre, c, cumulative, dims):
    if cumulative:
        shape = (1, c) + (1,) * dims
        d[pre + ".gain"], d[pre + ".bias"] = shape, shape
    else:
        d[pre + ".weight"], d[pre + ".bias"] = (c,), (c,)


def _cts_codec(d, en_pre, de_pres, cin, cumulative):
    for i in range(5):
        ci, kf = (cin, 5) if i == 0 else (64, 3)
        for n in ("conv", "gate_conv"):
            d[f"{en_pre}.{i}.0.{n}.1.weight"] = (64, ci, 2, kf)
            d[f"{en_pre}.{i}.0.{n}.1.bias"] = (64,)
        _cts_norm_rows(d, f"{en_pre}.{i}.1", 64, cumulative, 2)
        d[f"{en_pre}.{i}.2.weight"] = (64,)
    for de_pre, fc in de_pres:
        for i in range(5):
            co, kf = (1, 5) if i == 4 else (64, 3)
            for n in ("conv", "gate_conv"):
                d[f"{de_pre}.{i}.0.{n}.0.weight"] = (128, co, 2, kf)
                d[f"{de_pre}.{i}.0.{n}.0.bias"] = (co,)
            _cts_norm_rows(d, f"{de_pre}.{i}.1", co, cumulative, 2)
            d[f"{de_pre}.{i}.2.weight"] = (co,)
        d[fc + ".weight"], d[fc + ".bias"] = (161, 161), (161,)


def _cts_tcm(d, pre, j, branches, cumulative):
    d[f"{pre}.in_conv.weight"] = (64, 256, 1)
    for n in branches:
        d[f"{pre}.{n}.0.weight"] = (64,)
        _cts_norm_rows(d, f"{pre}.{n}.1", 64, cumulative, 1)
        d[f"{pre}.{n}.2.weight"] = (1, 1, 2 * 2 ** j - 1)
        d[f"{pre}.{n}.4.weight"] = (64, 64, 5)
    d[f"{pre}.out_conv.0.weight"] = (64,)
    _cts_norm_rows(d, f"{pre}.out_conv.1", 64, cumulative, 1)
    d[f"{pre}.out_conv.2.weight"] = (256, 64, 1)


def ctsnet_step1_template(cumulative=False):
    """CTSNet/Step1_network.py:12-19 (``Step1_net``); cumulative=True: CTSNet_new (gain/bias of the cLN)."""
    d = {}
    _cts_codec(d, "en.en", [("de.de", "de.de6.0")], 1, cumulative)
    # module registration order: en, de (de6 before the de list, Step1_network.py:110-115), tcm1..3
    keys = list(d)
    fc = [k for k in keys if k.startswith("de.de6")]
    rest = [k for k in keys if not k.startswith("de.de6")]
    first_de = next(i for i, k in enumerate(rest) if k.startswith("de.de."))
    order = rest[:first_de] + fc + rest[first_de:]
    d = {k: d[k] for k in order}
    for s in (1, 2, 3):
        for j in range(6):
            _cts_tcm(d, f"tcm{s}.tcm_list.{j}", j, ("left_conv", "right_conv"), cumulative)
    return d


def ctsnet_step2_template(X=6, R=3, cumulative=False):
    """CTSNet/Step2_network.py:13-21 (``Step2_net(X, R)``)."""
    d = {}
    _cts_codec(d, "en.en_module", [("de_r.de_list", "de_r.de6.0"), ("de_i.de_list", "de_i.de6.0")], 4, cumulative)
    for r in range(R):
        for j in range(X):
            _cts_tcm(d, f"tcm_list.{r}.glu_list.{j}", j, ("ori_conv", "att_ori"), cumulative)
    return d


def taylorsenet_template(cumulative=False):
    """TaylorSENet/TaylorSENet.py:8-64 in the configuration of taylorsenet_decode_vb.py:11-13 (811 entries, key order as
    ``TaylorSENet(...).state_dict()`` lists it; recorded from the reference module in <package>/taylor*_keys.json)."""
    import json
    import os
    name = "taylor_new_keys.json" if cumulative else "taylor_keys.json"
    with open(os.path.join(_KEYS_DIR, name)) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}


def g2net_template(cumulative=True):
    """G2Net_new/gaf_net_320.py:10-71 (cumulative LayerNorm) / G2Net_VB (InstanceNorm) in the configuration of
    com_decode.py:23 (825 entries; recorded from the reference module in <package>/g2net_*_keys.json)."""
    import json
    import os
    name = "g2net_new_keys.json" if cumulative else "g2net_vb_keys.json"
    with open(os.path.join(_KEYS_DIR, name)) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}
