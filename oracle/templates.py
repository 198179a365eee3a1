"""State-dict templates (key -> shape) of the reference models, so synthetic weights can be drawn
on a box that has neither the reference tree nor its checkpoints.  Shapes as listed by
``torch.load`` on CRN/BEST_MODEL/*.pth and LSTM/BEST_MODEL/*.pth.
TEST / BENCH INFRASTRUCTURE."""
from __future__ import annotations


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
