"""State-dict-compatible parameter containers.

The drop-in classes must ``load_state_dict()`` the reference's shipped ``BEST_MODEL/*.pth``
unchanged, i.e. expose exactly the reference's key namespace
(``en.en_module.0.1.weight`` ...).  Instead of restating the reference's nn.Sequential
definitions, the key list itself is the specification: nested bare ``nn.Module`` holders are
created from (key, shape, kind) rows, parameters where the reference has parameters, buffers
where it has buffers.  These holders carry no forward(); the arithmetic lives in the CUDA
library.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def build_param_tree(root: nn.Module, spec):
    """spec: iterable of (dotted_key, shape, kind) with kind in {'param', 'buffer', 'counter'}."""
    for key, shape, kind in spec:
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        leaf = parts[-1]
        if kind == "param":
            mod.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=False))
        elif kind == "buffer":
            mod.register_buffer(leaf, torch.zeros(shape))
        elif kind == "counter":
            mod.register_buffer(leaf, torch.zeros((), dtype=torch.int64))
        else:
            raise ValueError(kind)


def bn_rows(prefix, c):
    return [(f"{prefix}.weight", (c,), "param"), (f"{prefix}.bias", (c,), "param"),
            (f"{prefix}.running_mean", (c,), "buffer"), (f"{prefix}.running_var", (c,), "buffer"),
            (f"{prefix}.num_batches_tracked", (), "counter")]


def lstm_rows(prefix, input_size, hidden, layers):
    rows = []
    for l in range(layers):
        i = input_size if l == 0 else hidden
        rows += [(f"{prefix}.weight_ih_l{l}", (4 * hidden, i), "param"),
                 (f"{prefix}.weight_hh_l{l}", (4 * hidden, hidden), "param"),
                 (f"{prefix}.bias_ih_l{l}", (4 * hidden,), "param"),
                 (f"{prefix}.bias_hh_l{l}", (4 * hidden,), "param")]
    return rows
