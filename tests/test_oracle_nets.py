"""Pin oracle.nets / oracle.decode: against the committed fixtures (outputs of the unmodified
reference modules), against the textbook LSTM recurrence, and -- in the build container --
against the reference modules themselves."""
import os

import numpy as np
import pytest
import torch

from conftest import CKPT_DIR, GOLDEN
from oracle import decode, nets, ref_shims, synth, templates
from oracle.make_golden import sd_digest

CASES = {
    "crn_synth": (templates.crn_template, decode.enhance_crn, None),
    "lstm_synth": (templates.lstm_template, decode.enhance_lstm, None),
    "crn_ckpt": (templates.crn_template, decode.enhance_crn, "CRN__wsj0_si84_300h_crn_noncprs_model.pth"),
    "lstm_ckpt": (templates.lstm_template, decode.enhance_lstm, "LSTM__vb_lstm_noncprs_model.pth"),
    "fullsubnet_synth": (templates.fullsubnet_template, decode.enhance_fullsubnet, None),
    "fullsubnet_ckpt": (templates.fullsubnet_template, decode.enhance_fullsubnet,
                        "FullSubNet__wsj0_si84_300h_fullsubnet_cprs_model_512_256.pth"),
    "dccrn_synth": (templates.dccrn_template, decode.enhance_dccrn, None),
    "dccrn_ckpt": (templates.dccrn_template, decode.enhance_dccrn, "DCCRN__wsj0_si84_300h_dccrn_cprs_model.pth"),
    "gcrn_synth": (templates.gcrn_template, decode.enhance_gcrn, None),
    "gcrn_ckpt": (templates.gcrn_template, decode.enhance_gcrn, "GCRN__vb_gcrn_cprs_model.pth"),
    "dpcrn_synth": (templates.dpcrn_template, decode.enhance_dpcrn, None),
    "dpcrn_ckpt": (templates.dpcrn_template, decode.enhance_dpcrn, "DPCRN__vb_dpcrn_noncprs_model.pth"),
    "uformer_synth": (templates.uformer_template, decode.enhance_uformer, None),
    "uformer_ckpt": (templates.uformer_template, decode.enhance_uformer,
                     "Uformer__wsj0_si84_300h_uformer_noncprs_model.pth"),
}


def load_case(name):
    tmpl, enh, ckpt = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if ckpt is None:
        sd = synth.synthetic_state_dict(tmpl(), seed=0, gain=1.0 if name.startswith(("uformer", "dpcrn")) else 2.0)
    else:
        path = os.path.join(CKPT_DIR, ckpt)
        if not os.path.exists(path):
            pytest.skip("checkpoint copy not present (oracle/fetch_checkpoints.py)")
        sd = torch.load(path, map_location="cpu")
    return g, sd, enh


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name):
    g, sd, enh = load_case(name)
    assert sd_digest(sd) == str(g["digest"]), "weights differ from the ones the fixture was made with"
    assert float(g["ref_vs_oracle"]) < 1e-3
    for j in range(len(g["clip_ids"])):
        wav = synth.noisy_clip(int(g["clip_ids"][j]), int(g["nsamp"]))
        assert np.array_equal(wav, g[f"wav{j}"]), "synthetic clip generator is not reproducible"
        kw = {"p": float(g["p"])} if "p" in g.files else {}
        y, taps = enh(sd, wav.astype(np.float64), **kw)
        key = "mask" if "mask" in taps else "est"
        # Uformer fixtures come from the unmodified module (different op order than the restatement)
        tol = 5e-4 if name.startswith("uformer") else 2e-5
        assert np.abs(taps[key] - g[f"{key}{j}"]).max() < tol
        assert np.sqrt(np.mean((taps["y_norm"] - g[f"ynorm{j}"]) ** 2)) < 2e-6


def test_fused_lstm_equals_textbook_recurrence():
    sd = synth.synthetic_state_dict({k: v for k, v in templates.lstm_template().items() if k.startswith("lstm2")})
    x = torch.randn(2, 12, 1024, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a = nets.lstm(x, sd, "lstm2", 2)
        b = nets.lstm_manual(x, sd, "lstm2", 2)
    assert (a - b).abs().max() < 2e-6


@pytest.mark.needs_reference
@pytest.mark.parametrize("mdir,module,cls,tmpl,fwd", [
    ("CRN", "CRN", "crn_net", templates.crn_template, nets.crn_forward),
    ("LSTM", "LSTM", "lstm_net", templates.lstm_template, nets.lstm_net_forward)])
def test_oracle_equals_reference_module(mdir, module, cls, tmpl, fwd):
    mod = ref_shims.import_reference(mdir, module)
    net = getattr(mod, cls)().eval()
    sd = synth.synthetic_state_dict(tmpl(), seed=3)
    net.load_state_dict(sd)
    x = torch.rand(2, 23, 161, generator=torch.Generator().manual_seed(5)) * 4
    with torch.no_grad():
        assert (net(x) - fwd(sd, x)).abs().max() < 1e-6   # same ATen ops; threading may reassociate
