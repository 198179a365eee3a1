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
    "dccrn_snr_ckpt": (templates.dccrn_template, decode.enhance_dccrn, "DCCRN_SNR__wsj0_si84_300h_dccrn_snr_model.pth"),
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
        if "crop_first" in g.files:
            kw["crop_first"] = bool(g["crop_first"])
        y, taps = enh(sd, wav.astype(np.float64), **kw)
        key = "mask" if "mask" in taps else "est"
        # Uformer fixtures come from the unmodified module (different op order than the restatement)
        tol = 5e-4 if name.startswith("uformer") else 2e-5
        assert np.abs(taps[key] - g[f"{key}{j}"]).max() < tol
        assert np.sqrt(np.mean((taps["y_norm"] - g[f"ynorm{j}"]) ** 2)) < 2e-6


@pytest.mark.parametrize("name,nsamp", [("dccrn_ckpt", 64000), ("fullsubnet_ckpt", 160000)])
def test_oracle_reproduces_config_length_record(name, nsamp):
    """The clip at BASELINE configs[2] / configs[3] length (4 s DCCRN, 10 s FullSubNet) stored by make_golden.py, where
    the reference module and the oracle agreed to ``long_ref_vs_oracle``."""
    g, sd, enh = load_case(name)
    assert float(g["long_ref_vs_oracle"]) < 1e-4
    wav = synth.noisy_clip(int(g["long_clip_id"]), nsamp)
    _, taps = enh(sd, wav.astype(np.float64))
    assert np.sqrt(np.mean((taps["y_norm"] - g["long_ynorm"]) ** 2)) < 2e-6


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


@pytest.mark.needs_reference
@pytest.mark.parametrize("mode", ["E", "C", "R"])
def test_dccrn_oracle_mask_modes_equal_reference_module(mode):
    """All three masking modes of DCCRN.forward (DCCRN_cprs.py:206-224; the scripts only use 'E') against the
    unmodified class (run with the restated complexnn)."""
    mod = ref_shims.import_reference("DCCRN", "DCCRN_cprs")
    net = mod.DCCRN(rnn_units=256, masking_mode=mode, use_clstm=True, kernel_num=[32, 64, 128, 256, 256, 256]).eval()
    sd = synth.synthetic_state_dict(templates.dccrn_template(), seed=2)
    net.load_state_dict(sd)
    x = torch.randn(1, 2, 257, 7, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = net(x)
        got = nets.dccrn_forward(sd, x, masking_mode=mode)
    assert (ref - got).abs().max() < 1e-5 * max(1.0, ref.abs().max().item())


CTS_CKPTS = {"ctsnet_ckpt": ("CTSNet__step1_vb_cts_noncprs_model_final.pth", "CTSNet__step2_vb_cts_noncprs_model.pth"),
             "ctsnet_new_ckpt": ("CTSNet_new__step1_vb_cts_cprs_model_final.pth", "CTSNet_new__step2_vb_cts_cprs_model.pth")}


def load_cts_case(name):
    """(fixture, (sd1, sd2), p, cumulative) of a CTSNet fixture; skips when the checkpoint copy is absent."""
    from oracle.make_golden import cts_state_dicts
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cum = "_new" in name
    if name in CTS_CKPTS:
        paths = [os.path.join(CKPT_DIR, f) for f in CTS_CKPTS[name]]
        if not all(os.path.exists(p) for p in paths):
            pytest.skip("checkpoint copy not present (oracle/fetch_checkpoints.py)")
        sds = tuple(torch.load(p, map_location="cpu") for p in paths)
    else:
        sds = cts_state_dicts(None, None, None, cum)
    return g, sds, float(g["p"]), cum


@pytest.mark.parametrize("name", ["ctsnet_synth", "ctsnet_ckpt", "ctsnet_new_synth", "ctsnet_new_ckpt"])
def test_ctsnet_oracle_reproduces_golden(name):
    """Two-stage CTSNet (InstanceNorm) and CTSNet_new (cumulative LayerNorm): the restated decode loop reproduces
    the fixtures whose network outputs came from the UNMODIFIED Step1_net / Step2_net modules."""
    g, sds, p, cum = load_cts_case(name)
    assert sd_digest(sds[0]) + sd_digest(sds[1]) == str(g["digest"])
    assert float(g["ref_vs_oracle"]) < 1e-4
    for j in range(len(g["clip_ids"])):
        wav = synth.noisy_clip(int(g["clip_ids"][j]), int(g["nsamp"]))
        assert np.array_equal(wav, g[f"wav{j}"])
        y, taps = decode.enhance_ctsnet(sds, wav.astype(np.float64), p=p, cumulative=cum)
        assert np.abs(taps["est"] - g[f"est{j}"]).max() < 1e-4 * max(1.0, np.abs(g[f"est{j}"]).max())
        assert np.sqrt(np.mean((taps["y_norm"] - g[f"ynorm{j}"]) ** 2)) < 2e-6


@pytest.mark.needs_reference
@pytest.mark.parametrize("mdir,cum", [("CTSNet", False), ("CTSNet_new", True)])
def test_ctsnet_oracle_equals_reference_modules(mdir, cum):
    m1 = ref_shims.import_reference(mdir, "Step1_network").Step1_net().eval()
    m2 = ref_shims.import_reference(mdir, "Step2_network").Step2_net(X=6, R=3).eval()
    sd1 = synth.synthetic_state_dict(templates.ctsnet_step1_template(cum), seed=4, gain=1.0)
    sd2 = synth.synthetic_state_dict(templates.ctsnet_step2_template(cumulative=cum), seed=5, gain=1.0)
    m1.load_state_dict(sd1)
    m2.load_state_dict(sd2)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(2, 21, 161, generator=g) * 3
    z = torch.randn(2, 4, 21, 161, generator=g)
    with torch.no_grad():
        assert (m1(x) - nets.ctsnet_step1_forward(sd1, x, cum)).abs().max() < 1e-5
        assert (m2(z) - nets.ctsnet_step2_forward(sd2, z, cumulative=cum)).abs().max() < 1e-5


TAYLOR_CKPTS = {"taylor_ckpt": "TaylorSENet__vb_taylor_noncprs_model.pth",
                "taylor_new_ckpt": "TaylorSENet_new__vb_taylor_cprs_model.pth"}


def load_taylor_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cum = "_new" in name
    if name in TAYLOR_CKPTS:
        path = os.path.join(CKPT_DIR, TAYLOR_CKPTS[name])
        if not os.path.exists(path):
            pytest.skip("checkpoint copy not present (oracle/fetch_checkpoints.py)")
        sd = torch.load(path, map_location="cpu")
    else:
        sd = synth.synthetic_state_dict(templates.taylorsenet_template(cum), seed=0, gain=1.0)
    return g, sd, float(g["p"]), cum


@pytest.mark.parametrize("name", ["taylor_synth", "taylor_ckpt", "taylor_new_ckpt"])
def test_taylorsenet_oracle_reproduces_golden(name):
    g, sd, p, cum = load_taylor_case(name)
    assert sd_digest(sd) == str(g["digest"])
    assert float(g["ref_vs_oracle"]) < 1e-4
    for j in range(len(g["clip_ids"])):
        wav = synth.noisy_clip(int(g["clip_ids"][j]), int(g["nsamp"]))
        assert np.array_equal(wav, g[f"wav{j}"])
        y, taps = decode.enhance_taylorsenet(sd, wav.astype(np.float64), p=p, cumulative=cum)
        assert np.abs(taps["est"] - g[f"est{j}"]).max() < 1e-4 * max(1.0, np.abs(g[f"est{j}"]).max())
        assert np.sqrt(np.mean((taps["y_norm"] - g[f"ynorm{j}"]) ** 2)) < 2e-6


@pytest.mark.needs_reference
@pytest.mark.parametrize("mdir,cum", [("TaylorSENet", False), ("TaylorSENet_new", True)])
def test_taylorsenet_oracle_equals_reference_module(mdir, cum):
    from oracle.make_golden import TAYLOR_KW
    net = ref_shims.import_reference(mdir, "TaylorSENet").TaylorSENet(**TAYLOR_KW).eval()
    sd = synth.synthetic_state_dict(templates.taylorsenet_template(cum), seed=4, gain=1.0)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd)
    x = torch.randn(2, 2, 19, 161, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        assert (net(x) - nets.taylorsenet_forward(sd, x, cum)).abs().max() < 1e-5


G2_CKPTS = {"g2net_new_ckpt": "G2Net_new__vb_gaf_cprs_model.pth", "g2net_vb_ckpt": "G2Net_VB__vb_gaf_noncprs_model.pth"}


def load_g2net_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cum = "_vb" not in name
    if name in G2_CKPTS:
        path = os.path.join(CKPT_DIR, G2_CKPTS[name])
        if not os.path.exists(path):
            pytest.skip("checkpoint copy not present (oracle/fetch_checkpoints.py)")
        sd = torch.load(path, map_location="cpu")
    else:
        sd = synth.synthetic_state_dict(templates.g2net_template(cum), seed=0, gain=1.0)
    return g, sd, float(g["p"]), cum


@pytest.mark.parametrize("name", ["g2net_synth", "g2net_new_ckpt", "g2net_vb_ckpt"])
def test_g2net_oracle_reproduces_golden(name):
    g, sd, p, cum = load_g2net_case(name)
    assert sd_digest(sd) == str(g["digest"])
    assert float(g["ref_vs_oracle"]) < 1e-4
    for j in range(len(g["clip_ids"])):
        wav = synth.noisy_clip(int(g["clip_ids"][j]), int(g["nsamp"]))
        assert np.array_equal(wav, g[f"wav{j}"])
        y, taps = decode.enhance_g2net(sd, wav.astype(np.float64), p=p, cumulative=cum)
        assert np.abs(taps["est"] - g[f"est{j}"]).max() < 1e-4 * max(1.0, np.abs(g[f"est{j}"]).max())
        assert np.sqrt(np.mean((taps["y_norm"] - g[f"ynorm{j}"]) ** 2)) < 2e-6


@pytest.mark.needs_reference
@pytest.mark.parametrize("mdir,cum", [("G2Net_new", True), ("G2Net_VB", False)])
def test_g2net_oracle_equals_reference_module(mdir, cum):
    from oracle.make_golden import G2_ARGS, G2_KW
    net = ref_shims.import_reference(mdir, "gaf_net_320").gaf_base(*G2_ARGS, **G2_KW).eval()
    sd = synth.synthetic_state_dict(templates.g2net_template(cum), seed=4, gain=1.0)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd)
    x = torch.randn(2, 2, 19, 161, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        for a, b in zip(net(x), nets.g2net_forward(sd, x, cum)):
            assert (a - b).abs().max() < 1e-5
