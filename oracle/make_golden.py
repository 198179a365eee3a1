"""Generate the committed fixtures under tests/golden/ by running the UNMODIFIED reference
modules (imported from /root/reference through oracle.ref_shims) -- build container only.

    python -m oracle.make_golden

For every case the reference nn.Module produces the network output; the oracle restatement
(oracle.nets / oracle.decode) is checked against it on the spot and the max-abs difference is
stored in the fixture (``ref_vs_oracle``).  Cases:
  *_synth : weights from oracle.synth.synthetic_state_dict (reproducible anywhere);
  *_ckpt  : the shipped checkpoint the matching decode script loads (needs the checkpoint copy
            under checkpoints/_ref/ at test time -- see oracle/fetch_checkpoints.py).
"""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from . import decode, ref_shims, synth, templates

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, model dir, module, class, checkpoint, enhance fn, template, clip samples, clip ids
    ("crn_synth", "CRN", "CRN", "crn_net", None, decode.enhance_crn, templates.crn_template, 8000, (0, 1)),
    ("crn_ckpt", "CRN", "CRN", "crn_net", "wsj0_si84_300h_crn_noncprs_model.pth", decode.enhance_crn,
     templates.crn_template, 16000, (0, 1)),
    ("lstm_synth", "LSTM", "LSTM", "lstm_net", None, decode.enhance_lstm, templates.lstm_template, 8000, (2, 3)),
    ("lstm_ckpt", "LSTM", "LSTM", "lstm_net", "vb_lstm_noncprs_model.pth", decode.enhance_lstm,
     templates.lstm_template, 16000, (2, 3)),
]


FSN_ARGS = dict(sb_num_neighbors=15, fb_num_neighbors=0, num_freqs=257, look_ahead=2, sequence_model="LSTM",
                fb_output_activate_function="ReLU", sb_output_activate_function=None, fb_model_hidden_size=512,
                sb_model_hidden_size=384, weight_init=False, norm_type="offline_laplace_norm",
                num_groups_in_drop_band=2)     # FullSubNet/fullsubnet_sa_decode.py:11-24

FSN_CASES = [
    ("fullsubnet_synth", None, 8192, (4, 5), None),
    ("fullsubnet_ckpt", "wsj0_si84_300h_fullsubnet_cprs_model_512_256.pth", 16000, (4, 5), 41),
]


def make_fullsubnet():
    mod = ref_shims.import_reference("FullSubNet", "fullsubnet_net_sa.model")
    for name, ckpt, nsamp, clip_ids, long_id in FSN_CASES:
        net = mod.Model(**FSN_ARGS).eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.fullsubnet_template(), seed=0)
        else:
            sd = torch.load(ref_shims.checkpoint_path("FullSubNet", ckpt), map_location="cpu")
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp)}
        worst = 0.0
        if long_id is not None:
            # BASELINE configs[3] clip length (10 s, T = 626 + 2 look-ahead frames): see make_dccrn
            wav = synth.noisy_clip(long_id, 160000)
            _, taps = decode.enhance_fullsubnet(sd, wav.astype(np.float64))
            with torch.no_grad():
                mask_ref = net(torch.from_numpy(taps["mag"])[None, None]).squeeze(0).numpy()
            rec["long_clip_id"] = np.array(long_id)
            rec["long_ynorm"] = taps["y_norm"]
            rec["long_ref_vs_oracle"] = np.array(float(np.abs(mask_ref - taps["mask"]).max()))
            print(f"{name}: 10 s clip, ref_vs_oracle max-abs {float(rec['long_ref_vs_oracle']):.3e}")
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_fullsubnet(sd, wav.astype(np.float64))
            with torch.no_grad():     # B = 1, exactly as the script calls it (no drop_band)
                mask_ref = net(torch.from_numpy(taps["mag"])[None, None]).squeeze(0).numpy()
            worst = max(worst, float(np.abs(mask_ref - taps["mask"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"mag{j}"] = taps["mag"]
            rec[f"mask{j}"] = mask_ref.astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


# (fixture, reference dir, module, checkpoint, samples, clips, p, crop_first, clip id of the config-length record or None)
DCCRN_CASES = [
    ("dccrn_synth", "DCCRN", "DCCRN_cprs", None, 8000, (6, 7), 0.5, True, None),
    ("dccrn_ckpt", "DCCRN", "DCCRN_cprs", "wsj0_si84_300h_dccrn_cprs_model.pth", 16000, (6, 7), 0.5, True, 40),
    # DCCRN_SNR/DCCRN.py (decoder keeps [..., :-1]) with the checkpoint dccrn_decode_snr.py:13 loads
    ("dccrn_snr_ckpt", "DCCRN_SNR", "DCCRN", "wsj0_si84_300h_dccrn_snr_model.pth", 16000, (8, 9), 1.0, False, None),
]


def make_dccrn():
    for name, mdir, module, ckpt, nsamp, clip_ids, p, crop_first, long_id in DCCRN_CASES:
        mod = ref_shims.import_reference(mdir, module)    # runs with the restated complexnn injected
        kw = dict(masking_mode='E') if mdir == "DCCRN" else {}
        net = mod.DCCRN(rnn_units=256, use_clstm=True, kernel_num=[32, 64, 128, 256, 256, 256], **kw).eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.dccrn_template(), seed=0)
        else:
            sd = torch.load(ref_shims.checkpoint_path(mdir, ckpt), map_location="cpu")
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp),
               "p": np.array(p), "crop_first": np.array(crop_first)}
        worst = 0.0
        if long_id is not None:
            # BASELINE configs[2] clip length (4 s, T = 501): the oracle is pinned to the reference module at this
            # length too, and its decode of ONE clip is committed (the GPU test decodes a second one live)
            wav = synth.noisy_clip(long_id, 64000)
            _, taps = decode.enhance_dccrn(sd, wav.astype(np.float64), p=p, crop_first=crop_first)
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["feat"])[None]).squeeze(0).numpy()
            rec["long_clip_id"] = np.array(long_id)
            rec["long_ynorm"] = taps["y_norm"]
            rec["long_ref_vs_oracle"] = np.array(float(np.abs(est_ref - taps["est"]).max()))
            print(f"{name}: 4 s clip, ref_vs_oracle max-abs {float(rec['long_ref_vs_oracle']):.3e}")
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_dccrn(sd, wav.astype(np.float64), p=p, crop_first=crop_first)
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["feat"])[None]).squeeze(0).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"feat{j}"] = taps["feat"]
            rec[f"est{j}"] = est_ref.astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


GCRN_CASES = [
    ("gcrn_synth", None, 8000, (10, 11)),
    ("gcrn_ckpt", "vb_gcrn_cprs_model.pth", 16000, (10, 11)),
]


def make_gcrn():
    """Fixtures from the UNMODIFIED GCRN/GCRN_noncprs.py ``Net`` (synthetic weights + the shipped vb
    checkpoint, p = 0.5 as GCRN/gcrn_decode_vb.py runs it)."""
    mod = ref_shims.import_reference("GCRN", "GCRN_noncprs")
    for name, ckpt, nsamp, clip_ids in GCRN_CASES:
        net = mod.Net().eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.gcrn_template(), seed=0)
        else:
            sd = torch.load(ref_shims.checkpoint_path("GCRN", ckpt), map_location="cpu")
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_gcrn(sd, wav.astype(np.float64))
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["feat"])[None]).squeeze(0).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"feat{j}"] = taps["feat"]
            rec[f"est{j}"] = est_ref.astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


DPCRN_CASES = [
    ("dpcrn_synth", None, 8000, (12, 13), 0.5),
    ("dpcrn_ckpt", "vb_dpcrn_noncprs_model.pth", 16000, (12, 13), 1.0),       # as dpcrn_decode_vb.py runs it
]


def make_dpcrn():
    """Fixtures from the UNMODIFIED DPCRN/DPCRN.py ``dpcrn`` (synthetic weights with p = 0.5 as drcrn_decode.py,
    the shipped vb checkpoint with p = 1.0 as dpcrn_decode_vb.py)."""
    mod = ref_shims.import_reference("DPCRN", "DPCRN")
    for name, ckpt, nsamp, clip_ids, p in DPCRN_CASES:
        net = mod.dpcrn().eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.dpcrn_template(), seed=0, gain=1.0)
        else:
            sd = torch.load(ref_shims.checkpoint_path("DPCRN", ckpt), map_location="cpu")
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp),
               "p": np.array(p)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_dpcrn(sd, wav.astype(np.float64), p=p)
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["feat"])[None]).squeeze(0).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"feat{j}"] = taps["feat"]
            rec[f"est{j}"] = est_ref.astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


UF_CASES = [
    ("uformer_synth", None, 8000, (8, 9), 0),
    ("uformer_ckpt", "wsj0_si84_300h_uformer_noncprs_model.pth", 16000, (8, 9), 64000),
]


def make_uformer():
    """Fixtures from the UNMODIFIED reference Uformer; the functional restatement
    (oracle.nets.uformer_forward / oracle.decode.enhance_uformer) is checked against it on the spot."""
    from . import uformer_ref
    for name, ckpt, nsamp, clip_ids, long_n in UF_CASES:
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.uformer_template(), seed=0, gain=1.0)
        else:
            sd = torch.load(ref_shims.checkpoint_path("Uformer", ckpt), map_location="cpu")
        net = uformer_ref.build(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_uformer_ref(net, wav.astype(np.float64))
            _, t2 = decode.enhance_uformer(sd, wav.astype(np.float64))
            worst = max(worst, float(np.abs(t2["y_norm"] - taps["y_norm"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"est{j}"] = taps["est"].astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        if long_n:     # a 4 s clip exercises every dilation (up to 128 frames) and T = 401 attention
            wav = synth.noisy_clip(40, long_n)
            y, taps = decode.enhance_uformer_ref(net, wav.astype(np.float64))
            rec["long_clip_id"] = np.array(40)
            rec["long_ynorm"] = taps["y_norm"]
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle (waveform max-abs) {worst:.3e}, out rms {float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


CTS_CASES = [
    # name, model dir, stage-1 ckpt, stage-2 ckpt, samples, clip ids, p, cumulative
    ("ctsnet_synth", "CTSNet", None, None, 8000, (14, 15), 0.5, False),
    ("ctsnet_ckpt", "CTSNet", "step1_vb_cts_noncprs_model_final.pth", "step2_vb_cts_noncprs_model.pth", 16000, (14, 15),
     1.0, False),                                                       # as CTSNet/two_stage_com_decode_vb.py runs it
    ("ctsnet_new_synth", "CTSNet_new", None, None, 8000, (16, 17), 1.0, True),
    ("ctsnet_new_ckpt", "CTSNet_new", "step1_vb_cts_cprs_model_final.pth", "step2_vb_cts_cprs_model.pth", 16000, (16, 17),
     0.5, True),                                                        # as CTSNet_new/two_stage_com_decode_vb.py
]


def cts_state_dicts(mdir, ck1, ck2, cumulative, loader=None):
    if ck1 is None:
        return (synth.synthetic_state_dict(templates.ctsnet_step1_template(cumulative), seed=0, gain=1.0),
                synth.synthetic_state_dict(templates.ctsnet_step2_template(cumulative=cumulative), seed=1, gain=1.0))
    loader = loader or (lambda name: torch.load(ref_shims.checkpoint_path(mdir, name), map_location="cpu"))
    return loader(ck1), loader(ck2)


def make_ctsnet():
    """Fixtures from the UNMODIFIED CTSNet{,_new}/Step{1,2}_network.py modules run through the two-stage glue of
    two_stage_com_decode_vb.py (the oracle restatement supplies the DSP around them)."""
    for name, mdir, ck1, ck2, nsamp, clip_ids, p, cum in CTS_CASES:
        m1 = ref_shims.import_reference(mdir, "Step1_network").Step1_net().eval()
        m2 = ref_shims.import_reference(mdir, "Step2_network").Step2_net(X=6, R=3).eval()
        sd1, sd2 = cts_state_dicts(mdir, ck1, ck2, cum)
        m1.load_state_dict(sd1)
        m2.load_state_dict(sd2)
        rec = {"digest": np.array(sd_digest(sd1) + sd_digest(sd2)), "clip_ids": np.array(clip_ids),
               "nsamp": np.array(nsamp), "p": np.array(p)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_ctsnet((sd1, sd2), wav.astype(np.float64), p=p, cumulative=cum)
            with torch.no_grad():                                           # two_stage_com_decode_vb.py:79-84
                fx = torch.from_numpy(taps["feat"])[None]
                ph = torch.atan2(fx[:, 1], fx[:, 0])
                e1 = m1(torch.norm(fx, dim=1))
                s1 = torch.stack((e1 * torch.cos(ph), e1 * torch.sin(ph)), dim=1)
                est_ref = (m2(torch.cat((fx, s1), dim=1)) + s1).squeeze(0).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"feat{j}"] = taps["feat"]
            rec[f"est1{j}"] = e1.squeeze(0).numpy().astype(np.float32)
            rec[f"est{j}"] = est_ref.astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, |est| max {float(np.abs(rec['est0']).max()):.2f}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


TAYLOR_KW = dict(cin=2, k1=(1, 3), k2=(2, 3), c=64, kd1=5, cd1=64, d_feat=256, dilations=[1, 2, 5, 9], p=2, fft_num=320,
                 order_num=3, intra_connect='cat', inter_connect='cat', is_causal=True, is_conformer=False, is_u2=True,
                 is_param_share=False, is_encoder_share=False)     # TaylorSENet/taylorsenet_decode_vb.py:11-13
TAYLOR_CASES = [
    # name, model dir, checkpoint, samples, clip ids, p, cumulative
    ("taylor_synth", "TaylorSENet", None, 8000, (18, 19), 0.5, False),
    ("taylor_ckpt", "TaylorSENet", "vb_taylor_noncprs_model.pth", 16000, (18, 19), 1.0, False),
    ("taylor_new_ckpt", "TaylorSENet_new", "vb_taylor_cprs_model.pth", 16000, (20, 21), 0.5, True),
]


def make_taylor():
    """Fixtures from the UNMODIFIED TaylorSENet{,_new}/TaylorSENet.py module."""
    for name, mdir, ckpt, nsamp, clip_ids, p, cum in TAYLOR_CASES:
        net = ref_shims.import_reference(mdir, "TaylorSENet").TaylorSENet(**TAYLOR_KW).eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.taylorsenet_template(cum), seed=0, gain=1.0)
        else:
            sd = torch.load(ref_shims.checkpoint_path(mdir, ckpt), map_location="cpu")
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp), "p": np.array(p)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_taylorsenet(sd, wav.astype(np.float64), p=p, cumulative=cum)
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["feat"])[None]).squeeze(0).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"feat{j}"] = taps["feat"]
            rec[f"est{j}"] = est_ref.astype(np.float32)
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, |est| max {float(np.abs(rec['est0']).max()):.2f}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


G2_ARGS = (3, 64, 2, 4, 4, [1, 2, 5, 9], 256 + 161 * 2, 256, 256, (2, 3), (1, 3), 64, 'cat', 3)   # com_decode.py:23
G2_KW = dict(is_aux=False, encoder_type='U2Net', tcm_type='full-band')
G2_CASES = [
    # name, model dir, checkpoint, samples, clip ids, p, cumulative
    ("g2net_synth", "G2Net_new", None, 8000, (22, 23), 1.0, True),
    ("g2net_new_ckpt", "G2Net_new", "vb_gaf_cprs_model.pth", 16000, (22, 23), 0.5, True),       # G2Net_new/com_decode.py
    ("g2net_vb_ckpt", "G2Net_VB", "vb_gaf_noncprs_model.pth", 16000, (24, 25), 1.0, False),    # G2Net_VB/com_decode.py
]


def make_g2net():
    """Fixtures from the UNMODIFIED G2Net_{new,VB}/gaf_net_320.py ``gaf_base`` module."""
    for name, mdir, ckpt, nsamp, clip_ids, p, cum in G2_CASES:
        net = ref_shims.import_reference(mdir, "gaf_net_320").gaf_base(*G2_ARGS, **G2_KW).eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(templates.g2net_template(cum), seed=0, gain=1.0)
        else:
            sd = torch.load(ref_shims.checkpoint_path(mdir, ckpt), map_location="cpu")
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp), "p": np.array(p)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = decode.enhance_g2net(sd, wav.astype(np.float64), p=p, cumulative=cum)
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["feat"])[None])[-1].squeeze(0).permute(0, 2, 1).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"feat{j}"] = taps["feat"]
            rec[f"est{j}"] = est_ref.astype(np.float32)          # [2,T,F]
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, |est| max {float(np.abs(rec['est0']).max()):.2f}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, mdir, module, cls, ckpt, enh, tmpl, nsamp, clip_ids in CASES:
        mod = ref_shims.import_reference(mdir, module)
        net = getattr(mod, cls)().eval()
        if ckpt is None:
            sd = synth.synthetic_state_dict(tmpl(), seed=0)
        else:
            sd = torch.load(ref_shims.checkpoint_path(mdir, ckpt), map_location="cpu")
            assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(s) for k, s in tmpl().items()}
        net.load_state_dict(sd)
        rec = {"digest": np.array(sd_digest(sd)), "clip_ids": np.array(clip_ids), "nsamp": np.array(nsamp)}
        worst = 0.0
        for j, cid in enumerate(clip_ids):
            wav = synth.noisy_clip(cid, nsamp)
            y, taps = enh(sd, wav.astype(np.float64))
            with torch.no_grad():
                est_ref = net(torch.from_numpy(taps["mag"])[None]).squeeze(0).numpy()
            worst = max(worst, float(np.abs(est_ref - taps["est"]).max()))
            rec[f"wav{j}"] = wav
            rec[f"mag{j}"] = taps["mag"]
            rec[f"est{j}"] = est_ref.astype(np.float32)          # the REFERENCE module's output
            rec[f"ynorm{j}"] = taps["y_norm"]
            rec[f"y{j}"] = y
            rec[f"c{j}"] = np.array(taps["c"])
        rec["ref_vs_oracle"] = np.array(worst)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: ref_vs_oracle max-abs {worst:.3e}, out rms "
              f"{float(np.sqrt(np.mean(rec['ynorm0'] ** 2))):.4f}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 2 or sys.argv[1] == "mag":
        main()
    if len(sys.argv) < 2 or sys.argv[1] == "fullsubnet":
        make_fullsubnet()
    if len(sys.argv) < 2 or sys.argv[1] == "dccrn":
        make_dccrn()
    if len(sys.argv) < 2 or sys.argv[1] == "gcrn":
        make_gcrn()
    if len(sys.argv) < 2 or sys.argv[1] == "dpcrn":
        make_dpcrn()
    if len(sys.argv) < 2 or sys.argv[1] == "uformer":
        make_uformer()
    if len(sys.argv) < 2 or sys.argv[1] == "ctsnet":
        make_ctsnet()
    if len(sys.argv) < 2 or sys.argv[1] == "taylor":
        make_taylor()
    if len(sys.argv) < 2 or sys.argv[1] == "g2net":
        make_g2net()
