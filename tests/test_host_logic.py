"""CPU tests of the host side: C-ABI surface, state-dict drop-in, weight packing / tap tables /
layouts (through a torch mirror of the kernel semantics), sharding over gloo."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import emu_ops
import se_b200
from conftest import ROOT
from oracle import nets, synth, templates


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "se_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(se_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert "se_stft" in syms and "se_lstm_seq" in syms
    assert os.path.exists(se_b200._lib.LIB_PATH), "libse_b200.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(se_b200._lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/se_b200.h but not exported"
    assert set(se_b200._lib.PROTOTYPES) == set(syms), "ctypes prototypes out of sync with the header"
    assert se_b200._lib.load().se_abi_version() == 1


def test_no_cpu_fallback():
    m = se_b200.crn_net()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 4, 161))
    with pytest.raises(Exception):
        se_b200.ops.rms_scale(torch.zeros(1, 16))


@pytest.mark.parametrize("cls,tmpl", [(se_b200.crn_net, templates.crn_template),
                                      (se_b200.lstm_net, templates.lstm_template)])
def test_state_dict_namespace(cls, tmpl):
    m = cls()
    t = tmpl()
    sd = m.state_dict()
    assert list(sd.keys()) == list(t.keys())
    assert all(tuple(sd[k].shape) == tuple(t[k]) for k in t)
    m.load_state_dict(synth.synthetic_state_dict(t, seed=1))      # strict
    assert m.eval() is m


@pytest.mark.parametrize("seed", [0, 7])
def test_crn_host_logic_matches_oracle(monkeypatch, seed):
    emu_ops.install(se_b200.ops, monkeypatch)
    sd = synth.synthetic_state_dict(templates.crn_template(), seed=seed)
    m = se_b200.crn_net()
    m.load_state_dict(sd)
    x = torch.rand(3, 17, 161, generator=torch.Generator().manual_seed(seed)) * 4
    taps, rt = {}, {}
    y = m._forward_impl(x, taps)
    with torch.no_grad():
        yr = nets.crn_forward(sd, x, rt)
    for i in range(1, 6):
        assert (taps[f"en{i}"].permute(0, 3, 1, 2) - rt[f"en{i}"]).abs().max() < 1e-4
    for i in range(1, 5):
        assert (taps[f"de{i}"].permute(0, 3, 1, 2) - rt[f"de{i}"]).abs().max() < 1e-3
    assert (y - yr).abs().max() < 1e-3 * max(1.0, yr.abs().max().item())
    # re-loading other weights must invalidate the packed copies
    m.load_state_dict(synth.synthetic_state_dict(templates.crn_template(), seed=seed + 1))
    y2 = m._forward_impl(x)
    assert (y2 - y).abs().max() > 1e-3


def test_lstm_net_host_logic_matches_oracle(monkeypatch):
    emu_ops.install(se_b200.ops, monkeypatch)
    sd = synth.synthetic_state_dict(templates.lstm_template(), seed=2)
    m = se_b200.lstm_net()
    m.load_state_dict(sd)
    x = torch.rand(2, 9, 161, generator=torch.Generator().manual_seed(3)) * 4
    y = m._forward_impl(x)
    with torch.no_grad():
        yr = nets.lstm_net_forward(sd, x)
    assert (y - yr).abs().max() < 1e-4


def test_fullsubnet_host_logic_matches_oracle(monkeypatch):
    from se_b200.fullsubnet import Model
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.fullsubnet_template()
    sd = synth.synthetic_state_dict(t, seed=5)
    m = Model(num_freqs=257, look_ahead=2, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=15,
              fb_output_activate_function="ReLU", sb_output_activate_function=None, fb_model_hidden_size=512,
              sb_model_hidden_size=384)
    assert list(m.state_dict().keys()) == list(t.keys())
    m.load_state_dict(sd)
    x = torch.rand(2, 1, 257, 7, generator=torch.Generator().manual_seed(1)) * 3
    y = m._forward_impl(x)
    with torch.no_grad():
        yr = nets.fullsubnet_forward(sd, x)
    assert y.shape == yr.shape == (2, 2, 257, 7)
    assert (y - yr).abs().max() < 2e-4 * max(1.0, yr.abs().max().item())
    # >= 128 rows: the fp16-pair plan -- full-band H = 512 layers as the first block of a block-diagonal H = 1024 recurrence
    # (idle units: zero xproj columns and weights), fp16-pair projections / output layer / sub-band cells
    seen = []
    orig = se_b200.ops.lstm_seq
    monkeypatch.setattr(se_b200.ops, "lstm_seq", lambda xp, whh, hidden, out=None: (seen.append(hidden), orig(xp, whh, hidden, out))[1])
    x = torch.rand(2, 1, 257, 70, generator=torch.Generator().manual_seed(2)) * 3
    y = m._forward_impl(x)
    with torch.no_grad():
        yr = nets.fullsubnet_forward(sd, x)
    assert seen == [1024, 1024]
    assert (y - yr).abs().max() < 2e-4 * max(1.0, yr.abs().max().item())


def test_shard_range_partitions():
    from se_b200 import shard
    for b in (1, 7, 64, 256, 513):
        for w in (1, 2, 4, 8):
            spans = [shard.shard_range(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, batch, q):
    import torch.distributed as dist
    from se_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(batch * 5, dtype=torch.float32).view(batch, 5)
    s, e = shard.shard_range(batch, rank, world)
    out = shard.gather_waveforms(full[s:e].clone() * 1.0, batch)
    ok = bool(torch.equal(out, full))
    if batch % world == 0:
        # pipelined gather to rank 0: three batches in flight, results in submission order
        pipe = shard.GatherPipeline(batch, dst=0, depth=2)
        got = [pipe.submit(full[s:e] * float(k + 1)) for k in range(3)]
        pipe.drain()
        if rank == 0:
            ok = ok and all(torch.equal(torch.cat(g, 0), full * float(k + 1)) for k, g in enumerate(got))
        else:
            ok = ok and all(g is None for g in got)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_gather_over_gloo_world2(batch):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_dccrn_host_logic_matches_oracle(monkeypatch):
    """Complex->real stacked-channel packing, cLSTM block matrices, look-ahead decoder taps."""
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.dccrn_template()
    sd = synth.synthetic_state_dict(t, seed=4)
    m = se_b200.DCCRN(rnn_units=256, masking_mode='E', use_clstm=True, kernel_num=[32, 64, 128, 256, 256, 256])
    assert list(m.state_dict().keys()) == list(t.keys())
    m.load_state_dict(sd)
    x = torch.randn(2, 2, 257, 9, generator=torch.Generator().manual_seed(2))
    est = m._forward_nhwc(x.permute(0, 3, 2, 1).contiguous()).permute(0, 3, 2, 1)
    with torch.no_grad():
        ref = nets.dccrn_forward(sd, x)
    assert (est - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item())
    # DCCRN_SNR variant: crop the other side (DCCRN_SNR/DCCRN.py:159)
    m2 = se_b200.DCCRN(rnn_units=256, masking_mode='E', use_clstm=True, kernel_num=[32, 64, 128, 256, 256, 256],
                       crop_first=False)
    m2.load_state_dict(sd)
    est2 = m2._forward_nhwc(x.permute(0, 3, 2, 1).contiguous()).permute(0, 3, 2, 1)
    with torch.no_grad():
        ref2 = nets.dccrn_forward(sd, x, crop_first=False)
    assert (est2 - ref2).abs().max() < 2e-4 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("mode", ["C", "R"])
def test_dccrn_mask_modes_host_logic(monkeypatch, mode):
    """masking_mode 'C' / 'R' (DCCRN_cprs.py:221-224): same network, different mask rule."""
    emu_ops.install(se_b200.ops, monkeypatch)
    sd = synth.synthetic_state_dict(templates.dccrn_template(), seed=4)
    m = se_b200.DCCRN(rnn_units=256, masking_mode=mode, use_clstm=True, kernel_num=[32, 64, 128, 256, 256, 256])
    m.load_state_dict(sd)
    x = torch.randn(1, 2, 257, 6, generator=torch.Generator().manual_seed(8))
    est = m._forward_nhwc(x.permute(0, 3, 2, 1).contiguous()).permute(0, 3, 2, 1)
    with torch.no_grad():
        ref = nets.dccrn_forward(sd, x, masking_mode=mode)
    assert (est - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item())


def test_gcrn_host_logic_matches_oracle(monkeypatch):
    """Gated convs as [a | b] GEMMs, parity-class deconvs with output_padding, grouped-LSTM block weights and the
    three layout permutations (channels-last flatten, stack+flatten interleave, LayerNorm store index)."""
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.gcrn_template()
    sd = synth.synthetic_state_dict(t, seed=6)
    m = se_b200.gcrn.Net()
    assert list(m.state_dict().keys()) == list(t.keys())
    m.load_state_dict(sd)
    x = torch.randn(2, 2, 7, 161, generator=torch.Generator().manual_seed(3))
    taps, rtaps = {}, {}
    est = m._forward_impl(x, taps)
    with torch.no_grad():
        ref = nets.gcrn_forward(sd, x, rtaps)
    for k in ("e1", "e5"):
        assert (taps[k].permute(0, 3, 1, 2) - rtaps[k]).abs().max() < 1e-4 * max(1.0, rtaps[k].abs().max().item()), k
    assert (taps["glstm_nhwc"].permute(0, 3, 1, 2) - rtaps["glstm"]).abs().max() < 2e-4
    assert (est - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item())


def test_dpcrn_host_logic_matches_oracle(monkeypatch):
    """Bi-LSTM over F as strided cell steps (forward / reverse order, layer stacking), the inter-LSTM as a
    shared-weight multi-group launch, LayerNorm([4,128]) + residual, PReLU convs, de4 pad, the CRM multiply."""
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.dpcrn_template()
    sd = synth.synthetic_state_dict(t, seed=7, gain=1.0)
    m = se_b200.dpcrn()
    assert list(m.state_dict().keys()) == list(t.keys())
    m.load_state_dict(sd)
    x = torch.randn(2, 2, 6, 161, generator=torch.Generator().manual_seed(4))
    taps, rtaps = {}, {}
    est = m._forward_impl(x, taps)
    with torch.no_grad():
        ref = nets.dpcrn_forward(sd, x, rtaps)
    for k in ("en1", "en5", "dp1", "dp2", "de4", "de5"):
        a, r = taps[k].permute(0, 3, 1, 2), rtaps[k]
        assert (a - r).abs().max() < 2e-4 * max(1.0, r.abs().max().item()), k
    assert (est - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item())


def test_uformer_host_logic_matches_reference_fixture(monkeypatch):
    """668-entry state-dict drop-in + the whole Uformer orchestration (stacked complex convs, block QKV
    projection, signed head combination, gated dilated convs, fusion) against the fixture written by the
    UNMODIFIED reference module."""
    import os
    from conftest import GOLDEN
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.uformer_template()
    sd = synth.synthetic_state_dict(t, seed=0, gain=1.0)
    m = se_b200.Uformer()
    assert list(m.state_dict().keys()) == list(t.keys()) and len(t) == 668
    m.load_state_dict(sd)
    g = np.load(os.path.join(GOLDEN, "uformer_synth.npz"))
    wav = torch.from_numpy(g["wav0"])[None].double() * float(g["c0"])
    X = torch.stft(wav.float(), 512, 160, 400, torch.hann_window(400), return_complex=True)
    x = torch.view_as_real(X.permute(0, 2, 1).contiguous()).contiguous()
    est = m._network(x).permute(0, 3, 2, 1)[0]
    ref = torch.from_numpy(g["est0"])
    assert (est - ref).abs().max() < 5e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cum", [False, True])
def test_ctsnet_host_logic_matches_oracle(monkeypatch, cum):
    """CTSNet / CTSNet_new: gated convs as [a | b] GEMMs, k(2,5) / k(2,3) transposed-conv parity classes with the
    chomp folded into the taps, the two TCM branches as one 128-channel tensor + block-diagonal dilated conv,
    (c,f) <-> (f,c) flatten permutations, InstanceNorm vs cumulative LayerNorm statistics (per branch)."""
    emu_ops.install(se_b200.ops, monkeypatch)
    t1, t2 = templates.ctsnet_step1_template(cum), templates.ctsnet_step2_template(cumulative=cum)
    sd1 = synth.synthetic_state_dict(t1, seed=8, gain=1.0)
    sd2 = synth.synthetic_state_dict(t2, seed=9, gain=1.0)
    m1, m2 = se_b200.ctsnet.Step1_net(cumulative=cum), se_b200.ctsnet.Step2_net(X=6, R=3, cumulative=cum)
    assert list(m1.state_dict().keys()) == list(t1.keys()) and list(m2.state_dict().keys()) == list(t2.keys())
    m1.load_state_dict(sd1)
    m2.load_state_dict(sd2)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 37, 161, generator=g) * 2
    taps, rtaps = {}, {}
    est = m1._forward_impl(x, taps)
    with torch.no_grad():
        ref = nets.ctsnet_step1_forward(sd1, x, cum, rtaps)
    for k in ("e1", "e5"):
        assert (taps[k].permute(0, 3, 1, 2) - rtaps[k]).abs().max() < 1e-5 * max(1.0, rtaps[k].abs().max().item()), k
    assert (taps["tcm"].permute(0, 3, 2, 1).reshape(2, 256, 37) - rtaps["tcm"]).abs().max() < 5e-5
    assert (est - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max().item())
    z = torch.randn(2, 4, 37, 161, generator=g)
    with torch.no_grad():
        ref2 = nets.ctsnet_step2_forward(sd2, z, cumulative=cum)
    assert (m2._forward_impl(z) - ref2).abs().max() < 2e-5 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("cum", [False, True])
def test_taylorsenet_host_logic_matches_oracle(monkeypatch, cum):
    """TaylorSENet / TaylorSENet_new: U2-Net modules (gated in_conv, inner U-Net with 'cat' skips, residual), transposed
    convs as parity classes, the zeroth-order gain path, the high-order recursion on RI rows (split in_conv GEMMs,
    stacked real/imag residual GEMM, k * pre_term and 1/k! accumulation)."""
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.taylorsenet_template(cum)
    sd = synth.synthetic_state_dict(t, seed=8, gain=1.0)
    m = se_b200.TaylorSENet(cumulative=cum)
    assert list(m.state_dict().keys()) == list(t.keys())
    m.load_state_dict(sd)
    x = torch.randn(2, 2, 23, 161, generator=torch.Generator().manual_seed(5))
    taps, rtaps = {}, {}
    est = m._forward_impl(x, taps)
    with torch.no_grad():
        ref = nets.taylorsenet_forward(sd, x, cum, taps=rtaps)
    assert (taps["gain"] - rtaps["gain"]).abs().max() < 1e-5
    assert (taps["head"].permute(0, 3, 2, 1).reshape(2, 256, 23) - rtaps["head"]).abs().max() < 1e-5
    assert (est - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max().item())
    with pytest.raises(NotImplementedError):
        se_b200.TaylorSENet(order_num=2)


@pytest.mark.parametrize("cum", [True, False])
def test_g2net_host_logic_matches_oracle(monkeypatch, cum):
    """G2Net_new / G2Net_VB: U2-Net encoder with k(1,3) inner units, the split 578-wide in_conv GEMMs with stacked
    main | gate outputs, single-branch Glu TCMs, gain / residual heads written into RI rows, the stage update."""
    emu_ops.install(se_b200.ops, monkeypatch)
    t = templates.g2net_template(cum)
    sd = synth.synthetic_state_dict(t, seed=8, gain=1.0)
    m = se_b200.g2net.gaf_base(3, 64, 2, 4, 4, [1, 2, 5, 9], 256 + 161 * 2, 256, 256, (2, 3), (1, 3), 64, 'cat', 3,
                               is_aux=False, encoder_type='U2Net', tcm_type='full-band', cumulative=cum)
    assert list(m.state_dict().keys()) == list(t.keys())
    m.load_state_dict(sd)
    x = torch.randn(2, 2, 21, 161, generator=torch.Generator().manual_seed(5))
    taps, rtaps = {}, {}
    est = m._forward_impl(x, taps)
    with torch.no_grad():
        ref = nets.g2net_forward(sd, x, cum, taps=rtaps)
    assert (taps["feat"].permute(0, 3, 2, 1).reshape(2, 256, 21) - rtaps["feat"]).abs().max() < 1e-5
    assert len(est) == 3
    for a, r in zip(est, ref):
        assert a.shape == r.shape and (a - r).abs().max() < 2e-5 * max(1.0, r.abs().max().item())
    with pytest.raises(NotImplementedError):
        se_b200.g2net.gaf_base(is_aux=True)


def test_enhance_dir_batches_equal_lengths_and_writes_pcm16(tmp_path):
    """wav-dir -> wav-dir surface (CRN/crn_decode_vb.py:17-64): files are grouped by length, every file comes back under
    its own name as 16-bit PCM, and ``enhance(args)`` accepts both spellings of the output directory."""
    from types import SimpleNamespace
    from scipy.io import wavfile
    decode = se_b200.decode
    src, dst = tmp_path / "noisy", tmp_path / "out"
    src.mkdir()
    rng = np.random.default_rng(0)
    clips = {"a.wav": 800, "b.wav": 1200, "c.wav": 800, "d.wav": 800}
    ref = {}
    for name, n in clips.items():
        x = np.clip(rng.normal(0, 0.1, n), -0.9, 0.9)
        decode.write_wav(str(src / name), x, 16000)
        ref[name] = decode.read_wav(str(src / name), 16000)
        # libsndfile's asymmetric PCM_16 scaling (write x * 32767, read / 32768): up to |x| + 0.5 LSB
        assert np.abs(ref[name] - x).max() <= (0.5 + np.abs(x).max()) / 32768 + 1e-12
    seen = []

    def fake_enhance(model, wav, gain=1.0):
        seen.append(tuple(wav.shape))
        return wav * gain

    n = decode.enhance_dir(None, str(src), str(dst), fs=16000, batch=2, device="cpu", enhance_fn=fake_enhance, gain=0.5)
    assert n == 4 and sorted(seen) == [(1, 800), (1, 1200), (2, 800)]
    for name in clips:
        sr, y = wavfile.read(str(dst / name))
        assert sr == 16000 and y.dtype == np.int16
        assert np.abs(y / 32768.0 - 0.5 * ref[name]).max() <= 1.0 / 32768 + 1e-7
    args = SimpleNamespace(mix_file_path=str(src), esti_file_path=str(tmp_path / "out2"), fs=16000)
    assert decode.enhance(args, None, device="cpu", enhance_fn=fake_enhance) == 4
    with pytest.raises(ValueError):
        decode.read_wav(str(src / "a.wav"), 48000)
    assert decode.enhancer_for(se_b200.crn_net()) is decode.enhance_crn
    assert decode.enhancer_for((None, None)) is decode.enhance_ctsnet
    assert decode.enhancer_for(se_b200.gcrn.Net()) is decode.enhance_gcrn


def test_enhance_dir_resamples_mixed_rates(tmp_path, monkeypatch):
    """``resample_to=16000`` (the *_decode_vb.py front step, lstm_decode_vb.py:33-34): files of different rates are
    grouped by (rate, length), resampled before the decode loop, and written at ``fs``."""
    from scipy.io import wavfile
    from oracle import resample as R
    decode = se_b200.decode
    src, dst = tmp_path / "noisy", tmp_path / "out"
    src.mkdir()
    rng = np.random.default_rng(1)
    files = {"a.wav": (48000, 2400), "b.wav": (16000, 800), "c.wav": (48000, 2400), "d.wav": (44100, 2205)}
    for name, (sr, n) in files.items():
        decode.write_wav(str(src / name), np.clip(rng.normal(0, 0.1, n), -0.9, 0.9), sr)

    def fake_resample(x, sr_orig, sr_new, out=None):
        return torch.from_numpy(np.stack([R.librosa_resample(r.double().numpy(), sr_orig, sr_new) for r in x])).float()

    monkeypatch.setattr(se_b200.ops, "resample", fake_resample)
    seen = []

    def fake_enhance(model, wav):
        seen.append(tuple(wav.shape))
        return wav

    n = decode.enhance_dir(None, str(src), str(dst), fs=16000, batch=8, device="cpu", enhance_fn=fake_enhance,
                           resample_to=16000)
    assert n == 4 and sorted(seen) == [(1, 800), (1, 800), (2, 800)]
    for name, (sr, _) in files.items():
        out_sr, y = wavfile.read(str(dst / name))
        assert out_sr == 16000 and len(y) == 800
        x, _ = decode.read_wav_any(str(src / name))
        assert np.abs(y / 32768.0 - R.librosa_resample(x, sr, 16000)).max() <= 1.0 / 32768 + 1e-6


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one stdout line, valid JSON, the
    keys of the bench contract, nothing that needs a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "checkpoints", "_ref", "CRN__wsj0_si84_300h_crn_noncprs_model.pth")):
        pytest.skip("bench.py refuses to run without the shipped CRN checkpoint (python -m oracle.fetch_checkpoints)")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_enhance_dir_shards_batches_over_ranks(tmp_path):
    """rank / world: every rank builds the same batch list and takes batches rank, rank + world, ...; together the
    ranks write every file exactly once (utterances are independent: no collective)."""
    decode = se_b200.decode
    src = tmp_path / "noisy"
    src.mkdir()
    rng = np.random.default_rng(3)
    names = {f"u{i:02d}.wav": (800 if i % 3 else 1200) for i in range(11)}
    for name, n in names.items():
        decode.write_wav(str(src / name), np.clip(rng.normal(0, 0.1, n), -0.9, 0.9), 16000)
    written = []
    for r in range(3):
        dst = tmp_path / f"out{r}"
        n = decode.enhance_dir(None, str(src), str(dst), fs=16000, batch=2, device="cpu",
                               enhance_fn=lambda m, w: w, rank=r, world=3)
        got = sorted(os.listdir(dst))
        assert n == len(got)
        written += got
    assert sorted(written) == sorted(names)          # disjoint and complete
    with pytest.raises(ValueError):
        decode.enhance_dir(None, str(src), str(tmp_path / "bad"), device="cpu", enhance_fn=lambda m, w: w, rank=3, world=3)


def test_fp16_pair_product():
    """Numerics of the fp16-pair recurrence engine (csrc/lstm_f16.cu), emulated in numpy: x*S = hi + lo in fp16 with
    power-of-two scales, three-term product, the step tag forced into the LSB of h_lo.  The error must stay in the
    3xTF32 class (fp32-level), far below the 1e-4 waveform gate."""
    import math
    rng = np.random.default_rng(0)
    k = 1024
    w = (rng.standard_normal((256, k)) * 0.05).astype(np.float32)
    h = (np.tanh(rng.standard_normal((k, 64))) * rng.random((k, 64))).astype(np.float32)
    exact = w.astype(np.float64) @ h.astype(np.float64)

    def split16(x, scale, tag=None):
        xs = (x * np.float32(scale)).astype(np.float32)
        hi = xs.astype(np.float16)
        lo = (xs - hi.astype(np.float32)).astype(np.float16)
        if tag is not None:
            lo = ((lo.view(np.uint16) & 0xFFFE) | tag).astype(np.uint16).view(np.float16)
        return hi.astype(np.float64), lo.astype(np.float64)

    sw = 2.0 ** (13 - math.frexp(float(np.abs(w).max()))[1] + 1)     # ilogb = frexp exponent - 1
    assert 2 ** 13 <= np.abs(w).max() * sw < 2 ** 14
    sh = 1024.0
    for tag in (0, 1):
        whi, wlo = split16(w, sw)
        hhi, hlo = split16(h, sh, tag)
        d = (whi @ hhi + whi @ hlo + wlo @ hhi) / (sw * sh)
        assert np.abs(d - exact).max() < 1e-6


def test_plan_batches_buckets_by_length():
    """Length-bucketed batching of a directory: deterministic, every file once, bounded padding;
    exact-length grouping for decode loops without per-clip lengths."""
    decode = se_b200.decode
    rng = np.random.default_rng(1)
    infos = [(f"f{i:03d}.wav", 16000, int(n)) for i, n in enumerate(rng.integers(16000, 160000, 200))]
    infos += [("g48.wav", 48000, 50000), ("tiny.wav", 16000, 300)]
    plan = decode.plan_batches(infos, batch=16, ragged=True, pad_tolerance=0.25)
    names = [n for _, ns, _ in plan for n in ns]
    assert sorted(names) == sorted(i[0] for i in infos) and len(set(names)) == len(names)
    for sr, ns, lens in plan:
        assert len(ns) <= 16 and lens == sorted(lens)
        assert max(lens) <= 1.25 * min(lens) or len(ns) == 1 or min(lens) < 512
    assert len(plan) <= 40                      # 202 files: ~13 full batches + the splits the length bound forces
    assert decode.plan_batches(infos, 16, True) == plan
    exact = decode.plan_batches(infos[:20] + [("dup.wav", 16000, infos[0][2])], batch=16, ragged=False)
    assert all(len(set(lens)) == 1 for _, _, lens in exact) and max(len(ns) for _, ns, _ in exact) == 2


def test_enhance_dir_ragged_batches_and_bad_files(tmp_path):
    """enhance_dir with a decode loop that takes ``lengths``: files of different lengths share tail-padded batches, each
    comes back at its own length; a non-wav entry, a sub-directory and a truncated file are skipped and reported."""
    from scipy.io import wavfile
    decode = se_b200.decode
    src, dst = tmp_path / "noisy", tmp_path / "out"
    src.mkdir()
    (src / "sub").mkdir()
    (src / ".DS_Store").write_bytes(b"not audio")
    rng = np.random.default_rng(5)
    lens = {f"u{i}.wav": 1000 + 37 * i for i in range(9)}
    for name, n in lens.items():
        decode.write_wav(str(src / name), np.clip(rng.normal(0, 0.1, n), -0.9, 0.9), 16000)
    good = (src / "u0.wav").read_bytes()
    (src / "trunc.wav").write_bytes(good[:30])
    calls = []

    def fake(model, wav, lengths=None, gain=1.0):
        calls.append((tuple(wav.shape), None if lengths is None else lengths.tolist()))
        if lengths is not None:                       # the padding really is zero
            for i, n in enumerate(lengths.tolist()):
                assert float(wav[i, n:].abs().sum()) == 0.0
        return wav * gain

    rep = {}
    n = decode.enhance_dir(None, str(src), str(dst), fs=16000, batch=4, device="cpu", enhance_fn=fake, gain=0.5, report=rep)
    assert n == 9 and sorted(rep["written"]) == sorted(lens)
    assert set(rep["skipped"]) == {"sub", ".DS_Store", "trunc.wav"} and rep["ragged"] and rep["batches"] == 3
    assert [c[0][0] for c in calls] == [4, 4, 1] and all(c[1] is not None for c in calls[:2])
    for name, ln in lens.items():
        sr, y = wavfile.read(str(dst / name))
        x = decode.read_wav(str(src / name), 16000)
        assert sr == 16000 and len(y) == ln and np.abs(y / 32768.0 - 0.5 * x).max() <= 1.0 / 32768 + 1e-7


@pytest.mark.parametrize("m", [160, 256])
def test_fft_thread_algebra_on_cpu(m, tmp_path):
    """csrc/fft_thread.cuh (the per-thread radix-16 x radix-N2 FFT of the second-generation STFT / iSTFT kernels) is plain
    C++: tools/fft_thread_host_test.cpp drives it on the CPU exactly as the kernels do (two passes, exchange array,
    real-FFT split / merge) and the result is compared with numpy.fft."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fft_host_test")
    subprocess.run([cxx, "-O1", "-std=c++17", os.path.join(root, "tools", "fft_thread_host_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, str(m)], capture_output=True, text=True, check=True).stdout.splitlines()
    x = np.array(out[0].split(), dtype=np.float64)
    spec = np.array(out[1].split(), dtype=np.float64).reshape(-1, 2)
    back = np.array(out[2].split(), dtype=np.float64)
    ref = np.fft.rfft(x)
    assert len(x) == 2 * m and spec.shape[0] == m + 1
    assert np.abs(spec[:, 0] + 1j * spec[:, 1] - ref).max() < 2e-5
    assert abs(spec[0, 1]) < 1e-5 and abs(spec[m, 1]) < 1e-5
    assert np.abs(back - x).max() < 1e-6


def test_split_f16_pair_precision_range_and_saturation():
    """packing.split_f16 (the host twin of split_f16_dev / split_f16_kernel): x * 2^s = hi + lo to 22 significand bits
    wherever |x * 2^s| >= 2^-3, an absolute floor of 2^-25 / 2^s below, saturation (not inf) beyond the fp16 range, and
    the per-tensor scale that puts max|w| in [2^13, 2^14)."""
    from se_b200 import packing
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(4096, generator=g) * 3, torch.tensor([900.0, -1e-5, 3e-7, 0.0, 4093.0])])
    hi, lo, s = packing.split_f16(x, 4)
    assert s == 4 and hi.dtype == torch.float16 and lo.dtype == torch.float16
    back = (hi.double() + lo.double()) / 16.0
    err = (back - x.double()).abs()
    assert bool((err <= 2.0 ** -21 * x.double().abs() + 2.0 ** -29).all())
    hs, ls, _ = packing.split_f16(torch.tensor([1e6, -1e6]), 4)              # beyond +-4094: saturates, stays finite
    assert torch.isfinite(hs.float()).all() and torch.isfinite(ls.float()).all() and float(hs[0]) == 65504.0
    w = torch.randn(64, 96, generator=g) * 0.07
    wh, wl, ws = packing.split_f16(w)
    m = float(w.abs().max()) * 2.0 ** ws
    assert 2.0 ** 13 <= m < 2.0 ** 14
    assert ((wh.double() + wl.double()) * 2.0 ** -ws - w.double()).abs().max() < 2.0 ** -21 * float(w.abs().max())
    z = packing.split_f16(torch.zeros(8, 8))
    assert z[2] == 0 and float(z[0].abs().max()) == 0.0


def test_pack_conv_f16_pads_every_tap_and_source_block_to_64_channels():
    """The se_conv_f16x3 weight layout: K order (tap, [source 0 | source 1]); every block zero-padded to a multiple of 64
    channels; ConvWeights.f16 with activation channels padded to 8 (2-channel inputs) puts zero rows on the padding."""
    from se_b200 import packing
    from se_b200.conv_engine import ConvWeights, merge_parity
    g = torch.Generator().manual_seed(1)
    ntaps, c0, c1, co = 3, 96, 40, 16
    w = torch.randn(co, ntaps * (c0 + c1), generator=g)
    hi, lo, s = packing.pack_conv_f16(w, ntaps, c0, c1)
    p0, p1 = 128, 64
    assert hi.shape == (co, ntaps * (p0 + p1))
    full = ((hi.double() + lo.double()) * 2.0 ** -s).view(co, ntaps, p0 + p1)
    src = w.double().view(co, ntaps, c0 + c1)
    assert (full[:, :, :c0] - src[:, :, :c0]).abs().max() < 1e-6 and (full[:, :, p0:p0 + c1] - src[:, :, c0:]).abs().max() < 1e-6
    assert float(full[:, :, c0:p0].abs().max()) == 0.0 and float(full[:, :, p0 + c1:].abs().max()) == 0.0
    # 2 input channels carried in an 8-channel activation
    kn = torch.randn(10 * 2, 128, generator=g)
    cw = ConvWeights(kn, 128)
    h2, l2, s2 = cw.f16(10, 2, 0, 8, 0)
    f2 = ((h2.double() + l2.double()) * 2.0 ** -s2).view(128, 10, 64)
    assert (f2[:, :, :2] - kn.double().t().reshape(128, 10, 2)).abs().max() < 1e-6 and float(f2[:, :, 2:].abs().max()) == 0.0
    # both parity classes as one matrix: odd taps are a subset of the even taps, zero rows elsewhere
    ev, od = [(0, 0), (0, -1), (-1, 0), (-1, -1)], [(0, 0), (-1, 0)]
    we, wo = ConvWeights(torch.randn(4 * 8, 4, generator=g), 4), ConvWeights(torch.randn(2 * 8, 4, generator=g), 4)
    m = merge_parity(we, ev, wo, od)
    assert m.cout == 8 and torch.equal(m.kn[:, :4], we.kn[:, :4])
    assert torch.equal(m.kn[0:8, 4:8], wo.kn[0:8, :4]) and torch.equal(m.kn[16:24, 4:8], wo.kn[8:16, :4])
    assert float(m.kn[8:16, 4:8].abs().max()) == 0.0 and float(m.kn[24:32, 4:8].abs().max()) == 0.0
