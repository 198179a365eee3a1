"""GPU parity of the network building blocks (implicit-GEMM conv, LSTM recurrence) through the
C ABI against the torch mirror of their declared semantics (tests/emu_ops.py, CPU fp32/fp64)."""
import os

import numpy as np
import pytest
import torch

import emu_ops

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


DEFAULT_GEMM_ENGINE = int(os.environ.get("SE_GEMM_ENGINE", "5"))   # csrc/gemm_tc.cu: kDefaultGemmEngine

CONV_CASES = [
    # B, T, Fin, C0, C1, Cout, kind
    (2, 13, 80, 16, 0, 32, "conv"),
    (2, 13, 39, 32, 0, 64, "conv"),
    (1, 9, 19, 64, 0, 128, "conv"),
    (3, 7, 9, 128, 0, 256, "conv"),
    (2, 11, 4, 256, 256, 128, "deconv"),
    (2, 11, 9, 128, 128, 64, "deconv"),
    (1, 5, 19, 64, 64, 32, "deconv"),
    (2, 6, 39, 32, 32, 16, "deconv_shift"),
    # narrow outputs (Cout <= 4 and Fout >= 16: conv_rows_kernel, one CTA per output frame, input frames in smem)
    (2, 9, 39, 64, 64, 2, "deconv"),           # CTSNet / DPCRN style RI head with a skip source
    (1, 7, 80, 128, 0, 2, "conv"),
    (3, 5, 19, 32, 32, 1, "deconv_shift"),     # fill column + one channel
    (2, 4, 19, 256, 128, 4, "deconv"),         # uneven sources, three channel passes per lane
    (1, 3, 40, 36, 0, 3, "conv"),              # channel count that is not a multiple of 128
    (2, 4, 9, 64, 0, 2, "conv"),               # Fout < 16: stays on the tiled kernel
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_gemm_matches_semantics(case):
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    b, t, fin, c0, c1, co, kind = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x0 = torch.randn(b, t, fin, c0, generator=g)
    x1 = torch.randn(b, t, fin, c1, generator=g) if c1 else None
    ct = c0 + c1
    bias = torch.randn(co, generator=g)
    fill = torch.randn(co, generator=g)
    if kind == "conv":
        fout = (fin - 3) // 2 + 1
        runs = [(packing.CONV23_TAPS, 2, fout, 0, 1, -1)]
        dstF = fout
    else:
        shift = 1 if kind == "deconv_shift" else 0
        dstF = 2 * fin + 1 + shift
        runs = [(packing.DECONV_EVEN_TAPS, 1, fin + 1, shift, 2, -1),
                (packing.DECONV_ODD_TAPS, 1, fin, shift + 1, 2, 0 if shift else -1)]
    ref = torch.zeros(b, t, dstF, co, dtype=torch.float64)
    got = torch.zeros(b, t, dstF, co, device=dev)
    for taps, sf, fout, f0, fstep, fill_f in runs:
        w = torch.randn(len(taps) * ct, co, generator=g) / np.sqrt(len(taps) * ct)
        emu_ops.conv_gemm(x0.double(), None if x1 is None else x1.double(), b, t, fin, fout, taps, sf, w.double(),
                          bias.double(), co, "elu", ref, dstF, f0, fstep, fill_f, fill.double())
        ops.conv_gemm(x0.to(dev), None if x1 is None else x1.to(dev), b, t, fin, fout, taps, sf,
                      packing.pad_cols(w).to(dev), bias.to(dev), co, "elu", got, dstF, f0, fstep, fill_f,
                      fill.to(dev) if fill_f >= 0 else None)
    err = (got.cpu().double() - ref).abs().max().item()
    print(f"conv_gemm {case}: max err {err:.3e}")
    assert err < 2e-5


@pytest.mark.parametrize("m,k,n,act", [(37, 161, 4096, "none"), (401, 1024, 161, "softplus"), (130, 1024, 4096, "none"),
                                       (5, 48, 20, "relu")])
def test_linear_ragged_shapes(m, k, n, act):
    dev = _dev()
    import se_b200
    from se_b200 import packing
    g = torch.Generator().manual_seed(m + k + n)
    x = torch.randn(m, k, generator=g)
    w = torch.randn(k, n, generator=g) / np.sqrt(k)
    bias = torch.randn(n, generator=g)
    ref = emu_ops.linear(x.double(), w.double(), bias.double(), n, act)
    got = se_b200.ops.linear(x.to(dev), packing.pad_cols(w).to(dev), bias.to(dev), n, act)
    err = (got.cpu().double() - ref).abs().max().item()
    print(f"linear {m}x{k}x{n}: max err {err:.3e}")
    assert err < 2e-5


def test_conv_in1_and_deconv_out1():
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(4)
    b, t = 3, 10
    x = torch.rand(b, t, 161, generator=g) * 3
    w = torch.randn(6, 16, generator=g) * 0.4
    bias = torch.randn(16, generator=g) * 0.1
    ref = emu_ops.conv_in1(x.double(), w.double(), bias.double(), 16, "elu", 80)
    got = ops.conv_in1(x.to(dev), w.to(dev), bias.to(dev), 16, "elu", 80)
    e1 = (got.cpu().double() - ref).abs().max().item()
    s0 = torch.randn(b, t, 80, 16, generator=g)
    s1 = torch.randn(b, t, 80, 16, generator=g)
    w2 = torch.randn(6, 32, generator=g) * 0.2
    ref2 = emu_ops.deconv_out1(s0.double(), s1.double(), w2.double(), 0.3, "softplus")
    got2 = ops.deconv_out1(s0.to(dev), s1.to(dev), w2.to(dev), 0.3, "softplus")
    e2 = (got2.cpu().double() - ref2).abs().max().item()
    s2 = torch.randn(2, 7, 39, 64, generator=g)                 # one source, 64 channels, T not a multiple of the row tile
    w3 = torch.randn(6, 64, generator=g) * 0.2
    ref3 = emu_ops.deconv_out1(s2.double(), None, w3.double(), -0.1, "none")
    got3 = ops.deconv_out1(s2.to(dev), None, w3.to(dev), -0.1, "none")
    e3 = (got3.cpu().double() - ref3).abs().max().item()
    print(f"conv_in1 {e1:.3e} deconv_out1 {e2:.3e} {e3:.3e}")
    assert e1 < 1e-5 and e2 < 1e-5 and e3 < 1e-5


@pytest.mark.parametrize("b,t,h", [(1, 6, 1024), (5, 9, 1024), (64, 12, 1024), (70, 5, 1024), (3, 7, 256)])
def test_lstm_seq_matches_recurrence(b, t, h):
    dev = _dev()
    import se_b200
    g = torch.Generator().manual_seed(b * 100 + t)
    xp = torch.randn(b, t, 4 * h, generator=g)
    whh = torch.randn(h // 8, h, 32, generator=g) / np.sqrt(h)
    ref = emu_ops.lstm_seq(xp.double(), whh.double(), h)
    got = se_b200.ops.lstm_seq(xp.to(dev), whh.to(dev), h)
    torch.cuda.synchronize()
    err = (got.cpu().double() - ref).abs().max().item()
    print(f"lstm_seq B={b} T={t} H={h}: max err {err:.3e}")
    assert err < 2e-5


@pytest.mark.parametrize("m,k,n,act", [(128, 32, 128, "none"), (300, 1024, 256, "none"), (1000, 256, 4096, "softplus"),
                                       (77, 64, 161 + 3, "none")])
def test_gemm_tf32x3_fp32_accuracy(m, k, n, act):
    """tcgen05 3xTF32 GEMM: error must be fp32-class (<< single-pass TF32's ~5e-4 relative)."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g) / np.sqrt(k)
    bias = torch.randn(n, generator=g)
    ref = emu_ops._act(a.double() @ w.double().t() + bias.double(), act)
    a_hi, a_lo = ops.split_tf32(a.to(dev))
    w_hi, w_lo = ops.split_tf32(w.to(dev))
    assert ((a_hi + a_lo).cpu() - a).abs().max().item() < 2e-6
    got = ops.gemm_tf32x3(a_hi, a_lo, w_hi, w_lo, bias.to(dev), n, act)
    torch.cuda.synchronize()
    err = (got.cpu().double() - ref).abs().max().item()
    one_pass = ((a_hi.cpu().double() @ w_hi.cpu().double().t() + bias.double()) -
                (a.double() @ w.double().t() + bias.double())).abs().max().item()
    print(f"gemm_tf32x3 {m}x{k}x{n}: max err {err:.3e} (single-pass TF32 would be {one_pass:.3e})")
    assert err < 1e-5


@pytest.mark.parametrize("m,kx,h,steps", [(300, 32, 384, 3), (130, 64, 128, 2)])
def test_lstm_cell_tf32x3_matches_recurrence(m, kx, h, steps):
    """Fused per-step LSTM cell GEMM (FullSubNet sub-band regime) vs the textbook recurrence."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m + h)
    w_ih = torch.randn(4 * h, kx, generator=g) / np.sqrt(kx)
    w_hh = torch.randn(4 * h, h, generator=g) / np.sqrt(h)
    b_ih = torch.randn(4 * h, generator=g) * 0.1
    b_hh = torch.randn(4 * h, generator=g) * 0.1
    xs = torch.randn(steps, m, kx, generator=g)
    P = packing.pack_lstm_cell(w_ih, w_hh, b_ih, b_hh)
    P = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in P.items()}
    c = torch.zeros(m, h, device=dev)
    hbuf = [(torch.zeros(m, h, device=dev), torch.zeros(m, h, device=dev)) for _ in range(2)]
    hout = torch.empty(m, h, device=dev)
    # reference (float64)
    hr = torch.zeros(m, h, dtype=torch.float64)
    cr = torch.zeros(m, h, dtype=torch.float64)
    for t in range(steps):
        x_hi, x_lo = ops.split_tf32(xs[t].to(dev))
        src, dst = hbuf[t & 1], hbuf[(t + 1) & 1]
        ops.lstm_cell_tf32x3(x_hi, x_lo, src[0], src[1], P["w_hi"], P["w_lo"], P["bias"], c, dst[0], dst[1], hout)
        gates = xs[t].double() @ w_ih.double().t() + hr @ w_hh.double().t() + (b_ih + b_hh).double()
        i, f, gg, o = gates.chunk(4, dim=1)
        cr = torch.sigmoid(f) * cr + torch.sigmoid(i) * torch.tanh(gg)
        hr = torch.sigmoid(o) * torch.tanh(cr)
    torch.cuda.synchronize()
    e_h = (hout.cpu().double() - hr).abs().max().item()
    e_c = (c.cpu().double() - cr).abs().max().item()
    e_split = ((dst[0] + dst[1]).cpu().double() - hr).abs().max().item()
    print(f"lstm_cell M={m} Kx={kx} H={h} steps={steps}: h err {e_h:.3e} c err {e_c:.3e} hi+lo err {e_split:.3e}")
    assert e_h < 1e-5 and e_c < 1e-5 and e_split < 1e-5


@pytest.mark.parametrize("m,k,n,act", [(256, 32, 256, "none"), (300, 1024, 256, "none"), (1000, 256, 4096, "softplus"),
                                       (2309, 2048, 512 + 164, "none"), (25664, 1024, 4096, "none")])
@pytest.mark.parametrize("engine", [1, 2, 3, 4])
def test_gemm_pair_engine_matches_fp64_and_single_cta(m, k, n, act, engine):
    """CTA-pair (engine 1: cta_group::2, 256x256 tiles) and multicast-cluster (engines 2 / 3 / 4: 2x2, 1x2, 2x1 CTAs
    sharing TMA-multicast operand tiles) engines: fp32-class error vs fp64 on a row sample (first / last rows and a
    random draw), and agreement with the one-CTA engine on the WHOLE output (all add the same products in the same
    k order, promoted to fp32 registers every 4 k-blocks)."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g) / np.sqrt(k)
    bias = torch.randn(n, generator=g)
    a_hi, a_lo = ops.split_tf32(a.to(dev))
    w_hi, w_lo = ops.split_tf32(w.to(dev))
    try:
        ops.set_gemm_engine(0)
        one = ops.gemm_tf32x3(a_hi, a_lo, w_hi, w_lo, bias.to(dev), n, act)
        ops.set_gemm_engine(engine)
        two = ops.gemm_tf32x3(a_hi, a_lo, w_hi, w_lo, bias.to(dev), n, act)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)
    rows = torch.unique(torch.cat([torch.arange(0, min(m, 300)), torch.arange(max(0, m - 300), m),
                                   torch.randint(0, m, (256,), generator=g)]))
    ref = emu_ops._act(a[rows].double() @ w.double().t() + bias.double(), act)
    err = (two.cpu()[rows].double() - ref).abs().max().item()
    diff = (two - one).abs().max().item()
    print(f"gemm engine {engine} {m}x{k}x{n}: max err vs fp64 {err:.3e}, vs one-CTA engine {diff:.3e}")
    assert err < 1e-5 and diff < 1e-5


@pytest.mark.parametrize("m,kx,h,steps", [(300, 32, 384, 3), (8224, 32, 384, 2), (1031, 64, 128, 2)])
def test_lstm_cell_pair_engine_matches_single_cta(m, kx, h, steps):
    """Fused LSTM cell epilogue on the CTA-pair engine vs the one-CTA engine (same packing, 64-column gate groups)."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m + h)
    P = packing.pack_lstm_cell(torch.randn(4 * h, kx, generator=g) / np.sqrt(kx), torch.randn(4 * h, h, generator=g) / np.sqrt(h),
                               torch.randn(4 * h, generator=g) * 0.1, torch.randn(4 * h, generator=g) * 0.1)
    P = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in P.items()}
    xs = torch.randn(steps, m, kx, generator=g).to(dev)
    res = []
    try:
        for eng in (0, 1, 2, 3, 4):
            ops.set_gemm_engine(eng)
            c = torch.zeros(m, h, device=dev)
            hbuf = [(torch.zeros(m, h, device=dev), torch.zeros(m, h, device=dev)) for _ in range(2)]
            hout = torch.empty(m, h, device=dev)
            for t in range(steps):
                x_hi, x_lo = ops.split_tf32(xs[t])
                src, dst = hbuf[t & 1], hbuf[(t + 1) & 1]
                ops.lstm_cell_tf32x3(x_hi, x_lo, src[0], src[1], P["w_hi"], P["w_lo"], P["bias"], c, dst[0], dst[1], hout)
            torch.cuda.synchronize()
            res.append((hout.clone(), c.clone(), (dst[0] + dst[1]).clone()))
    finally:
        ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)
    for eng in (1, 2, 3, 4):
        for a, b, nm in zip(res[0], res[eng], ("h", "c", "hi+lo")):
            d = (a - b).abs().max().item()
            print(f"lstm_cell engine {eng} M={m} H={h}: {nm} diff {d:.3e}")
            assert d < 1e-5


TC_CONV_CASES = [
    # B, T, Fin, C0, C1, Cout, kind
    (2, 37, 19, 64, 0, 128, "crn_conv"),       # Fout 9  -> Tbox 14
    (2, 70, 9, 128, 0, 256, "crn_conv"),       # Fout 4  -> Tbox 32, two n tiles
    (1, 21, 39, 32, 0, 64, "crn_conv"),        # Fout 19 -> Tbox 6
    (2, 33, 4, 256, 256, 128, "crn_deconv"),
    (1, 17, 19, 64, 64, 32, "crn_deconv"),
    (2, 9, 39, 32, 32, 16, "crn_deconv"),
    (1, 11, 256, 32, 0, 64, "dccrn_conv"),     # Fout 128 -> Tbox 1, stride-2 box of 255
    (2, 19, 8, 256, 0, 256, "dccrn_conv"),
    (2, 13, 4, 256, 256, 256, "dccrn_deconv"),
    (1, 7, 64, 64, 64, 32, "dccrn_deconv"),
]


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("case", TC_CONV_CASES + [(3, 5, 9, 128, 0, 256, "crn_conv"),     # odd tile count, one 256 tile
                                                  (1, 1, 8, 256, 0, 192, "dccrn_conv")])  # a single activation tile
def test_conv_tf32x3_matches_semantics(case, engine):
    """Tensor-core implicit-GEMM conv (4-D TMA A tiles) vs the declared conv semantics in fp64; engine 1 = CTA pairs
    (cta_group::2: two activation tiles per MMA, 256-wide tiles for Cout > 128)."""
    import se_b200
    try:
        se_b200.ops.set_gemm_engine(engine)
        _conv_tf32x3_case(case)
    finally:
        se_b200.ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)


def _conv_tf32x3_case(case):
    dev = _dev()
    import se_b200
    from se_b200 import packing
    from se_b200.dccrn import DEC_EVEN, DEC_ODD, ENC_TAPS
    ops = se_b200.ops
    b, t, fin, c0, c1, co, kind = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 1000)
    x0 = torch.randn(b, t, fin, c0, generator=g)
    x1 = torch.randn(b, t, fin, c1, generator=g) if c1 else None
    ct = c0 + c1
    bias = torch.randn(co, generator=g)
    if kind == "crn_conv":
        fout = (fin - 3) // 2 + 1
        runs = [(packing.CONV23_TAPS, 2, fout, 0, 1)]
        dstF = fout
    elif kind == "crn_deconv":
        dstF = 2 * fin + 1
        runs = [(packing.DECONV_EVEN_TAPS, 1, fin + 1, 0, 2), (packing.DECONV_ODD_TAPS, 1, fin, 1, 2)]
    elif kind == "dccrn_conv":
        runs = [(ENC_TAPS, 2, fin // 2, 0, 1)]
        dstF = fin // 2
    else:
        dstF = 2 * fin
        runs = [(DEC_EVEN, 1, fin, 0, 2), (DEC_ODD, 1, fin, 1, 2)]
    ref = torch.zeros(b, t, dstF, co, dtype=torch.float64)
    got = torch.zeros(b, t, dstF, co, device=dev)
    got_hi, got_lo = torch.zeros_like(got), torch.zeros_like(got)
    s0 = ops.split_tf32(x0.to(dev))
    s1 = ops.split_tf32(x1.to(dev)) if x1 is not None else None
    for taps, sf, fout, f0, fstep in runs:
        w = torch.randn(len(taps) * ct, co, generator=g) / np.sqrt(len(taps) * ct)
        emu_ops.conv_gemm(x0.double(), None if x1 is None else x1.double(), b, t, fin, fout, taps, sf, w.double(),
                          bias.double(), co, "prelu", ref, dstF, f0, fstep, -1, None, 0.2)
        w_hi, w_lo = packing.split_tf32(w.t().contiguous())
        ops.conv_tf32x3(s0, s1, b, t, fin, fout, taps, sf, w_hi.to(dev), w_lo.to(dev), bias.to(dev), co, "prelu", dstF,
                        f0, fstep, act_param=0.2, out=got, out_pair=(got_hi, got_lo))
    torch.cuda.synchronize()
    err = (got.cpu().double() - ref).abs().max().item()
    err_pair = ((got_hi + got_lo).cpu().double() - ref).abs().max().item()
    print(f"conv_tf32x3 {case}: max err {err:.3e} (hi+lo {err_pair:.3e})")
    assert err < 2e-5 and err_pair < 2e-5


def test_lstm_seq_multi_equals_separate_launches():
    dev = _dev()
    import se_b200
    g = torch.Generator().manual_seed(11)
    b, t, h, ng = 5, 9, 128, 4
    xp = torch.randn(b, t, ng * 4 * h, generator=g).to(dev)
    whh = (torch.randn(ng, h // 8, h, 32, generator=g) / np.sqrt(h)).to(dev)
    out = torch.empty(b, t, ng * h, device=dev)
    se_b200.ops.lstm_seq_multi(xp, whh, h, ng, out)
    for k in range(ng):
        ref = emu_ops.lstm_seq(xp[:, :, k * 4 * h:(k + 1) * 4 * h].cpu().double(), whh[k].cpu().double(), h)
        err = (out[:, :, k * h:(k + 1) * h].cpu().double() - ref).abs().max().item()
        assert err < 2e-5, (k, err)


def test_uformer_glue_kernels():
    """se_uf_prep / se_uf_fusion / se_group_layernorm / se_uf_mask / extended GEMM epilogue vs torch mirrors."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(3)
    b, t, f = 2, 7, 257
    x = torch.randn(b, t, f, 2, generator=g)
    x[0, 0, 5] = 0.0
    got = ops.uf_prep(x.to(dev))
    ref = emu_ops.uf_prep(x.double())
    for a, r in zip(got, ref):
        assert (a.cpu().double() - r).abs().max() < 1e-5
    c = torch.randn(b, t, 4, 64, generator=g)
    m = torch.randn(b, t, 4, 32, generator=g)
    gc, gm = ops.uf_fusion(c.to(dev), m.to(dev))
    rc, rm = emu_ops.uf_fusion(c.double(), m.double())
    assert (gc.cpu().double() - rc).abs().max() < 1e-6 and (gm.cpu().double() - rm).abs().max() < 1e-6
    for groups, cc, post in ((2, 128, "prelu"), (1, 128, "none"), (2, 16, "none"), (2, 32, "swish"), (1, 32, "swish")):
        v = torch.randn(50, groups * cc, generator=g) * 3
        gate = torch.randn(50, groups * cc, generator=g)
        res = torch.randn(50, groups * cc, generator=g)
        gamma, beta = torch.rand(cc, generator=g) + 0.5, torch.randn(cc, generator=g)
        o, pair = ops.group_layernorm(v.to(dev), groups, gamma.to(dev), beta.to(dev), gate=gate.to(dev), post=post,
                                      slope=0.2, res=res.to(dev), want_f32=True, want_pair=True)
        r, _ = emu_ops.group_layernorm(v.double(), groups, gamma.double(), beta.double(), gate=gate.double(), post=post,
                                       slope=0.2, res=res.double())
        assert (o.cpu().double() - r).abs().max() < 2e-5
        assert ((pair[0] + pair[1]).cpu().double() - r).abs().max() < 2e-5
    cm = torch.randn(b, t, f - 1, 2, generator=g)
    md = torch.randn(b, t, f - 1, 1, generator=g)
    mag, ph = torch.rand(b, t, f, generator=g) * 3, (torch.rand(b, t, f, generator=g) - 0.5) * 6
    ge = ops.uf_mask(cm.to(dev), md.to(dev), mag.to(dev), ph.to(dev))
    re_ = emu_ops.uf_mask(cm.double(), md.double(), mag.double(), ph.double())
    assert (ge.cpu().double() - re_).abs().max() < 2e-5
    # extended GEMM epilogue: alpha, residual, PReLU, split output
    a = torch.randn(300, 64, generator=g)
    w = torch.randn(160, 64, generator=g) / 8
    bias = torch.randn(160, generator=g)
    res = torch.randn(300, 160, generator=g)
    ap = ops.split_tf32(a.to(dev))
    wp = ops.split_tf32(w.to(dev))
    o, pair = ops.gemm_tf32x3_ex(ap, wp[0], wp[1], bias.to(dev), 160, act="prelu", act_param=0.3, alpha=0.5,
                                 res=res.to(dev), want_f32=True, want_pair=True)
    y = a.double() @ w.double().t() + bias.double()
    r = torch.where(y >= 0, y, 0.3 * y) * 0.5 + res.double()
    assert (o.cpu().double() - r).abs().max() < 1e-5 and ((pair[0] + pair[1]).cpu().double() - r).abs().max() < 1e-5


@pytest.mark.parametrize("over_t", [True, False])
@pytest.mark.parametrize("cplx", [True, False])
def test_attention_matches_softmax(over_t, cplx):
    dev = _dev()
    import se_b200
    from se_b200.uformer import CPLX_HEADS
    g = torch.Generator().manual_seed(5)
    b, t, f = 2, 150, 4
    nheads = 8 if cplx else 1
    qkv = torch.randn(b * t * f, nheads * 48, generator=g)
    ho = [o for _, o, _ in CPLX_HEADS] if cplx else [0]
    hs = [s for _, _, s in CPLX_HEADS] if cplx else [1.0]
    nout = 2 if cplx else 1
    args = (t, f, b, t * f, f, 1) if over_t else (f, 1, b * t, f, 1, 0)
    got = se_b200.ops.attention(qkv.to(dev), nheads, ho, hs, nout, *args)
    ref = emu_ops.attention(qkv.double(), nheads, ho, hs, nout, *args)
    err = (got.cpu().double() - ref).abs().max().item()
    print(f"attention over_{'t' if over_t else 'f'} cplx={cplx}: max err {err:.3e}")
    assert err < 2e-5


DEFAULT_LSTM_ENGINE = 4


@pytest.mark.parametrize("h,b,t", [(1024, 64, 20), (1024, 5, 3), (1024, 64, 401), (512, 7, 9), (128, 33, 15), (128, 64, 401),
                                   (128, 1, 2)])
def test_lstm_engines_agree(h, b, t):
    """fp32 FMA / mma.sync 3xTF32 / tcgen05 3xTF32 (H = 1024 only, else the FMA kernel) / 3 (tcgen05 at H = 1024, the
    sequence-parallel kernel at H = 128) / default (4: the fp16-pair tagged-state tcgen05 kernel at H = 1024) recurrences
    vs fp64."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(h + b)
    xp = torch.randn(b, t, 4 * h, generator=g)
    whh = torch.randn(h // 8, h, 32, generator=g) / np.sqrt(h)
    ref = emu_ops.lstm_seq(xp.double(), whh.double(), h)
    errs = []
    try:
        for eng in (0, 1, 2, 3, 4, 5):
            ops.set_lstm_engine(eng)
            got = ops.lstm_seq(xp.to(dev), whh.to(dev), h)
            torch.cuda.synchronize()
            errs.append((got.cpu().double() - ref).abs().max().item())
    finally:
        ops.set_lstm_engine(DEFAULT_LSTM_ENGINE)
    print(f"lstm engines H={h} B={b} T={t}: fma err {errs[0]:.3e}, mma err {errs[1]:.3e}, tcgen05 err {errs[2]:.3e}, "
          f"engine 3 err {errs[3]:.3e}, fp16 pairs err {errs[4]:.3e}, fp16 pairs / two chains err {errs[5]:.3e}")
    assert max(errs) < 2e-5


@pytest.mark.parametrize("engine", [4, 5])
@pytest.mark.parametrize("scale", [1.0, 1e-3, 6.0])
def test_lstm_f16_engine_relaunch_and_weight_scales(scale, engine):
    """Engine 4 (csrc/lstm_f16.cu): (a) back-to-back launches on the same work buffer with DIFFERENT inputs and lengths --
    the tagged state words of one launch must never validate in the next; (b) bit-identical repeats (no race in the
    tag protocol shows up as run-to-run differences); (c) weight magnitudes from 1e-3 to 6 x 1/sqrt(H) exercise the
    per-block power-of-two scale; (d) a W_hh with all-zero blocks."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    h = 1024
    g = torch.Generator().manual_seed(11)
    whh = torch.randn(h // 8, h, 32, generator=g) / np.sqrt(h) * scale
    whh[3:9] = 0.0
    outs = []
    ops.set_lstm_engine(engine)
    # large recurrent weights make the recurrence chaotic (any fp32 rounding difference grows): few steps there
    for b, t in ([(64, 37), (17, 50), (64, 37), (64, 38)] if scale <= 1.0 else [(64, 4), (17, 5), (64, 4), (64, 3)]):
        gg = torch.Generator().manual_seed(1000 + b + t)
        xp = torch.randn(b, t, 4 * h, generator=gg)
        got = ops.lstm_seq(xp.to(dev), whh.to(dev), h)
        again = ops.lstm_seq(xp.to(dev), whh.to(dev), h)
        torch.cuda.synchronize()
        assert torch.equal(got, again)
        ref = emu_ops.lstm_seq(xp.double(), whh.double(), h)
        err = (got.cpu().double() - ref).abs().max().item()
        outs.append(err)
        assert err < 2e-5, (b, t, err)
    ops.set_lstm_engine(DEFAULT_LSTM_ENGINE)
    print(f"lstm f16 engine {engine}, weight scale {scale}: errs {['%.2e' % e for e in outs]}")


@pytest.mark.parametrize("c", [1, 16, 130])
def test_pointwise_kernels_match_semantics(c):
    """se_glu_affine_act (vector and scalar paths, fp32 + TF32-split outputs), se_unary, se_cmul."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(c)
    x = torch.randn(3, 7, 5, 2 * c, generator=g)
    sc, sh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    got, pair = ops.glu_affine_act(x.to(dev), sc.to(dev), sh.to(dev), "elu", want_f32=True, want_pair=True)
    ref, _ = emu_ops.glu_affine_act(x.double(), sc.double(), sh.double(), "elu")
    assert (got.cpu().double() - ref).abs().max() < 1e-5
    assert ((pair[0] + pair[1]).cpu().double() - ref).abs().max() < 1e-5
    u, up = ops.unary(x.to(dev), "elu", want_f32=True, want_pair=True)
    uref = torch.where(x > 0, x, torch.expm1(x))
    assert (u.cpu() - uref).abs().max() < 1e-6 and ((up[0] + up[1]).cpu() - uref).abs().max() < 1e-5
    a, m = torch.randn(4, 9, 161, 2, generator=g), torch.randn(4, 9, 161, 2, generator=g)
    z = ops.cmul(a.to(dev), m.to(dev)).cpu()
    zr = torch.view_as_real(torch.view_as_complex(a) * torch.view_as_complex(m))
    assert (z - zr).abs().max() < 1e-5


def test_group_layernorm_wide_with_store_index():
    """C = 1024 (GCRN GLSTM LayerNorm) with the out_index permutation, and C = 512 + residual (DPCRN ln1/ln2)."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(37, 1024, generator=g) * 3 + 1
    gam, bet = torch.randn(1024, generator=g), torch.randn(1024, generator=g)
    idx = torch.randperm(1024, generator=g).to(torch.int32)
    y, pair = ops.group_layernorm(x.to(dev), 1, gam.to(dev), bet.to(dev), want_f32=True, want_pair=True,
                                  out_index=idx.to(dev))
    ref, _ = emu_ops.group_layernorm(x.double(), 1, gam.double(), bet.double(), out_index=idx)
    assert (y.cpu().double() - ref).abs().max() < 2e-5
    assert ((pair[0] + pair[1]).cpu().double() - ref).abs().max() < 2e-5
    x2, r2 = torch.randn(50, 512, generator=g), torch.randn(50, 512, generator=g)
    y2, _ = ops.group_layernorm(x2.to(dev), 1, gam[:512].contiguous().to(dev), bet[:512].contiguous().to(dev),
                                res=r2.to(dev))
    ref2, _ = emu_ops.group_layernorm(x2.double(), 1, gam[:512].double(), bet[:512].double(), res=r2.double())
    assert (y2.cpu().double() - ref2).abs().max() < 2e-5


@pytest.mark.parametrize("m", [200, 1031])
def test_lstm_cell_ex_strided_bidirectional(m):
    """se_lstm_cell_tf32x3_ex as DPCRN's intra Bi-LSTM uses it: strided x / h views of [M, 4, 128] buffers, first
    step without state, forward and reverse position order -- against a float64 bidirectional LSTM."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m)
    x = torch.randn(m, 4, 128, generator=g)
    out_ref = torch.zeros(m, 4, 128, dtype=torch.float64)
    dst = (torch.zeros(m, 4, 128, device=dev), torch.zeros(m, 4, 128, device=dev))
    src = ops.split_tf32(x.to(dev))
    cst = torch.empty(m, 64, device=dev)
    for d in range(2):
        w_ih = torch.randn(256, 128, generator=g) / np.sqrt(128)
        w_hh = torch.randn(256, 64, generator=g) / 8
        b_ih, b_hh = torch.randn(256, generator=g) * 0.1, torch.randn(256, generator=g) * 0.1
        P = packing.pack_lstm_cell(w_ih, w_hh, b_ih, b_hh)
        order = (0, 1, 2, 3) if d == 0 else (3, 2, 1, 0)
        prev = None
        hr = torch.zeros(m, 64, dtype=torch.float64)
        cr = torch.zeros(m, 64, dtype=torch.float64)
        for f in order:
            hh, hl = dst[0][:, f, 64 * d:64 * d + 64], dst[1][:, f, 64 * d:64 * d + 64]
            ops.lstm_cell_tf32x3_ex((src[0][:, f], src[1][:, f]), prev, P["w_hi"].to(dev), P["w_lo"].to(dev),
                                    P["bias"].to(dev), cst, hh, hl)
            prev = (hh, hl)
            gates = x[:, f].double() @ w_ih.double().t() + hr @ w_hh.double().t() + (b_ih + b_hh).double()
            i, ff, gg, o = gates.chunk(4, dim=1)
            cr = torch.sigmoid(ff) * cr + torch.sigmoid(i) * torch.tanh(gg)
            hr = torch.sigmoid(o) * torch.tanh(cr)
            out_ref[:, f, 64 * d:64 * d + 64] = hr
    torch.cuda.synchronize()
    err = ((dst[0] + dst[1]).cpu().double() - out_ref).abs().max().item()
    print(f"lstm_cell_ex M={m}: max err {err:.3e}")
    assert err < 2e-5


@pytest.mark.parametrize("engine", [0, 3])
@pytest.mark.parametrize("b", [9, 37, 64])       # engine 3 at H = 128: 1 / 1 / 2 sequences per CTA with 4 groups
def test_lstm_seq_multi_shared_weights(b, engine):
    """Groups that share one W_hh (DPCRN inter-LSTM: the 4 frequency positions of a frame)."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(5)
    t, h, ng = 12, 128, 4
    xp = torch.randn(b, t, ng * 4 * h, generator=g)
    whh = torch.randn(h // 8, h, 32, generator=g) / np.sqrt(h)
    out = torch.empty(b, t, ng * h, device=dev)
    try:
        ops.set_lstm_engine(engine)
        ops.lstm_seq_multi(xp.to(dev), whh.to(dev), h, ng, out)
    finally:
        ops.set_lstm_engine(DEFAULT_LSTM_ENGINE)
    ref = emu_ops.lstm_seq_multi(xp.double(), whh.double(), h, ng, torch.empty(b, t, ng * h, dtype=torch.float64))
    assert (out.cpu().double() - ref).abs().max() < 2e-5


def test_lstm_seq_small_eight_groups_four_sequences_per_cta():
    """Sequence-parallel H = 128 kernel with 8 groups x 64 sequences = 512 sequences: 4 per CTA (128 CTAs)."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(6)
    b, t, h, ng = 64, 7, 128, 8
    xp = torch.randn(b, t, ng * 4 * h, generator=g)
    whh = torch.randn(ng, h // 8, h, 32, generator=g) / np.sqrt(h)
    out = torch.empty(b, t, ng * h, device=dev)
    ops.lstm_seq_multi(xp.to(dev), whh.to(dev), h, ng, out)
    for k in range(ng):
        ref = emu_ops.lstm_seq(xp[:, :, k * 4 * h:(k + 1) * 4 * h].double(), whh[k].double(), h)
        assert (out[:, :, k * h:(k + 1) * h].cpu().double() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("b,rows,c,pre", [(3, 401 * 79, 64, "glu"), (2, 401, 128, "prelu"), (5, 161 * 50, 1, "glu"),
                                          (2, 700, 64, "glu_prelu"), (1, 33, 64, "none")])
def test_chan_stats_and_norm_instance(b, rows, c, pre):
    """InstanceNorm statistics + normalise/PReLU pass vs the fp64 mirror."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(rows % 97)
    cin = {"glu": 2 * c, "glu_prelu": 2 * c, "prelu": c // 2 if c > 1 else c, "none": c}[pre]
    x = torch.randn(b, rows, cin, generator=g) * 2 + 0.3
    slope = torch.rand(c, generator=g) * 0.5
    gamma, beta, ps = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.2, torch.rand(c, generator=g)
    m_ref, r_ref = emu_ops.chan_stats(x, b, rows, c, pre, slope)
    y_ref, _ = emu_ops.chan_norm(x.double(), b, rows, c, m_ref.double(), r_ref.double(), gamma.double(), beta.double(),
                                 pre=pre, pre_slope=slope.double(), post="prelu", post_slope=ps.double())
    xd, sd_ = x.to(dev), slope.to(dev)
    for rep in range(2):                      # the second call checks that the workspace ticket was left at zero
        m, r = ops.chan_stats(xd, b, rows, c, pre, sd_)
    y, pair = ops.chan_norm(xd, b, rows, c, m, r, gamma.to(dev), beta.to(dev), pre=pre, pre_slope=sd_, post="prelu",
                            post_slope=ps.to(dev), want_f32=True, want_pair=True)
    em, er = (m.cpu() - m_ref).abs().max().item(), ((r.cpu() - r_ref).abs() / r_ref).max().item()
    ey = (y.cpu().double() - y_ref).abs().max().item()
    ep = ((pair[0] + pair[1]).cpu().double() - y_ref).abs().max().item()
    print(f"chan_stats/norm B={b} rows={rows} C={c} pre={pre}: mean err {em:.2e}, rstd rel err {er:.2e}, y err {ey:.2e}")
    assert em < 1e-6 and er < 1e-6 and ey < 2e-5 and ep < 2e-5


@pytest.mark.parametrize("k", [1, 3, 63])
@pytest.mark.parametrize("cumulative", [False, True])
def test_chan_norm_fir_and_cumulative(k, cumulative):
    """TCM branch prologue: two PReLU/norm/ShareSepConv branches read one 64-channel tensor -> 128 channels."""
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    b, t, c = 3, 150, 128
    g = torch.Generator().manual_seed(k)
    x = torch.randn(b, t, 64, generator=g)
    slope, gamma, beta = torch.rand(c, generator=g) * 0.5, torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    fir = torch.randn(2, k, generator=g) * 0.5
    xd = x.to(dev)
    if cumulative:
        m_ref, r_ref = emu_ops.cum_stats(x, b, t, 1, c, "prelu", slope, groups=2)
        m, r = ops.cum_stats(xd, b, t, 1, c, "prelu", slope.to(dev), groups=2)
    else:
        m_ref, r_ref = emu_ops.chan_stats(x, b, t, c, "prelu", slope)
        m, r = ops.chan_stats(xd, b, t, c, "prelu", slope.to(dev))
    kw = dict(pre="prelu", cumulative=cumulative, rows_per_t=1, stat_groups=2, post="fir", fir_groups=2)
    y_ref, _ = emu_ops.chan_norm(x.double(), b, t, c, m_ref.double(), r_ref.double(), gamma.double(), beta.double(),
                                 pre_slope=slope.double(), fir_w=fir.double(), **kw)
    _, pair = ops.chan_norm(xd, b, t, c, m, r, gamma.to(dev), beta.to(dev), pre_slope=slope.to(dev), fir_w=fir.to(dev),
                            want_f32=False, want_pair=True, **kw)
    es = ((r.cpu() - r_ref).abs() / r_ref).max().item()
    ey = ((pair[0] + pair[1]).cpu().double() - y_ref).abs().max().item()
    print(f"chan_norm FIR k={k} cumulative={cumulative}: rstd rel err {es:.2e}, y err {ey:.2e} (|y| max {y_ref.abs().max():.2f})")
    assert es < 1e-5 and ey < 1e-5 * max(1.0, y_ref.abs().max().item())


def test_cum_stats_2d_and_cts_glue():
    dev = _dev()
    import se_b200
    ops = se_b200.ops
    g = torch.Generator().manual_seed(11)
    b, t, f, c = 2, 60, 19, 64
    x = torch.randn(b, t, f, 2 * c, generator=g)
    m_ref, r_ref = emu_ops.cum_stats(x, b, t, f, c, "glu")
    m, r = ops.cum_stats(x.to(dev), b, t, f, c, "glu")
    gamma, beta, ps = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g)
    y_ref, _ = emu_ops.chan_norm(x, b, t * f, c, m_ref, r_ref, gamma, beta, pre="glu", cumulative=True, rows_per_t=f,
                                 post="prelu", post_slope=ps)
    y, _ = ops.chan_norm(x.to(dev), b, t * f, c, m, r, gamma.to(dev), beta.to(dev), pre="glu", cumulative=True,
                         rows_per_t=f, post="prelu", post_slope=ps.to(dev))
    assert ((r.cpu() - r_ref).abs() / r_ref).max() < 1e-5 and (y.cpu() - y_ref).abs().max() < 2e-5
    xr = torch.randn(b, t, 161, 2, generator=g)
    xr[0, 0, 0] = 0.0                                          # atan2(0, 0) = 0 -> (cos, sin) = (1, 0)
    e = torch.rand(b, t, 161, generator=g)
    s2 = ops.cts_glue1(xr.to(dev), e.to(dev))
    assert (s2.cpu() - emu_ops.cts_glue1(xr, e)).abs().max() < 1e-6
    o_r, o_i = torch.randn(b, t, 161, generator=g), torch.randn(b, t, 161, generator=g)
    est = ops.cts_glue2(o_r.to(dev), o_i.to(dev), s2)
    assert (est.cpu() - emu_ops.cts_glue2(o_r, o_i, s2.cpu())).abs().max() < 1e-6
    a, bb = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    s, pr = ops.add(a.to(dev), bb.to(dev), want_pair=True)
    assert torch.equal(s.cpu(), a + bb) and ((pr[0] + pr[1]).cpu() - (a + bb)).abs().max() < 1e-6   # hi+lo: 21+ bits
    z, _ = ops.axpby(a.to(dev), bb.to(dev), 1.0, 1.0 / 6.0)
    assert (z.cpu() - (a + bb / 6.0)).abs().max() < 1e-6
    gain = torch.rand(b, t, 161, generator=g)
    rows, rp = ops.taylor_zero(xr.to(dev), gain.to(dev), 352)
    ref_rows, _ = emu_ops.taylor_zero(xr, gain, 352)
    assert (rows.cpu() - ref_rows).abs().max() < 2e-6 and ((rp[0] + rp[1]).cpu() - ref_rows).abs().max() < 2e-6
    assert rows[:, 322:].abs().max().item() == 0.0
    # G2Net stage update on RI rows (re at 0, im at 176): relayout of the channels-last input, then gain / residual
    n = b * t
    r0, p0 = ops.gaf_update(xr.to(dev), xr.to(dev)[..., 1], 322, 2, None, None, n, 161, 352, 176)
    e0, _ = emu_ops.gaf_update(xr, xr[..., 1], 322, 2, None, None, n, 161, 352, 176)
    assert torch.equal(r0.cpu(), e0)
    resi = torch.zeros(n, 352)
    resi[:, :161], resi[:, 176:337] = torch.randn(n, 161, generator=g), torch.randn(n, 161, generator=g)
    r1, p1 = ops.gaf_update(r0, r0[:, 176:], 352, 1, gain.to(dev).view(n, 161), resi.to(dev), n, 161, 352, 176)
    e1, _ = emu_ops.gaf_update(e0, e0[:, 176:], 352, 1, gain.view(n, 161), resi, n, 161, 352, 176)
    assert (r1.cpu() - e1).abs().max() < 2e-6 and ((p1[0] + p1[1]).cpu() - e1).abs().max() < 2e-6


@pytest.mark.parametrize("mode", ["E", "C", "R"])
def test_dccrn_mask_modes(mode):
    """se_dccrn_mask_ex vs the three branches of DCCRN.forward (DCCRN_cprs.py:206-224) on random tensors."""
    dev = _dev()
    import se_b200
    g = torch.Generator().manual_seed(11)
    b, t, f = 2, 9, 257
    m = torch.randn(b, t, f - 1, 2, generator=g)
    x = torch.randn(b, t, f, 2, generator=g)
    e = torch.empty(b, t, f, 2, device=dev)
    se_b200.ops.dccrn_mask(m.to(dev), x[..., 0].to(dev), x[..., 1].to(dev), e[..., 0], e[..., 1], mode=mode)
    er, ei = torch.empty(b, t, f), torch.empty(b, t, f)
    emu_ops.dccrn_mask(m, x[..., 0], x[..., 1], er, ei, mode=mode)
    assert (e[..., 0].cpu() - er).abs().max() < 2e-6 and (e[..., 1].cpu() - ei).abs().max() < 2e-6


def test_pad_split_tf32_and_odd_k_projection():
    """se_pad_split_tf32: rows of 161 floats -> 192 with a zero tail, split hi / lo; the LSTM-net layer-0 projection
    (K = 161, LSTM/LSTM.py:17) through the tensor-core GEMM on the padded operands vs fp64."""
    dev = _dev()
    import se_b200
    from se_b200 import lstm_engine, packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(21)
    m, k, h = 401 * 3, 161, 1024
    x = torch.randn(m, k, generator=g)
    hi, lo = ops.pad_split_tf32(x.to(dev), 192)
    assert hi.shape == (m, 192) and (hi[:, k:] == 0).all() and (lo[:, k:] == 0).all()
    assert ((hi + lo)[:, :k].cpu() - x).abs().max() < 2e-6
    w_ih = torch.randn(4 * h, k, generator=g) / np.sqrt(k)
    w_hh = torch.randn(4 * h, h, generator=g) / np.sqrt(h)
    b_ih, b_hh = torch.randn(4 * h, generator=g), torch.randn(4 * h, generator=g)
    layer = packing.pack_lstm_layer(w_ih, w_hh, b_ih, b_hh)
    assert layer["wih_hi"].shape == (4 * h, 192) and layer["kin"] == k
    layer = {kk: (v.to(dev) if torch.is_tensor(v) else v) for kk, v in layer.items()}
    got = lstm_engine.input_projection(x.to(dev), layer)
    rows = packing.slice_rows(h)
    ref = x.double() @ w_ih.double()[rows].t() + (b_ih + b_hh).double()[rows]
    assert (got.cpu().double() - ref).abs().max() < 1e-5


@pytest.mark.parametrize("ch", [1, 8, 32, 130])
def test_uf_fusion_ex_outputs(ch):
    """se_uf_fusion_ex: fp32 and TF32-pair outputs of the cross-branch fusion (vector path for C % 4 == 0, scalar else)."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(ch)
    c = torch.randn(3, 5, 7, 2 * ch, generator=g)
    m = torch.randn(3, 5, 7, ch, generator=g)
    rc, rm = emu_ops.uf_fusion(c.double(), m.double())
    cf, cp, mf, mp = ops.uf_fusion_ex(c.to(dev), m.to(dev), c_f32=True, c_pair=True, m_f32=True, m_pair=True)
    assert (cf.cpu().double() - rc).abs().max() < 2e-6 and (mf.cpu().double() - rm).abs().max() < 2e-6
    for f32, pair in ((cf, cp), (mf, mp)):
        hi, lo = packing.split_tf32(f32.cpu())
        assert torch.equal(pair[0].cpu(), hi) and torch.equal(pair[1].cpu(), lo)
    cf2, cp2, mf2, mp2 = ops.uf_fusion_ex(c.to(dev), m.to(dev), c_f32=False, c_pair=True, m_f32=True, m_pair=False)
    assert cf2 is None and mp2 is None and torch.equal(cp2[0], cp[0]) and torch.equal(mf2, mf)



@pytest.mark.parametrize("case", [(2, 9, 39, 32, 0, 64, "conv"), (1, 7, 19, 64, 64, 128, "deconv"),
                                  (2, 5, 9, 128, 128, 256, "deconv")])
def test_conv_tf32x3_gated_epilogue(case):
    """Gated conv fused into the tensor-core epilogue (se_conv_tc_desc.glu): columns (2j, 2j+1) = (conv1, conv2),
    out = ELU((conv1 * sigmoid(conv2)) * scale + shift) with Cout / 2 channels, fp32 and TF32-pair outputs."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    from se_b200.gcrn import DEC_EVEN, DEC_ODD, ENC_TAPS
    ops = se_b200.ops
    b, t, fin, c0, c1, co2, kind = case
    g = torch.Generator().manual_seed(co2 + fin)
    x0 = torch.randn(b, t, fin, c0, generator=g)
    x1 = torch.randn(b, t, fin, c1, generator=g) if c1 else None
    ct = c0 + c1
    bias = torch.randn(co2, generator=g)
    scale = torch.rand(co2 // 2, generator=g) + 0.5
    shift = torch.randn(co2 // 2, generator=g)
    if kind == "conv":
        fout = (fin - 3) // 2 + 1
        runs, dstF = [(ENC_TAPS, 2, fout, 0, 1)], fout
    else:
        dstF = 2 * fin + 1
        runs = [(DEC_EVEN, 1, fin + 1, 0, 2), (DEC_ODD, 1, fin, 1, 2)]
    ref = torch.zeros(b, t, dstF, co2 // 2, dtype=torch.float64)
    got = torch.zeros(b, t, dstF, co2 // 2, device=dev)
    got_hi, got_lo = torch.zeros_like(got), torch.zeros_like(got)
    s0 = ops.split_tf32(x0.to(dev))
    s1 = ops.split_tf32(x1.to(dev)) if x1 is not None else None
    for taps, sf, fo, f0, fstep in runs:
        w = torch.randn(len(taps) * ct, co2, generator=g) / np.sqrt(len(taps) * ct)
        w_hi, w_lo = packing.split_tf32(w.t().contiguous())
        emu_ops.conv_tf32x3((x0.double(), torch.zeros_like(x0).double()),
                            None if x1 is None else (x1.double(), torch.zeros_like(x1).double()), b, t, fin, fo, taps, sf,
                            w.t().contiguous().double(), torch.zeros(co2, len(taps) * ct, dtype=torch.float64),
                            bias.double(), co2, "elu", dstF, f0, fstep, out=ref, glu=(scale.double(), shift.double()))
        ops.conv_tf32x3(s0, s1, b, t, fin, fo, taps, sf, w_hi.to(dev), w_lo.to(dev), bias.to(dev), co2, "elu", dstF, f0,
                        fstep, out=got, out_pair=(got_hi, got_lo), glu=(scale.to(dev), shift.to(dev)))
    torch.cuda.synchronize()
    err = (got.cpu().double() - ref).abs().max().item()
    err_pair = ((got_hi + got_lo).cpu().double() - ref).abs().max().item()
    print(f"gated conv {case}: max err {err:.3e} (hi+lo {err_pair:.3e})")
    assert err < 2e-5 and err_pair < 2e-5


@pytest.mark.parametrize("m,k,n,act", [(256, 64, 256, "none"), (300, 1024, 256, "none"), (1000, 192, 4096, "softplus"),
                                       (2309, 2048, 512 + 164, "none"), (25664, 1024, 4096, "none"),
                                       (130, 40, 24, "relu")])
@pytest.mark.parametrize("engine", [0, 1])
def test_gemm_f16x3_fp32_accuracy(m, k, n, act, engine):
    """fp16-pair GEMM (tcgen05 kind::f16, three products of scaled hi/lo halves): fp32-class error vs fp64, at least as
    good as the 3xTF32 engine on the same data; also the fp16-pair / TF32-pair copies of the output it can emit."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g) * 3.0
    a[0, 0] = 900.0                        # far-out activations stay inside the fp16 range at scale 2^4
    a[1, : min(k, 8)] = 1e-5               # and tiny ones keep an absolute floor of 2^-29
    w = torch.randn(n, k, generator=g) / np.sqrt(k)
    bias = torch.randn(n, generator=g)
    w_hi, w_lo, ws = packing.split_f16(w)
    assert ((w_hi.double() + w_lo.double()) * 2.0 ** -ws - w.double()).abs().max().item() < 2e-7 * w.abs().max().item()
    a_pair = ops.split_f16(a.to(dev))
    back = (a_pair[0].double() + a_pair[1].double()).cpu()[:, :k] / 16.0
    assert bool(((back - a.double()).abs() <= 2.0 ** -21 * a.double().abs() + 2.0 ** -28).all())   # 22 bits or the floor
    rows = torch.unique(torch.cat([torch.arange(0, min(m, 300)), torch.arange(max(0, m - 300), m),
                                   torch.randint(0, m, (256,), generator=g)]))
    ref = emu_ops._act(a[rows].double() @ w.double().t() + bias.double(), act)
    try:
        ops.set_gemm_engine(engine)
        got, p32, p16 = ops.gemm_f16x3(a_pair, (w_hi.to(dev), w_lo.to(dev)), ws, bias.to(dev), n, act, pair_out=True,
                                       pair16_out=True)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)
    err = (got.cpu()[rows].double() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    e32 = ((p32[0] + p32[1]).cpu()[rows].double() - got.cpu()[rows].double()).abs().max().item()
    e16 = ((p16[0].double() + p16[1].double()).cpu()[rows] / 16.0 - got.cpu()[rows].double()).abs().max().item()
    print(f"gemm_f16x3 engine {engine} {m}x{k}x{n}: max err vs fp64 {err:.3e} (|ref| max {scale:.1f}), "
          f"TF32-pair copy {e32:.2e}, fp16-pair copy {e16:.2e}")
    assert err < 1e-5 * scale and e32 < 1e-6 * scale and e16 < 1e-6 * scale


@pytest.mark.parametrize("m,kx,h,steps", [(300, 32, 384, 4), (8224, 32, 384, 3), (1031, 64, 128, 3)])
@pytest.mark.parametrize("engine", [0, 1])
def test_lstm_cell_f16x3_matches_recurrence(m, kx, h, steps, engine):
    """Fused LSTM cell on fp16 operand pairs vs the float64 recurrence (and hence vs the TF32 cell's 1e-5 gate)."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    ops = se_b200.ops
    g = torch.Generator().manual_seed(m + h)
    w_ih = torch.randn(4 * h, kx, generator=g) / np.sqrt(kx)
    w_hh = torch.randn(4 * h, h, generator=g) / np.sqrt(h)
    b_ih = torch.randn(4 * h, generator=g) * 0.1
    b_hh = torch.randn(4 * h, generator=g) * 0.1
    xs = torch.randn(steps, m, kx, generator=g) * 2.0
    P = packing.pack_lstm_cell_f16(w_ih, w_hh, b_ih, b_hh)
    P = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in P.items()}
    c = torch.zeros(m, h, device=dev)
    z16 = lambda: torch.zeros(m, h, device=dev, dtype=torch.float16)   # noqa: E731
    hbuf = [(z16(), z16()), (z16(), z16())]
    hout = torch.empty(m, h, device=dev)
    hr = torch.zeros(m, h, dtype=torch.float64)
    cr = torch.zeros(m, h, dtype=torch.float64)
    try:
        ops.set_gemm_engine(engine)
        for t in range(steps):
            x_pair = ops.split_f16(xs[t].to(dev))
            src, dst = hbuf[t & 1], hbuf[(t + 1) & 1]
            ops.lstm_cell_f16x3(x_pair, None if t == 0 else src, P, c, dst[0], dst[1], hout)
            gates = xs[t].double() @ w_ih.double().t() + hr @ w_hh.double().t() + (b_ih + b_hh).double()
            i, f, gg, o = gates.chunk(4, dim=1)
            cr = torch.sigmoid(f) * cr + torch.sigmoid(i) * torch.tanh(gg)
            hr = torch.sigmoid(o) * torch.tanh(cr)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)
    e_h = (hout.cpu().double() - hr).abs().max().item()
    e_c = (c.cpu().double() - cr).abs().max().item()
    e_split = ((dst[0].double() + dst[1].double()).cpu() / 16.0 - hr).abs().max().item()
    print(f"lstm_cell_f16 engine {engine} M={m} Kx={kx} H={h} steps={steps}: h err {e_h:.3e} c err {e_c:.3e} "
          f"hi+lo err {e_split:.3e}")
    assert e_h < 1e-5 and e_c < 1e-5 and e_split < 1e-5


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("case", TC_CONV_CASES + [(3, 5, 9, 128, 0, 256, "crn_conv"), (1, 1, 8, 256, 0, 192, "dccrn_conv"),
                                                  (2, 23, 80, 16, 0, 32, "crn_conv"),       # 16 of a 64-channel k-block
                                                  (2, 9, 39, 96, 40, 16, "crn_deconv")])    # ragged slices, two sources
def test_conv_f16x3_matches_semantics(case, engine):
    """The tensor-core conv on fp16 operand pairs (64-channel k-blocks, zero-filled / zero-padded ragged slices) vs the
    declared conv semantics in fp64; fp32, TF32-pair and fp16-pair outputs."""
    dev = _dev()
    import se_b200
    from se_b200 import packing
    from se_b200.dccrn import DEC_EVEN, DEC_ODD, ENC_TAPS
    ops = se_b200.ops
    b, t, fin, c0, c1, co, kind = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 1000)
    x0 = torch.randn(b, t, fin, c0, generator=g) * 2.0
    x1 = torch.randn(b, t, fin, c1, generator=g) if c1 else None
    ct = c0 + c1
    bias = torch.randn(co, generator=g)
    if kind == "crn_conv":
        fout = (fin - 3) // 2 + 1
        runs = [(packing.CONV23_TAPS, 2, fout, 0, 1)]
        dstF = fout
    elif kind == "crn_deconv":
        dstF = 2 * fin + 1
        runs = [(packing.DECONV_EVEN_TAPS, 1, fin + 1, 0, 2), (packing.DECONV_ODD_TAPS, 1, fin, 1, 2)]
    elif kind == "dccrn_conv":
        runs = [(ENC_TAPS, 2, fin // 2, 0, 1)]
        dstF = fin // 2
    else:
        dstF = 2 * fin
        runs = [(DEC_EVEN, 1, fin, 0, 2), (DEC_ODD, 1, fin, 1, 2)]
    ref = torch.zeros(b, t, dstF, co, dtype=torch.float64)
    got = torch.zeros(b, t, dstF, co, device=dev)
    got_hi, got_lo = torch.zeros_like(got), torch.zeros_like(got)
    g16_hi, g16_lo = torch.zeros_like(got, dtype=torch.float16), torch.zeros_like(got, dtype=torch.float16)

    def pair16(x):
        hi, lo = ops.split_f16(x.to(dev).view(-1, x.shape[-1]))
        return hi.view(x.shape), lo.view(x.shape)
    s0 = pair16(x0)
    s1 = pair16(x1) if x1 is not None else None
    try:
        ops.set_gemm_engine(engine)
        for taps, sf, fout, f0, fstep in runs:
            w = torch.randn(len(taps) * ct, co, generator=g) / np.sqrt(len(taps) * ct)
            emu_ops.conv_gemm(x0.double(), None if x1 is None else x1.double(), b, t, fin, fout, taps, sf, w.double(),
                              bias.double(), co, "prelu", ref, dstF, f0, fstep, -1, None, 0.2)
            w_hi, w_lo, ws = packing.pack_conv_f16(w.t().contiguous(), len(taps), c0, c1)
            ops.conv_f16x3(s0, s1, b, t, fin, fout, taps, sf, w_hi.to(dev), w_lo.to(dev), ws, bias.to(dev), co, "prelu",
                           dstF, f0, fstep, act_param=0.2, out=got, out_pair=(got_hi, got_lo), out_pair16=(g16_hi, g16_lo))
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)
    err = (got.cpu().double() - ref).abs().max().item()
    err_pair = ((got_hi + got_lo).cpu().double() - ref).abs().max().item()
    err_16 = ((g16_hi.double() + g16_lo.double()).cpu() / 16.0 - ref).abs().max().item()
    print(f"conv_f16x3 engine {engine} {case}: max err {err:.3e} (TF32 pair {err_pair:.3e}, fp16 pair {err_16:.3e})")
    assert err < 2e-5 and err_pair < 2e-5 and err_16 < 2e-5


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("case", [(2, 33, 4, 256, 256, 128, "crn"), (1, 17, 19, 64, 64, 32, "crn"), (2, 9, 39, 32, 32, 16, "crn"),
                                  (2, 13, 4, 256, 256, 256, "dccrn"), (1, 7, 64, 64, 64, 32, "dccrn")])
def test_conv_f16x3_two_parity_classes_in_one_launch(case, engine):
    """se_conv_f16x3 with ncls = 2 (both output-column parity classes of a stride-2 transposed conv in one launch, the odd
    class's taps a subset of the even class's) vs the declared semantics of the two classes in fp64."""
    dev = _dev()
    import se_b200
    from se_b200 import conv_engine, packing
    from se_b200.conv_engine import Act, ConvWeights
    from se_b200.dccrn import DEC_EVEN, DEC_ODD
    ops = se_b200.ops
    b, t, fin, c0, c1, co, kind = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 1000)
    x0 = torch.randn(b, t, fin, c0, generator=g)
    x1 = torch.randn(b, t, fin, c1, generator=g)
    bias = torch.randn(co, generator=g)
    ev, od = (packing.DECONV_EVEN_TAPS, packing.DECONV_ODD_TAPS) if kind == "crn" else (DEC_EVEN, DEC_ODD)
    fe, fo_, dstF = (fin + 1, fin, 2 * fin + 1) if kind == "crn" else (fin, fin, 2 * fin)
    ct = c0 + c1
    we = torch.randn(len(ev) * ct, co, generator=g) / np.sqrt(len(ev) * ct)
    wo = torch.randn(len(od) * ct, co, generator=g) / np.sqrt(len(od) * ct)
    ref = torch.zeros(b, t, dstF, co, dtype=torch.float64)
    emu_ops.conv_gemm(x0.double(), x1.double(), b, t, fin, fe, ev, 1, we.double(), bias.double(), co, "elu", ref, dstF, 0, 2)
    emu_ops.conv_gemm(x0.double(), x1.double(), b, t, fin, fo_, od, 1, wo.double(), bias.double(), co, "elu", ref, dstF, 1, 2)
    wm = conv_engine.merge_parity(ConvWeights(we.to(dev), co), ev, ConvWeights(wo.to(dev), co), od)
    src, skip = Act(x0.to(dev)), Act(x1.to(dev))
    out = conv_engine.new_act(b, t, dstF, co, dev, want_f32=True, want_pair=True, f16=True)
    out.f32.zero_()
    assert conv_engine.parity2_eligible(src, skip, wm, fe) == (co <= 64)     # the models merge narrow layers only
    try:
        ops.set_gemm_engine(engine)
        conv_engine.conv_parity2(src, skip, b, t, fin, fe, fo_, ev, wm, bias.to(dev), "elu", out, dstF)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_engine(DEFAULT_GEMM_ENGINE)
    err = (out.f32.cpu().double() - ref).abs().max().item()
    err16 = ((out.pair[0].double() + out.pair[1].double()).cpu() / 16.0 - ref).abs().max().item()
    print(f"conv_f16x3 ncls=2 engine {engine} {case}: max err {err:.3e} (fp16 pair {err16:.3e})")
    assert err < 2e-5 and err16 < 2e-5
