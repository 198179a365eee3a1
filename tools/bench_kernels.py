"""Kernel-level timings on the GPU box (CUDA events, warm, best of N): the building blocks of the
CRN step at the bench shapes.  Development tool; numbers feed DESIGN.md / profiles/."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, packing


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda")
    B, T, H = 64, 401, 1024
    M = B * T
    res = {}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, 1024, generator=g).to(dev)
    w_nk = (torch.randn(4096, 1024, generator=g) / 32).to(dev)
    bias = torch.randn(4096, generator=g).to(dev)
    w_kn = packing.pad_cols(w_nk.t().contiguous())
    out = torch.empty(M, 4096, device=dev)
    fl = 2.0 * M * 1024 * 4096
    ms = timeit(lambda: ops.linear(x, w_kn, bias, 4096, out=out))
    res["linear_simt_25664x1024x4096"] = {"ms": ms, "tflops": fl / ms / 1e9}
    ref = out.clone()
    if "--no-tc" not in sys.argv:
        w_hi, w_lo = ops.split_tf32(w_nk)
        ms_split = timeit(lambda: ops.split_tf32(x))
        x_hi, x_lo = ops.split_tf32(x)
        out2 = torch.empty(M, 4096, device=dev)
        ms = timeit(lambda: ops.gemm_tf32x3(x_hi, x_lo, w_hi, w_lo, bias, 4096, out=out2))
        res["gemm_tf32x3_25664x1024x4096"] = {"ms": ms, "tflops_fp32_equiv": fl / ms / 1e9, "split_ms": ms_split,
                                              "max_abs_diff_vs_simt": (out2 - ref).abs().max().item()}
        # other engines on the same shape: 1 = CTA pairs (cta_group::2, 256x256 tiles), 2 / 3 / 4 = multicast clusters
        for eng in (1, 2, 3, 4):
            ops.set_gemm_engine(eng)
            out3 = torch.empty(M, 4096, device=dev)
            ms = timeit(lambda: ops.gemm_tf32x3(x_hi, x_lo, w_hi, w_lo, bias, 4096, out=out3))
            res[f"gemm_tf32x3_engine{eng}_25664x1024x4096"] = {"ms": ms, "tflops_fp32_equiv": fl / ms / 1e9,
                                                             "max_abs_diff_vs_one_cta": (out3 - out2).abs().max().item()}
        for eng in (0, 1, 2, 3, 4):
            ops.set_gemm_engine(eng)
            m2, kx, h2 = 32 * 257, 384, 384
            Pc = packing.pack_lstm_cell(torch.randn(4 * h2, kx, generator=g) / 20, torch.randn(4 * h2, h2, generator=g) / 20,
                                        torch.zeros(4 * h2), torch.zeros(4 * h2))
            Pc = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in Pc.items()}
            xh, xl = ops.split_tf32(torch.randn(m2, kx, generator=g).to(dev))
            hh, hl = ops.split_tf32(torch.randn(m2, h2, generator=g).to(dev) * 0.1)
            c = torch.zeros(m2, h2, device=dev)
            oh, ol, ho = (torch.empty(m2, h2, device=dev) for _ in range(3))
            ms = timeit(lambda: ops.lstm_cell_tf32x3(xh, xl, hh, hl, Pc["w_hi"], Pc["w_lo"], Pc["bias"], c, oh, ol, ho), iters=10)
            res[f"lstm_cell_tf32x3_engine{eng}_M{m2}_K{kx + h2}_H{h2}"] = {
                "ms": ms, "tflops_fp32_equiv": 2.0 * m2 * (kx + h2) * 4 * h2 / ms / 1e9}
        ops.set_gemm_engine(int(os.environ.get("SE_GEMM_ENGINE", "5")))
    xp = torch.randn(B, T, 4 * H, generator=g).to(dev)
    whh = (torch.randn(H // 8, H, 32, generator=g) / 32).to(dev)
    hs = torch.empty(B, T, H, device=dev)
    for eng, nm in ((0, "fma"), (1, "mma"), (2, "tcgen05")):
        ops.set_lstm_engine(eng)
        ms = timeit(lambda: ops.lstm_seq(xp, whh, H, out=hs), iters=3, warm=1)
        res[f"lstm_seq_{nm}_B64_T401_H1024"] = {"ms": ms, "us_per_step": 1e3 * ms / T,
                                                "tflops": 2.0 * B * (T - 1) * 4 * H * H / ms / 1e9}
        if eng == 0:
            ref_h = hs.clone()
        else:
            res[f"lstm_seq_{nm}_B64_T401_H1024"]["max_abs_diff_vs_fma"] = (hs - ref_h).abs().max().item()
    # H = 128 recurrences: DPCRN inter-chunk LSTM (4 frequency positions x 64 clips, shared weights) and DCCRN's four real
    # LSTMs (32 clips): slice kernel with a device-wide barrier per step (engine 0) vs the sequence-parallel kernel (3)
    for nm, ng, bb, tt, shared in (("dpcrn_inter", 4, 64, 401, True), ("dccrn_clstm", 4, 32, 501, False)):
        xp2 = torch.randn(bb, tt, ng * 512, generator=g).to(dev)
        w2 = (torch.randn(*(() if shared else (ng,)), 16, 128, 32, generator=g) / 11.3).to(dev)
        o2 = torch.empty(bb, tt, ng * 128, device=dev)
        for eng in (0, 3):
            ops.set_lstm_engine(eng)
            ms = timeit(lambda: ops.lstm_seq_multi(xp2, w2, 128, ng, o2), iters=3, warm=1)
            res[f"lstm_seq_multi_{nm}_engine{eng}"] = {"ms": ms, "us_per_step": 1e3 * ms / tt}
            if eng == 0:
                ref2 = o2.clone()
            else:
                res[f"lstm_seq_multi_{nm}_engine{eng}"]["max_abs_diff_vs_engine0"] = (o2 - ref2).abs().max().item()
    ops.set_lstm_engine(3)
    # conv layers of CRN
    for (fin, c0, c1, co, kind) in [(9, 128, 0, 256, "conv"), (19, 64, 0, 128, "conv"), (4, 256, 256, 128, "deconv"),
                                    (9, 128, 128, 64, "deconv")]:
        s0 = torch.randn(B, T, fin, c0, generator=g).to(dev)
        s1 = torch.randn(B, T, fin, c1, generator=g).to(dev) if c1 else None
        ct = c0 + c1
        if kind == "conv":
            fo = (fin - 3) // 2 + 1
            w = torch.randn(6 * ct, co, generator=g).to(dev)
            o = torch.empty(B, T, fo, co, device=dev)
            fn = lambda: ops.conv_gemm(s0, s1, B, T, fin, fo, packing.CONV23_TAPS, 2, w, None, co, "elu", o, fo)
            flc = 2.0 * B * T * fo * 6 * ct * co
        else:
            fo = 2 * fin + 1
            we = torch.randn(4 * ct, co, generator=g).to(dev)
            wo = torch.randn(2 * ct, co, generator=g).to(dev)
            o = torch.empty(B, T, fo, co, device=dev)
            def fn():
                ops.conv_gemm(s0, s1, B, T, fin, fin + 1, packing.DECONV_EVEN_TAPS, 1, we, None, co, "elu", o, fo, 0, 2)
                ops.conv_gemm(s0, s1, B, T, fin, fin, packing.DECONV_ODD_TAPS, 1, wo, None, co, "elu", o, fo, 1, 2)
            flc = 2.0 * B * T * fin * 6 * ct * co
        ms = timeit(fn)
        res[f"{kind}_F{fin}_C{ct}_to_{co}"] = {"ms": ms, "tflops": flc / ms / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
