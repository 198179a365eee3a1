#!/usr/bin/env python
"""bench.py -- enhanced STFT frames/s of the decode path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (headline: configs[1], CRN)
    python bench.py --config dccrn|fullsubnet|uformer ...    # another BASELINE config as the headline
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one pass of the whole decode hot path over one batch of synthetic 16 kHz clips per GPU:
RMS scale -> STFT -> model.forward -> recombination -> iSTFT -> 1/c (CRN/crn_decode.py:38-57 and its siblings).
The headline (configs[1]) is CRN, 64 x 4 s clips per GPU at the 320/320/160 geometry its checkpoints are wired
to.  Weak scaling: every rank decodes its own clips; for N > 1 the enhanced waveforms are
gathered on rank 0 (the one collective of the path) on a side stream behind the next
step's compute.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same metric from
pinned host memory to host memory through the public API.  The line also carries
  * `configs`: BASELINE configs[2..4] (DCCRN 256 clips / 8 GPUs, FullSubNet 128 x 10 s / 4, Uformer 512 / 8) at their
    per-GPU shard (32 / 32 / 64 clips per rank) on the launched N, each with frames/s, ms/step, dominant kernel;
  * `dsp_only`: STFT -> identity -> iSTFT at the literal 512-FFT / 256-hop of the metric string (and the other three
    geometries) with the achieved fraction of the measured HBM peak;
  * `gpu_eager_baseline`: the oracle network on the SAME GPU through ATen / cuDNN (what the reference's own scripts
    do with the model, CRN/crn_decode.py:22,47-48), batch 1 with host DSP and batch 64 -- context only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16000
N_INPUT_SETS_BYTES = 150e6                # rotate inputs over more than the 126 MB L2
CKPT_DIR = os.path.join(ROOT, "checkpoints", "_ref")
FSN_ARGS = dict(num_freqs=257, look_ahead=2, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=15,
                fb_output_activate_function="ReLU", sb_output_activate_function=None, fb_model_hidden_size=512,
                sb_model_hidden_size=384)

# BASELINE.json configs[1..4] (SURVEY.md section 8(d) "Configs as concrete inputs").  per_rank = the per-GPU shard of
# the split BASELINE names; flops_frame = algorithmic FLOPs per STFT frame.
CONFIGS = {
    "crn": dict(idx=1, label="CRN/crn_decode.py IRM / magnitude mapping", ckpt="CRN__wsj0_si84_300h_crn_noncprs_model.pth",
                seconds=4, geom=(320, 320, 160), per_rank=64, split=(64, 1), kw=dict(p=1.0), flops_frame=43.1e6,
                oracle="enhance_crn"),
    "dccrn": dict(idx=2, label="DCCRN/dccrn_decode.py complex-ratio mask (DCCRN-E, cLSTM)",
                  ckpt="DCCRN__wsj0_si84_300h_dccrn_cprs_model.pth", seconds=4, geom=(512, 512, 128), per_rank=32,
                  split=(256, 8), kw=dict(p=0.5), flops_frame=106.7e6, oracle="enhance_dccrn"),
    "fullsubnet": dict(idx=3, label="FullSubNet/fullsubnet_sa_decode.py",
                       ckpt="FullSubNet__wsj0_si84_300h_fullsubnet_cprs_model_512_256.pth", seconds=10,
                       geom=(512, 512, 256), per_rank=32, split=(128, 4), kw=dict(p=0.5), flops_frame=943e6,
                       oracle="enhance_fullsubnet"),
    "uformer": dict(idx=4, label="Uformer/uformer_decode.py dual-path complex conformer",
                    ckpt="Uformer__wsj0_si84_300h_uformer_noncprs_model.pth", seconds=4, geom=(512, 400, 160), per_rank=64,
                    split=(512, 8), kw=dict(), flops_frame=68.7e6, oracle="enhance_uformer"),
}


def frames_per_clip(cfg):
    return 1 + cfg["seconds"] * FS // cfg["geom"][2]


def load_weights(name):
    """The shipped checkpoint the matching decode script loads.  There is NO random-init fallback: a bench line on
    other weights than the ones SURVEY.md section 8(d) names would not be the configuration BASELINE.json quotes."""
    path = os.path.join(CKPT_DIR, CONFIGS[name]["ckpt"])
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run `python -m oracle.fetch_checkpoints` in the build container "
                                "(the copy travels to the GPU box with the snapshot)")
    return torch.load(path, map_location="cpu"), "shipped checkpoint " + CONFIGS[name]["ckpt"].split("__", 1)[1]


def build_model(name, sd):
    import se_b200
    if name == "crn":
        model = se_b200.crn_net()
    elif name == "dccrn":
        model = se_b200.DCCRN(rnn_units=256, masking_mode='E', use_clstm=True, kernel_num=[32, 64, 128, 256, 256, 256])
    elif name == "fullsubnet":
        model = se_b200.fullsubnet.Model(**FSN_ARGS)
    else:
        model = se_b200.Uformer()
    model.load_state_dict(sd)
    model.eval().cuda()
    return model, getattr(se_b200.decode, "enhance_" + name)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def crn_flops_per_frame():
    """Algorithmic FLOPs (2*MAC) per STFT frame, from the layer shapes."""
    enc = [(16, 80, 1), (32, 39, 16), (64, 19, 32), (128, 9, 64), (256, 4, 128)]     # Cout, Fout, Cin
    dec = [(4, 512, 128), (9, 256, 64), (19, 128, 32), (39, 64, 16), (80, 32, 1)]    # Fin, Cin, Cout
    conv = sum(co * fo * ci * 6 for co, fo, ci in enc) + sum(fi * ci * co * 6 for fi, ci, co in dec)
    lstm_in = 2 * 4096 * 1024
    lstm_rec = 2 * 4096 * 1024
    return {"conv": 2 * conv, "lstm_in": 2 * lstm_in, "lstm_rec": 2 * lstm_rec,
            "total": 2 * (conv + lstm_in + lstm_rec)}


# ---- CPU arm: the oracle port of the matching decode script, one clip at a time --------------------------------
def oracle_enhance(name):
    from oracle import decode as odecode
    return getattr(odecode, CONFIGS[name]["oracle"])


def pick_cpu_threads(fn, sd, clip, kw):
    """torch CPU ops do not scale to every core of a big host (128 threads were 100x slower than 8 on the B200 box):
    time one clip at a few thread counts and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = None, None
    x = clip.astype(np.float64)
    for c in cands:
        torch.set_num_threads(c)
        fn(sd, x, **kw)
        t0 = time.perf_counter()
        fn(sd, x, **kw)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline(name, sd, clips, repeats=1):
    """Returns (frames/s, ms/clip, threads used)."""
    fn, kw = oracle_enhance(name), CONFIGS[name]["kw"]
    threads = pick_cpu_threads(fn, sd, clips[0], kw)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for x in clips:
            fn(sd, x.astype(np.float64), **kw)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(clips) * frames_per_clip(CONFIGS[name]) / best, 1e3 * best / len(clips), threads


def metric_name(name):
    cfg = CONFIGS[name]
    n_fft, _, hop = cfg["geom"]
    return (f"enhanced STFT frames/s ({name.upper() if name != 'fullsubnet' else 'FullSubNet'} decode, 16 kHz, "
            f"{n_fft}-FFT/{hop}-hop, {cfg['seconds']} s clips)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    name = args.config
    cfg = CONFIGS[name]
    sd, wdesc = load_weights(name)
    nclip = 4 if name == "crn" else 1
    clips = synth.noisy_batch(nclip, cfg["seconds"] * FS)
    fn, kw = oracle_enhance(name), cfg["kw"]
    threads = pick_cpu_threads(fn, sd, clips[0], kw)
    for _ in range(max(1, args.warmup)):
        fn(sd, clips[0].astype(np.float64), **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for x in clips:
            fn(sd, x.astype(np.float64), **kw)
    dt = time.perf_counter() - t0
    ms = 1e3 * dt / args.steps
    fps = nclip * frames_per_clip(cfg) * args.steps / dt
    sample = (f"{nclip} clip(s) x {cfg['seconds']} s per step, batch 1 loop as in crn_decode.py:37, torch CPU ops, {threads} of "
              f"{os.cpu_count()} host threads (fastest setting)")
    line = {
        "impl": "reference", "metric": metric_name(name),
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{cfg['label']}, {cfg['seconds']} s clips, {'/'.join(map(str, cfg['geom']))} STFT",
                   "weights": wdesc},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything that libraries print on fd 1 (NCCL's version banner, ...) goes to stderr; the ONE JSON line is written
    to the saved descriptor by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


class Runner:
    """One BASELINE config on this rank: model, rotating device-resident inputs, the step function (decode + pipelined
    gather to rank 0) and the timed loop (barrier + synchronize on both sides, CUDA events, max over ranks)."""

    def __init__(self, name, rank, world, dev):
        import se_b200
        from oracle import synth
        self.se = se_b200
        self.name, self.cfg, self.rank, self.world, self.dev = name, CONFIGS[name], rank, world, dev
        self.sd, self.wdesc = load_weights(name)
        self.model, self.enhance = build_model(name, self.sd)
        self.per_rank = self.cfg["per_rank"]
        self.n = self.cfg["seconds"] * FS
        self.frames = frames_per_clip(self.cfg)
        # synthetic noisy speech: distinct clips per rank; rotated copies (rolled in time, so every buffer is different
        # data) keep a step's input out of L2 between steps
        nbase = min(self.per_rank, 16)
        base = torch.from_numpy(synth.noisy_batch(nbase, self.n, first_index=rank * self.per_rank))
        base = torch.cat([torch.roll(base, 131 * k, dims=1) for k in range(self.per_rank // nbase)], 0)
        self.nsets = max(2, int(N_INPUT_SETS_BYTES // (self.per_rank * self.n * 4)) + 1)
        self.host_sets = [torch.roll(base, shifts=997 * i, dims=1).contiguous().pin_memory() for i in range(self.nsets)]
        self.dev_sets = [h.to(dev) for h in self.host_sets]
        self.pipe = se_b200.shard.GatherPipeline(self.per_rank * world, dst=0, depth=2) if world > 1 else None

    def step(self, i):
        y = self.enhance(self.model, self.dev_sets[i % self.nsets], **self.cfg["kw"])
        if self.pipe is not None:
            self.pipe.submit(y)
        return y

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            tt = torch.tensor([ms], device=self.dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    def timed(self, steps, warmup, record=True):
        for i in range(warmup):
            self.step(i)
        if self.pipe is not None:
            self.pipe.drain()
        self.barrier()
        launches0 = self.se.ops.launch_count()
        if record:
            self.se.ops.start_recording()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for i in range(steps):
            self.step(i)
        if self.pipe is not None:
            self.pipe.drain()             # the compute stream waits for the last gathers: they are inside the timed region
        ev1.record()
        self.barrier()
        per_op = self.se.ops.stop_recording() if record else {}
        launches = self.se.ops.launch_count() - launches0
        ms_total = self.max_over_ranks(ev0.elapsed_time(ev1))
        return ms_total, per_op, launches

    def e2e(self, steps):
        """pinned host -> device -> enhance -> pinned host through decode.enhance_host_stream: every step uploads its own
        clips and downloads its enhanced clips inside the timed region; copies of neighbouring steps overlap the decode."""
        def host_batches(n):
            for i in range(n):
                yield self.host_sets[i % self.nsets]

        def run(n):
            got = 0
            for _y in self.se.decode.enhance_host_stream(self.model, host_batches(n), self.enhance, **self.cfg["kw"]):
                got += 1
            assert got == n

        run(2)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)                        # returns after the last download has completed (event sync on the host)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def release(self):
        self.model = self.dev_sets = self.host_sets = None
        torch.cuda.empty_cache()


def se_f16_pairs():
    import se_b200
    return bool(se_b200.lstm_engine.USE_F16_PAIRS and se_b200.lstm_engine.USE_TENSOR_CORES)


def step_roofline(cfg, frames_step, ms_step, peaks, peak_src):
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    tf = cfg["flops_frame"] * frames_step / (ms_step * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
            "what": "whole step: algorithmic FLOPs per frame (SURVEY.md 8(d)) x frames/s per GPU",
            "peak_source": f"{peak_src} bf16 dense, sustained"}


def lstm_roofline(per_op, ms_total, ms_step, peaks, peak_src, per_rank, frames):
    """Dominant kernel of the CRN step: the H = 1024 recurrence (2 launches per step)."""
    fl = crn_flops_per_frame()
    dom = max(per_op.items(), key=lambda kv: kv[1][1]) if per_op else ("none", (1, 1.0))
    share = {k: round(v[1] / ms_total, 4) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][1])}
    rec_cnt, rec_ms = per_op.get("lstm_seq", (1, 1e9))
    rec_flops_per_launch = 2.0 * per_rank * (frames - 1) * 4096 * 1024   # 2*B*(T-1)*4H*H
    rec_avg_ms = rec_ms / max(rec_cnt, 1)
    achieved = rec_flops_per_launch / (rec_avg_ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    kname = "lstm_seq_f16_kernel"
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")     # written by tools/ncu_summarize.py --traffic
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(kname)
        if rec:
            traffic, traffic_src = rec["dram_bytes_per_launch"], "committed ncu --set full capture: " + rec["source"]
    return {
        "kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": f"{peak_src} bf16 dense, sustained",
        "pipe": "tcgen05 kind::f16 on fp16 hi/lo operand pairs (3 tensor flops per algorithmic flop: hi*hi, hi*lo, lo*hi; a "
                "single-pass half/TF32 product breaks the 1e-4 gate); W_hh hi part in tensor memory, K split over a 4-CTA "
                "cluster, state exchanged as self-validating tagged words through L2.  The step is a latency chain "
                "(publish -> poll -> load -> MMA -> DSMEM reduce -> cell), not pipe-bound: profiles/lstm_f16_phases_*_r02.json",
        "pipe_peak": peak / 3.0, "frac_of_pipe": achieved / (peak / 3.0),
        "algorithmic_flops_per_launch": rec_flops_per_launch,
        "avg_launch_ms": rec_avg_ms, "dominant_by_time": dom[0], "time_share": share,
        "whole_step_tflops": fl["total"] * per_rank * frames / (ms_step * 1e-3) / 1e12,
    }


def dsp_only(dev, peaks):
    """STFT -> identity -> iSTFT on 64 x 4 s clips per geometry: per-kernel time, frames/s and the
    fraction of the measured HBM peak its algorithmic bytes reach.  512/512/256 is the metric string's geometry."""
    from oracle import synth
    peak = float(peaks["hbm_gbs"])
    n = 4 * FS
    base = torch.from_numpy(synth.noisy_batch(16, n))
    result = {"workload": "rms -> STFT -> identity -> iSTFT on 4 s clips, device-resident, inputs rotated over > L2; algorithmic "
                          "bytes per frame: 4 hop + 8 F each way.  batch 64 = the headline batch (a 30-100 us kernel: launch and "
                          "wave effects count); batch 512 = BASELINE configs[4]'s batch", "hbm_peak_GBps": peak}
    for bsz in (64, 512):
        result[f"batch{bsz}"] = _dsp_only_batch(dev, peak, base, n, bsz)
    result["geometries"] = result["batch64"]          # the key round-1 readers looked at
    return result


def _dsp_only_batch(dev, peak, base, n, bsz):
    from se_b200 import ops
    from se_b200._lib import ISTFT_SPEC
    wav0 = torch.cat([base] * (bsz // 16), 0).to(dev)
    nsets = max(2, int(N_INPUT_SETS_BYTES // (bsz * n * 4)) + 1)
    wavs = [torch.roll(wav0, 997 * i, dims=1).contiguous() for i in range(nsets)]
    out = {}

    def best_of(fns, iters=4):
        for fn in fns:
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for fn in fns:
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / len(fns))
        return best

    for gname, (n_fft, win, hop) in {"512/512/256": (512, 512, 256), "320/320/160": (320, 320, 160),
                                     "512/512/128": (512, 512, 128), "512/400/160": (512, 400, 160)}.items():
        t, f = 1 + n // hop, n_fft // 2 + 1
        c, inv_c = ops.rms_scale(wavs[0])
        specs = [torch.empty(bsz, t, f, 2, device=dev) for _ in range(nsets)]
        outs = [torch.empty(bsz, n, device=dev) for _ in range(nsets)]
        ms_s = best_of([(lambda w=w, s=s: ops.stft(w, c, n_fft, win, hop, re=s[..., 0], im=s[..., 1]))
                        for w, s in zip(wavs, specs)])
        ms_i = best_of([(lambda s=s, o=o: ops.istft(ISTFT_SPEC, s[..., 0], s[..., 1], None, None, n_fft, win, hop, o, n,
                                                    out_scale=inv_c)) for s, o in zip(specs, outs)])
        err = (outs[0] - wavs[0]).abs().max().item()
        frames = bsz * t
        b_stft = frames * (4 * hop + 8 * f)          # audio in + complex spectrum out
        b_istft = frames * (8 * f + 4 * hop)         # complex spectrum in + audio out
        out[gname] = {"frames_per_s": frames / ((ms_s + ms_i) * 1e-3), "roundtrip_max_abs_err": err,
                      "stft_us": 1e3 * ms_s, "stft_GBps": b_stft / ms_s / 1e6, "stft_frac_of_hbm_peak": b_stft / ms_s / 1e6 / peak,
                      "istft_us": 1e3 * ms_i, "istft_GBps": b_istft / ms_i / 1e6,
                      "istft_frac_of_hbm_peak": b_istft / ms_i / 1e6 / peak}
        del specs, outs
    del wavs
    torch.cuda.empty_cache()
    return out


def gpu_eager_baseline(sd, dev):
    """The reference's own division of labour on THIS box: host DSP (numpy float64, as librosa) + the network on the GPU
    through ATen / cuDNN (CRN/crn_decode.py:22,47-48), one clip at a time; and the same eager network at batch 64 with
    device DSP left out.  TF32 is off (fp32 parity).  Context for the headline, never part of it."""
    from oracle import decode as odecode
    from oracle import nets, synth
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.to(dev) for k, v in sd.items()}
    t = 401
    clips = synth.noisy_batch(4, 4 * FS)

    def fwd(_sd, feat):
        return nets.crn_forward(sdc, feat.to(dev)).cpu()

    odecode._mag_mapping_320(sd, fwd, clips[0].astype(np.float64), 1.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for x in clips:
        odecode._mag_mapping_320(sd, fwd, x.astype(np.float64), 1.0)
    torch.cuda.synchronize()
    b1 = (time.perf_counter() - t0) / len(clips)
    mag = torch.rand(64, t, 161, device=dev)
    with torch.no_grad():
        nets.crn_forward(sdc, mag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            nets.crn_forward(sdc, mag)
        e1.record()
        torch.cuda.synchronize()
    b64 = e0.elapsed_time(e1) / 3
    return {"what": "oracle.nets.crn_forward on cuda (ATen / cuDNN LSTM + conv, TF32 off)",
            "batch1_decode_loop": {"ms_per_clip": 1e3 * b1, "frames_per_s": t / b1,
                                   "note": "host numpy STFT/iSTFT + H2D/D2H per clip, as crn_decode.py does"},
            "batch64_network_only": {"ms_per_batch": b64, "frames_per_s": 64 * t / (b64 * 1e-3),
                                     "note": "forward only, no DSP; the reference never batches"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="crn", choices=list(CONFIGS), help="headline workload (default: configs[1], CRN)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs / dsp_only / gpu_eager_baseline records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    quiet_stdout()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist

    import se_b200
    from oracle import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    peaks, peak_src = measured_peaks()

    # ---- headline ---------------------------------------------------------------------------------------------
    name = args.config
    cfg = CONFIGS[name]
    run = Runner(name, rank, world, dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, per_op, launches = run.timed(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    frames_step = run.per_rank * world * run.frames
    value = frames_step / (ms_step * 1e-3)
    e2e_ms = run.e2e(args.steps)
    e2e_value = frames_step / (e2e_ms / args.steps * 1e-3)
    sd_headline = run.sd
    per_rank, frames, n_samples, nsets = run.per_rank, run.frames, run.n, run.nsets
    wdesc = run.wdesc
    run.release()

    # ---- the other BASELINE configs at their per-GPU shard, on the launched N ------------------------------------
    extra = []
    if not args.no_extras:
        for other in CONFIGS:
            if other == name:
                continue
            ocfg = CONFIGS[other]
            try:
                r = Runner(other, rank, world, dev)
            except FileNotFoundError as e:
                extra.append({"config": other, "unavailable": str(e)})
                continue
            osteps = max(3, min(args.steps, 6 if other == "fullsubnet" else 10))
            oms_total, oper_op, olaunch = r.timed(osteps, 3)
            oms = oms_total / osteps
            ofr = r.per_rank * world * r.frames
            tot = sum(v[1] for v in oper_op.values()) or 1.0
            top = sorted(oper_op.items(), key=lambda kv: -kv[1][1])[:5]
            total, ngpu = ocfg["split"]
            extra.append({
                "config": other, "baseline_config": f"configs[{ocfg['idx']}]: {ocfg['label']}, batch={total} x {ocfg['seconds']} s "
                                                    f"clips sharded across {ngpu} x B200",
                "clips_per_gpu": r.per_rank, "global_batch": r.per_rank * world, "n_gpus": world,
                "is_baseline_split": world == ngpu, "frames_per_clip": r.frames,
                "value": ofr / (oms * 1e-3), "unit": "frames/s", "ms_per_step": oms, "steps": osteps, "warmup": 3,
                "rtf": (oms * 1e-3) / (r.per_rank * world * ocfg["seconds"]), "weights": r.wdesc,
                "gpu_launches_per_step": int(olaunch // osteps),
                "dominant_kernel": top[0][0] if top else None,
                "time_share": {k: round(v[1] / tot, 3) for k, v in top},
                "roofline": step_roofline(ocfg, r.per_rank * r.frames, oms, peaks, peak_src),
            })
            r.release()

    if rank == 0:
        if name == "crn":
            roofline = lstm_roofline(per_op, ms_total, ms_step, peaks, peak_src, per_rank, frames)
        else:
            roofline = step_roofline(cfg, per_rank * frames, ms_step, peaks, peak_src)
            tot = sum(v[1] for v in per_op.values()) or 1.0
            roofline["time_share"] = {k: round(v[1] / tot, 4) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][1])}
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # the CPU arm is timed on rank 0 at N=1 only
            nclips = 8 if name == "crn" else 2
            clips = synth.noisy_batch(nclips, n_samples)
            fps, ms_clip, threads = cpu_baseline(name, sd_headline, list(clips), repeats=3 if name == "crn" else 1)
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": f"{nclips} clips x {cfg['seconds']} s, batch 1 loop, best pass ({ms_clip:.1f} ms/clip), oracle port "
                             f"of the decode script on torch CPU ops, fastest of the tried thread counts "
                             f"({threads} of {os.cpu_count()})"}
        line = {
            "metric": metric_name(name),
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['label']}, batch={per_rank} x {cfg['seconds']} s clips per GPU, "
                                   f"{'/'.join(map(str, cfg['geom']))} STFT (the geometry the checkpoint is wired to; SURVEY.md 0.1)",
                       "global_batch": per_rank * world, "frames_per_clip": frames, "weights": wdesc,
                       "parallelism": (f"dp{world} (utterance shards; gather of the enhanced clips to rank 0 on a side "
                                       "stream behind the next step)") if world > 1 else "single GPU",
                       "l2": f"inputs rotate over {nsets} batches ({nsets * per_rank * n_samples * 4 / 1e6:.0f} MB > 126 MB L2); "
                             "per-step activations stream through L2",
                       "arithmetic": ("fp32 results; dense contractions as three tensor-core products of fp16 hi/lo operand "
                                      "pairs (tcgen05 kind::f16, fp32 accumulation in tensor memory): the 22-bit products of "
                                      "3xTF32, within the 1e-4 RMS gate (a single half / TF32 pass is not)")
                                     if se_f16_pairs() else
                                     "fp32 results; dense contractions as 3xTF32 on tcgen05 (SE_F16_PAIRS=0)"},
            "rtf": (ms_step * 1e-3) / (per_rank * world * cfg["seconds"]),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "h2d_bytes_per_step": per_rank * n_samples * 4, "d2h_bytes_per_step": per_rank * n_samples * 4,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "se_b200.decode.enhance_host_stream (pinned host in -> pinned host out, copies of step i+1 / "
                           "i-1 overlap the decode loop of step i)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "configs": extra,
        }
        if not args.no_extras:
            line["dsp_only"] = dsp_only(dev, peaks)
            if world == 1 and name == "crn":
                try:
                    line["gpu_eager_baseline"] = gpu_eager_baseline(sd_headline, dev)
                except Exception as e:           # context only: never fail the bench line over it
                    line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
