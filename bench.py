#!/usr/bin/env python
"""bench.py -- enhanced STFT frames/s of the CRN decode path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one pass of the whole decode hot path over one batch of 64 synthetic 4 s / 16 kHz
clips per GPU:  RMS scale -> STFT (320/320/160, the geometry the CRN checkpoints are wired to,
SURVEY.md section 0.1) -> crn_net.forward -> magnitude x noisy phase -> iSTFT -> 1/c
(CRN/crn_decode.py:38-57).  Weak scaling: every rank decodes its own 64 clips; for N > 1 the
enhanced waveforms are all-gathered (the one collective of the path).

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the
same metric from pinned host memory to host memory through the public API.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16000
CLIP_SECONDS = 4
N_SAMPLES = FS * CLIP_SECONDS
BATCH_PER_GPU = 64
N_FFT, WIN, HOP = 320, 320, 160
T_FRAMES = 1 + N_SAMPLES // HOP          # 401
CKPT = os.path.join(ROOT, "checkpoints", "_ref", "CRN__wsj0_si84_300h_crn_noncprs_model.pth")
N_INPUT_SETS = 9                          # 9 x 16.4 MB of inputs > 126 MB L2
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # nominal CUDA-core fp32 (FFMA) peak


def load_weights():
    from oracle import synth, templates
    if os.path.exists(CKPT):
        return torch.load(CKPT, map_location="cpu"), "shipped checkpoint wsj0_si84_300h_crn_noncprs_model.pth"
    return synth.synthetic_state_dict(templates.crn_template(), seed=0), "random-init (seeded synthetic)"


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


def pick_cpu_threads(sd, clip):
    """torch CPU ops do not scale to every core of a big host (128 threads were 100x slower than
    8 on the B200 box): time one clip at a few thread counts and keep the fastest."""
    from oracle import decode as odecode
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = None, None
    x = clip.astype(np.float64)
    for c in cands:
        torch.set_num_threads(c)
        odecode.enhance_crn(sd, x)
        t0 = time.perf_counter()
        odecode.enhance_crn(sd, x)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline(sd, clips, repeats=1):
    """The reference's CPU path (oracle port of CRN/crn_decode.py), one clip at a time, best of
    `repeats` passes.  Returns (frames/s, ms/clip, threads used)."""
    from oracle import decode as odecode
    threads = pick_cpu_threads(sd, clips[0])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for x in clips:
            odecode.enhance_crn(sd, x.astype(np.float64))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(clips) * T_FRAMES / best, 1e3 * best / len(clips), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    sd, wdesc = load_weights()
    nclip = 4
    clips = synth.noisy_batch(nclip, N_SAMPLES)
    threads = pick_cpu_threads(sd, clips[0])
    from oracle import decode as odecode
    for _ in range(max(1, args.warmup)):
        odecode.enhance_crn(sd, clips[0].astype(np.float64))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for x in clips:
            odecode.enhance_crn(sd, x.astype(np.float64))
    dt = time.perf_counter() - t0
    ms = 1e3 * dt / args.steps
    fps = nclip * T_FRAMES * args.steps / dt
    sample = f"{nclip} clips x 4 s per step, batch 1 loop as in crn_decode.py:37, torch CPU ops, {threads} of {os.cpu_count()} host threads (fastest setting)"
    line = {
        "impl": "reference", "metric": "enhanced STFT frames/s (CRN decode, 16 kHz, 320-FFT/160-hop, 4 s clips)",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "CRN/crn_decode.py magnitude mapping, 4 s clips, 320/320/160 STFT", "weights": wdesc},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    sd, wdesc = load_weights()
    model = se_b200.crn_net()
    model.load_state_dict(sd)
    model.eval().cuda()

    # synthetic noisy speech: 64 distinct clips per rank; N_INPUT_SETS rotated copies (rolled in
    # time, so every buffer is different data) keep the step's input out of L2 between steps
    base = torch.from_numpy(synth.noisy_batch(BATCH_PER_GPU, N_SAMPLES, first_index=rank * BATCH_PER_GPU))
    host_sets = [torch.roll(base, shifts=997 * i, dims=1).contiguous().pin_memory() for i in range(N_INPUT_SETS)]
    dev_sets = [h.to(dev) for h in host_sets]
    batch_total = BATCH_PER_GPU * world

    def step(i):
        y = se_b200.decode.enhance_crn(model, dev_sets[i % N_INPUT_SETS])
        if world > 1:
            y = se_b200.shard.gather_waveforms(y, batch_total)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- timed region: device-resident inputs ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = se_b200.ops.launch_count()
    se_b200.ops.start_recording()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    per_op = se_b200.ops.stop_recording()
    launches = se_b200.ops.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    frames_step = batch_total * T_FRAMES
    value = frames_step / (ms_step * 1e-3)

    # ---- e2e: pinned host -> device -> enhance -> pinned host, through the public API ------------------
    # decode.enhance_host_stream: every step uploads its own 64 clips from pinned host memory and downloads its enhanced
    # clips inside the timed region; the copies of neighbouring steps overlap the decode loop (copy streams).
    def host_batches(n):
        for i in range(n):
            yield host_sets[i % N_INPUT_SETS]

    def e2e_run(n):
        got = 0
        for y in se_b200.decode.enhance_host_stream(model, host_batches(n), se_b200.decode.enhance_crn):
            got += 1                      # y: pinned host tensor [64, N] with this step's enhanced clips
        assert got == n

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)                   # returns after the last download has completed (event sync on the host)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = frames_step / (e2e_ms / args.steps * 1e-3)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fl = crn_flops_per_frame()
        # dominant kernel by measured device time inside the timed region
        dom = max(per_op.items(), key=lambda kv: kv[1][1]) if per_op else ("none", (1, 1.0))
        share = {k: round(v[1] / ms_total, 4) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][1])}
        rec_cnt, rec_ms = per_op.get("lstm_seq", (1, 1e9))
        rec_flops_per_launch = 2.0 * BATCH_PER_GPU * (T_FRAMES - 1) * 4096 * 1024   # 2*B*(T-1)*4H*H
        rec_avg_ms = rec_ms / max(rec_cnt, 1)
        achieved = rec_flops_per_launch / (rec_avg_ms * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        traffic = None
        kname = "lstm_seq_tc_kernel"
        # dram__bytes_read.sum + dram__bytes_write.sum (MB per launch) of the committed `ncu --set full` capture
        for ncu_json in ("ncu_lstm_tc_r01b.json", "ncu_full_r01b.json", "ncu_full_r01.json"):
            path = os.path.join(ROOT, "profiles", ncu_json)
            if traffic is None and os.path.exists(path):
                for k in json.load(open(path))["kernels"]:
                    if kname in k["kernel"]:
                        traffic = (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) * 1e6
                        break
        # the recurrence computes every product three times (3xTF32: hi*hi, hi*lo, lo*hi) on the TF32 tensor pipe
        tf32_peak = peak / 2.0
        roofline = {
            "kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": f"{peak_src} bf16 dense, sustained",
            "pipe": "tcgen05 kind::tf32, 3xTF32 split (3 tensor flops per algorithmic flop; single-pass TF32 breaks the "
                    "1e-4 gate); W_hh hi part in tensor memory, K split over a 4-CTA cluster.  The step is a latency "
                    "chain (device-wide barrier -> TMA -> MMA -> DSMEM reduce -> cell -> publish), not pipe-bound: "
                    "profiles/lstm_tc_phases_r01.json",
            "pipe_peak": tf32_peak / 3.0, "frac_of_pipe": achieved / (tf32_peak / 3.0),
            "algorithmic_flops_per_launch": rec_flops_per_launch,
            "avg_launch_ms": rec_avg_ms, "dominant_by_time": dom[0], "time_share": share,
            "whole_step_tflops": fl["total"] * BATCH_PER_GPU * T_FRAMES / (ms_step * 1e-3) / 1e12,
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # the CPU arm is timed on rank 0 at N=1 only
            clips = synth.noisy_batch(8, N_SAMPLES)
            fps, ms_clip, threads = cpu_baseline(sd, list(clips), repeats=3)
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": f"8 clips x 4 s, batch 1 loop, best of 3 passes ({ms_clip:.1f} ms/clip), oracle port of "
                             f"CRN/crn_decode.py on torch CPU ops, fastest of the tried thread counts "
                             f"({threads} of {os.cpu_count()})"}
        line = {
            "metric": "enhanced STFT frames/s (CRN decode, 16 kHz, 320-FFT/160-hop, 4 s clips)",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CRN/crn_decode.py magnitude mapping, batch=64 x 4 s clips per GPU, "
                                   "320/320/160 STFT (native CRN geometry; SURVEY.md 0.1)",
                       "global_batch": batch_total, "frames_per_clip": T_FRAMES, "weights": wdesc,
                       "parallelism": f"dp{world} (utterance shards, final all_gather)" if world > 1 else "single GPU",
                       "l2": f"inputs rotate over {N_INPUT_SETS} batches (147 MB > 126 MB L2); per-step activations "
                             "~1.5 GB stream through L2"},
            "rtf": (ms_step * 1e-3) / (batch_total * CLIP_SECONDS),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "h2d_bytes_per_step": BATCH_PER_GPU * N_SAMPLES * 4, "d2h_bytes_per_step": BATCH_PER_GPU * N_SAMPLES * 4,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "se_b200.decode.enhance_host_stream (pinned host in -> pinned host out, copies of step i+1 / "
                           "i-1 overlap the decode loop of step i)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
