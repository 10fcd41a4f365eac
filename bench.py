#!/usr/bin/env python
"""bench.py -- clips/sec for 30 s audio -> beats (final0-shaped checkpoint) on N B200 GPUs.

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path (default)
    python bench.py --impl reference --gpus 1 --steps K ...  # the reference algorithm on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of `--batch` synthetic 30 s clips
per GPU (BASELINE config 2: Audio2Frames final0, batch 64, bf16; the minimal peak picker is
included so that the step is audio -> beats).  Prints ONE JSON line (rank 0).

value  : device-resident throughput (audio already in HBM; CUDA events; max over ranks).
e2e    : the same step through the public API with a pinned HOST buffer: H2D copy of the
         audio, all kernels, D2H of the beat/downbeat timestamp arrays.
roofline: the dominant kernel (time-direction flash attention, tcgen05) -- algorithmic
         FLOPs / CUDA-event time measured live in the timed region (bt_profile_*).
cpu_baseline: the CPU oracle port of the reference timed on this box's host cores.
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

# stdout carries exactly ONE line, the JSON result: everything else that writes to file descriptor 1 -- Python
# prints, the NCCL version banner (printed from C at NCCL_DEBUG >= VERSION), library chatter -- goes to stderr.
_REAL_STDOUT = os.dup(1)
sys.stdout.flush()
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

SR = 22050
METRIC = "clips/sec (30 s audio->beats, final0)"
CACHE = os.environ.get("BT_TEST_CACHE", "/tmp/beat_this_b200_cache")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def synth_batch(n_clips: int, seconds: float, seed0: int):
    from beat_this_b200 import synthetic

    # a handful of distinct seeded clips, tiled: generation of 64 x 30 s clips in numpy is slow and
    # the kernels are data independent (no early exit anywhere on the path)
    base = [synthetic.synth_clip(seed0 + i, seconds).astype(np.float32) for i in range(min(n_clips, 8))]
    return [base[i % len(base)] for i in range(n_clips)]


def pick_threads(sd) -> int:
    """Use all the host threads torch can profit from: more threads than physical/cgroup cores
    makes the CPU path slower (measured 10x on a 128-thread box), so the count is calibrated on
    a short forward pass and the fastest setting is used for the timed run."""
    from oracle import beat_this_oracle as O

    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    cands = sorted({c for c in (n, n // 2, 32, 16, 8) if 1 <= c <= n}, reverse=True)
    x = torch.rand(1, 300, 128) * 7
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.inference_mode():
            O.forward(sd, x)
            t0 = time.perf_counter()
            O.forward(sd, x)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, world):
    """Reference arm: the reference's own algorithm (CPU oracle port of its PyTorch forward,
    oracle/beat_this_oracle.py -- /root/reference is not present on the GPU box and the
    reference has no compiled code to build) on the host cores, same config and metric."""
    if rank != 0:
        return
    from beat_this_b200 import synthetic
    from oracle import beat_this_oracle as O

    ckpt = synthetic.write_checkpoint(os.path.join(CACHE, "final0_s0.ckpt"), "final0", 0)
    sd = O.strip_prefix(torch.load(ckpt, weights_only=True)["state_dict"])
    threads = pick_threads(sd)
    clips_per_step = args.ref_clips_per_step
    clips = [synthetic.synth_clip(1000 + i, args.seconds) for i in range(clips_per_step)]
    for _ in range(args.warmup):
        O.audio2beats(sd, clips[0])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for c in clips:
            O.audio2beats(sd, c)
    dt = time.perf_counter() - t0
    value = clips_per_step * args.steps / dt
    sample = f"{clips_per_step} clip(s) of {args.seconds:g} s per step, fp32, torch CPU ops, {threads} threads (calibrated; {os.cpu_count()} logical CPUs)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Audio2Beats final0-shaped synthetic checkpoint, {args.seconds:g} s clips @22.05 kHz mono, minimal peak picking",
                   "batch_per_gpu": clips_per_step, "bounded_sample": sample},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(seconds: float, budget_s: float = 20.0):
    from beat_this_b200 import synthetic
    from oracle import beat_this_oracle as O

    ckpt = synthetic.write_checkpoint(os.path.join(CACHE, "final0_s0.ckpt"), "final0", 0)
    sd = O.strip_prefix(torch.load(ckpt, weights_only=True)["state_dict"])
    threads = pick_threads(sd)
    x = synthetic.synth_clip(1000, seconds)
    O.audio2beats(sd, x)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        O.audio2beats(sd, x)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 16:
            break
    return {"value": n / el, "unit": "clips/s", "cores": threads, "kind": "port",
            "sample": f"{n} x {seconds:g} s clip(s), oracle port of the reference forward (fp32 torch CPU ops), {el:.1f} s"}


def run_ours(args, rank, world, local):
    import torch.distributed as dist

    from beat_this_b200 import synthetic
    from beat_this_b200.distributed import load_model_distributed
    from beat_this_b200.inference import Audio2Beats

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ckpt = synthetic.write_checkpoint(os.path.join(CACHE, "final0_s0.ckpt"), "final0", 0) if rank == 0 else None
    bf16 = not args.float32
    a2b = Audio2Beats.__new__(Audio2Beats)  # assemble around a broadcast model (one NCCL broadcast of the weights)
    a2b.device, a2b.float16 = dev, bf16
    a2b.model = load_model_distributed(ckpt, dev, float16=bf16, wave_chunks=args.wave)
    from beat_this_b200.postprocessor import Postprocessor
    from beat_this_b200.preprocessing import LogMelSpect

    a2b.spect = LogMelSpect(device=dev, _engine=a2b.model.engine)
    a2b._pinned = None
    a2b.frames2beats = Postprocessor("minimal", engine=a2b.model.engine)
    eng = a2b.model.engine

    clips = synth_batch(args.batch, args.seconds, 2000 + 100 * rank)
    so = [0]
    for c in clips:
        so.append(so[-1] + len(c))
    host = torch.empty(so[-1], dtype=torch.float32).pin_memory()
    hn = host.numpy()
    for i, c in enumerate(clips):
        hn[so[i]:so[i + 1]] = c
    audio_dev = host.to(dev)
    fo = eng.frame_offsets(so)
    n_chunks = sum(int(eng.lib.bt_plan_chunks(fo[i + 1] - fo[i], None, None, 0)) for i in range(len(clips)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        beat, down, f = eng.audio2frames_cat(audio_dev, so)
        return eng.peakpick_cat(beat, down, f)

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    eng.profile_reset()
    eng.profile_enable(True)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - l0
    eng.profile_enable(False)
    prof = eng.profile_results()
    clocks = sampler.stop()
    # ---- end-to-end timing: pinned host audio in, numpy timestamps out, every step ------------
    # BeatPipeline double-buffers: the H2D copy of step i+1 overlaps the kernels of step i.
    from beat_this_b200.pipeline import BeatPipeline

    pipe = BeatPipeline(a2b, depth=2)
    hosts = [host, host.clone().pin_memory()]
    for i in range(max(2, args.warmup // 2)):
        pipe.submit(hosts[i % 2], so)
        if len(pipe.inflight) == 2:
            pipe.collect()
    while pipe.inflight:
        pipe.collect()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(args.steps):
        if not pipe.free:
            res = pipe.collect()
        h = pipe.submit(hosts[i % 2], so)
        d2h = h.d2h_bytes
    while pipe.inflight:
        res = pipe.collect()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1000.0
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])
    if rank != 0:
        return
    peaks, peak_src = load_peaks()
    total_clips = args.batch * world
    value = total_clips * args.steps / (ms / 1000.0)
    e2e_value = total_clips * args.steps / e2e_s
    # roofline of the dominant kernel: time-direction attention (frontend + main layers).
    # algorithmic FLOPs = 4 * L^2 * d per (sequence, head) (QK^T and PV, 2 flops per MAC), L = chunk length
    D = eng.hparams.get("transformer_dim", 512)
    headseqs_per_chunk = 3 * 32 + eng.hparams.get("n_layers", 6) * (D // 32)
    Lc = 1500 if args.seconds >= 29.76 else fo[1] + 12
    attn_flops = 4.0 * Lc * Lc * 32 * headseqs_per_chunk * n_chunks * args.steps
    key = "attn_time_tc" if bf16 else "attn_time_simt"
    a_ms, a_n = prof.get(key, (0.0, 0))
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    # DRAM traffic of the kernel from the committed ncu --set full capture of the same configuration
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "attn_traffic.json")
    if bf16 and os.path.exists(tpath) and args.batch == 64 and args.seconds == 30.0:
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch_mean"), tj.get("source")
    roof = {"bound": "tensor", "kernel": key, "achieved": (attn_flops / (a_ms / 1000.0) / 1e12) if a_ms > 0 else None,
            "peak": peak_tf, "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
            "unit": "TFLOP/s", "traffic": traffic, "traffic_source": traffic_src, "launches": a_n, "avg_launch_ms": (a_ms / a_n) if a_n else None,
            "algorithmic_flops_per_launch": attn_flops / a_n if a_n else None}
    roof["frac"] = (roof["achieved"] / peak_tf) if roof["achieved"] else None
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    shares = {k: {"ms_per_step": round(v[0] / args.steps, 3), "launches_per_step": v[1] // max(1, args.steps), "share": round(v[0] / tot_ms, 4)}
              for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if bf16 else "f32", "data": "synthetic",
        "config": {"workload": f"Audio2Beats final0-shaped synthetic checkpoint (seeded random weights), batch {args.batch} x {args.seconds:g} s clips "
                               f"@22.05 kHz mono per GPU ({n_chunks} chunks of {Lc} frames), log-mel + BeatThis forward + minimal peak picking",
                   "batch_per_gpu": args.batch, "global_batch": total_clips, "parallelism": f"dp{world} (clips sharded, weights broadcast once over NCCL)",
                   "wave_chunks": args.wave, "l2_policy": f"inputs larger than L2: {so[-1] * 4 / 1e6:.0f} MB audio per step per GPU; activations stream through HBM",
                   "profile": "per-kernel CUDA events recorded inside the timed region (bt_profile_*)"},
        "roofline": roof, "kernel_time_shares": shares,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": so[-1] * 4, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000.0 * e2e_s / args.steps, "api": "beat_this_b200.pipeline.BeatPipeline over Audio2Beats (pinned host fp32 audio in, numpy timestamps out; H2D of step i+1 overlaps the kernels of step i)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.seconds)
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--wave", type=int, default=128, help="chunks per wave (one wave = one launch of every kernel)")
    ap.add_argument("--float32", action="store_true", help="fp32 CUDA-core path instead of bf16 tcgen05")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-clips-per-step", type=int, default=2)
    args = ap.parse_args()
    from beat_this_b200.distributed import init_from_env

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = init_from_env("nccl")
    try:
        run_ours(args, rank, world, local)
    finally:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
