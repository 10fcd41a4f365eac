#!/usr/bin/env python3
"""``beat_this`` command line tool on the B200 engine (reference beat_this/cli.py:22-191: same options, same
output naming, same ``.beats`` / ``.npy`` files), re-organised around the batched device path: the work list is
built first, files are decoded and pushed through ``Audio2Frames`` / the device peak picker ``--batch`` files at
a time (files of one sample rate share a launch), and under ``torchrun`` every rank takes every WORLD_SIZE-th
task (tasks are independent: no collective).

    python -m beat_this_b200.cli song.wav                     # -> song.beats
    python -m beat_this_b200.cli music_dir -o out --float16    # directory tree -> out/.../*.beats
    torchrun --nproc-per-node 8 -m beat_this_b200.cli music_dir -o out --skip-existing --touch-first
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np

from .utils import save_beat_tsv


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="beat_this_b200", description="Beat and downbeat times for audio files (Beat This! model on the B200 engine).")
    add = ap.add_argument
    add("inputs", nargs="+", help="audio files and/or directories that are searched recursively")
    add("--model", default="final0", help="checkpoint name or path [%(default)s]")
    add("--output", "-o", default=None, help="result file (one input file) or result directory; default: next to each input")
    add("--suffix", "-s", default=".beats", help="extension of the result files [%(default)s]")
    add("--append", action="store_true", help="keep the audio extension and add the suffix after it")
    add("--skip-existing", action="store_true", help="leave results that already exist untouched")
    add("--touch-first", action="store_true", help="create the (empty) result file before working on it: with --skip-existing, "
                                                   "several processes can split one directory between them")
    add("--dbn", default=False, action=argparse.BooleanOptionalAction, help="DBN post-processing on the host instead of peak picking")
    add("--gpu", type=int, default=None, help="CUDA device index [LOCAL_RANK or 0]; a GPU is required")
    add("--float16", action="store_true", help="bf16 tensor-core kernels (fast path) instead of fp32")
    add("--activations", action="store_true", help="also write the frame activations as <result>.npy (2 x frames)")
    add("--batch", type=int, default=32, help="files per device batch [%(default)s]")
    return ap


def output_path_for(src: Path, suffix: str, append: bool, out_dir: Path | None = None, root: Path | None = None) -> Path:
    """Where the result of `src` goes: next to it, or under `out_dir` keeping the path relative to the directory
    `root` that was named on the command line; the suffix replaces the old one unless `append`."""
    if out_dir is None:
        dst = src
    else:
        dst = out_dir / (src.relative_to(root) if root is not None else src.name)
    return dst.parent / (dst.name + suffix) if append else dst.with_suffix(suffix)


def collect_tasks(inputs, output, suffix, append, skip_existing):
    """[(audio file, output file)] for the command line; a single plain file may name its output file directly."""
    inputs = [Path(p) for p in inputs]
    output = Path(output) if output is not None else None
    if len(inputs) == 1 and not inputs[0].is_dir():
        dst = output
        if dst is None or dst.is_dir():
            dst = output_path_for(inputs[0], suffix, append, dst)
        return [(inputs[0], dst)], True
    tasks = []
    for item in inputs:
        if item.is_dir():
            for fn in sorted(item.rglob("*")):
                if fn.is_dir() or fn.name.endswith(suffix):
                    continue
                dst = output_path_for(fn, suffix, append, output, root=item)
                if not skip_existing or not dst.exists():
                    tasks.append((fn, dst))
        else:
            tasks.append((item, output_path_for(item, suffix, append, output)))
    return tasks, False


def _claim(dst: Path, skip_existing: bool, touch_first: bool) -> bool:
    """Reference cli.py:171-177: with --touch-first the empty output file is the lock."""
    if touch_first:
        try:
            dst.parent.mkdir(parents=True, exist_ok=True)
            dst.touch(exist_ok=not skip_existing)
        except FileExistsError:
            return False
        return True
    return not (skip_existing and dst.exists())


def run(inputs, model="final0", output=None, suffix=".beats", append=False, skip_existing=False, touch_first=False,
        dbn=False, gpu=None, float16=False, activations=False, batch=32) -> int:
    from .inference import Audio2Beats
    from .preprocessing import load_audio

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if gpu is None:
        gpu = int(os.environ.get("LOCAL_RANK", "0"))
    if gpu < 0:
        raise SystemExit("beat_this_b200 has no CPU path: --gpu must name a CUDA device")
    tasks, single = collect_tasks(inputs, output, suffix, append, skip_existing)
    tasks = tasks[rank::world]
    a2b = Audio2Beats(model, f"cuda:{gpu}", float16, dbn)
    failed = 0
    for b0 in range(0, len(tasks), max(1, batch)):
        group = []
        for src, dst in tasks[b0 : b0 + max(1, batch)]:
            if not single and not _claim(dst, skip_existing, touch_first):
                continue
            try:
                wav, sr = load_audio(src)
                group.append((src, dst, wav, sr))
            except Exception as e:
                if single:
                    raise
                failed += 1
                print(f'Could not load "{src}": {e}', file=sys.stderr)
        for sr in sorted({g[3] for g in group}):
            same = [g for g in group if g[3] == sr]
            try:
                beat, down, fo = a2b._frames_batch([g[2] for g in same], sr)
                times = a2b.frames2beats.batch_cat(beat, down, fo)
            except Exception as e:
                if single:
                    raise
                failed += len(same)
                print(f"Could not process a batch of {len(same)} file(s) at {sr} Hz ({e}); rerun them one by one for details.", file=sys.stderr)
                continue
            beat_h, down_h = beat.cpu().numpy(), down.cpu().numpy()
            for i, (src, dst, _, _) in enumerate(same):
                beats, downbeats = times[i]
                if activations:  # reference cli.py:139-147: vstack([beat, downbeat]) next to the .beats file
                    dst.parent.mkdir(parents=True, exist_ok=True)
                    np.save(dst.with_suffix(".npy"), np.vstack([beat_h[fo[i] : fo[i + 1]], down_h[fo[i] : fo[i + 1]]]))
                save_beat_tsv(beats, downbeats, dst)
    return 1 if failed else 0


def main(argv=None) -> int:
    return run(**vars(build_parser().parse_args(argv)))


if __name__ == "__main__":
    sys.exit(main())
