"""Host <-> device pipeline behind ``Audio2Frames.batch`` / ``Audio2Beats.batch`` / ``File2Beats.batch``.

The reference handles one clip at a time on one stream (reference beat_this/inference.py:215,269-281: numpy mono mix,
``torch.tensor(signal, device=...)``, model, ``.cpu()``).  Here a call is cut into *groups* of clips (one group =
one pass of every kernel); for every group

  host threads   mono mix + fp32 cast of all clips straight into a pinned ring slot (``bt_stage_audio`` /
                 ``bt_stage_wav_files``: C++, GIL released)
  copy stream    one H2D copy of the slot
  compute stream log-mel -> BeatThis forward -> (peak picking -> D2H of the timestamps | D2H of the logits for the DBN)

and group g+1 is staged and copied while the kernels of group g run; results are collected in order.  Nothing in the
enqueue path waits for the GPU (the C library keeps its small tables in a ring of pinned slots), so the device queue
stays one group ahead of the host.
"""
from __future__ import annotations

import ctypes
import time
from collections import deque
from ctypes import c_void_p

import numpy as np
import torch

from . import _lib

SIG_F32, SIG_F64, SIG_I16 = 0, 1, 2
_NP2SIG = {np.dtype(np.float32): SIG_F32, np.dtype(np.float64): SIG_F64, np.dtype(np.int16): SIG_I16}


def as_signal_array(signal) -> np.ndarray:
    """What Audio2Frames.signal2spect accepts (reference inference.py:269-273): 1-D or 2-D (time, channels) array-like.
    Returns a C-contiguous float32 / float64 / int16 ndarray without copying when the input already is one."""
    if isinstance(signal, torch.Tensor):
        signal = signal.detach().cpu().numpy()
    a = np.asarray(signal)
    if a.ndim not in (1, 2):
        raise ValueError(f"Expected 1D or 2D signal, got shape {a.shape}")
    if a.dtype not in _NP2SIG:
        a = a.astype(np.float64)  # ints other than int16, float16, ...: the reference's mean(1) works in float64 too
    return np.ascontiguousarray(a)


def plan_groups(costs, max_cost: int, max_clips: int):
    """Consecutive clips -> groups of at most `max_clips` clips / `max_cost` total cost (a clip above the limit is a
    group of its own).  Returns a list of (first, last+1) index pairs."""
    groups, lo, acc = [], 0, 0
    for i, n in enumerate(costs):
        if i > lo and (acc + n > max_cost or i - lo >= max_clips):
            groups.append((lo, i))
            lo, acc = i, 0
        acc += int(n)
    if lo < len(costs):
        groups.append((lo, len(costs)))
    return groups


def chunk_cost(n_samples: int, sr: int = 22050) -> int:
    """1500-frame model passes a clip of n_samples at `sr` Hz needs: ceil(frames / 1488) (split_piece, inference.py:119-125)."""
    frames = 1 + (int(n_samples) * 22050 // max(1, int(sr))) // 441
    return max(1, -(-frames // 1488))


class _Slot:
    def __init__(self):
        self.host = None      # pinned fp32 staging buffer
        self.dev = None       # device copy
        self.copied = None
        self.peak = {}        # reusable buffers of Engine.peakpick_async
        self.logits_h = None  # pinned logits (DBN path)
        self.done = None
        self.t0 = None        # events around the group's kernels (stats: GPU busy time)
        self.t1 = None


class BeatPipeline:
    """Ring of `depth` slots; `submit_*` enqueues one group, `collect` returns the oldest group's result."""

    # staging slots are sized once for a full group (128 chunks = 128 * 1488 frames of 441 samples, 336 MB of fp32):
    # page-locking hundreds of MB takes ~0.1 s, which must not recur while batches stream through
    SLOT_SAMPLES = 128 * 1488 * 441 + 6 * 441

    def __init__(self, engine, depth: int = 3, host_threads: int | None = None):
        self.engine = engine
        self.lib = engine.lib
        self.device = engine.device
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.slots = [_Slot() for _ in range(depth)]
        self.free = deque(range(depth))
        self.inflight = deque()
        if host_threads is None:
            import os

            try:
                n = len(os.sched_getaffinity(0))
            except AttributeError:
                n = os.cpu_count() or 1
            host_threads = max(1, min(32, n))
        self.host_threads = int(host_threads)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        # host seconds spent staging (mono mix / decode into pinned memory), enqueueing and waiting for results
        self.stats = {"stage_s": 0.0, "enqueue_s": 0.0, "collect_wait_s": 0.0, "gpu_busy_s": 0.0, "groups": 0}

    # ---- staging --------------------------------------------------------------------------------
    def _slot(self, n_samples: int):
        if not self.free:
            raise RuntimeError("pipeline full: collect() a result first")
        idx = self.free.popleft()
        s = self.slots[idx]
        if s.host is None or s.host.numel() < n_samples:
            cap = max(int(n_samples), self.SLOT_SAMPLES)
            s.host = torch.empty(cap, dtype=torch.float32, pin_memory=True)
            s.dev = torch.empty(cap, dtype=torch.float32, device=self.device)
            s.copied = torch.cuda.Event()
        return idx, s

    def stage_signals(self, arrays, dst: torch.Tensor):
        """Mono mix + fp32 cast of C-contiguous ndarrays (see as_signal_array) into `dst` (host fp32 tensor);
        returns the sample offsets."""
        n = len(arrays)
        so = [0]
        for a in arrays:
            so.append(so[-1] + a.shape[0])
        ptrs = (c_void_p * n)(*[a.ctypes.data for a in arrays])
        dts = (ctypes.c_int32 * n)(*[_NP2SIG[a.dtype] for a in arrays])
        frames = (ctypes.c_int64 * n)(*[a.shape[0] for a in arrays])
        chans = (ctypes.c_int32 * n)(*[1 if a.ndim == 1 else a.shape[1] for a in arrays])
        offs = (ctypes.c_int64 * (n + 1))(*so)
        code = self.lib.bt_stage_audio(ptrs, dts, frames, chans, n, c_void_p(dst.data_ptr()), offs, self.host_threads)
        if code != 0:
            raise _lib.BTError(f"bt_stage_audio failed ({code}): bad signal array")
        return so

    def _enqueue(self, idx, s, so, sr, want):
        n = so[-1]
        with torch.cuda.stream(self.copy_stream):
            s.dev[:n].copy_(s.host[:n], non_blocking=True)
            s.copied.record(self.copy_stream)
        self.h2d_bytes += n * 4
        eng = self.engine
        if s.t0 is None:
            s.t0, s.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(s.copied)
            s.t0.record(self.compute_stream)
            audio, offs = s.dev[:n], so
            if sr != 22050:
                audio, offs = eng.resample_cat(audio, so, sr)
            beat, down, fo = eng.audio2frames_cat(audio, offs)
            if want == "beats":
                handle = eng.peakpick_async(beat, down, fo, s.peak)
                self.d2h_bytes += handle.d2h_bytes
                payload = ("beats", handle)
            elif want == "logits_host":
                total = fo[-1]
                if s.logits_h is None or s.logits_h.shape[1] < total:
                    s.logits_h = torch.empty((2, max(int(total * 1.25), 1024)), dtype=torch.float32, pin_memory=True)
                s.logits_h[0, :total].copy_(beat, non_blocking=True)
                s.logits_h[1, :total].copy_(down, non_blocking=True)
                s.done = torch.cuda.Event()
                s.done.record(self.compute_stream)
                self.d2h_bytes += total * 8
                payload = ("logits_host", (s, fo, (beat, down)))
            else:  # device logits
                s.done = torch.cuda.Event()
                s.done.record(self.compute_stream)
                payload = ("frames", (s, beat, down, fo))
            s.t1.record(self.compute_stream)
        self.inflight.append((idx, payload))

    def submit_signals(self, arrays, sr: int = 22050, want: str = "beats"):
        idx, s = self._slot(sum(a.shape[0] for a in arrays))
        try:
            t0 = time.perf_counter()
            so = self.stage_signals(arrays, s.host)
            t1 = time.perf_counter()
            self._enqueue(idx, s, so, int(sr), want)
            self.stats["stage_s"] += t1 - t0
            self.stats["enqueue_s"] += time.perf_counter() - t1
            self.stats["groups"] += 1
        except Exception:
            self.free.append(idx)
            raise

    def submit_wavs(self, paths, infos, sr: int, want: str = "beats"):
        """paths: list of str; infos: ctypes array of bt_wav_info (all `sr` Hz) from bt_wav_probe."""
        n = len(paths)
        so = [0]
        for i in range(n):
            so.append(so[-1] + int(infos[i].frames))
        idx, s = self._slot(so[-1])
        try:
            cpaths = (ctypes.c_char_p * n)(*[str(p).encode() for p in paths])
            offs = (ctypes.c_int64 * (n + 1))(*so)
            status = (ctypes.c_int32 * n)()
            t0 = time.perf_counter()
            code = self.lib.bt_stage_wav_files(cpaths, infos, n, c_void_p(s.host.data_ptr()), offs, self.host_threads, status)
            if code != 0:
                bad = [str(paths[i]) for i in range(n) if status[i] != 0]
                raise RuntimeError(f"Could not load audio from {bad}")
            t1 = time.perf_counter()
            self._enqueue(idx, s, so, int(sr), want)
            self.stats["stage_s"] += t1 - t0
            self.stats["enqueue_s"] += time.perf_counter() - t1
            self.stats["groups"] += 1
        except Exception:
            self.free.append(idx)
            raise

    def submit_pinned(self, audio_host: torch.Tensor, sample_offsets, sr: int = 22050, want: str = "beats"):
        """Mono fp32 audio already in one pinned host tensor (clips back to back)."""
        so = [int(v) for v in sample_offsets]
        idx, s = self._slot(so[-1])
        s.host[: so[-1]].copy_(audio_host[: so[-1]])
        self._enqueue(idx, s, so, int(sr), want)

    # ---- results ----------------------------------------------------------------------------------
    def collect(self):
        """Oldest group: list of (beat_times, downbeat_times) ["beats"], (beat, down, fo) host arrays
        ["logits_host"] or device tensors ["frames"]."""
        idx, (kind, p) = self.inflight.popleft()
        t0 = time.perf_counter()
        try:
            if kind == "beats":
                return p.result()
            if kind == "logits_host":
                s, fo, _keep = p
                s.done.synchronize()
                total = fo[-1]
                return s.logits_h[0, :total].numpy().copy(), s.logits_h[1, :total].numpy().copy(), fo
            s, beat, down, fo = p
            s.done.synchronize()
            return beat, down, fo
        finally:
            self.stats["collect_wait_s"] += time.perf_counter() - t0
            try:
                self.slots[idx].t1.synchronize()
                self.stats["gpu_busy_s"] += self.slots[idx].t0.elapsed_time(self.slots[idx].t1) / 1000.0
            except Exception:
                pass
            self.free.append(idx)

    def run(self, n_groups: int, submit):
        """Call submit(g) for g in range(n_groups), keeping the ring full; yields the results in order."""
        for g in range(n_groups):
            if not self.free:
                yield self.collect()
            submit(g)
        while self.inflight:
            yield self.collect()

    def drain(self):
        while self.inflight:
            try:
                self.collect()
            except Exception:
                pass
