"""Double-buffered host -> device -> host pipeline around ``Audio2Beats``: the H2D copy of batch
i+1 (copy stream) overlaps the kernels of batch i (compute stream); results come back through
pinned memory without blocking the enqueue loop.  The reference processes one clip at a time on
one stream (reference inference.py:215); this is the batched serving loop of the B200 path."""
from __future__ import annotations

from collections import deque

import torch


class BeatPipeline:
    def __init__(self, a2b, depth: int = 2):
        self.a2b = a2b
        self.engine = a2b.model.engine
        self.device = self.engine.device
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.slots = [dict(peak={}, audio=None, copied=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(depth)]
        self.free = deque(range(depth))
        self.inflight = deque()

    def submit(self, audio_host: torch.Tensor, sample_offsets):
        """audio_host: pinned fp32 host tensor holding all clips back to back (mono, 22.05 kHz)."""
        if not self.free:
            raise RuntimeError("pipeline full: collect() a result first")
        idx = self.free.popleft()
        slot = self.slots[idx]
        n = int(sample_offsets[-1])
        if slot["audio"] is None or slot["audio"].numel() < n:
            slot["audio"] = torch.empty(n, dtype=torch.float32, device=self.device)
        dev_audio = slot["audio"][:n]
        with torch.cuda.stream(self.copy_stream):
            dev_audio.copy_(audio_host[:n], non_blocking=True)
            slot["copied"].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(slot["copied"])
            beat, down, fo = self.engine.audio2frames_cat(dev_audio, list(sample_offsets))
            handle = self.engine.peakpick_async(beat, down, fo, slot["peak"])
        self.inflight.append((idx, handle))
        return handle

    def collect(self):
        """Results of the oldest submitted batch: list of (beat_times, downbeat_times)."""
        idx, handle = self.inflight.popleft()
        res = handle.result()
        self.free.append(idx)
        return res

    def run(self, batches):
        """Iterate over (audio_host, sample_offsets) batches, yielding results in order."""
        for audio_host, so in batches:
            if not self.free:
                yield self.collect()
            self.submit(audio_host, so)
        while self.inflight:
            yield self.collect()
