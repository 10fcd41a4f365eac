"""BASELINE config 4 (Audio2Beats final0 --dbn, DBN on the host): clips/s of the device part + host DBN.
usage: python tools/config4_dbn.py [n_clips=256] [batch=64]   -> one JSON line"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from beat_this_b200 import synthetic
from beat_this_b200.inference import Audio2Beats

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ckpt = synthetic.write_checkpoint("/tmp/beat_this_b200_cache/final0_s0.ckpt", "final0", 0)
a2b = Audio2Beats(ckpt, "cuda:0", True, True)
clips = [synthetic.synth_clip(3000 + (i % 16), 30.0).astype(np.float32) for i in range(batch)]
a2b.batch(clips, 22050)  # warm-up
torch.cuda.synchronize()
t_dev = t_dbn = 0.0
n_beats = 0
t0 = time.perf_counter()
for _ in range(n_clips // batch):
    ta = time.perf_counter()
    beat, down, fo = a2b._frames_batch(clips, 22050)
    torch.cuda.synchronize()
    tb = time.perf_counter()
    res = a2b.frames2beats.batch_cat(beat, down, fo)
    tc = time.perf_counter()
    t_dev += tb - ta
    t_dbn += tc - tb
    n_beats += sum(len(r[0]) for r in res)
total = time.perf_counter() - t0
done = (n_clips // batch) * batch
print(json.dumps({"config": "Audio2Beats final0-shaped synthetic ckpt, --dbn, 30 s clips, 1 GPU, bf16", "clips": done, "batch": batch,
                  "clips_per_s": done / total, "device_s": t_dev, "host_dbn_s": t_dbn, "host_threads": os.cpu_count(),
                  "dbn_impl": type(a2b.frames2beats.dbn).__module__, "beats_per_clip": n_beats / done}))
