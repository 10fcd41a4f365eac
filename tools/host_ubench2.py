import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from beat_this_b200.dbn import DBNDownBeatTracker
from beat_this_b200.postprocessor import Postprocessor
pp = Postprocessor.__new__(Postprocessor)
pp.type, pp.fps = "dbn", 50
pp.dbn = DBNDownBeatTracker(beats_per_bar=[3, 4], min_bpm=55.0, max_bpm=215.0, fps=50, transition_lambda=100)
rng = np.random.default_rng(0)
T = 1501
fo = [i * T for i in range(65)]
t = np.arange(T)
beat = np.concatenate([4 * np.sin(2 * np.pi * t / (18 + i % 7) + i) - 1 + 0.3 * rng.standard_normal(T) for i in range(64)]).astype(np.float32)
down = np.concatenate([4 * np.sin(2 * np.pi * t / (72 + 4 * (i % 7)) + i) - 2.5 + 0.3 * rng.standard_normal(T) for i in range(64)]).astype(np.float32)
pp.batch_host(beat, down, fo)
for rep in range(3):
    t0 = time.perf_counter(); out = pp.batch_host(beat, down, fo); t1 = time.perf_counter()
    print("batch_host 64 x 1501 frames: %.1f ms; beats in piece 0: %d" % (1e3 * (t1 - t0), len(out[0][0])))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); pp.batch_host(beat, down, fo); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
