"""Time variants of the time-direction attention kernel on the main-layer / frontend shapes of one bench step.
usage (GPU box): python tools/attn_ubench.py [variant ...]   variant = template parameter V of attn_tc48_kernel"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from beat_this_b200.engine import Engine


def main():
    variants = [int(v, 0) for v in sys.argv[1:]] or [-1]
    eng = Engine(None, None, device="cuda:0", half=True)  # no weights: only the context is used
    for name, (seqs, L, heads) in {"main": (128, 1500, 16), "front_f32": (128 * 32, 1500, 1)}.items():
        for v in variants:
            ms = ctypes.c_float()
            code = eng.lib.bt_debug_attention_time(eng.ctx, seqs, L, heads, v, 5, ctypes.byref(ms))
            print(f"{name} variant {v}: " + (f"{ms.value:.4f} ms" if code == 0 else f"error {code}"), flush=True)


if __name__ == "__main__":
    main()
