"""Small host helpers of the inference path with the reference's names and behaviour
(reference beat_this/utils.py: infer_beat_numbers :26-76, save_beat_tsv :79-102, replace_state_dict_key :105-111),
written for array input: the bar positions are computed with numpy index arithmetic instead of a Python loop."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np


def replace_state_dict_key(state_dict: dict, old: str, new: str) -> dict:
    """Rename, in place, every key that contains `old` (used to strip the 'model.' / '_orig_mod.' prefixes)."""
    renamed = [(k, k.replace(old, new)) for k in state_dict if old in k]
    for src, dst in renamed:
        state_dict[dst] = state_dict.pop(src)
    return state_dict


def _warn(msg: str) -> None:
    print("WARNING: " + msg, file=sys.stdout)


def infer_beat_numbers(beats: np.ndarray, downbeats: np.ndarray) -> np.ndarray:
    """Position of every beat inside its bar: downbeats are 1 and the count runs up to the next downbeat.  Beats
    before the first downbeat (the pickup) continue the count of a bar as long as the first complete one when
    that is possible, otherwise they simply count up from 2 (same rule and same warnings as the reference)."""
    beats = np.asarray(beats)
    downbeats = np.asarray(downbeats)
    is_down = np.isin(beats, downbeats)
    if int(is_down.sum()) != len(np.unique(downbeats)):
        raise ValueError("Not all downbeats are beats.")
    start = 1  # the counter value "before" the first beat
    down_idx = np.flatnonzero(is_down)
    if len(down_idx) >= 2:
        pickup, first_bar = int(down_idx[0]), int(down_idx[1] - down_idx[0])
        if pickup < first_bar:
            start = first_bar - pickup
        else:
            _warn("There are more beats in the pickup measure than in the first measure. "
                  "The beat count will start from 2 without trying to estimate the length of the pickup measure.")
    else:
        _warn("There are less than two downbeats in the predictions. Something may be wrong. "
              "The beat count will start from 2 without trying to estimate the length of the pickup measure.")
    idx = np.arange(len(beats))
    # index of the most recent downbeat at or before every beat (-1 inside the pickup)
    last_down = np.maximum.accumulate(np.where(is_down, idx, -1)) if len(beats) else idx
    numbers = np.where(last_down >= 0, idx - last_down + 1, start + 1 + idx)
    return numbers.astype(np.int64)


def save_beat_tsv(beats: np.ndarray, downbeats: np.ndarray, outpath: str) -> None:
    """The `.beats` text format: one line per beat, "<time in seconds><TAB><position in the bar>", 1 = downbeat.
    A file cut short by Ctrl-C is removed rather than left half written."""
    rows = zip(np.asarray(beats).tolist(), infer_beat_numbers(beats, downbeats).tolist())
    text = "".join(f"{t}\t{n}\n" for t, n in rows)
    target = Path(outpath)
    target.parent.mkdir(parents=True, exist_ok=True)
    try:
        target.write_text(text)
    except KeyboardInterrupt:
        target.unlink(missing_ok=True)
        raise
