"""Host utilities mirrored from reference beat_this/utils.py (the parts the inference path
uses: replace_state_dict_key :105-111, infer_beat_numbers :26-76, save_beat_tsv :79-102)."""
from __future__ import annotations

from itertools import chain
from pathlib import Path

import numpy as np


def replace_state_dict_key(state_dict: dict, old: str, new: str):
    """Replaces `old` in all keys of `state_dict` with `new`."""
    for key in list(state_dict.keys()):
        if old in key:
            state_dict[key.replace(old, new)] = state_dict.pop(key)
    return state_dict


def infer_beat_numbers(beats: np.ndarray, downbeats: np.ndarray) -> np.ndarray:
    """Number every beat so that downbeats get 1 and the beats in between count upwards; the
    pickup measure is counted back from the first full measure (reference utils.py:26-76)."""
    if not np.all(np.isin(downbeats, beats)):
        raise ValueError("Not all downbeats are beats.")
    if len(downbeats) >= 2:
        first_downbeat, second_downbeat = np.searchsorted(beats, downbeats[:2])
        beats_in_first_measure = second_downbeat - first_downbeat
        pickup_beats = first_downbeat
        if pickup_beats < beats_in_first_measure:
            start_counter = beats_in_first_measure - pickup_beats
        else:
            print(
                "WARNING: There are more beats in the pickup measure than in the first measure. "
                "The beat count will start from 2 without trying to estimate the length of the pickup measure."
            )
            start_counter = 1
    else:
        print(
            "WARNING: There are less than two downbeats in the predictions. Something may be wrong. "
            "The beat count will start from 2 without trying to estimate the length of the pickup measure."
        )
        start_counter = 1
    numbers = []
    counter = start_counter
    downbeats = chain(downbeats, [-1])
    next_downbeat = next(downbeats)
    for beat in beats:
        if beat == next_downbeat:
            counter = 1
            next_downbeat = next(downbeats)
        else:
            counter += 1
        numbers.append(counter)
    return np.asarray(numbers)


def save_beat_tsv(beats: np.ndarray, downbeats: np.ndarray, outpath: str) -> None:
    """Write the standard .beats format: "<seconds>\\t<beat number>" per line, 1 = downbeat
    (reference utils.py:79-102)."""
    numbers = infer_beat_numbers(beats, downbeats)
    outpath = Path(outpath)
    outpath.parent.mkdir(parents=True, exist_ok=True)
    try:
        with open(outpath, "w") as f:
            f.writelines(f"{beat}\t{number}\n" for beat, number in zip(beats, numbers))
    except KeyboardInterrupt:
        outpath.unlink()  # avoid half-written files
