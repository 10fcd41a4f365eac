"""Drop-in mirror of the reference ``beat_this.inference`` API (reference
beat_this/inference.py:16-315): same function and class names, constructor and call
signatures, return types and exceptions -- with everything between "audio samples" and
"beat timestamps" executed by the sm_100a CUDA library.

Differences a user can observe:
* ``device`` must be a CUDA device (default ``"cuda"``); ``device="cpu"`` raises.
* ``float16=False`` -> fp32 CUDA-core kernels (reference-exact numerics, <=1e-3 on logits);
  ``float16=True`` -> bf16 tcgen05 tensor-core kernels with fp32 accumulation.
* every class has a ``batch(...)`` method that processes many clips per call (the
  reference is strictly one clip, one chunk at a time: inference.py:215).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .engine import Engine
from .postprocessor import Postprocessor
from .preprocessing import LogMelSpect, load_audio
from .utils import replace_state_dict_key, save_beat_tsv
from .weights import filter_hparams, pack_parameters

CHECKPOINT_URL = "https://cloud.cp.jku.at/public.php/dav/files/7ik4RrBKTS273gp"


def load_checkpoint(checkpoint_path: str, device: str | torch.device = "cpu") -> dict:
    """Load a BeatThis checkpoint as a dictionary (reference inference.py:16-53): local file,
    else a short name / URL fetched through torch.hub (needs network)."""
    try:
        return torch.load(checkpoint_path, map_location=device, weights_only=True)
    except FileNotFoundError:
        try:
            if not (str(checkpoint_path).startswith("https://") or str(checkpoint_path).startswith("http://")):
                checkpoint_url = f"{CHECKPOINT_URL}/{checkpoint_path}.ckpt"
                file_name = f"beat_this-{checkpoint_path}.ckpt"
            else:
                checkpoint_url = checkpoint_path
                file_name = None
            return torch.hub.load_state_dict_from_url(checkpoint_url, file_name=file_name, map_location=device)
        except Exception:
            raise ValueError("Could not load the checkpoint given the provided name", checkpoint_path)


class BeatThisB200:
    """What ``load_model`` returns in place of the reference ``BeatThis`` nn.Module: the packed
    weights living on one GPU inside a ``bt_ctx``."""

    def __init__(self, hparams: dict, packed: dict, device, float16: bool = False, wave_chunks: int | None = None):
        self.hparams = filter_hparams(hparams)
        self.engine = Engine(packed, self.hparams, device, bf16=float16, wave_chunks=wave_chunks)
        self.device = self.engine.device
        self.float16 = float16

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("beat_this_b200 models live on a CUDA device; there is no CPU fallback")
        return self


def load_model(checkpoint_path: str | dict | None = "final0", device: str | torch.device = "cuda", float16: bool = False,
               wave_chunks: int | None = None) -> BeatThisB200:
    """Load a BeatThis model from a checkpoint (reference inference.py:56-87).  Accepts the
    reference ``.ckpt`` layout unchanged (``hyper_parameters`` + ``state_dict`` with the
    ``model.`` prefix).  ``checkpoint_path`` may also be an already loaded checkpoint dict."""
    from .engine import _cuda_device

    _cuda_device(device)  # fail before touching the checkpoint: there is no CPU path
    if checkpoint_path is None:
        raise ValueError("beat_this_b200 needs a checkpoint (the reference's random-init BeatThis() has no use here)")
    checkpoint = checkpoint_path if isinstance(checkpoint_path, dict) else load_checkpoint(checkpoint_path, "cpu")
    hparams = filter_hparams(checkpoint["hyper_parameters"])
    state_dict = replace_state_dict_key(dict(checkpoint["state_dict"]), "model.", "")
    packed = pack_parameters(state_dict, hparams)
    return BeatThisB200(hparams, packed, device, float16, wave_chunks)


def zeropad(spect: torch.Tensor, left: int = 0, right: int = 0):
    """reference inference.py:90-97"""
    if left == 0 and right == 0:
        return spect
    return F.pad(spect, (0, 0, left, right), "constant", 0)


def split_piece(spect: torch.Tensor, chunk_size: int, border_size: int = 6, avoid_short_end: bool = True):
    """Host mirror of reference inference.py:100-135 (the CUDA path plans chunks natively in
    bt_plan_chunks and gathers them inside the stem kernel; this function is kept for API
    compatibility and as the test oracle of the native planner)."""
    starts = np.arange(-border_size, len(spect) - border_size, chunk_size - 2 * border_size)
    if avoid_short_end and len(spect) > chunk_size - 2 * border_size:
        starts[-1] = len(spect) - (chunk_size - border_size)
    chunks = [
        zeropad(
            spect[max(start, 0) : min(start + chunk_size, len(spect))],
            left=max(0, -start),
            right=max(0, min(border_size, start + chunk_size - len(spect))),
        )
        for start in starts
    ]
    return chunks, starts


def aggregate_prediction(pred_chunks: list, starts: list, full_size: int, chunk_size: int, border_size: int,
                         overlap_mode: str, device: str | torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """Host mirror of reference inference.py:138-185 (the CUDA head kernel scatters with the
    same keep_first ownership rule)."""
    if border_size > 0:
        pred_chunks = [
            {"beat": p["beat"][border_size:-border_size], "downbeat": p["downbeat"][border_size:-border_size]}
            for p in pred_chunks
        ]
    beat = torch.full((full_size,), -1000.0, device=device)
    downbeat = torch.full((full_size,), -1000.0, device=device)
    if overlap_mode == "keep_first":
        pred_chunks = reversed(list(pred_chunks))
        starts = reversed(list(starts))
    for start, p in zip(starts, pred_chunks):
        beat[start + border_size : start + chunk_size - border_size] = p["beat"]
        downbeat[start + border_size : start + chunk_size - border_size] = p["downbeat"]
    return beat, downbeat


class Spect2Frames:
    """Framewise beat / downbeat logits from a spectrogram (reference inference.py:233-257)."""

    def __init__(self, checkpoint_path="final0", device="cuda", float16=False):
        super().__init__()
        self.device = torch.device(device)
        self.float16 = float16
        self.model = load_model(checkpoint_path, self.device, float16)
        self.device = self.model.device

    def spect2frames(self, spect):
        spect = torch.as_tensor(spect, dtype=torch.float32, device=self.device).contiguous()
        if spect.ndim != 2 or spect.shape[1] != 128:
            raise ValueError(f"Expected a (time, 128) spectrogram, got shape {tuple(spect.shape)}")
        beat, down = self.model.engine.spect2frames_cat(spect, [0, spect.shape[0]])
        return beat, down

    def spects2frames(self, spects):
        """Batched variant: list of [T_i,128] tensors -> list of (beat, downbeat)."""
        spects = [torch.as_tensor(s, dtype=torch.float32, device=self.device) for s in spects]
        fo = [0]
        for s in spects:
            fo.append(fo[-1] + s.shape[0])
        beat, down = self.model.engine.spect2frames_cat(torch.cat(spects).contiguous(), fo)
        return [(beat[fo[i] : fo[i + 1]], down[fo[i] : fo[i + 1]]) for i in range(len(spects))]

    def __call__(self, spect):
        return self.spect2frames(spect)


def _mono(signal):
    """Channel mix of Audio2Frames.signal2spect (inference.py:269-273): float64 mean over axis 1."""
    signal = np.asarray(signal) if not isinstance(signal, (np.ndarray, torch.Tensor)) else signal
    if isinstance(signal, torch.Tensor):
        signal = signal.detach().cpu().numpy()
    if signal.ndim == 2:
        signal = signal.mean(1)
    elif signal.ndim != 1:
        raise ValueError(f"Expected 1D or 2D signal, got shape {signal.shape}")
    return signal


def _soxr_resample(signal, sr):
    """The reference's own host resampler (inference.py:274-275), when the package is installed."""
    try:
        import soxr
    except ImportError as e:
        raise RuntimeError("resampler='soxr' needs the `soxr` package; the default resampler='device' does not") from e
    return soxr.resample(signal, in_rate=sr, out_rate=22050)


class Audio2Frames(Spect2Frames):
    """Framewise logits from an audio signal (reference inference.py:260-281).

    Audio that is not at 22.05 kHz is resampled on the device (`resampler="device"`, a Kaiser-windowed-sinc
    polyphase FIR designed to soxr-HQ-like targets, see preprocessing.resample_filter_bank) or, with
    `resampler="soxr"`, by the reference's own host library when it is installed."""

    def __init__(self, checkpoint_path="final0", device="cuda", float16=False, resampler="device"):
        super().__init__(checkpoint_path, device, float16)
        if resampler not in ("device", "soxr"):
            raise ValueError("resampler must be 'device' or 'soxr'")
        self.resampler = resampler
        self.spect = LogMelSpect(device=self.device, _engine=self.model.engine)
        self._pinned = None

    def signal2spect(self, signal, sr):
        audio, so = self._stage([signal], sr)
        spect, _ = self.model.engine.logmel_cat(audio, so)
        return spect

    def __call__(self, signal, sr):
        beat, down, _ = self._frames_batch([signal], sr)
        return beat, down

    # ---- batched path ----------------------------------------------------------------------
    def _stage(self, signals, sr):
        """mono fp32 signals -> one pinned host buffer -> device (async) -> 22.05 kHz on the device."""
        sr = int(sr)
        host_resample = sr != 22050 and self.resampler == "soxr"
        mono = [_mono(s) for s in signals]
        if host_resample:
            mono = [_soxr_resample(m, sr) for m in mono]
        mono = [np.ascontiguousarray(m, dtype=np.float32) for m in mono]
        so = [0]
        for m in mono:
            so.append(so[-1] + m.shape[0])
        if self._pinned is None or self._pinned.numel() < so[-1]:
            self._pinned = torch.empty(max(so[-1], 1), dtype=torch.float32).pin_memory()
        host = self._pinned[: so[-1]]
        hn = host.numpy()
        for i, m in enumerate(mono):
            hn[so[i] : so[i + 1]] = m
        audio = host.to(self.device, non_blocking=True)
        if sr != 22050 and not host_resample:
            audio, so = self.model.engine.resample_cat(audio, so, sr)
        return audio, so

    def _frames_batch(self, signals, sr):
        audio, so = self._stage(signals, sr)
        return self.model.engine.audio2frames_cat(audio, so)

    def frames_from_device(self, audio: torch.Tensor, sample_offsets):
        """Audio already on the device (flat fp32 tensor + offsets)."""
        return self.model.engine.audio2frames_cat(audio, list(sample_offsets))

    def batch(self, signals, sr=22050):
        """list of signals -> list of (beat_logits, downbeat_logits) device tensors."""
        beat, down, fo = self._frames_batch(signals, sr)
        return [(beat[fo[i] : fo[i + 1]], down[fo[i] : fo[i + 1]]) for i in range(len(signals))]


class Audio2Beats(Audio2Frames):
    """Beat / downbeat positions in seconds from an audio signal (reference
    inference.py:284-303)."""

    def __init__(self, checkpoint_path="final0", device="cuda", float16=False, dbn=False, resampler="device"):
        super().__init__(checkpoint_path, device, float16, resampler)
        self.frames2beats = Postprocessor(type="dbn" if dbn else "minimal", engine=self.model.engine)

    def __call__(self, signal, sr):
        beat_logits, downbeat_logits = super().__call__(signal, sr)
        return self.frames2beats(beat_logits, downbeat_logits)

    def batch(self, signals, sr=22050):
        """list of signals -> list of (beat_times, downbeat_times) numpy float64 arrays."""
        beat, down, fo = self._frames_batch(signals, sr)
        return self.frames2beats.batch_cat(beat, down, fo)

    def batch_from_pinned(self, audio_host: torch.Tensor, sample_offsets):
        """One pinned fp32 host tensor holding all clips back to back -> beats.  This is the
        end-to-end call bench.py times (H2D copy, all kernels, D2H of the timestamps)."""
        audio = audio_host.to(self.device, non_blocking=True)
        beat, down, fo = self.model.engine.audio2frames_cat(audio, list(sample_offsets))
        return self.frames2beats.batch_cat(beat, down, fo)


class File2Beats(Audio2Beats):
    def __call__(self, audio_path):
        signal, sr = load_audio(audio_path)
        return super().__call__(signal, sr)

    def batch(self, audio_paths):
        """Many files per call; files of equal sample rate share one launch."""
        loaded = [load_audio(p) for p in audio_paths]
        out = [None] * len(loaded)
        for sr in sorted({sr for _, sr in loaded}):
            idx = [i for i, (_, s) in enumerate(loaded) if s == sr]
            res = Audio2Beats.batch(self, [loaded[i][0] for i in idx], sr)
            for i, r in zip(idx, res):
                out[i] = r
        return out


class File2File(File2Beats):
    def __call__(self, audio_path, output_path):
        downbeats, beats = super().__call__(audio_path)
        save_beat_tsv(downbeats, beats, output_path)
