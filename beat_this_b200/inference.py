"""Drop-in mirror of the reference ``beat_this.inference`` API (reference beat_this/inference.py:16-315): same
function and class names, constructor and call signatures, return types and exceptions -- with everything between
"audio samples" and "beat timestamps" executed by the sm_100a CUDA library.

Differences a user can observe:
* ``device`` must be a CUDA device (default ``"cuda"``); ``device="cpu"`` raises.
* ``float16=False`` -> fp32 CUDA-core kernels (reference-exact numerics, <=1e-3 on logits);
  ``float16=True`` -> fp16-operand tcgen05 tensor-core kernels with fp32 accumulation and an fp32 residual stream
  (the reference autocasts to fp16 here as well, inference.py:245-246).
* every class has a ``batch(...)`` method that processes many clips per call through the host/device pipeline of
  ``beat_this_b200.pipeline`` (the reference is strictly one clip, one chunk at a time: inference.py:215).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from .engine import Engine
from .pipeline import BeatPipeline, as_signal_array, chunk_cost, plan_groups
from .postprocessor import Postprocessor
from .preprocessing import LogMelSpect, load_audio
from .utils import replace_state_dict_key, save_beat_tsv
from .weights import filter_hparams, pack_parameters

CHECKPOINT_URL = "https://cloud.cp.jku.at/public.php/dav/files/7ik4RrBKTS273gp"

# one group = one pass of every kernel over up to GROUP_CHUNKS model passes (chunks of 1500 frames): 64 clips of 30 s
# (two chunks each, inference.py:119-125), one full wave of the C library
GROUP_CHUNKS = 128
GROUP_CLIPS = 256


def _checkpoint_source(name) -> tuple[str, str | None]:
    """(url, cache file name) for a checkpoint that is not a local file: a full URL is taken as is, anything else is
    a short name (``final0``, ``small1`` ...) below CHECKPOINT_URL (reference inference.py:34-45)."""
    name = str(name)
    if name.startswith(("https://", "http://")):
        return name, None
    return f"{CHECKPOINT_URL}/{name}.ckpt", f"beat_this-{name}.ckpt"


def load_checkpoint(checkpoint_path: str, device: str | torch.device = "cpu") -> dict:
    """Checkpoint dictionary from a local file, a short name or a URL (reference inference.py:16-53; names and URLs
    go through the torch.hub cache and need network).  ``ValueError`` when nothing can be loaded."""
    try:
        return torch.load(checkpoint_path, map_location=device, weights_only=True)
    except FileNotFoundError:
        pass
    url, file_name = _checkpoint_source(checkpoint_path)
    try:
        return torch.hub.load_state_dict_from_url(url, file_name=file_name, map_location=device)
    except Exception:
        raise ValueError("Could not load the checkpoint given the provided name", checkpoint_path)


class BeatThisB200:
    """What ``load_model`` returns in place of the reference ``BeatThis`` nn.Module: the packed
    weights living on one GPU inside a ``bt_ctx``."""

    def __init__(self, hparams: dict, packed: dict, device, float16: bool = False, wave_chunks: int | None = None):
        self.hparams = filter_hparams(hparams)
        self.engine = Engine(packed, self.hparams, device, half=float16, wave_chunks=wave_chunks)
        self.device = self.engine.device
        self.float16 = float16

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("beat_this_b200 models live on a CUDA device; there is no CPU fallback")
        return self

    def __call__(self, spect: torch.Tensor) -> dict:
        """BeatThis.forward (reference beat_tracker.py:188-192) for a batch of equal-length spectrogram chunks
        [B, T<=1500, 128]: every chunk runs as its own piece without borders being cut."""
        if spect.ndim != 3 or spect.shape[2] != 128 or spect.shape[1] > 1500:
            raise ValueError(f"Expected [B, T<=1500, 128] chunks, got {tuple(spect.shape)}")
        B, T, _ = spect.shape
        beat, down = self.engine.forward_chunks(spect.to(self.device, torch.float32).contiguous())
        return {"beat": beat.view(B, T), "downbeat": down.view(B, T)}


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


# ------------------------------------------------------------------------------------------------------------
# Chunking helpers of the reference API (inference.py:90-230).  The CUDA path never calls them -- bt_plan_chunks
# plans natively and the stem / head kernels gather and scatter in place -- they serve users of the reference's
# function-level API and are pinned to the reference by tests/golden/chunking.npz.
# ------------------------------------------------------------------------------------------------------------
def chunk_starts(n_frames: int, chunk_size: int, border_size: int = 6, avoid_short_end: bool = True) -> np.ndarray:
    """First frame of every chunk: windows advance by chunk_size - 2*border_size from -border_size; with
    `avoid_short_end` the last window is pulled back so that it ends border_size frames past the piece
    (same plan as bt_plan_chunks for 1500 / 6)."""
    step = chunk_size - 2 * border_size
    if step <= 0:
        raise ValueError("chunk_size must exceed twice the border")
    count = max(0, math.ceil(n_frames / step))
    starts = step * np.arange(count, dtype=np.int64) - border_size
    if avoid_short_end and count and n_frames > step:
        starts[-1] = n_frames + border_size - chunk_size
    return starts


def zeropad(spect: torch.Tensor, left: int = 0, right: int = 0) -> torch.Tensor:
    """`left` / `right` zero frames around a [T, F] spectrogram (reference inference.py:90-97)."""
    if left <= 0 and right <= 0:
        return spect
    T = spect.shape[0]
    out = spect.new_zeros((left + T + right,) + tuple(spect.shape[1:]))
    out[left : left + T] = spect
    return out


def split_piece(spect: torch.Tensor, chunk_size: int, border_size: int = 6, avoid_short_end: bool = True):
    """Chunks (zero padded by up to border_size frames at the piece boundaries) and their start frames, as reference
    inference.py:100-135 returns them; all chunks are views of ONE padded copy of the piece."""
    T = len(spect)
    starts = chunk_starts(T, chunk_size, border_size, avoid_short_end)
    padded = zeropad(spect, border_size, border_size)  # frame t lives at row t + border_size
    chunks = [padded[s + border_size : min(s + chunk_size, T + border_size) + border_size] for s in starts.tolist()]
    return chunks, starts


def aggregate_prediction(pred_chunks: list, starts: list, full_size: int, chunk_size: int, border_size: int,
                         overlap_mode: str, device: str | torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """Piece-level logits from chunk-level ones (reference inference.py:138-185): every chunk loses border_size frames
    on both sides, overlaps go to the earlier ("keep_first") or later ("keep_last") chunk, frames no chunk covers
    stay at -1000.  Every frame is written once, from the chunk that owns it."""
    if overlap_mode not in ("keep_first", "keep_last"):
        raise ValueError("overlap_mode must be 'keep_first' or 'keep_last'")
    beat = torch.full((full_size,), -1000.0, device=device)
    downbeat = torch.full((full_size,), -1000.0, device=device)
    starts = [int(s) for s in starts]
    spans = [(s + border_size, s + len(p["beat"]) - border_size) for s, p in zip(starts, pred_chunks)]
    order = range(len(spans)) if overlap_mode == "keep_first" else range(len(spans) - 1, -1, -1)
    claimed_lo, claimed_hi = None, None  # frames already owned: one interval, chunks are visited in order
    for i in order:
        lo, hi = max(spans[i][0], 0), min(spans[i][1], full_size)
        if claimed_lo is not None:
            if overlap_mode == "keep_first":
                lo = max(lo, claimed_hi)
            else:
                hi = min(hi, claimed_lo)
        if hi > lo:
            off = lo - starts[i]
            beat[lo:hi] = pred_chunks[i]["beat"][off : off + hi - lo]
            downbeat[lo:hi] = pred_chunks[i]["downbeat"][off : off + hi - lo]
            claimed_lo = lo if claimed_lo is None else min(claimed_lo, lo)
            claimed_hi = hi if claimed_hi is None else max(claimed_hi, hi)
    return beat, downbeat


def split_predict_aggregate(spect: torch.Tensor, chunk_size: int, border_size: int, overlap_mode: str,
                            model) -> dict:
    """Reference inference.py:188-230: chunk the piece, run `model` on every chunk, stitch.  A ``BeatThisB200`` model
    with the standard 1500 / 6 / keep_first setting runs as ONE call of the CUDA path (all chunks batched, chunking
    and stitching inside the kernels); anything else goes chunk by chunk through the functions above."""
    if isinstance(model, BeatThisB200) and (chunk_size, border_size, overlap_mode) == (1500, 6, "keep_first"):
        spect = torch.as_tensor(spect, dtype=torch.float32, device=model.device).contiguous()
        beat, down = model.engine.spect2frames_cat(spect, [0, spect.shape[0]])
        return {"beat": beat, "downbeat": down}
    chunks, starts = split_piece(spect, chunk_size, border_size=border_size, avoid_short_end=True)
    preds = []
    for chunk in chunks:
        out = model(chunk.unsqueeze(0))
        preds.append({"beat": out["beat"][0], "downbeat": out["downbeat"][0]})
    beat, down = aggregate_prediction(preds, starts, spect.shape[0], chunk_size, border_size, overlap_mode, spect.device)
    return {"beat": beat, "downbeat": down}


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

    _want = "frames"

    def __init__(self, checkpoint_path="final0", device="cuda", float16=False, resampler="device"):
        super().__init__(checkpoint_path, device, float16)
        self._init_front(resampler)

    def _init_front(self, resampler="device"):
        if resampler not in ("device", "soxr"):
            raise ValueError("resampler must be 'device' or 'soxr'")
        self.resampler = resampler
        self.spect = LogMelSpect(device=self.device, _engine=self.model.engine)
        self._pipe = None

    @classmethod
    def from_model(cls, model: BeatThisB200, **kw):
        """Wrap an already loaded model (e.g. one whose weights arrived by broadcast, distributed.py)."""
        self = cls.__new__(cls)
        self.model, self.device, self.float16 = model, model.device, model.float16
        self._init_front(kw.pop("resampler", "device"))
        self._init_post(**kw)
        return self

    def _init_post(self):
        pass

    @property
    def pipeline(self) -> BeatPipeline:
        if self._pipe is None:
            self._pipe = BeatPipeline(self.model.engine)
        return self._pipe

    # ---- one clip (the reference's call signatures) ---------------------------------------------------
    def _prepare(self, signals, sr):
        arrays = [as_signal_array(s) for s in signals]
        if int(sr) != 22050 and self.resampler == "soxr":  # the reference's order: mono mix, then soxr (inference.py:270-275)
            arrays = [np.ascontiguousarray(_soxr_resample(a if a.ndim == 1 else a.mean(1), sr)) for a in arrays]
            sr = 22050
        return arrays, int(sr)

    def signal2spect(self, signal, sr):
        arrays, sr = self._prepare([signal], sr)
        n = arrays[0].shape[0]
        host = torch.empty(max(n, 1), dtype=torch.float32, pin_memory=True)
        so = self.pipeline.stage_signals(arrays, host)
        audio = host[:n].to(self.device)
        if sr != 22050:
            audio, so = self.model.engine.resample_cat(audio, so, sr)
        spect, _ = self.model.engine.logmel_cat(audio, so)
        return spect

    def __call__(self, signal, sr):
        (beat, down), = Audio2Frames.batch(self, [signal], sr)
        return beat, down

    # ---- many clips ---------------------------------------------------------------------------------------
    def _run_groups(self, arrays, sr, want):
        """Generator over groups: (first index, last index + 1, pipeline result)."""
        pipe = self.pipeline
        groups = plan_groups([chunk_cost(a.shape[0], sr) for a in arrays], GROUP_CHUNKS, GROUP_CLIPS)
        try:
            results = pipe.run(len(groups), lambda g: pipe.submit_signals(arrays[groups[g][0] : groups[g][1]], sr, want))
            for (lo, hi), res in zip(groups, results):
                yield lo, hi, res
        finally:
            pipe.drain()

    def batch(self, signals, sr=22050):
        """list of signals (1-D or (time, channels) arrays, `sr` Hz) -> list of (beat_logits, downbeat_logits) device
        tensors; staging, copies and kernels of consecutive groups of clips overlap."""
        arrays, sr = self._prepare(signals, sr)
        out = [None] * len(arrays)
        for lo, hi, (beat, down, fo) in self._run_groups(arrays, sr, "frames"):
            for i in range(lo, hi):
                out[i] = (beat[fo[i - lo] : fo[i - lo + 1]], down[fo[i - lo] : fo[i - lo + 1]])
        return out

    def frames_from_device(self, audio: torch.Tensor, sample_offsets):
        """Audio already on the device (flat mono fp32 tensor at 22.05 kHz + offsets)."""
        return self.model.engine.audio2frames_cat(audio, list(sample_offsets))


class Audio2Beats(Audio2Frames):
    """Beat / downbeat positions in seconds from an audio signal (reference
    inference.py:284-303)."""

    def __init__(self, checkpoint_path="final0", device="cuda", float16=False, dbn=False, resampler="device"):
        super().__init__(checkpoint_path, device, float16, resampler)
        self._init_post(dbn)

    def _init_post(self, dbn=False):
        self.frames2beats = Postprocessor(type="dbn" if dbn else "minimal", engine=self.model.engine)

    def __call__(self, signal, sr):
        return Audio2Beats.batch(self, [signal], sr)[0]

    def _finish(self, res):
        """pipeline result of one group -> list of (beat_times, downbeat_times)."""
        if self.frames2beats.type == "minimal":
            return res
        import time

        beat_h, down_h, fo = res
        t0 = time.perf_counter()
        out = self.frames2beats.batch_host(beat_h, down_h, fo)
        st = self.pipeline.stats
        st["post_s"] = st.get("post_s", 0.0) + time.perf_counter() - t0
        return out

    @property
    def _want_beats(self):
        return "beats" if self.frames2beats.type == "minimal" else "logits_host"

    def batch(self, signals, sr=22050):
        """list of signals -> list of (beat_times, downbeat_times) numpy float64 arrays.  With the DBN, the host
        Viterbi of group g runs while the GPU works on group g+1."""
        arrays, sr = self._prepare(signals, sr)
        out = [None] * len(arrays)
        if self.frames2beats.type == "minimal":
            for lo, hi, res in self._run_groups(arrays, sr, "beats"):
                out[lo:hi] = res
            return out
        # DBN: the host Viterbi of a group runs on a worker thread (numpy and the C++ tracker release the GIL), so the
        # calling thread is free to stage and enqueue the next groups
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=1) as post:
            pending = [(lo, hi, post.submit(self._finish, res)) for lo, hi, res in self._run_groups(arrays, sr, "logits_host")]
            for lo, hi, fut in pending:
                out[lo:hi] = fut.result()
        return out


class File2Beats(Audio2Beats):
    def __call__(self, audio_path):
        signal, sr = load_audio(audio_path)
        return super().__call__(signal, sr)

    def probe(self, audio_paths):
        """bt_wav_probe on every path: (ctypes array of bt_wav_info, list of ok flags).  Files that are not plain
        WAV are decoded by load_audio's backend chain instead."""
        from ._lib import bt_wav_info

        lib = self.model.engine.lib
        infos = (bt_wav_info * len(audio_paths))()
        ok = [lib.bt_wav_probe(str(p).encode(), ctypes.byref(infos[i])) == 0 and infos[i].frames > 0
              for i, p in enumerate(audio_paths)]
        return infos, ok

    def batch(self, audio_paths, on_error: str = "raise"):
        """Many files per call.  WAV files are read, mixed to mono and cast by the native host threads straight into
        the pinned staging ring (no numpy round trip); other containers go through load_audio.  Files of equal
        sample rate share groups.  on_error: "raise", or "skip" (a file that cannot be loaded or processed yields
        None instead of aborting the call -- the behaviour of the reference's per-file loop, cli.py:185-190)."""
        from ._lib import bt_wav_info

        if on_error not in ("raise", "skip"):
            raise ValueError("on_error must be 'raise' or 'skip'")
        paths = [str(p) for p in audio_paths]
        out = [None] * len(paths)
        infos, is_wav = self.probe(paths)
        for i in range(len(paths)):  # a clip needs more than 512 samples at 22.05 kHz (reflect padding of the STFT)
            if is_wav[i] and infos[i].frames * 22050 // max(1, infos[i].sample_rate) <= 512:
                if on_error == "raise":
                    raise ValueError(f'"{paths[i]}" is too short ({infos[i].frames} samples)')
                is_wav[i] = False
                infos[i].frames = -1  # marks "known bad": not retried through load_audio either
        pipe = self.pipeline
        want = self._want_beats
        for sr in sorted({infos[i].sample_rate for i in range(len(paths)) if is_wav[i]}):
            idx = [i for i in range(len(paths)) if is_wav[i] and infos[i].sample_rate == sr]
            groups = plan_groups([chunk_cost(infos[i].frames, sr) for i in idx], GROUP_CHUNKS, GROUP_CLIPS)

            def submit(g, idx=idx, groups=groups, sr=sr):
                sel = idx[groups[g][0] : groups[g][1]]
                sub = (bt_wav_info * len(sel))(*[infos[i] for i in sel])
                pipe.submit_wavs([paths[i] for i in sel], sub, sr, want)

            try:
                for (lo, hi), res in zip(groups, pipe.run(len(groups), submit)):
                    for k, r in zip(idx[lo:hi], self._finish(res)):
                        out[k] = r
            except Exception:
                pipe.drain()
                if on_error == "raise":
                    raise
                for i in idx:  # isolate the failure: one file at a time
                    if out[i] is None:
                        try:
                            out[i] = File2Beats.__call__(self, paths[i])
                        except Exception:
                            out[i] = None
        rest = [i for i in range(len(paths)) if not is_wav[i] and infos[i].frames >= 0]
        loaded = {}
        for i in rest:
            try:
                loaded[i] = load_audio(paths[i])
            except Exception:
                if on_error == "raise":
                    raise
        for sr in sorted({s for _, s in loaded.values()}):
            idx = [i for i in loaded if loaded[i][1] == sr]
            try:
                res = Audio2Beats.batch(self, [loaded[i][0] for i in idx], sr)
            except Exception:
                if on_error == "raise":
                    raise
                res = []
                for i in idx:
                    try:
                        res.append(Audio2Beats.__call__(self, *loaded[i]))
                    except Exception:
                        res.append(None)
            for i, r in zip(idx, res):
                out[i] = r
        return out


class File2File(File2Beats):
    def __call__(self, audio_path, output_path):
        beats, downbeats = File2Beats.__call__(self, audio_path)
        save_beat_tsv(beats, downbeats, output_path)
