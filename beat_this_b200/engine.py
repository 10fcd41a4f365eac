"""Thin Python owner of one ``bt_ctx`` (one per GPU / model).  PyTorch is used only for
device memory, streams and pinned host buffers; every FLOP runs in libbeatthis_sm100.so."""
from __future__ import annotations

import ctypes
from ctypes import c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import BT_DTYPE_H16, BT_DTYPE_F32, bt_hparams, i64_array


def _cuda_device(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(
            f"beat_this_b200 runs on an sm_100a CUDA device only (got device={device!r}); "
            "there is no CPU fallback"
        )
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class _PeakHandle:
    def __init__(self, slot, n):
        self.slot, self.n = slot, n

    def result(self):
        self.slot["event"].synchronize()
        cnt = self.slot["cnt_h"].numpy()
        times = self.slot["times_h"].numpy()
        out = [(times[0, i, : cnt[0, i]].copy(), times[1, i, : cnt[1, i]].copy()) for i in range(self.n)]
        self.slot["keep"] = None
        return out

    @property
    def d2h_bytes(self):
        return self.slot["times_h"].numel() * 8 + self.slot["cnt_h"].numel() * 4


class Engine:
    def __init__(self, packed: dict | None, hparams: dict | None, device="cuda", half: bool = False, wave_chunks: int | None = None):
        self.lib = _lib.load()
        self.device = _cuda_device(device)
        self.half = bool(half)  # 16-bit tcgen05 path (fp16 operands; see bt_act_dtype) instead of fp32 CUDA cores
        self.act_dtype = self.lib.bt_act_dtype().decode() if half else "f32"
        hp = hparams or {}
        self.hparams = dict(hp)
        chp = bt_hparams(
            int(hp.get("spect_dim", 128)),
            int(hp.get("transformer_dim", 512)),
            int(hp.get("ff_mult", 4)),
            int(hp.get("n_layers", 6)),
            int(hp.get("head_dim", 32)),
            int(hp.get("stem_dim", 32)),
            int(bool(hp.get("sum_head", True))),
            int(bool(hp.get("partial_transformers", True))),
        )
        ctx = c_void_p()
        code = self.lib.bt_create(ctypes.byref(ctx), self.device.index, ctypes.byref(chp), BT_DTYPE_H16 if half else BT_DTYPE_F32)
        _lib.check(self.lib, None, code)
        self.ctx = ctx
        self._model_ready = False
        self._resample_banks = {}  # (sr_in, sr_out) -> (coef device tensor, L, M, K)
        if packed is not None:
            self.set_params(packed)
        if wave_chunks:
            self.set_wave_chunks(wave_chunks)

    # ---- construction -----------------------------------------------------------------------
    @classmethod
    def mel_only(cls, device="cuda"):
        from .preprocessing import mel_constants

        eng = cls(None, None, device, False)
        for k, v in mel_constants().items():
            eng._set_param(k, v)
        return eng

    def _set_param(self, name: str, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        code = self.lib.bt_set_param(self.ctx, name.encode(), arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), arr.size)
        _lib.check(self.lib, self.ctx, code)

    def set_params(self, packed: dict):
        for k, v in packed.items():
            self._set_param(k, v)
        _lib.check(self.lib, self.ctx, self.lib.bt_finalize(self.ctx))
        self._model_ready = True

    def set_wave_chunks(self, n: int):
        _lib.check(self.lib, self.ctx, self.lib.bt_set_wave_chunks(self.ctx, int(n)))

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.bt_destroy(self.ctx)
            self.ctx = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.bt_launch_count(self.ctx))

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- hot path ---------------------------------------------------------------------------
    @staticmethod
    def frame_offsets(sample_offsets):
        fo = [0]
        for a, b in zip(sample_offsets[:-1], sample_offsets[1:]):
            fo.append(fo[-1] + 1 + (int(b) - int(a)) // 441)
        return fo

    def resample_cat(self, audio: torch.Tensor, sample_offsets, sr: int, sr_out: int = 22050):
        """Concatenated fp32 device audio at `sr` Hz -> (audio at `sr_out` Hz, new sample offsets): the device
        stand-in for soxr.resample (reference inference.py:274-275), see preprocessing.resample_filter_bank."""
        from . import preprocessing as P

        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()
        key = (int(sr), int(sr_out))
        if key not in self._resample_banks:
            coef, L, M, K = P.resample_filter_bank(*key)
            self._resample_banks[key] = (torch.from_numpy(coef).to(self.device).contiguous(), L, M, K)
        coef, L, M, K = self._resample_banks[key]
        so = [int(v) for v in sample_offsets]
        oo = [0]
        for i in range(len(so) - 1):
            oo.append(oo[-1] + P.resampled_length(so[i + 1] - so[i], L, M))
        out = torch.empty(max(oo[-1], 1), dtype=torch.float32, device=self.device)[: oo[-1]]
        code = self.lib.bt_resample(self.ctx, c_void_p(audio.data_ptr()), i64_array(so), len(so) - 1, c_void_p(coef.data_ptr()),
                                    L, M, K, c_void_p(out.data_ptr()), i64_array(oo), self._stream())
        _lib.check(self.lib, self.ctx, code)
        return out, oo

    def logmel_cat(self, audio: torch.Tensor, sample_offsets):
        """audio: flat fp32 device tensor; returns (spect [total_frames,128], frame_offsets)."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()
        fo = self.frame_offsets(sample_offsets)
        spect = torch.empty((fo[-1], 128), dtype=torch.float32, device=self.device)
        code = self.lib.bt_logmel(self.ctx, c_void_p(audio.data_ptr()), i64_array(sample_offsets), len(sample_offsets) - 1,
                                  c_void_p(spect.data_ptr()), i64_array(fo), self._stream())
        _lib.check(self.lib, self.ctx, code)
        return spect, fo

    def logmel(self, signals):
        sigs = [torch.as_tensor(s, dtype=torch.float32, device=self.device).contiguous() for s in signals]
        so = [0]
        for s in sigs:
            so.append(so[-1] + s.numel())
        spect, fo = self.logmel_cat(torch.cat(sigs) if len(sigs) > 1 else sigs[0], so)
        return [spect[fo[i] : fo[i + 1]] for i in range(len(sigs))]

    def spect2frames_cat(self, spect: torch.Tensor, frame_offsets):
        assert self._model_ready, "model parameters not loaded"
        assert spect.is_cuda and spect.dtype == torch.float32 and spect.is_contiguous()
        total = int(frame_offsets[-1])
        beat = torch.empty(total, dtype=torch.float32, device=self.device)
        down = torch.empty(total, dtype=torch.float32, device=self.device)
        code = self.lib.bt_spect2frames(self.ctx, c_void_p(spect.data_ptr()), i64_array(frame_offsets), len(frame_offsets) - 1,
                                        c_void_p(beat.data_ptr()), c_void_p(down.data_ptr()), self._stream())
        _lib.check(self.lib, self.ctx, code)
        return beat, down

    def forward_chunks(self, chunks: torch.Tensor):
        """BeatThis.forward on [B, T<=1500, 128] chunks (no chunk planning, no borders cut): flat (beat, downbeat)."""
        assert self._model_ready, "model parameters not loaded"
        assert chunks.is_cuda and chunks.dtype == torch.float32 and chunks.is_contiguous() and chunks.ndim == 3
        B, T, _ = chunks.shape
        beat = torch.empty(B * T, dtype=torch.float32, device=self.device)
        down = torch.empty(B * T, dtype=torch.float32, device=self.device)
        code = self.lib.bt_forward_chunks(self.ctx, c_void_p(chunks.data_ptr()), B, T, c_void_p(beat.data_ptr()),
                                          c_void_p(down.data_ptr()), self._stream())
        _lib.check(self.lib, self.ctx, code)
        return beat, down

    def tap_chunks(self, name: str, chunks: torch.Tensor, capacity: int):
        """Test hook: activation `name` of a forward_chunks call (single wave)."""
        buf = torch.zeros(capacity, dtype=torch.float32, device=self.device)
        _lib.check(self.lib, self.ctx, self.lib.bt_debug_request_tap(self.ctx, name.encode(), c_void_p(buf.data_ptr()), capacity))
        try:
            out = self.forward_chunks(chunks)
            torch.cuda.synchronize(self.device)
            n = int(self.lib.bt_debug_tap_count(self.ctx))
        finally:
            self.lib.bt_debug_request_tap(self.ctx, b"", None, 0)
        return buf[:n], out

    def audio2frames_cat(self, audio: torch.Tensor, sample_offsets):
        assert self._model_ready, "model parameters not loaded"
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()
        fo = self.frame_offsets(sample_offsets)
        beat = torch.empty(fo[-1], dtype=torch.float32, device=self.device)
        down = torch.empty(fo[-1], dtype=torch.float32, device=self.device)
        code = self.lib.bt_audio2frames(self.ctx, c_void_p(audio.data_ptr()), i64_array(sample_offsets), len(sample_offsets) - 1,
                                        c_void_p(beat.data_ptr()), c_void_p(down.data_ptr()), i64_array(fo), self._stream())
        _lib.check(self.lib, self.ctx, code)
        return beat, down, fo

    def peakpick_cat(self, beat: torch.Tensor, down: torch.Tensor, frame_offsets):
        """Minimal postprocessor on device; returns a list of (beat_times, downbeat_times)
        float64 numpy arrays, one pair per clip."""
        n = len(frame_offsets) - 1
        if n == 0:
            return []
        max_peaks = max(1, max(int(frame_offsets[i + 1]) - int(frame_offsets[i]) for i in range(n)))
        bt_t = torch.empty((n, max_peaks), dtype=torch.float64, device=self.device)
        dn_t = torch.empty((n, max_peaks), dtype=torch.float64, device=self.device)
        cnt = torch.zeros((2, n), dtype=torch.int32, device=self.device)
        code = self.lib.bt_peakpick(self.ctx, c_void_p(beat.data_ptr()), c_void_p(down.data_ptr()), i64_array(frame_offsets), n,
                                    c_void_p(bt_t.data_ptr()), c_void_p(cnt[0].data_ptr()), c_void_p(dn_t.data_ptr()),
                                    c_void_p(cnt[1].data_ptr()), max_peaks, self._stream())
        _lib.check(self.lib, self.ctx, code)
        cnt_h = cnt.cpu().numpy()
        width = int(cnt_h.max()) if cnt_h.size else 0
        if width > max_peaks:
            raise _lib.BTError("peak buffer overflow (internal error)")
        bt_h = bt_t[:, :width].cpu().numpy()
        dn_h = dn_t[:, :width].cpu().numpy()
        return [(bt_h[i, : cnt_h[0, i]].copy(), dn_h[i, : cnt_h[1, i]].copy()) for i in range(n)]

    def peakpick_async(self, beat: torch.Tensor, down: torch.Tensor, frame_offsets, slot: dict | None = None):
        """Like peakpick_cat but without a host synchronisation: the timestamp arrays are copied to
        pinned host memory asynchronously on the current stream; call ``.result()`` on the returned
        handle (it waits on a CUDA event) to get the per-clip numpy arrays."""
        n = len(frame_offsets) - 1
        max_peaks = max(1, max(int(frame_offsets[i + 1]) - int(frame_offsets[i]) for i in range(n)))
        slot = slot if slot is not None else {}
        key = (n, max_peaks)
        if slot.get("key") != key:
            slot["key"] = key
            slot["times"] = torch.empty((2, n, max_peaks), dtype=torch.float64, device=self.device)
            slot["cnt"] = torch.zeros((2, n), dtype=torch.int32, device=self.device)
            slot["times_h"] = torch.empty((2, n, max_peaks), dtype=torch.float64).pin_memory()
            slot["cnt_h"] = torch.empty((2, n), dtype=torch.int32).pin_memory()
            slot["event"] = torch.cuda.Event()
        t, cnt = slot["times"], slot["cnt"]
        code = self.lib.bt_peakpick(self.ctx, c_void_p(beat.data_ptr()), c_void_p(down.data_ptr()), i64_array(frame_offsets), n,
                                    c_void_p(t[0].data_ptr()), c_void_p(cnt[0].data_ptr()), c_void_p(t[1].data_ptr()),
                                    c_void_p(cnt[1].data_ptr()), max_peaks, self._stream())
        _lib.check(self.lib, self.ctx, code)
        slot["times_h"].copy_(t, non_blocking=True)
        slot["cnt_h"].copy_(cnt, non_blocking=True)
        slot["event"].record(torch.cuda.current_stream(self.device))
        slot["keep"] = (beat, down)  # keep the logits alive until the kernels have run
        return _PeakHandle(slot, n)

    # ---- per-kernel-class timing (bench.py roofline) -------------------------------------------
    def profile_enable(self, on: bool = True):
        _lib.check(self.lib, self.ctx, self.lib.bt_profile_enable(self.ctx, int(on)))

    def profile_reset(self):
        _lib.check(self.lib, self.ctx, self.lib.bt_profile_reset(self.ctx))

    def profile_results(self) -> dict:
        """{kernel class: (total device ms, launches)} accumulated since the last reset."""
        _lib.check(self.lib, self.ctx, self.lib.bt_profile_collect(self.ctx))
        out = {}
        for i in range(int(self.lib.bt_profile_count(self.ctx))):
            name = ctypes.create_string_buffer(64)
            ms, n = ctypes.c_double(), ctypes.c_int64()
            self.lib.bt_profile_get(self.ctx, i, name, 64, ctypes.byref(ms), ctypes.byref(n))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    # ---- test hooks --------------------------------------------------------------------------
    def tap(self, name: str, spect: torch.Tensor, frame_offsets, capacity: int):
        buf = torch.zeros(capacity, dtype=torch.float32, device=self.device)
        _lib.check(self.lib, self.ctx, self.lib.bt_debug_request_tap(self.ctx, name.encode(), c_void_p(buf.data_ptr()), capacity))
        try:
            out = self.spect2frames_cat(spect, frame_offsets)
            torch.cuda.synchronize(self.device)
            n = int(self.lib.bt_debug_tap_count(self.ctx))
        finally:
            self.lib.bt_debug_request_tap(self.ctx, b"", None, 0)
        return buf[:n], out

    def debug_gemm(self, a: torch.Tensor, w: torch.Tensor):
        M, K = a.shape
        N = w.shape[0]
        d = torch.empty((M, N), dtype=torch.float32, device=self.device)
        code = self.lib.bt_debug_gemm(self.ctx, c_void_p(a.data_ptr()), c_void_p(w.data_ptr()), c_void_p(d.data_ptr()), M, N, K, self._stream())
        _lib.check(self.lib, self.ctx, code)
        return d

    def debug_attention(self, q, k, v):
        seqs, L, C = q.shape
        o = torch.empty_like(q)
        code = self.lib.bt_debug_attention(self.ctx, c_void_p(q.data_ptr()), c_void_p(k.data_ptr()), c_void_p(v.data_ptr()),
                                           c_void_p(o.data_ptr()), seqs, L, C // 32, self._stream())
        _lib.check(self.lib, self.ctx, code)
        return o
