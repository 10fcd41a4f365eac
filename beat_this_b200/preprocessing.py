"""Audio loading and the log-mel frontend (mirror of reference beat_this/preprocessing.py).

``LogMelSpect`` keeps the reference's constructor defaults and call signature
(preprocessing.py:27-59) but runs the fused sm_100a kernel (frame -> Hann -> 1024-point FFT
-> |.| -> 128-band slaney mel -> log1p(1000 x)) through ``bt_logmel``.
``load_audio`` (preprocessing.py:6-24) walks the reference's decoder chain (torchaudio, soundfile, madmom -- whichever
is installed) and then two dependency-free WAV readers; the batched File2Beats path reads WAV files natively
(``bt_stage_wav_files``) and only falls back to this function for other containers.
"""
from __future__ import annotations

import math
import wave

import numpy as np
import torch

SAMPLE_RATE = 22050
N_FFT = 1024
HOP = 441
N_MELS = 128
F_MIN = 30.0
F_MAX = 11000.0


def _decode_torchaudio(path, dtype):
    import torchaudio

    waveform, sr = torchaudio.load(str(path), channels_first=False)
    return np.asanyarray(waveform.squeeze().numpy(), dtype=dtype), int(sr)


def _decode_soundfile(path, dtype):
    import soundfile

    wav, sr = soundfile.read(str(path), dtype=dtype)
    return wav, int(sr)


def _decode_madmom(path, dtype):
    import madmom.io

    wav, sr = madmom.io.load_audio_file(str(path), dtype=dtype)
    return wav, int(sr)


def _decode_wav_scipy(path, dtype):
    from scipy.io import wavfile

    sr, data = wavfile.read(str(path))
    full_scale = {np.dtype(np.int16): 32768.0, np.dtype(np.int32): 2147483648.0}
    if data.dtype in full_scale:
        wav = data.astype(dtype) / full_scale[data.dtype]
    elif data.dtype == np.uint8:
        wav = (data.astype(dtype) - 128.0) / 128.0
    else:
        wav = data.astype(dtype)
    return wav, int(sr)


def _decode_wav_stdlib(path, dtype):
    with wave.open(str(path), "rb") as w:
        sr, nch, width = w.getframerate(), w.getnchannels(), w.getsampwidth()
        raw = np.frombuffer(w.readframes(w.getnframes()), dtype=np.uint8)
    if width == 1:
        wav = (raw.astype(dtype) - 128.0) / 128.0
    else:  # little-endian signed PCM of `width` bytes: assemble in int64, sign-extend
        b = raw.reshape(-1, width).astype(np.int64)
        v = sum(b[:, k] << (8 * k) for k in range(width))
        v = np.where(v >= 1 << (8 * width - 1), v - (1 << (8 * width)), v)
        wav = v.astype(dtype) / float(1 << (8 * width - 1))
    return (wav.reshape(-1, nch) if nch > 1 else wav), int(sr)


# tried in this order; the first three are the reference's chain (preprocessing.py:6-24), the last two need nothing
# beyond scipy / the standard library and keep WAV input working where none of those packages has a decoder
AUDIO_BACKENDS = (("torchaudio", _decode_torchaudio), ("soundfile", _decode_soundfile), ("madmom", _decode_madmom),
                  ("scipy.io.wavfile", _decode_wav_scipy), ("wave", _decode_wav_stdlib))


def load_audio(path, dtype="float64"):
    """(waveform [time] or [time, channels] in [-1, 1), sample rate) like reference preprocessing.py:6-24: torchaudio,
    then soundfile, then madmom (each only if it is installed and can decode the file), then two WAV-only readers."""
    tried = []
    for name, decode in AUDIO_BACKENDS:
        try:
            return decode(path, dtype)
        except Exception as e:  # missing package, missing codec, unreadable file: next backend
            tried.append(f"{name}: {type(e).__name__}")
    raise RuntimeError(f'Could not load audio from "{path}". (' + "; ".join(tried) + ")")


# ------------------------------------------------------------------------------------------
# constants of the fused log-mel kernel
# ------------------------------------------------------------------------------------------


def _hz_to_mel_slaney(freq: float) -> float:
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    if freq >= min_log_hz:
        mels = min_log_hz / f_sp + math.log(freq / min_log_hz) / (math.log(6.4) / 27.0)
    return mels


def mel_filterbank(n_freqs=N_FFT // 2 + 1, f_min=F_MIN, f_max=F_MAX, n_mels=N_MELS, sample_rate=SAMPLE_RATE):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='slaney') restated with the
    same fp32 torch ops so that the coefficients are bit-identical to what the reference's
    MelSpectrogram holds (reference preprocessing.py:43-53)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2)
    f_sp = 200.0 / 3
    f_pts = f_sp * m_pts
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = m_pts >= min_log_mel
    f_pts[log_t] = min_log_hz * torch.exp(logstep * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))  # [n_freqs, n_mels]


def mel_constants() -> dict:
    """Window, FFT twiddles and the filterbank in CSR form (each mel band is one contiguous
    run of FFT bins) as packed parameters ``mel.*``."""
    fb = mel_filterbank().numpy()  # [513, 128]
    starts, ptr, w = [], [0], []
    for m in range(fb.shape[1]):
        nz = np.nonzero(fb[:, m])[0]
        if len(nz) == 0:
            starts.append(0)
        else:
            lo, hi = int(nz[0]), int(nz[-1]) + 1
            starts.append(lo)
            w.extend(fb[lo:hi, m].tolist())
        ptr.append(len(w))
    k = np.arange(512, dtype=np.float64)
    tw = np.stack([np.cos(2 * np.pi * k / N_FFT), -np.sin(2 * np.pi * k / N_FFT)], axis=1)
    return {
        "mel.window": torch.hann_window(N_FFT, periodic=True).numpy(),
        "mel.twiddle": tw.astype(np.float32).reshape(-1),
        "mel.fb_start": np.asarray(starts, dtype=np.float32),
        "mel.fb_ptr": np.asarray(ptr, dtype=np.float32),
        "mel.fb_w": np.asarray(w, dtype=np.float32),
    }


# ------------------------------------------------------------------------------------------------
# Resampler constants (device stand-in for ``soxr.resample(signal, in_rate=sr, out_rate=22050)``,
# reference inference.py:274-275).  soxr is a third-party C library that is absent offline, so this
# is a *restated published method* (band-limited interpolation with a Kaiser-windowed sinc, J. O.
# Smith, "Digital Audio Resampling"), designed to soxr's documented HQ targets: flat (+-1e-5 dB) to
# 0.93 of the output Nyquist, <= -124 dB from the Nyquist on.  Parity with soxr itself is UNPINNED
# (DESIGN.md section 2); the CUDA kernel is checked against a float64 direct-form evaluation.
#   y[n] = sum_j x[j] * g * h(s * (n*M/L - j)),  h(t) = rho * sinc(rho t) * kaiser_beta(t / Z), |t| <= Z
#   L/M = sr_out/sr_in in lowest terms, s = g = min(1, L/M)
RESAMPLE_ZERO_CROSSINGS = 94
RESAMPLE_BETA = 12.8
RESAMPLE_ROLLOFF = 0.9565


def resample_ratio(sr_in: int, sr_out: int = SAMPLE_RATE):
    """(L, M) with sr_out / sr_in = L / M in lowest terms."""
    sr_in, sr_out = int(sr_in), int(sr_out)
    if sr_in <= 0 or sr_out <= 0:
        raise ValueError("sample rates must be positive integers")
    g = math.gcd(sr_in, sr_out)
    return sr_out // g, sr_in // g


def resampled_length(n: int, L: int, M: int) -> int:
    """Number of output samples for n input samples: round-half-up of n * L / M."""
    return (2 * n * L + M) // (2 * M)


def resample_kernel(t):
    """h(t) of the header comment, float64, t in units of output-band zero crossings."""
    t = np.asarray(t, dtype=np.float64)
    u = np.clip(1.0 - (t / RESAMPLE_ZERO_CROSSINGS) ** 2, 0.0, None)
    w = np.i0(RESAMPLE_BETA * np.sqrt(u)) / np.i0(RESAMPLE_BETA)
    h = RESAMPLE_ROLLOFF * np.sinc(RESAMPLE_ROLLOFF * t) * w
    return np.where(np.abs(t) <= RESAMPLE_ZERO_CROSSINGS, h, 0.0)


def resample_filter_bank(sr_in: int, sr_out: int = SAMPLE_RATE):
    """Polyphase bank for the device kernel: (coef float32 [L, K], L, M, K).  Output n reads the K input
    samples q - K/2 + 1 + k (k = 0..K-1, q = floor(n M / L)) with the row phase = (n M) mod L."""
    L, M = resample_ratio(sr_in, sr_out)
    s = min(1.0, L / M)
    K = 2 * int(math.ceil(RESAMPLE_ZERO_CROSSINGS / s))
    if L * K > (1 << 26):
        raise ValueError(f"resampling {sr_in} -> {sr_out} Hz needs a {L} x {K} polyphase bank; use a rate with a "
                         "larger common divisor with the target rate")
    phase = np.arange(L, dtype=np.float64)[:, None] / L
    k = np.arange(K, dtype=np.float64)[None, :]
    coef = s * resample_kernel(s * (phase + (K // 2 - 1) - k))
    return coef.astype(np.float32), L, M, K


class LogMelSpect(torch.nn.Module):
    """Drop-in for the reference class (preprocessing.py:27-59).  Only the reference's
    default analysis parameters are implemented in the kernel; anything else raises."""

    def __init__(
        self,
        sample_rate=22050,
        n_fft=1024,
        hop_length=441,
        f_min=30,
        f_max=11000,
        n_mels=128,
        mel_scale="slaney",
        normalized="frame_length",
        power=1,
        log_multiplier=1000,
        device="cuda",
        _engine=None,
    ):
        super().__init__()
        given = (sample_rate, n_fft, hop_length, f_min, f_max, n_mels, mel_scale, normalized, power, log_multiplier)
        if given != (22050, 1024, 441, 30, 11000, 128, "slaney", "frame_length", 1, 1000):
            raise NotImplementedError("the sm_100a log-mel kernel implements the reference defaults only")
        from .engine import Engine

        self.engine = _engine if _engine is not None else Engine.mel_only(device)

    def forward(self, x):
        """Input is a waveform as a monodimensional array of shape T,
        output is a 2D log mel spectrogram of shape (F,128)."""
        return self.engine.logmel([x])[0]
