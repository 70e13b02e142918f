"""Oracle for the mel / spectral front-end.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two reference call sites are restated here in numpy:

* `MelResidualEncoder.forward`, mel transform only -- /root/reference/neural_audio_tokenizer.py:2281-2290:
  `torchaudio.transforms.MelSpectrogram(sample_rate, n_fft=2048, hop_length, n_mels=128, normalized=True)`.
  The arithmetic lives in a third-party dependency that is NOT under /root/reference: torchaudio (this image pins
  2.11.0+cu128; the reference's requirements.txt:4-5 only say torch>=1.12 / torchaudio>=0.12).  Its published
  algorithm (`functional.spectrogram` + `functional.melscale_fbanks`, HTK scale, norm=None) is:
  reflect-pad n_fft//2 each side, frames at hop, periodic Hann, one-sided DFT, divide by sqrt(sum(w^2)),
  |.|^2, then [T, n_freqs] @ fb[n_freqs, n_mels].
* `SemanticAudioEncoder._spectral_fallback`, STFT/centroid/bandwidth part -- nat.py:2395-2433.

Arithmetic is float64 inside and float32 at the boundary: the reference computes in float32 (MKL/pocketfft), so a
float64 restatement is the tighter arbiter; `tests/test_oracle_golden.py` pins it to the real torchaudio / reference
outputs stored in tests/golden/ with the tolerance written there.
"""
from __future__ import annotations

import math

import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def mel_filterbank(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0, f_max: float | None = None
                   ) -> np.ndarray:
    """HTK triangular filterbank [n_fft//2+1, n_mels] as torchaudio.functional.melscale_fbanks(norm=None) builds it.

    float32 steps mirror torch: linspace in fp32, slopes and min/max in fp32.
    """
    n_freqs = n_fft // 2 + 1
    if f_max is None:
        f_max = float(sample_rate // 2)
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=np.float64).astype(np.float32)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2, dtype=np.float64).astype(np.float32)
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0)) - np.float32(1.0))
             ).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)


def stft_power(wave: np.ndarray, n_fft: int, hop: int, center: bool = True, window_norm: bool = True) -> np.ndarray:
    """wave [S] -> power spectrogram [n_freqs, T] (float64)."""
    x = np.asarray(wave, dtype=np.float64)
    if center:
        x = np.pad(x, n_fft // 2, mode="reflect")
    T = 1 + (len(x) - n_fft) // hop
    w = hann_periodic(n_fft)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    spec = np.fft.rfft(x[idx] * w[None, :], axis=1)
    if window_norm:
        spec = spec / math.sqrt(float((w * w).sum()))
    return (spec.real ** 2 + spec.imag ** 2).T


def mel_power(wave: np.ndarray, sample_rate: int, n_fft: int = 2048, hop: int = 512, n_mels: int = 128) -> np.ndarray:
    """wave [B, S] or [S] -> mel power [B, n_mels, 1 + S//hop] float32 (nat.py:2281-2290)."""
    wave = np.asarray(wave)
    squeeze = wave.ndim == 1
    if squeeze:
        wave = wave[None]
    fb = mel_filterbank(sample_rate, n_fft, n_mels).astype(np.float64)
    out = np.stack([(stft_power(w, n_fft, hop).T @ fb).T for w in wave]).astype(np.float32)
    return out[0] if squeeze else out


def log_mel_db(mel: np.ndarray, amin: float = 1e-10) -> np.ndarray:
    """The optional log output: 10*log10(max(mel, amin)), power_to_db as the evaluator uses it (nat.py:3813)."""
    return (10.0 * np.log10(np.maximum(mel.astype(np.float64), amin))).astype(np.float32)


def spectral_stats(wave: np.ndarray, sample_rate: int, n_fft: int = 2048, hop: int = 512) -> np.ndarray:
    """wave [S] -> [2, T] (centroid, bandwidth) float32, following nat.py:2395-2433 step by step."""
    x = np.asarray(wave, dtype=np.float64).reshape(-1)
    S = len(x)
    T = 1 + (S - n_fft) // hop if S >= n_fft else 1                  # nat.py:2400-2403
    w = hann_periodic(n_fft)
    frames = np.zeros((T, n_fft), dtype=np.float64)
    for i in range(T):                                               # nat.py:2407-2415 (zero-padded tail)
        seg = x[i * hop:i * hop + n_fft]
        frames[i, :len(seg)] = seg
    mag = np.abs(np.fft.rfft(frames * w[None, :], axis=1)).T + 1e-12  # [freq, time], nat.py:2418
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sample_rate)[:, None]       # nat.py:2421
    total = mag.sum(axis=0) + 1e-8                                   # nat.py:2425
    centroid = (mag * freqs).sum(axis=0) / total                     # nat.py:2426
    bandwidth = np.sqrt((mag * (freqs - centroid[None, :]) ** 2).sum(axis=0) / total)   # nat.py:2429-2430
    return np.stack([centroid, bandwidth]).astype(np.float32)
