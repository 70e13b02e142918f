"""Drop-in front-end: mel power spectrogram and the spectral-fallback statistics, on libnat_b200.so.

Mirrors (nat.py = /root/reference/neural_audio_tokenizer.py):
  * the transform `MelResidualEncoder.forward` builds and calls at nat.py:2281-2290,
    `torchaudio.transforms.MelSpectrogram(sample_rate, n_fft, hop_length, n_mels, normalized=True)`: an object with a
    `sample_rate` attribute (the host rebuilds it when that mismatches, nat.py:2277-2279), `.to(device)`, callable on
    `[B, S]` fp32 and returning `[B, n_mels, 1 + S // hop]`;
  * the STFT / centroid / bandwidth part of `SemanticAudioEncoder._spectral_fallback`, nat.py:2395-2433.
No CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


def htk_filterbank(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0, f_max: Optional[float] = None
                   ) -> torch.Tensor:
    """[n_fft//2+1, n_mels] triangular HTK filterbank, the published torchaudio `melscale_fbanks(norm=None)` recipe
    evaluated with the same float32 torch ops so the weights are the ones the reference multiplies by."""
    n_freqs = n_fft // 2 + 1
    f_max = float(sample_rate // 2) if f_max is None else f_max
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class MelSpectrogram(nn.Module):
    """`T.MelSpectrogram(..., normalized=True)` as the reference configures it, computed by one fused CUDA kernel."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, hop_length: Optional[int] = None,
                 n_mels: int = 128, normalized: bool = True, log_mel: bool = False, **unsupported):
        super().__init__()
        if unsupported:
            raise TypeError(f"unsupported MelSpectrogram options for the B200 path: {sorted(unsupported)}")
        if n_fft != 2048:
            raise ValueError("the B200 front-end implements n_fft=2048, the value the reference hard-codes "
                             "(nat.py:2233)")
        if not normalized:
            raise ValueError("the reference always passes normalized=True (nat.py:2286)")
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.hop_length = hop_length if hop_length is not None else n_fft // 2
        self.n_mels = n_mels
        self.normalized = normalized
        self.log_mel = log_mel
        self.register_buffer("fb", htk_filterbank(sample_rate, n_fft, n_mels), persistent=False)
        self.last_log_mel = None
        self._banded = {}            # device -> (filterbank version, banded form prepared once by the library)

    def _banded_filterbank(self, dev):
        """The filterbank is a constant of the transform: its banded form (what the kernel reads) is derived once per
        device, again only if someone writes into `fb`."""
        lib = _lib.load()
        fb = self.fb if self.fb.device == dev else self.fb.to(dev)
        key = (fb.data_ptr(), fb._version)
        hit = self._banded.get(dev)
        if hit is None or hit[0] != key:
            buf = torch.empty(lib.nat_mel_filterbank_bytes(self.n_mels), dtype=torch.uint8, device=dev)
            fbc = fb.contiguous()
            _lib.check(lib.nat_mel_filterbank_prepare(fbc.data_ptr(), self.n_mels, buf.data_ptr(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
            hit = (key, buf, fb)
            self._banded[dev] = hit
        return hit[1]

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        if not waveform.is_cuda:
            raise RuntimeError(f"waveform is on {waveform.device}: the B200 front-end has no CPU fallback")
        if waveform.dtype != torch.float32:
            raise TypeError(f"expected float32 waveform, got {waveform.dtype}")
        lead = waveform.shape[:-1]
        S = waveform.shape[-1]
        w = waveform.reshape(-1, S).contiguous()
        B = w.shape[0]
        if S <= self.n_fft // 2:
            raise RuntimeError(f"reflect padding of {self.n_fft // 2} needs a longer input than {S} samples")
        dev = w.device
        T = lib.nat_mel_num_frames(S, self.hop_length)
        mel = torch.empty((B, self.n_mels, T), dtype=torch.float32, device=dev)
        logmel = torch.empty_like(mel) if self.log_mel else None
        with torch.cuda.device(dev):
            banded = self._banded_filterbank(dev)
            _lib.check(lib.nat_mel_power_banded_f32(w.data_ptr(), B, S, self.sample_rate, self.n_fft, self.hop_length,
                                                    self.n_mels, banded.data_ptr(), mel.data_ptr(),
                                                    logmel.data_ptr() if logmel is not None else None,
                                                    torch.cuda.current_stream(dev).cuda_stream))
        self.last_log_mel = logmel.reshape(lead + (self.n_mels, T)) if logmel is not None else None
        return mel.reshape(lead + (self.n_mels, T))


def spectral_stats(waveform: torch.Tensor, sample_rate: int, n_fft: int = 2048, hop: int = 512) -> torch.Tensor:
    """[2, T] (spectral centroid, bandwidth) of a mono waveform, the loop of nat.py:2405-2433 as one kernel."""
    lib = _lib.load()
    if not waveform.is_cuda:
        raise RuntimeError(f"waveform is on {waveform.device}: the B200 front-end has no CPU fallback")
    w = waveform.squeeze() if waveform.dim() > 1 else waveform       # nat.py:2388-2391
    if w.dim() != 1:
        raise ValueError(f"expected a mono waveform, got shape {tuple(waveform.shape)}")
    w = w.to(torch.float32).contiguous()
    S = w.shape[0]
    T = lib.nat_spectral_num_frames(S, n_fft, hop)
    out = torch.empty((2, T), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.nat_spectral_stats_f32(w.data_ptr(), S, sample_rate, n_fft, hop, out.data_ptr(),
                                              torch.cuda.current_stream(w.device).cuda_stream))
    return out
