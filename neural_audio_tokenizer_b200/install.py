"""Graft the B200 path into a live reference tokenizer, or into the reference module's namespace.

The reference has no plugin registry; its seams are attributes and module-level names:
  * `NeuralAudioTokenizer.semantic_quantizer / acoustic_quantizer` are called as `rvq(x)` at nat.py:3239-3240;
  * `MelResidualEncoder.forward` builds `T.MelSpectrogram(...)` from the module-level alias `T` (nat.py:2281) and
    rebuilds it whenever `mel_transform.sample_rate` mismatches, so the mel seam is the name `T.MelSpectrogram`;
  * `SemanticAudioEncoder._spectral_fallback` (nat.py:2382-2442) is the semantic path whenever Wav2Vec2 is
    unavailable (always, offline: nat.py:2336-2353): its per-frame `rfft` loop is the method itself.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from .frontend import MelSpectrogram, spectral_stats
from .quantizers import ResidualVectorQuantizer


def convert_rvq(ref_rvq, force_argmin: bool = False, keep_delegate: bool = True) -> ResidualVectorQuantizer:
    """Build a drop-in from a reference `ResidualVectorQuantizer`, sharing nothing but copying every buffer.
    Does not consume the global RNG stream beyond construction (SURVEY.md "RNG stream hygiene"): the RNG state is
    saved and restored around the throw-away randn codebooks the constructor draws."""
    q0 = ref_rvq.quantizers[0]
    cpu_state = torch.get_rng_state()
    cuda_states = torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None
    new = ResidualVectorQuantizer(ref_rvq.input_dim, ref_rvq.codebook_size, ref_rvq.num_quantizers,
                                  commitment_weight=ref_rvq.commitment_weight, ema_decay=q0.ema_decay,
                                  temperature=q0.temperature, use_stochastic=q0.use_stochastic)
    torch.set_rng_state(cpu_state)
    if cuda_states is not None:
        torch.cuda.set_rng_state_all(cuda_states)
    new = new.to(q0.codebook.device)
    with torch.no_grad():
        for dst, src in zip(new.quantizers, ref_rvq.quantizers):
            dst.codebook.copy_(src.codebook)
            dst.ema_count.copy_(src.ema_count)
            dst.ema_weight.copy_(src.ema_weight)
            dst.use_stochastic = False if force_argmin else src.use_stochastic
            dst.temperature = src.temperature
    new.train(ref_rvq.training)
    new._reference_class = type(ref_rvq)           # codebook initializers keep running the reference's own code
    if keep_delegate and not force_argmin:
        new.stochastic_delegate = ref_rvq          # sampling modes keep running the unmodified reference module
    return new


def make_spectral_fallback(original):
    """`SemanticAudioEncoder._spectral_fallback` with the STFT / centroid / bandwidth loop (nat.py:2395-2433: one
    `torch.fft.rfft` launch per frame from Python) replaced by `spectral_stats` (one kernel); the lazily created
    `fallback_proj` and its draw from the global generator (nat.py:2436-2437) are kept as they are. Host tensors keep
    going to the original method."""
    def _spectral_fallback(self, waveform, sample_rate):
        if not waveform.is_cuda:
            return original(self, waveform, sample_rate)
        device = waveform.device
        features = spectral_stats(waveform, int(sample_rate), n_fft=2048, hop=512)          # [2, time]
        if self.fallback_proj is None or self.fallback_proj.weight.device != device:
            self.fallback_proj = nn.Linear(2, self.target_dim).to(device)
        projected = self.fallback_proj(features.transpose(0, 1)).transpose(0, 1)            # [target_dim, time]
        return projected.unsqueeze(0)
    return _spectral_fallback


def install(tokenizer, force_argmin: bool = False, codes_on_cpu: bool = True, mel: bool = True,
            spectral: bool = True):
    """Replace the two RVQ stacks (and the mel transform factory) of a reference `NeuralAudioTokenizer` in place.

    force_argmin=True sets use_stochastic=False on all 8 layers: the documented deviation BASELINE.json's argmin
    contract forces; without it the sampling default keeps delegating to the reference modules.
    mel=True swaps the mel transform factory (nat.py:2277-2290), spectral=True the STFT loop of the semantic
    fallback encoder (nat.py:2395-2433). Returns the tokenizer.
    """
    tokenizer.semantic_quantizer = convert_rvq(tokenizer.semantic_quantizer, force_argmin)
    tokenizer.acoustic_quantizer = convert_rvq(tokenizer.acoustic_quantizer, force_argmin)
    tokenizer.semantic_quantizer.codes_on_cpu = codes_on_cpu
    tokenizer.acoustic_quantizer.codes_on_cpu = codes_on_cpu
    if mel:
        enc = tokenizer.acoustic_encoder
        enc.mel_transform = None                    # rebuilt lazily at nat.py:2277-2287 through the patched factory

        def forward(self, waveform, sample_rate: int):
            if (self.mel_transform is None or getattr(self.mel_transform, "sample_rate", None) != sample_rate):
                self.mel_transform = MelSpectrogram(sample_rate=sample_rate, n_fft=self.n_fft,
                                                    hop_length=self.hop_length, n_mels=self.n_mels,
                                                    normalized=True).to(waveform.device)
            mel_spec = self.mel_transform(waveform)
            if mel_spec.dim() == 3:
                mel_spec = mel_spec.unsqueeze(1)
            encoded = self.proj(self.encoder(mel_spec))
            return encoded.mean(dim=2)

        enc.forward = types.MethodType(forward, enc)
    sem = getattr(tokenizer, "semantic_encoder", None)
    if spectral and sem is not None and hasattr(sem, "_spectral_fallback"):
        sem._spectral_fallback = types.MethodType(make_spectral_fallback(type(sem)._spectral_fallback), sem)
    return tokenizer


def patch_reference_module(nat_module) -> None:
    """Rebind the class names inside the imported reference module so tokenizers constructed afterwards use the B200
    path from the start. Seeded construction draws the same randn codebooks in the same order (nat.py:2115)."""
    if nat_module.ResidualVectorQuantizer is not ResidualVectorQuantizer:
        ResidualVectorQuantizer._reference_class = nat_module.ResidualVectorQuantizer   # codebook initializers
    nat_module.ResidualVectorQuantizer = ResidualVectorQuantizer
    sem_cls = getattr(nat_module, "SemanticAudioEncoder", None)
    if sem_cls is not None and not getattr(sem_cls._spectral_fallback, "_nat_b200", False):
        fb = make_spectral_fallback(sem_cls._spectral_fallback)
        fb._nat_b200 = True
        sem_cls._spectral_fallback = fb
    shim = types.SimpleNamespace(**{k: getattr(nat_module.T, k) for k in dir(nat_module.T) if not k.startswith("_")})
    shim.MelSpectrogram = MelSpectrogram
    nat_module.T = shim
    from . import ndjson
    ndjson.install(nat_module)                      # StreamingProtocol.create_ndjson_stream -> native emitter
    from . import align
    align.install(nat_module)                       # the time-base alignment F.interpolate (nat.py:3230-3236)
