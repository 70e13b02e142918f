"""Graft the B200 path into a live reference tokenizer, or into the reference module's namespace.

The reference has no plugin registry; its seams are attributes and module-level names:
  * `NeuralAudioTokenizer.semantic_quantizer / acoustic_quantizer` are called as `rvq(x)` at nat.py:3239-3240;
  * `MelResidualEncoder.forward` builds `T.MelSpectrogram(...)` from the module-level alias `T` (nat.py:2281) and
    rebuilds it whenever `mel_transform.sample_rate` mismatches, so the mel seam is the name `T.MelSpectrogram`.
"""
from __future__ import annotations

import types

import torch

from .frontend import MelSpectrogram
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
    if keep_delegate and not force_argmin:
        new.stochastic_delegate = ref_rvq          # sampling modes keep running the unmodified reference module
    return new


def install(tokenizer, force_argmin: bool = False, codes_on_cpu: bool = True, mel: bool = True):
    """Replace the two RVQ stacks (and the mel transform factory) of a reference `NeuralAudioTokenizer` in place.

    force_argmin=True sets use_stochastic=False on all 8 layers: the documented deviation BASELINE.json's argmin
    contract forces; without it the sampling default keeps delegating to the reference modules.
    Returns the tokenizer.
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
    return tokenizer


def patch_reference_module(nat_module) -> None:
    """Rebind the class names inside the imported reference module so tokenizers constructed afterwards use the B200
    path from the start. Seeded construction draws the same randn codebooks in the same order (nat.py:2115)."""
    nat_module.ResidualVectorQuantizer = ResidualVectorQuantizer
    shim = types.SimpleNamespace(**{k: getattr(nat_module.T, k) for k in dir(nat_module.T) if not k.startswith("_")})
    shim.MelSpectrogram = MelSpectrogram
    nat_module.T = shim
    from . import ndjson
    ndjson.install(nat_module)                      # StreamingProtocol.create_ndjson_stream -> native emitter
    from . import align
    align.install(nat_module)                       # the time-base alignment F.interpolate (nat.py:3230-3236)
