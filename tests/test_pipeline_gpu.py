"""BASELINE.json config 1 on the device: the UNMODIFIED reference pipeline (carried as the git-ignored baseline/_ref,
imported through oracle/ref_shim) runs `test_simple.wav` on cuda:0 with the drop-in installed -- both RVQ stacks, the
mel transform factory and the spectral-fallback STFT all go through libnat_b200.so -- and emits NDJSON.

What is asserted: (1) the grafted pipeline runs end to end and emits the reference's frame events; (2) the index
streams equal the oracle's on the very tensors the reference handed to the quantisers (captured with forward
pre-hooks), bit for bit up to counted near-ties; (3) against the CPU-minted golden NDJSON of the same clip the frame
count is equal and token differences are reported, not hidden: the features reach the quantisers through cuDNN
convolutions and a device STFT, whose fp32 rounding differs from the CPU run that minted the golden file (mel
tolerance 2e-5 of the clip maximum, tests/test_frontend_parity_gpu.py), so a token may legitimately move."""
import io
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rvq_oracle
from oracle.ref_shim import load_reference, reference_available

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_available(), reason="reference not carried (baseline/_ref) or mounted")]

CFG = dict(semantic_dim=64, acoustic_dim=64, codebook_size=128, num_quantizers=8, n_mels=128, hop_length=512)


@pytest.fixture(scope="module")
def nat(tmp_path_factory):
    os.environ.setdefault("HOME", str(tmp_path_factory.mktemp("home")))
    return load_reference()


def _wav(tmp_path):
    from oracle.make_golden import sine_fixture, write_wav
    wav = str(tmp_path / "test_simple.wav")
    write_wav(wav, sine_fixture(), 22050)
    return wav


def _pipeline(nat):
    return nat.AudioTokenizationPipeline(sample_rate=22050, model_config=dict(CFG), device="cuda",
                                         enable_reconstruction=False, deterministic=True, deterministic_seed=42,
                                         codebook_init_method="random", enable_codebook_cache=False, codebook_size=128)


def _run(pipe, wav):
    captured = {}
    tok = pipe.tokenizer
    h1 = tok.semantic_quantizer.register_forward_pre_hook(lambda m, a: captured.__setitem__("sem_in", a[0].detach().clone()))
    h2 = tok.acoustic_quantizer.register_forward_pre_hook(lambda m, a: captured.__setitem__("ac_in", a[0].detach().clone()))
    old = sys.stdout
    sys.stdout = io.StringIO()
    try:
        result = pipe.process_audio(wav, ndjson_streaming=True)
    finally:
        sys.stdout = old
        h1.remove()
        h2.remove()
    lines = [l for l in result["ndjson_output"].splitlines() if '"event":"frame"' in l]
    return [json.loads(l) for l in lines], lines, captured


def _check_against_oracle(tok, frames, captured):
    S = np.array([f["S"] for f in frames]).T                       # [4, frames]
    A = np.array([f["A"] for f in frames]).T
    for name, rvq, x, got in (("S", tok.semantic_quantizer, captured["sem_in"], S),
                              ("A", tok.acoustic_quantizer, captured["ac_in"], A)):
        cbs = [q.codebook.detach().cpu() for q in rvq.quantizers]
        xc = x.cpu()
        ref = np.stack([c.reshape(-1).numpy() for c in rvq_oracle.rvq_forward(xc, cbs)[1]])
        rep = rvq_oracle.classify_mismatches(xc[0].T.numpy(), [c.numpy() for c in cbs], ref[:, :got.shape[1]], got)
        assert rep["real_mismatches"] == 0, (name, rep["flips"][:3])


def test_config1_reference_pipeline_on_cuda_with_the_dropin_installed(nat, tmp_path):
    import neural_audio_tokenizer_b200 as b200
    from neural_audio_tokenizer_b200 import _lib
    g = load_golden("pipeline_tone_argmin")
    pipe = _pipeline(nat)
    b200.install(pipe.tokenizer, force_argmin=True, codes_on_cpu=True, mel=True, spectral=True)
    tok = pipe.tokenizer
    assert isinstance(tok.semantic_quantizer, b200.ResidualVectorQuantizer)
    # seeded construction: the codebooks are the ones the CPU run drew
    np.testing.assert_array_equal(np.stack([q.codebook.cpu().numpy() for q in tok.semantic_quantizer.quantizers]),
                                  g["sem_codebooks"])
    lib = _lib.load()
    n0 = lib.nat_launch_count()
    frames, lines, captured = _run(pipe, _wav(tmp_path))
    assert lib.nat_launch_count() - n0 >= 4                         # mel + spectral + two stacks, at the least
    assert isinstance(tok.acoustic_encoder.mel_transform, b200.MelSpectrogram)
    assert tok.semantic_encoder.using_fallback                      # offline: the spectral fallback is the semantic path
    assert len(frames) == len(g["S"]) == 3
    for f, l in zip(frames, lines):                                 # the reference's own frame events, compact JSON
        assert set(f) >= {"event", "fi", "ts", "dur", "S", "A"} and len(f["S"]) == 4 and len(f["A"]) == 4
        assert l == json.dumps(f, separators=(",", ":"))
    _check_against_oracle(tok, frames, captured)
    moved = int((np.array([f["S"] for f in frames]) != g["S"]).sum() + (np.array([f["A"] for f in frames]) != g["A"]).sum())
    feat_err = max(float(np.abs(captured["sem_in"].cpu().numpy() - g["sem_in"]).max()),
                   float(np.abs(captured["ac_in"].cpu().numpy() - g["ac_in"]).max()))
    print(f"config 1 on cuda: {len(frames)} frames, {moved} of {8 * len(frames)} tokens differ from the CPU-minted golden "
          f"stream, quantiser inputs differ by at most {feat_err:.3e} (device convolutions / STFT vs CPU)")
    if feat_err < 1e-6:
        assert moved == 0


def test_patched_module_builds_the_same_tokenizer(nat, tmp_path):
    """patch_reference_module() before construction: the pipeline is born with the drop-in classes; same streams as
    the installed one."""
    import neural_audio_tokenizer_b200 as b200
    wav = _wav(tmp_path)
    pipe_a = _pipeline(nat)
    b200.install(pipe_a.tokenizer, force_argmin=True)
    frames_a, _, _ = _run(pipe_a, wav)
    saved = {k: getattr(nat, k) for k in ("ResidualVectorQuantizer", "T", "F")}
    saved_fb = nat.SemanticAudioEncoder._spectral_fallback
    saved_nd = nat.StreamingProtocol.create_ndjson_stream
    try:
        b200.patch_reference_module(nat)
        pipe_b = _pipeline(nat)
        tok = pipe_b.tokenizer
        assert isinstance(tok.semantic_quantizer, b200.ResidualVectorQuantizer)
        for rvq in (tok.semantic_quantizer, tok.acoustic_quantizer):
            for q in rvq.quantizers:
                q.use_stochastic = False
            rvq.codes_on_cpu = True
        frames_b, _, captured = _run(pipe_b, wav)
        _check_against_oracle(tok, frames_b, captured)
        assert [f["S"] for f in frames_b] == [f["S"] for f in frames_a]
        assert [f["A"] for f in frames_b] == [f["A"] for f in frames_a]
    finally:
        for k, v in saved.items():
            setattr(nat, k, v)
        nat.SemanticAudioEncoder._spectral_fallback = saved_fb
        nat.StreamingProtocol.create_ndjson_stream = saved_nd
