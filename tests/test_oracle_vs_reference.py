"""Container-only cross-checks against the UNMODIFIED reference (skipped where /root/reference is absent, i.e. on
the GPU box). 1) the oracle is bit-identical to the reference classes on CPU; 2) the drop-in classes graft into the
reference pipeline and leave its NDJSON output unchanged -- the device call is replaced by an oracle-backed test
double here, because this container has no GPU (the real device path is covered by the -m gpu tests on the same
fixtures)."""
import io
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mel_oracle, rvq_oracle
from oracle.ref_shim import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not mounted")


@pytest.fixture(scope="module")
def nat(tmp_path_factory):
    os.environ.setdefault("HOME", str(tmp_path_factory.mktemp("home")))
    return load_reference()


def _ref_rvq(nat, seed, D, K, L):
    torch.manual_seed(seed)
    rvq = nat.ResidualVectorQuantizer(D, K, L).eval()
    for q in rvq.quantizers:
        q.use_stochastic = False
    return rvq


@pytest.mark.parametrize("D,K,L,B,T", [(64, 128, 4, 1, 50), (80, 300, 3, 2, 37), (768, 1024, 4, 1, 400)])
def test_oracle_is_bit_identical_to_reference_rvq(nat, D, K, L, B, T):
    rvq = _ref_rvq(nat, 3, D, K, L)
    x = torch.randn(B, D, T, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        q_ref, codes_ref, losses_ref = rvq(x)
        dec_ref = rvq.decode(codes_ref)
        enc_ref = rvq.encode(x)
    cbs = [q.codebook for q in rvq.quantizers]
    q, codes, losses = rvq_oracle.rvq_forward(x, cbs)
    assert torch.equal(q, q_ref) and float(losses["vq_loss"]) == float(losses_ref["vq_loss"])
    for a, b, c in zip(codes, codes_ref, enc_ref):
        assert torch.equal(a, b) and torch.equal(a, c)
    assert torch.equal(rvq_oracle.rvq_decode(codes, cbs), dec_ref)
    # 2-D input: the stack keeps the batch dim it added (nat.py:1382-1383), the single layer squeezes it (2175-2177)
    with torch.no_grad():
        q2, c2, _ = rvq(x[0])
        v_q, v_c, _ = rvq.quantizers[0](x[0])
    assert q2.dim() == 3 and c2[0].shape == (1, T) and v_q.shape == (D, T) and v_c.shape == (T,)


def test_reference_raises_value_error_like_the_mirror(nat):
    rvq = _ref_rvq(nat, 1, 8, 16, 2)
    with pytest.raises(ValueError):
        rvq(torch.randn(8))
    with pytest.raises(ValueError):
        rvq(torch.randn(1, 9, 4))


def test_spectral_oracle_vs_reference_fallback(nat):
    enc = nat.SemanticAudioEncoder(target_dim=2)
    enc.fallback_proj = torch.nn.Linear(2, 2)
    with torch.no_grad():
        enc.fallback_proj.weight.copy_(torch.eye(2))
        enc.fallback_proj.bias.zero_()
        w = torch.randn(1, 9000, generator=torch.Generator().manual_seed(2)) * 0.2
        feats = enc._spectral_fallback(w, 16000)[0].numpy()
    np.testing.assert_allclose(mel_oracle.spectral_stats(w[0].numpy(), 16000), feats, rtol=2e-4, atol=1e-2)


def test_install_into_reference_pipeline_keeps_ndjson(nat, tmp_path, monkeypatch):
    """BASELINE.json config 1 through the reference's own pipeline with the drop-in classes installed."""
    import neural_audio_tokenizer_b200 as b200
    from neural_audio_tokenizer_b200 import quantizers
    from oracle.make_golden import sine_fixture, write_wav

    def fake_native_encode(pack, codebooks, x_bct, commitment_weight, want_quantized, want_loss, exact_scan=False,
                           stats=None, code_dtype=torch.int64):
        q, codes, losses = rvq_oracle.rvq_forward(x_bct, list(codebooks), commitment_weight)
        per_layer = []
        r = x_bct
        for cb in codebooks:
            ql, _, ll = rvq_oracle.vq_layer(r, cb, commitment_weight)
            per_layer.append(ll)
            r = r - ql
        return torch.stack(codes).to(code_dtype), (q if want_quantized else None), \
            (torch.stack(per_layer) if want_loss else None)

    monkeypatch.setattr(quantizers, "_native_encode", fake_native_encode)
    g = load_golden("pipeline_tone_argmin")
    wav = str(tmp_path / "test_simple.wav")
    write_wav(wav, sine_fixture(), 22050)
    cfg = dict(semantic_dim=64, acoustic_dim=64, codebook_size=128, num_quantizers=8, n_mels=128, hop_length=512)
    pipe = nat.AudioTokenizationPipeline(sample_rate=22050, model_config=cfg, device="cpu",
                                         enable_reconstruction=False, deterministic=True, deterministic_seed=42,
                                         codebook_init_method="random", enable_codebook_cache=False, codebook_size=128)
    rng_before = torch.get_rng_state()
    b200.install(pipe.tokenizer, force_argmin=True, codes_on_cpu=True, mel=False)     # mel kernel needs the GPU
    assert torch.equal(rng_before, torch.get_rng_state()), "install() consumed global RNG (SURVEY.md RNG hygiene)"
    assert isinstance(pipe.tokenizer.semantic_quantizer, b200.ResidualVectorQuantizer)
    np.testing.assert_array_equal(
        np.stack([q.codebook.numpy() for q in pipe.tokenizer.semantic_quantizer.quantizers]), g["sem_codebooks"])
    old = sys.stdout
    sys.stdout = io.StringIO()
    try:
        result = pipe.process_audio(wav, ndjson_streaming=True)
    finally:
        sys.stdout = old
    lines = [l for l in result["ndjson_output"].splitlines() if '"event":"frame"' in l]
    frames = [json.loads(l) for l in lines]
    np.testing.assert_array_equal(np.array([f["S"] for f in frames]), g["S"])
    np.testing.assert_array_equal(np.array([f["A"] for f in frames]), g["A"])
    assert lines == [str(s) for s in g["frame_lines"]]                 # byte-identical frame events


def test_patch_reference_module_rebinds_names(nat):
    import types
    import neural_audio_tokenizer_b200 as b200
    saved = (nat.ResidualVectorQuantizer, nat.T)
    try:
        b200.patch_reference_module(nat)
        assert nat.ResidualVectorQuantizer is b200.ResidualVectorQuantizer
        assert nat.T.MelSpectrogram is b200.MelSpectrogram and hasattr(nat.T, "Resample")
    finally:
        nat.ResidualVectorQuantizer, nat.T = saved


@pytest.mark.parametrize("rle", [False, True])
def test_native_ndjson_stream_equals_live_reference(nat, rle):
    """The drop-in create_ndjson_stream, bound to a live reference StreamingProtocol, returns the reference's own
    text (header and end events included) for the same token streams."""
    from neural_audio_tokenizer_b200 import ndjson as nd
    rng = np.random.default_rng(77)
    n = 900
    sem = [torch.from_numpy(np.repeat(rng.integers(0, 1024, n // 3), 3))[None] for _ in range(4)]
    ac = [torch.from_numpy(rng.integers(0, 1024, n))[None] for _ in range(4)]
    tokens = {"semantic_codes": sem, "acoustic_codes": ac}
    kw = dict(sample_rate=22050, hop_length=512, rle_mode=rle, codebook_size=1024, keyframe_interval_seconds=2.0)
    want = nat.StreamingProtocol(**kw).create_ndjson_stream(tokens, metadata={"k": 1}, processing_stats={"n": n},
                                                            duration_seconds=n * 512 / 22050)
    got = nd.create_ndjson_stream(nat.StreamingProtocol(**kw), tokens, metadata={"k": 1}, processing_stats={"n": n},
                                  duration_seconds=n * 512 / 22050)
    assert got == want


def test_ndjson_layer_count_mismatch_behaves_like_the_reference(nat, capsys):
    """Fewer / more code streams than the protocol's layer count: the reference pads or truncates each frame with a
    printed warning (nat.py:2731-2744); the drop-in hands that malformed-input case to the reference's own method."""
    from neural_audio_tokenizer_b200 import ndjson as nd
    rng = np.random.default_rng(5)
    sem = [torch.from_numpy(rng.integers(0, 64, 20))[None] for _ in range(3)]      # protocol expects 4 + 4
    ac = [torch.from_numpy(rng.integers(0, 64, 20))[None] for _ in range(5)]
    tokens = {"semantic_codes": sem, "acoustic_codes": ac}
    kw = dict(sample_rate=22050, hop_length=512, rle_mode=False, codebook_size=64)
    want = nat.StreamingProtocol(**kw).create_ndjson_stream(tokens)
    warned = capsys.readouterr().err                       # the reference redirects print to stderr (nat.py:161-455)
    got = nd.create_ndjson_stream(nat.StreamingProtocol(**kw), tokens)
    assert got == want and capsys.readouterr().err == warned and "Warning: Expected 4 semantic tokens" in warned


def test_oracle_training_forward_is_bit_identical_to_reference(nat):
    """Training mode: every layer samples and then updates its EMA statistics and codebook (nat.py:2150, 2179-2181,
    2205-2221). The oracle restatement must leave the same codes and the same buffers behind."""
    torch.manual_seed(11)
    rvq = nat.ResidualVectorQuantizer(16, 32, 3).train()
    cbs = [q.codebook.clone() for q in rvq.quantizers]
    cnt = [q.ema_count.clone() for q in rvq.quantizers]
    wgt = [q.ema_weight.clone() for q in rvq.quantizers]
    x = torch.randn(2, 16, 40, generator=torch.Generator().manual_seed(12))
    torch.manual_seed(99)
    q_ref, codes_ref, losses_ref = rvq(x)
    torch.manual_seed(99)
    q, codes, losses = rvq_oracle.rvq_forward_training(x, cbs, cnt, wgt)
    assert torch.equal(q, q_ref.detach()) and float(losses["vq_loss"]) == float(losses_ref["vq_loss"])
    for a, b in zip(codes, codes_ref):
        assert torch.equal(a, b)
    for l, ql in enumerate(rvq.quantizers):
        assert torch.equal(cbs[l], ql.codebook) and torch.equal(cnt[l], ql.ema_count) and torch.equal(wgt[l], ql.ema_weight)
