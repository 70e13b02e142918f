"""Mint the golden vectors in tests/golden/ by running the UNMODIFIED reference in the authoring container.

TEST INFRASTRUCTURE ONLY.  Run as `python -m oracle.make_golden` from the repo root (needs /root/reference; about
a minute, dominated by the reference import).  The reference ships no golden vectors for this path,
so these files are the pin: every array below is produced by the reference's own classes
(`ResidualVectorQuantizer`, `VectorQuantizer`, `MelResidualEncoder`'s `T.MelSpectrogram`,
`SemanticAudioEncoder._spectral_fallback`, `AudioTokenizationPipeline`) with `use_stochastic=False` forced on every
layer and the module in eval(), which is the argmin contract of BASELINE.json.

Large codebooks are not stored: they are redrawn from the recorded torch seed in construction order
(nat.py:2115, `torch.randn(K, D)` per layer) and checked against a stored float64 checksum.
"""
from __future__ import annotations

import io
import json
import os
import sys
import tempfile
import wave as wave_mod

import numpy as np
import torch

from oracle.ref_shim import load_reference

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sine_fixture(sample_rate: int = 22050, freq: float = 440.0, amp: float = 0.5) -> np.ndarray:
    """The reference's own test tone (test_output_behavior.py:146-149), quantised to int16 like the shipped wav.

    floor(x * 32768) reproduces /root/reference/test_simple.wav sample for sample (checked in main()).
    """
    t = np.linspace(0, 1.0, int(sample_rate * 1.0))
    audio = np.sin(2 * np.pi * freq * t) * amp
    return np.clip(np.floor(audio * 32768.0), -32768, 32767).astype(np.int16)


def write_wav(path: str, pcm16: np.ndarray, sample_rate: int) -> None:
    with wave_mod.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm16.tobytes())


def build_ref_rvq(nat, seed: int, D: int, K: int, L: int):
    torch.manual_seed(seed)
    rvq = nat.ResidualVectorQuantizer(D, K, L).eval()
    for q in rvq.quantizers:
        q.use_stochastic = False
    return rvq


def rvq_case(nat, name: str, seed: int, D: int, K: int, L: int, B: int, T: int, store_codebooks: bool,
             x_seed: int = 1234, x_scale: float = 1.0, duplicate_rows: bool = False):
    rvq = build_ref_rvq(nat, seed, D, K, L)
    if duplicate_rows:                      # exact ties: rows 3 and 7 of layer 0 are identical, lower index must win
        rvq.quantizers[0].codebook[7].copy_(rvq.quantizers[0].codebook[3])
    g = torch.Generator().manual_seed(x_seed)
    x = torch.randn(B, D, T, generator=g) * x_scale
    if duplicate_rows:                      # a frame that IS a code vector: distance clamps to 0 (decode->encode)
        x[0, :, 0] = rvq.quantizers[0].codebook[3]
    with torch.no_grad():
        quantized, codes, losses = rvq(x)
        decoded = rvq.decode(codes)
    cbs = np.stack([q.codebook.numpy() for q in rvq.quantizers])
    out = {
        "seed": seed, "x_seed": x_seed, "D": D, "K": K, "L": L, "B": B, "T": T, "x_scale": x_scale,
        "duplicate_rows": int(duplicate_rows),
        "codes": np.stack([c.numpy() for c in codes]),                            # [L, B, T] int64
        "vq_loss": np.float32(losses["vq_loss"].item()),
        "codebook_checksum": np.float64(cbs.astype(np.float64).sum()),
        "x_checksum": np.float64(x.double().sum().item()),
        "quantized_checksum": np.float64(quantized.double().sum().item()),
        "quantized_head": quantized[:, :8, :8].numpy(),
    }
    if store_codebooks:                     # small cases carry every tensor; big ones are redrawn from the seeds
        out["codebooks"] = cbs
        out["x"] = x.numpy()
        out["quantized"] = quantized.numpy()
        out["decoded"] = decoded.numpy()
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)
    print(f"{name}: codes[0,0,:8]={out['codes'][0, 0, :8].tolist()} vq_loss={out['vq_loss']:.6f}")


def mel_case(nat, name: str, wave: np.ndarray, sr: int, hop: int):
    import torchaudio.transforms as T
    enc = nat.MelResidualEncoder(n_mels=128, n_fft=2048, hop_length=hop, target_dim=64)
    # exactly the constructor call at nat.py:2281-2287
    mel_t = T.MelSpectrogram(sample_rate=sr, n_fft=enc.n_fft, hop_length=enc.hop_length, n_mels=enc.n_mels,
                             normalized=True)
    w = torch.from_numpy(wave.astype(np.float32))[None]
    with torch.no_grad():
        mel = mel_t(w)
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), wave=wave.astype(np.float32), sr=sr, hop=hop,
                        n_fft=2048, n_mels=128, mel=mel.numpy(), fb=mel_t.mel_scale.fb.numpy())
    print(f"{name}: mel {tuple(mel.shape)} max={mel.max().item():.5g}")


def spectral_case(nat, name: str, wave: np.ndarray, sr: int):
    enc = nat.SemanticAudioEncoder(target_dim=2)          # offline: Wav2Vec2 load fails -> spectral fallback
    assert not enc.available
    enc.fallback_proj = torch.nn.Linear(2, 2)
    with torch.no_grad():
        enc.fallback_proj.weight.copy_(torch.eye(2))
        enc.fallback_proj.bias.zero_()
        feats = enc._spectral_fallback(torch.from_numpy(wave.astype(np.float32))[None], sr)   # [1, 2, T]
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), wave=wave.astype(np.float32), sr=sr,
                        stats=feats[0].numpy())
    print(f"{name}: stats {tuple(feats.shape)} first={feats[0, :, 0].tolist()}")


def pipeline_case(nat, name: str):
    """BASELINE.json config 1 at small dims: test tone -> full reference pipeline -> NDJSON, argmin mode."""
    pcm = sine_fixture()
    tmp = tempfile.mkdtemp(prefix="nat_golden_")
    wav = os.path.join(tmp, "test_simple.wav")
    write_wav(wav, pcm, 22050)
    cfg = dict(semantic_dim=64, acoustic_dim=64, codebook_size=128, num_quantizers=8, n_mels=128, hop_length=512)
    pipe = nat.AudioTokenizationPipeline(sample_rate=22050, model_config=cfg, device="cpu",
                                         enable_reconstruction=False, deterministic=True, deterministic_seed=42,
                                         codebook_init_method="random", enable_codebook_cache=False,
                                         codebook_size=128)
    tok = pipe.tokenizer
    for rvq in (tok.semantic_quantizer, tok.acoustic_quantizer):
        for q in rvq.quantizers:
            q.use_stochastic = False
    captured = {}
    h1 = tok.semantic_quantizer.register_forward_pre_hook(lambda m, a: captured.__setitem__("sem_in", a[0].clone()))
    h2 = tok.acoustic_quantizer.register_forward_pre_hook(lambda m, a: captured.__setitem__("ac_in", a[0].clone()))
    old_stdout = sys.stdout
    sys.stdout = io.StringIO()
    try:
        result = pipe.process_audio(wav, ndjson_streaming=True)
    finally:
        sys.stdout = old_stdout
    h1.remove(); h2.remove()
    lines = [l for l in result["ndjson_output"].splitlines() if l.strip()]
    frames = [json.loads(l) for l in lines if '"event":"frame"' in l]
    audio, sr = pipe.load_audio(wav)
    with torch.no_grad():
        mel = tok.acoustic_encoder.mel_transform(torch.from_numpy(audio).float()[None])
    np.savez_compressed(
        os.path.join(GOLDEN, f"{name}.npz"),
        pcm16=pcm, audio=audio.astype(np.float32), sr=sr, mel=mel.numpy(),
        sem_in=captured["sem_in"].numpy(), ac_in=captured["ac_in"].numpy(),
        sem_codebooks=np.stack([q.codebook.numpy() for q in tok.semantic_quantizer.quantizers]),
        ac_codebooks=np.stack([q.codebook.numpy() for q in tok.acoustic_quantizer.quantizers]),
        S=np.array([f["S"] for f in frames], dtype=np.int64), A=np.array([f["A"] for f in frames], dtype=np.int64),
        frame_lines=np.array([l for l in lines if '"event":"frame"' in l]),
        header_line=np.array([l for l in lines if '"event":"header"' in l][:1]),
    )
    print(f"{name}: {len(frames)} frames, first S={frames[0]['S']} A={frames[0]['A']}")


def rvq_sampling_case(nat, name: str, seed: int, D: int, K: int, L: int, T: int, noise_seed: int,
                      argmin_layers=()) -> None:
    """The reference in its DEFAULT mode (eval, use_stochastic=True, nat.py:2150-2154) with a seeded global generator:
    the known answer for the sampling path. `argmin_layers` switches individual layers to use_stochastic=False."""
    torch.manual_seed(seed)
    rvq = nat.ResidualVectorQuantizer(D, K, L).eval()
    for i in argmin_layers:
        rvq.quantizers[i].use_stochastic = False
    x = torch.randn(1, D, T, generator=torch.Generator().manual_seed(1234))
    torch.manual_seed(noise_seed)
    with torch.no_grad():
        quantized, codes, losses = rvq(x)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                        x=x.numpy(), codebooks=np.stack([q.codebook.numpy() for q in rvq.quantizers]),
                        codes=np.stack([c.numpy() for c in codes]), quantized=quantized.numpy(),
                        vq_loss=np.float32(losses["vq_loss"]), noise_seed=np.int64(noise_seed),
                        temperatures=np.array([q.temperature if q.use_stochastic else 0.0 for q in rvq.quantizers],
                                              dtype=np.float32))
    print(name, "codes[:, 0, :6] =", np.stack([c.numpy() for c in codes])[:, 0, :6].tolist())


NDJSON_CASES = [
    # name, sr, hop, rle, per_layer_encoding, keyframe interval, frames, stickiness of (semantic, acoustic) streams
    ("dense_22050", 22050, 512, False, None, 5.0, 200, (0.0, 0.0)),
    ("dense_16000_exact_ms", 16000, 320, False, None, 5.0, 60, (0.0, 0.0)),
    ("rle_default_keyframes", 22050, 512, True, None, 5.0, 1500, (0.9, 0.0)),
    ("rle_per_layer_24000", 24000, 320, True, {"S0": "rle", "S1": "rle", "S2": "dense", "S3": "dense", "A0": "rle",
                                               "A1": "dense", "A2": "dense", "A3": "dense"}, 1.0, 400, (0.8, 0.7)),
    ("rle_all_layers_constant", 22050, 512, True, {f"{k}{i}": "rle" for k in "SA" for i in range(4)}, 2.0, 300,
     (1.0, 1.0)),
    ("rle_all_layers_sticky", 44100, 441, True, {f"{k}{i}": "rle" for k in "SA" for i in range(4)}, 0.5, 350,
     (0.95, 0.9)),
    ("rle_single_frame", 22050, 512, True, None, 5.0, 1, (0.0, 0.0)),
    ("dense_no_frames", 22050, 512, False, None, 5.0, 0, (0.0, 0.0)),
]


def sticky_stream(rng, n, vocab, stick):
    """Token stream whose value repeats with probability `stick` (runs exercise the RLE duration bookkeeping)."""
    out = np.empty(n, dtype=np.int64)
    cur = int(rng.integers(vocab))
    for i in range(n):
        if i == 0 or rng.random() >= stick:
            cur = int(rng.integers(vocab))
        out[i] = cur
    return out


def ndjson_cases(nat) -> None:
    """Known-answer NDJSON bodies from the reference's own StreamingProtocol.create_ndjson_stream (nat.py:4452)."""
    cases = {}
    for name, sr, hop, rle, enc, key_s, n, (stick_s, stick_a) in NDJSON_CASES:
        rng = np.random.default_rng(sum(map(ord, name)))
        sem = [sticky_stream(rng, n, 1024, stick_s) for _ in range(4)]
        ac = [sticky_stream(rng, n, 1024, stick_a) for _ in range(4)]
        proto = nat.StreamingProtocol(sample_rate=sr, hop_length=hop, rle_mode=rle, codebook_size=1024,
                                      num_semantic_layers=4, num_acoustic_layers=4,
                                      per_layer_encoding=None if enc is None else dict(enc),
                                      keyframe_interval_seconds=key_s)
        tokens = {"semantic_codes": [torch.from_numpy(s)[None] for s in sem],
                  "acoustic_codes": [torch.from_numpy(a)[None] for a in ac]}
        text = proto.create_ndjson_stream(tokens, metadata={"case": name}, processing_stats={"n": n},
                                          duration_seconds=n * hop / sr)
        lines = text.split("\n")
        assert json.loads(lines[0])["event"] == "header" and json.loads(lines[-1])["event"] == "end"
        cases[name] = {"sr": sr, "hop": hop, "rle": rle, "per_layer_encoding": enc, "keyframe_interval_seconds": key_s,
                       "semantic": [s.tolist() for s in sem], "acoustic": [a.tolist() for a in ac],
                       "header": lines[0], "body": lines[1:-1], "end": lines[-1]}
    with open(os.path.join(GOLDEN, "ndjson_cases.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("ndjson_cases:", {k: len(v["body"]) for k, v in cases.items()})


def token_stats_cases(nat) -> None:
    """Diversity / entropy / mutual information from the reference's own TokenizationEvaluator on seeded streams."""
    ev = nat.TokenizationEvaluator(sample_rate=22050)
    cases = {}
    specs = [("uniform_1024", 1024, 5000, 0.0), ("sticky_1024", 1024, 4000, 0.9), ("narrow_40", 40, 3000, 0.5),
             ("constant", 1024, 500, 1.0), ("two_values", 2, 777, 0.3)]
    for name, vocab, n, stick in specs:
        rng = np.random.default_rng(sum(map(ord, name)))
        sem = [sticky_stream(rng, n, vocab, stick) for _ in range(4)]
        ac = [sticky_stream(rng, n - 7 * i, vocab, stick * 0.5) for i in range(4)]        # ragged lengths
        sem_t = [torch.from_numpy(s)[None] for s in sem]
        ac_t = [torch.from_numpy(a)[None] for a in ac]
        all_s = torch.cat([c.flatten().long() for c in sem_t])
        all_a = torch.cat([c.flatten().long() for c in ac_t])
        cases[name] = {
            "vocab": vocab, "semantic": [s.tolist() for s in sem], "acoustic": [a.tolist() for a in ac],
            "semantic_diversity": len(torch.unique(all_s)) / len(all_s),          # nat.py:4916
            "acoustic_diversity": len(torch.unique(all_a)) / len(all_a),
            "semantic_entropy": ev._calculate_entropy(all_s), "acoustic_entropy": ev._calculate_entropy(all_a),
            "mutual_information": ev._calculate_mutual_information(all_s, all_a)}
    with open(os.path.join(GOLDEN, "token_stats.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("token_stats:", {k: (round(v["semantic_entropy"], 4), round(v["mutual_information"], 4)) for k, v in cases.items()})


INTERP_CASES = [(44, 3), (3, 44), (100, 97), (2250, 2249), (1000, 333), (7, 7), (513, 1024), (1, 5), (4410, 4307)]


def interp_cases(nat) -> None:
    """The reference's time-base alignment call (nat.py:3230: F.interpolate(..., mode='linear', align_corners=False))
    on seeded [2, 2, T] features, through the reference module's own `F`."""
    out = {}
    for n, (t_in, t_out) in enumerate(INTERP_CASES):
        x = torch.randn(2, 2, t_in, generator=torch.Generator().manual_seed(100 + n))
        y = nat.F.interpolate(x, size=t_out, mode="linear", align_corners=False)
        out[f"x_{t_in}_{t_out}"] = x.numpy()
        out[f"y_{t_in}_{t_out}"] = y.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "interp_cases.npz"), **out)
    print("interp_cases:", len(INTERP_CASES))


def sampling_cases(nat) -> None:
    rvq_sampling_case(nat, "rvq_sampling_small", seed=7, D=64, K=128, L=4, T=300, noise_seed=99)
    rvq_sampling_case(nat, "rvq_sampling_mixed", seed=8, D=48, K=100, L=4, T=120, noise_seed=5, argmin_layers=(1, 3))
    rvq_sampling_case(nat, "rvq_sampling_wide", seed=9, D=200, K=300, L=2, T=64, noise_seed=6)


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    nat = load_reference()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    if "ndjson" in sys.argv[1:]:                              # `python -m oracle.make_golden ndjson`: only that file
        ndjson_cases(nat)
        return
    if "sampling" in sys.argv[1:]:
        sampling_cases(nat)
        return
    if "stats" in sys.argv[1:]:
        token_stats_cases(nat)
        return
    if "interp" in sys.argv[1:]:
        interp_cases(nat)
        return
    rvq_case(nat, "rvq_small", seed=7, D=64, K=128, L=4, B=1, T=50, store_codebooks=True)
    rvq_case(nat, "rvq_ragged", seed=11, D=80, K=300, L=3, B=2, T=37, store_codebooks=True)
    rvq_case(nat, "rvq_ties", seed=13, D=64, K=64, L=2, B=1, T=9, store_codebooks=True, duplicate_rows=True)
    rvq_case(nat, "rvq_single_frame", seed=17, D=128, K=256, L=4, B=1, T=1, store_codebooks=True)
    rvq_case(nat, "rvq_768x1024", seed=42, D=768, K=1024, L=4, B=1, T=1000, store_codebooks=False)
    rvq_case(nat, "rvq_512x4096", seed=43, D=512, K=4096, L=4, B=1, T=300, store_codebooks=False)
    rvq_case(nat, "rvq_1024x1024", seed=44, D=1024, K=1024, L=4, B=1, T=300, store_codebooks=False, x_scale=3.0)

    import wave as _w
    with _w.open(os.path.join(os.environ.get("NAT_REFERENCE_DIR", "/root/reference"), "test_simple.wav"), "rb") as f:
        shipped = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16)
    assert np.array_equal(shipped, sine_fixture()), "sine_fixture() no longer reproduces test_simple.wav"
    tone = sine_fixture().astype(np.float32) / 32768.0
    rng = np.random.default_rng(5)
    noise = (rng.standard_normal(24000) * 0.1).astype(np.float32)
    mel_case(nat, "mel_tone_22050_hop512", tone, 22050, 512)
    mel_case(nat, "mel_noise_24000_hop320", noise, 24000, 320)
    spectral_case(nat, "spectral_tone_22050", tone, 22050)
    spectral_case(nat, "spectral_noise_24000", noise, 24000)
    spectral_case(nat, "spectral_short", noise[:1000], 22050)         # shorter than n_fft: one zero-padded frame
    pipeline_case(nat, "pipeline_tone_argmin")
    ndjson_cases(nat)
    sampling_cases(nat)
    token_stats_cases(nat)
    interp_cases(nat)


if __name__ == "__main__":
    main()
