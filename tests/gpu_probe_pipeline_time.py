"""Ad-hoc probe (not a test): wall time of the UNMODIFIED reference pipeline (baseline/_ref) on cuda:0 for one clip,
stock and with the drop-in installed (both RVQ stacks, mel transform, spectral fallback). The conv encoders and
everything else the reference does stay as they are. PROBE_SECS sets the clip length."""
import io, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.ref_shim import load_reference
from oracle.make_golden import write_wav
os.environ.setdefault("HOME", tempfile.mkdtemp())
nat = load_reference()
import neural_audio_tokenizer_b200 as b200

secs = float(os.environ.get("PROBE_SECS", 60))
sr = 22050
rng = np.random.default_rng(3)
t = np.arange(int(sr * secs)) / sr
audio = (0.3 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 1333 * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
wav = os.path.join(tempfile.mkdtemp(), "clip.wav")
write_wav(wav, audio, sr)
CFG = dict(semantic_dim=768, acoustic_dim=768, codebook_size=1024, num_quantizers=8, n_mels=128, hop_length=512)

def pipeline():
    return nat.AudioTokenizationPipeline(sample_rate=sr, model_config=dict(CFG), device="cuda", enable_reconstruction=False,
                                         deterministic=True, deterministic_seed=42, codebook_init_method="random",
                                         enable_codebook_cache=False, codebook_size=1024)

def run(pipe, reps=3):
    times = []
    for _ in range(reps):
        old = sys.stdout; sys.stdout = io.StringIO()
        try:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            result = pipe.process_audio(wav, ndjson_streaming=True)
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        finally:
            sys.stdout = old
    n = sum(1 for l in result["ndjson_output"].splitlines() if '"event":"frame"' in l)
    return times, n

for label, graft in (("stock reference on cuda (argmin)", None), ("drop-in installed (argmin)", dict(force_argmin=True)),
                     ("stock reference on cuda (default sampling)", "stock_sampling"),
                     ("drop-in installed (sampling, philox)", dict(force_argmin=False))):
    pipe = pipeline()
    tok = pipe.tokenizer
    if graft is None:
        for rvq in (tok.semantic_quantizer, tok.acoustic_quantizer):
            for q in rvq.quantizers:
                q.use_stochastic = False
    elif isinstance(graft, dict):
        b200.install(tok, codes_on_cpu=True, mel=True, spectral=True, **graft)
        if not graft["force_argmin"]:
            tok.semantic_quantizer.sampling_mode = "philox"; tok.acoustic_quantizer.sampling_mode = "philox"
    times, n = run(pipe)
    print(f"{label}: {secs:.0f} s clip, {n} frames: " + ", ".join(f"{x:.3f}" for x in times) + " s per process_audio call", flush=True)
