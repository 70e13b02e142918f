"""Ad-hoc probe (not a test): the BASELINE.json configs that are not the bench line, with the current build.

config 4  1024-d features, codebook 1024 / 4096 (and 768-d x 4096): 4-layer stack kernel time, both stacks per step;
          mel front-end (24 kHz, hop 320, one hour) followed by the acoustic stack on one stream without a host sync
config 5  1 s chunks (75 frames x 768-d, 4 + 4 layers, batch 1), 1000 chunks after warm-up: p50 / p99 of the device
          time (CUDA events), of the boundary call including the bulk D2H of the 8 index streams, and of a CUDA-graph
          replay of the same chunk path
Output goes to stdout (profiles/r2_configs.log is this output)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_audio_tokenizer_b200 import _lib, MelSpectrogram, ResidualVectorQuantizer, encode_stacks
lib = _lib.load()
N = 270000
st = torch.cuda.current_stream().cuda_stream
PEAK = float(os.environ.get("PROBE_PEAK_TFLOPS", 1644.8))


def stacks_for(D, K):
    torch.manual_seed(42)
    return [ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda() for _ in range(2)]


print("== config 4: stack kernel per shape (one 4-layer stack over 270 000 frames; both stacks per step)")
for D, K in ((768, 1024), (1024, 1024), (1024, 4096), (768, 4096)):
    stacks = stacks_for(D, K)
    x = torch.randn(1, D, N, device="cuda")
    hs = [s._pack.get(s._codebooks()) for s in stacks]
    harr = (ctypes.c_void_p * 2)(*[h.value for h in hs])
    xarr = (ctypes.c_void_p * 2)(x.data_ptr(), x.data_ptr())
    wsb = lib.nat_rvq_stacks_workspace_bytes(harr, 2, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    codes = torch.empty((8, N), dtype=torch.int16, device="cuda")
    prof = (ctypes.c_float * 8)()
    best, prep = 1e9, 0.0
    for rep in range(4):
        _lib.check(lib.nat_rvq_encode_stacks_profile_f32(harr, 2, xarr, 0, 1, N, codes.data_ptr(), 2, ws.data_ptr(), wsb, 0, st, prof))
        if rep and prof[1] / max(prof[6], 1) < best: best, prep = prof[1] / max(prof[6], 1), prof[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(lib.nat_rvq_encode_stacks_f32(harr, 2, xarr, 0, 1, N, codes.data_ptr(), 2, ws.data_ptr(), wsb, 0, st))
    e1.record(); torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 5
    tf = 2.0 * K * D * N * 4 / (best * 1e-3) / 1e12
    print(f"D={D} K={K}: stack kernel {best:.3f} ms per 4-layer launch = {tf:.0f} TFLOP/s = {tf / PEAK:.3f} of burst bf16; "
          f"prep {prep:.3f} ms; 8-layer step {step:.3f} ms = {N / step / 1e3:.1f} M frames/s", flush=True)
    del x, ws, codes, stacks

print("== config 4: mel front-end then the acoustic stack (1024-d, K 1024), one stream, no host sync in between")
sr, hop = 24000, 320
S = sr * 3600
wave = torch.randn(1, S, device="cuda") * 0.1
mt = MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128).cuda()
T = 1 + S // hop
ac = stacks_for(1024, 1024)[1]
feats = torch.randn(1, 1024, T, device="cuda")           # stands in for the conv encoders between mel and RVQ
def pair():
    mel = mt(wave)
    return mel, ac.encode(feats)
pair(); torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
ms_mel, ms_both = [], []
for _ in range(5):
    e0.record(); mel = mt(wave); e1.record(); c = ac.encode(feats); e2.record(); torch.cuda.synchronize()
    ms_mel.append(e0.elapsed_time(e1)); ms_both.append(e0.elapsed_time(e2))
print(f"mel {T} frames: {min(ms_mel):.3f} ms; mel + acoustic 4-layer stack back to back: {min(ms_both):.3f} ms "
      f"({T / min(ms_both) / 1e3:.1f} M frames/s through both)", flush=True)
del wave, feats

print("== config 5: 1 s chunks (75 frames x 768-d, 4 + 4 layers, codebook 1024, batch 1), 1000 chunks")
D, K, Tc = 768, 1024, 75
stacks = stacks_for(D, K)
chunks = [torch.randn(1, D, Tc, device="cuda") for _ in range(16)]
out = torch.empty((8, 1, Tc), dtype=torch.int16, device="cuda")
host = torch.empty((8, 1, Tc), dtype=torch.int16, pin_memory=True)
for _ in range(20):
    encode_stacks(stacks, chunks[0], torch.int16, out=out)
torch.cuda.synchronize()
dev, wall = [], []
for i in range(1000):
    x = chunks[i % 16]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    encode_stacks(stacks, x, torch.int16, out=out)
    e1.record()
    host.copy_(out, non_blocking=True)
    torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
    dev.append(e0.elapsed_time(e1))
dev, wall = np.array(dev), np.array(wall)
print(f"eager: device ms p50={np.percentile(dev, 50):.4f} p99={np.percentile(dev, 99):.4f} | boundary call + D2H of the 8 streams "
      f"ms p50={np.percentile(wall, 50):.4f} p99={np.percentile(wall, 99):.4f}", flush=True)
x = chunks[0].clone()
hs = [s._pack.get(s._codebooks()) for s in stacks]
harr = (ctypes.c_void_p * 2)(*[h.value for h in hs])
ws = torch.empty(lib.nat_rvq_stacks_workspace_bytes(harr, 2, Tc), dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        encode_stacks(stacks, x, torch.int16, out=out, workspace=ws)
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    encode_stacks(stacks, x, torch.int16, out=out, workspace=ws)
g.replay(); torch.cuda.synchronize()
dev, wall = [], []
for i in range(1000):
    x.copy_(chunks[i % 16])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); g.replay(); e1.record()
    host.copy_(out, non_blocking=True)
    torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
    dev.append(e0.elapsed_time(e1))
ref = encode_stacks(stacks, x, torch.int16)
torch.cuda.synchronize()
dev, wall = np.array(dev), np.array(wall)
print(f"CUDA graph replay: device ms p50={np.percentile(dev, 50):.4f} p99={np.percentile(dev, 99):.4f} | replay + D2H ms "
      f"p50={np.percentile(wall, 50):.4f} p99={np.percentile(wall, 99):.4f} | replay tracks new input: {bool(torch.equal(ref, out))}", flush=True)
