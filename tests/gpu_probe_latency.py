"""Ad-hoc probe (not a test): BASELINE config 5, per-chunk tokenize latency for 1 s chunks (75 frames x 768-d, 4+4 layers,
K = 1024, batch 1): p50 / p99 over 1000 chunks of (i) device time (CUDA events) and (ii) the boundary call including
the bulk D2H of the 8 index streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
torch.manual_seed(42)
D, K, T = 768, 1024, int(os.environ.get("PROBE_T", 75))
stacks = [ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda() for _ in range(2)]
for s in stacks:
    s.codes_on_cpu = False
chunks = [torch.randn(1, D, T, device="cuda") for _ in range(16)]
def run(x):
    return [s.encode(x) for s in stacks]
for _ in range(20):
    run(chunks[0])
torch.cuda.synchronize()
dev, wall = [], []
for i in range(1000):
    x = chunks[i % 16]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    codes = run(x)
    e1.record()
    host = torch.stack([c for cs in codes for c in cs]).to(torch.int16).cpu()      # one bulk D2H of [8, 1, T]
    wall.append((time.perf_counter() - t0) * 1e3)
    dev.append(e0.elapsed_time(e1))
dev, wall = np.array(dev), np.array(wall)
print(f"T={T} device ms p50={np.percentile(dev, 50):.4f} p99={np.percentile(dev, 99):.4f} | boundary call + D2H ms "
      f"p50={np.percentile(wall, 50):.4f} p99={np.percentile(wall, 99):.4f}")
