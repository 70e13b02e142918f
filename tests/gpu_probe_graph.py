"""Ad-hoc probe (not a test): the chunk path (75 frames, 4 + 4 layers) captured in a CUDA graph and replayed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
torch.manual_seed(42)
D, K, T = 768, 1024, int(os.environ.get("PROBE_T", 75))
stacks = [ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda() for _ in range(2)]
x = torch.randn(1, D, T, device="cuda")
eager = [s.encode(x) for s in stacks]                       # builds the codebook handles outside the capture
torch.cuda.synchronize()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        [s.encode(x) for s in stacks]
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = [s.encode(x) for s in stacks]
g.replay(); torch.cuda.synchronize()
same = all(torch.equal(a, b) for sa, sb in zip(eager, out) for a, b in zip(sa, sb))
dev = []
for i in range(1000):
    x.copy_(torch.randn(1, D, T, device="cuda"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    dev.append(e0.elapsed_time(e1))
dev = np.array(dev)
ref = [s.encode(x) for s in stacks]; torch.cuda.synchronize()
same2 = all(torch.equal(a, b) for sa, sb in zip(ref, out) for a, b in zip(sa, sb))
print(f"graph replay T={T}: device ms p50={np.percentile(dev, 50):.4f} p99={np.percentile(dev, 99):.4f} same_as_eager={same} replay_tracks_new_input={same2}")
