"""Ad-hoc probe (not a test): per-chunk latency (75 frames, 8 layers) of the fused path vs the exact-scan path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
torch.manual_seed(42)
D, K = 768, 1024
for T in (75, 8, 128, 256, 512):
    for exact in (False, True):
        stacks = [ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda() for _ in range(2)]
        for s in stacks:
            s.exact_scan = exact
        x = torch.randn(1, D, T, device="cuda")
        for _ in range(20):
            [s.encode(x) for s in stacks]
        torch.cuda.synchronize()
        dev = []
        for i in range(300):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); [s.encode(x) for s in stacks]; e1.record(); torch.cuda.synchronize()
            dev.append(e0.elapsed_time(e1))
        dev = np.array(dev)
        print(f"T={T} exact_scan={exact} device ms p50={np.percentile(dev, 50):.4f} p99={np.percentile(dev, 99):.4f}", flush=True)
