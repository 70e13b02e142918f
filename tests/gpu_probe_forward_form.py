"""Ad-hoc probe (not a test): the forward form rvq(x) -> (quantised sum, codes, loss) of one 4-layer stack over
270 000 frames, the form install() grafts into the reference (nat.py:3239-3240); the command ncu wraps for its
launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
torch.manual_seed(42)
D, K, N = 768, 1024, int(os.environ.get("PROBE_N", 270000))
rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
x = torch.randn(1, D, N, device="cuda")
for _ in range(2):
    out = rvq(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = rvq(x)
e1.record(); torch.cuda.synchronize()
print(f"forward form, one stack: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
codes = out[1]
for _ in range(2):
    dec = rvq.decode(codes)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    dec = rvq.decode(codes)
e1.record(); torch.cuda.synchronize()
print(f"decode, one stack ([1, {D}, {N}] out): {e0.elapsed_time(e1) / 5:.3f} ms  ({N * D * 4 / (e0.elapsed_time(e1) / 5) / 1e6:.0f} GB/s of output)", flush=True)
