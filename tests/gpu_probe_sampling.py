"""Ad-hoc probe (not a test): throughput of the sampling mode (nat_rvq_sample_f32) on the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
torch.manual_seed(42)
D, K = 768, 1024
rvq = ResidualVectorQuantizer(D, K, 4).eval().cuda()                  # reference default: use_stochastic=True, T=0.5
for mode, N in (("philox", 270000), ("host_noise", 16384)):
    rvq.sampling_mode = mode
    x = torch.randn(1, D, N, device="cuda")
    rvq.encode(x); torch.cuda.synchronize()
    t0 = time.perf_counter()
    codes = rvq.encode(x); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"sampling_mode={mode} frames={N} 4 layers: {dt * 1e3:.1f} ms -> {N / dt:.3e} frames/s (distinct codes layer 0: {codes[0].unique().numel()})", flush=True)
# per-kernel split of the Philox path (CUDA events around every launch)
import ctypes
from neural_audio_tokenizer_b200 import _lib
lib = _lib.load()
N = 270000
x = torch.randn(1, D, N, device="cuda")
h = rvq._pack.get(rvq._codebooks())
wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
temps = (ctypes.c_float * 4)(0.5, 0.5, 0.5, 0.5)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.nat_rvq_sample_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, temps, None, 1, 0, ws.data_ptr(), wsb, 0, st))
    e1.record(); torch.cuda.synchronize()
    print(f"nat_rvq_sample_f32 philox, 270000 frames, 4 layers: {e0.elapsed_time(e1):.2f} ms", flush=True)
