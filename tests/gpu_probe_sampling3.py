"""Ad-hoc probe (not a test): three sampling calls (philox) on the bench workload, the command ncu wraps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer, _lib
lib = _lib.load()
torch.manual_seed(42)
D, K, N = 768, 1024, 270000
rvq = ResidualVectorQuantizer(D, K, 4).eval().cuda()
x = torch.randn(1, D, N, device="cuda")
h = rvq._pack.get(rvq._codebooks())
wsb = lib.nat_rvq_workspace_bytes(h, N)
temps = (ctypes.c_float * 4)(0.5, 0.5, 0.5, 0.5)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
for d in range(3):
    _lib.check(lib.nat_rvq_sample_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, temps, None, 1, d, ws.data_ptr(), wsb, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", int(codes.long().sum()))
