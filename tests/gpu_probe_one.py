"""Ad-hoc probe (not a test): a few launches of one 4-layer stack encode, the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer
lib = _lib.load()
torch.manual_seed(42)
D = int(os.environ.get("PROBE_D", 768)); K = int(os.environ.get("PROBE_K", 1024)); N = int(os.environ.get("PROBE_N", 270000))
rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
x = torch.randn(1, D, N, device="cuda")
h = rvq._pack.get(rvq._codebooks())
wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
for _ in range(int(os.environ.get("PROBE_REPS", 3))):
    _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 0, st))
torch.cuda.synchronize()
print("ok", int(codes.long().sum()))
