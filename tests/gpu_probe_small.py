"""Ad-hoc probe (not a test): per-kernel-class device time of one 4-layer stack on a small chunk (CUDA events around
every launch), fused and per-layer paths."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer
lib = _lib.load()
torch.manual_seed(42)
D, K = 768, 1024
rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
h = rvq._pack.get(rvq._codebooks())
st = torch.cuda.current_stream().cuda_stream
for N in (75, 1000, 4000, 4736):
    x = torch.randn(1, D, N, device="cuda")
    wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
    for fused in ("1", "0"):
        os.environ["NAT_RVQ_FUSED"] = fused
        prof = (ctypes.c_float * 8)()
        for rep in range(5):
            _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
        print(f"N={N} fused={fused}", {n: round(prof[i] * 1e3, 1) for i, n in enumerate(_lib.PROF_NAMES)}, "(microseconds; launches count raw)", flush=True)
