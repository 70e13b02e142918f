"""Ad-hoc probe (not a test): per-class kernel time of one 4-layer stack with and without the two-lane overlap."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer
lib = _lib.load()
torch.manual_seed(42)
rvq = ResidualVectorQuantizer(768, 1024, 4, use_stochastic=False).eval().cuda()
N = 270000
x = torch.randn(1, 768, N, device="cuda")
h = rvq._pack.get(rvq._codebooks())
wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for flags in (0, 2):
    for rep in range(3):
        prof = (ctypes.c_float * 8)()
        _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, flags, st, prof))
    print("flags", flags, {n: round(prof[i], 3) for i, n in enumerate(_lib.PROF_NAMES)})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, flags, st))
    e1.record(); torch.cuda.synchronize()
    print("flags", flags, "ms per stack", e0.elapsed_time(e1) / 10)
