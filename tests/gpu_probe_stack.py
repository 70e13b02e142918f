"""Ad-hoc probe (not a test): ms per 4-layer stack (prep and fused kernel separately) for tile groupings, A/B-able
across library builds with NAT_B200_LIB; checks all variants emit the same indices."""
import ctypes, os, sys
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
base = None
for group in os.environ.get("PROBE_GROUPS", "2,3").split(","):
    os.environ["NAT_RVQ_GROUP"] = group
    codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
    prof = (ctypes.c_float * 8)()
    best = 1e9
    for rep in range(int(os.environ.get("PROBE_REPS", 6))):
        _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
        if rep: best = min(best, prof[1])
    if base is None:
        base = codes.clone()
    print(f"D={D} K={K} group={group} stack_ms={best:.3f} prep_ms={prof[0]:.3f} same_codes={bool(torch.equal(codes, base))} checksum={int(codes.long().sum())}", flush=True)
