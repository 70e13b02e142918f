"""Ad-hoc probe (not a test): both stacks of a tokenizer in one launch (nat_rvq_encode_stacks_f32) against two
single-stack calls: kernel times per class, same index streams."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer
lib = _lib.load()
D, N = int(os.environ.get("PROBE_D", 768)), int(os.environ.get("PROBE_N", 270000))
st = torch.cuda.current_stream().cuda_stream
for K in [int(k) for k in os.environ.get("PROBE_KS", "1024").split(",")]:
    torch.manual_seed(42)
    stacks = [ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda() for _ in range(2)]
    x = torch.randn(1, D, N, device="cuda")
    hs = [s._pack.get(s._codebooks()) for s in stacks]
    harr = (ctypes.c_void_p * 2)(*[h.value for h in hs])
    xarr = (ctypes.c_void_p * 2)(x.data_ptr(), x.data_ptr())
    wsb = lib.nat_rvq_stacks_workspace_bytes(harr, 2, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    single = torch.empty((8, N), dtype=torch.int16, device="cuda")
    prof = (ctypes.c_float * 8)()
    t_single = [1e9, 1e9]
    for rep in range(4):
        for i, h in enumerate(hs):
            _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, single[4 * i:].data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
            if rep: t_single[i] = min(t_single[i], prof[1])
    both = torch.empty((8, N), dtype=torch.int16, device="cuda")
    best, prep = 1e9, 0
    for rep in range(5):
        _lib.check(lib.nat_rvq_encode_stacks_profile_f32(harr, 2, xarr, 0, 1, N, both.data_ptr(), 2, ws.data_ptr(), wsb, 0, st, prof))
        if rep and prof[1] < best: best, prep, launches = prof[1], prof[0], prof[6]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(lib.nat_rvq_encode_stacks_f32(harr, 2, xarr, 0, 1, N, both.data_ptr(), 2, ws.data_ptr(), wsb, 0, st))
    e1.record(); torch.cuda.synchronize()
    tf = 2.0 * K * D * N * 8 / (best * 1e-3) / 1e12
    print(f"K={K} single stack_ms={t_single[0]:.3f}+{t_single[1]:.3f}  two-stack launch: stack_ms={best:.3f} ({launches:.0f} launch) prep_ms={prep:.3f} "
          f"step_ms={e0.elapsed_time(e1) / 5:.3f} TFLOP/s={tf:.1f} same_codes={bool(torch.equal(both, single))} checksum={int(both.long().sum())}", flush=True)
