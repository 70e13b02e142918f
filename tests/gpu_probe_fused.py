"""Ad-hoc probe (not a test): ms per 4-layer stack (270k frames x 768, K=1024) for the fused kernel at several tile
groupings and for the per-layer kernels; checks that all of them emit identical indices."""
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
variants = [("0", "2", "0")] + [("1", g, hh) for g in os.environ.get("PROBE_GROUPS", "1,2,3,4,1000000").split(",") for hh in os.environ.get("PROBE_HINTS", "0").split(",")]
for fused, group, hints in variants:
    os.environ["NAT_RVQ_FUSED"] = fused; os.environ["NAT_RVQ_GROUP"] = group; os.environ["NAT_RVQ_L2_HINTS"] = hints
    codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
    prof = (ctypes.c_float * 8)()
    for rep in range(2):
        _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 0, st))
    e1.record(); torch.cuda.synchronize()
    stats = torch.zeros((4, 4), dtype=torch.int64, device="cuda")
    _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, stats.data_ptr(), ws.data_ptr(), wsb, 0, st))
    torch.cuda.synchronize()
    print("   stats", stats[:, :3].cpu().tolist())
    if base is None:
        base = codes.clone()
    same = bool(torch.equal(codes, base))
    print(f"fused={fused} group={group} hints={hints} ms/stack={e0.elapsed_time(e1) / 10:.3f} same_codes={same}",
          {n: round(prof[i], 3) for i, n in enumerate(_lib.PROF_NAMES)}, flush=True)

# where the roles of the fused kernel wait (cycles, averaged over CTAs)
import numpy as np
names = ["tma_wait_ready", "tma_wait_empty", "mma_wait_tempty", "mma_wait_full", "epi_wait_tfull", "epi_wait_cempty",
         "epi_total", "upd_wait_cfull", "upd_total", "kernel_total"]
for group in os.environ.get("PROBE_GROUPS", "2").split(","):
    os.environ["NAT_RVQ_FUSED"] = "1"; os.environ["NAT_RVQ_GROUP"] = group
    nc, ns = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.nat_debug_stack_counters(h, 1, None, 0, ctypes.byref(nc), ctypes.byref(ns)))
    _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 0, st))
    buf = np.zeros((nc.value, ns.value), dtype=np.uint64)
    _lib.check(lib.nat_debug_stack_counters(h, 0, buf.ctypes.data_as(ctypes.c_void_p), nc.value, None, None))
    avg = buf.astype(np.float64).mean(axis=0)
    print(f"group={group} kcycles/CTA:", {n: round(avg[i] / 1e3, 1) for i, n in enumerate(names)}, flush=True)
