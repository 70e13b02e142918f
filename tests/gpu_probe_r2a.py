"""Ad-hoc probe (not a test): where every role of the fused stack kernel waits, for several codebook sizes and with
parts of the work switched off (NAT_RVQ_DBG_MODE), plus the production kernel's time beside it."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer
lib = _lib.load()
names = ["tma_wait_ready", "tma_wait_empty", "mma_wait_tempty", "mma_wait_full", "epi_wait_tfull", "epi_wait_cempty",
         "epi_total", "upd_wait_cfull", "upd_total", "kernel_total", "epi_wait_ld", "epi_events", "upd_decide", "upd_resid", "upd_fence"]
D, N = int(os.environ.get("PROBE_D", 768)), int(os.environ.get("PROBE_N", 270000))
st = torch.cuda.current_stream().cuda_stream
for K in [int(k) for k in os.environ.get("PROBE_KS", "1024,4096").split(",")]:
    torch.manual_seed(42)
    rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
    x = torch.randn(1, D, N, device="cuda")
    h = rvq._pack.get(rvq._codebooks())
    wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
    for group, mask in [(g, m) for g in os.environ.get("PROBE_GROUPS", "3").split(",") for m in os.environ.get("PROBE_MASKS", "").split(",")]:
        os.environ["NAT_RVQ_GROUP"] = group
        if mask:
            os.environ["NAT_RVQ_STORE_MASK_SET"] = mask
            group = f"{group} store_mask={mask}"
        os.environ["NAT_RVQ_DBG_MODE"] = "0"
        prof = (ctypes.c_float * 8)()
        best = 1e9
        for rep in range(5):
            _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
            if rep: best = min(best, prof[1])
        print(f"K={K} group={group} PRODUCTION stack_ms={best:.3f} prep_ms={prof[0]:.3f} checksum={int(codes.long().sum())}", flush=True)
        for mode in os.environ.get("PROBE_MODES", "0,3,6,7,16,19").split(","):
            os.environ["NAT_RVQ_DBG_MODE"] = mode
            nc, ns = ctypes.c_int(), ctypes.c_int()
            for rep in range(2):
                _lib.check(lib.nat_debug_stack_counters(h, 1, None, 0, ctypes.byref(nc), ctypes.byref(ns)))
                _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
                buf = np.zeros((nc.value, ns.value), dtype=np.uint64)
                _lib.check(lib.nat_debug_stack_counters(h, 0, buf.ctypes.data_as(ctypes.c_void_p), nc.value, None, None))
            avg = buf.astype(np.float64).mean(axis=0)
            print(f"K={K} group={group} mode={mode} stack_ms={prof[1]:.3f} kcycles/CTA:",
                  {n: int(avg[i] / 1e3) for i, n in enumerate(names)}, flush=True)
        os.environ["NAT_RVQ_DBG_MODE"] = "0"
