"""Ad-hoc probe (not a test): timing experiments on the fused stack kernel with parts of the work switched off."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer
lib = _lib.load()
torch.manual_seed(42)
D, K, N = 768, 1024, 270000
rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
x = torch.randn(1, D, N, device="cuda")
h = rvq._pack.get(rvq._codebooks())
wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
names = ["tma_wait_ready", "tma_wait_empty", "mma_wait_tempty", "mma_wait_full", "epi_wait_tfull", "epi_wait_cempty",
         "epi_total", "upd_wait_cfull", "upd_total", "kernel_total", "epi_wait_ld", "epi_events_lane0_x1000", "upd_decide", "upd_resid", "upd_fence"]
os.environ["NAT_RVQ_FUSED"] = "1"
for group in os.environ.get("PROBE_GROUPS", "2").split(","):
    for mode in os.environ.get("PROBE_MODES", "0,1,2,3").split(","):
        os.environ["NAT_RVQ_GROUP"] = group; os.environ["NAT_RVQ_DBG_MODE"] = mode
        nc, ns = ctypes.c_int(), ctypes.c_int()
        for rep in range(2):
            _lib.check(lib.nat_debug_stack_counters(h, 1, None, 0, ctypes.byref(nc), ctypes.byref(ns)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 0, st))
            e1.record(); torch.cuda.synchronize()
            buf = np.zeros((nc.value, ns.value), dtype=np.uint64)
            _lib.check(lib.nat_debug_stack_counters(h, 0, buf.ctypes.data_as(ctypes.c_void_p), nc.value, None, None))
        avg = buf.astype(np.float64).mean(axis=0)
        print(f"group={group} mode={mode} ms(prep+stack)={e0.elapsed_time(e1):.3f} kcycles/CTA:",
              {n: int(avg[i] / 1e3) for i, n in enumerate(names)}, "events_per_frame_layer(thread 0)", round(avg[11] / (4 * ((N + 127) // 128 + 147) // 148), 2), flush=True)
