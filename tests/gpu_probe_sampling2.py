"""Ad-hoc probe (not a test): host time of the sampling and argmin calls against N."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer, _lib
lib = _lib.load()
torch.manual_seed(42)
D, K = 768, 1024
rvq = ResidualVectorQuantizer(D, K, 4).eval().cuda()
h = rvq._pack.get(rvq._codebooks())
temps = (ctypes.c_float * 4)(0.5, 0.5, 0.5, 0.5)
st = torch.cuda.current_stream().cuda_stream
for N in (270000, 65536, 16384):
    x = torch.randn(1, D, N, device="cuda")
    wsb = lib.nat_rvq_workspace_bytes(h, N)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
    def timed(fn):
        out = []
        for _ in range(6):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            out.append((round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t0), 2), round(e0.elapsed_time(e1), 2)))
        return out[2:]
    print(N, "workspace MB", wsb >> 20)
    print("  sample:", timed(lambda: _lib.check(lib.nat_rvq_sample_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, temps, None, 1, 0, ws.data_ptr(), wsb, 0, st))), flush=True)
    print("  argmin:", timed(lambda: _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 0, st))), flush=True)
    os.environ["NAT_RVQ_FUSED"] = "0"
    print("  argmin, per-layer kernels:", timed(lambda: _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 0, st))), flush=True)
    del os.environ["NAT_RVQ_FUSED"]
