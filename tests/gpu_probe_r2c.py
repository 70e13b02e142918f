"""Ad-hoc probe (not a test): production stack kernel time under env knobs (store mask, group), and the mel kernel with
and without a caller-supplied filterbank."""
import ctypes, os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import _lib, ResidualVectorQuantizer, MelSpectrogram
lib = _lib.load()
D, K, N = 768, 1024, 270000
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(42)
rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
x = torch.randn(1, D, N, device="cuda")
h = rvq._pack.get(rvq._codebooks())
wsb = lib.nat_rvq_workspace_bytes(h, N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
codes = torch.empty((4, N), dtype=torch.int16, device="cuda")
prof = (ctypes.c_float * 8)()
for mask, group in itertools.product(os.environ.get("PROBE_MASKS", "1,0,3").split(","), os.environ.get("PROBE_GROUPS", "3,2").split(",")):
    os.environ["NAT_RVQ_STORE_MASK_SET"] = mask; os.environ["NAT_RVQ_GROUP"] = group
    best = 1e9
    for rep in range(5):
        _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), 0, 1, N, codes.data_ptr(), 2, None, None, 0.25, None, ws.data_ptr(), wsb, 2, st, prof))
        if rep: best = min(best, prof[1])
    print(f"store_mask={mask} group={group} stack_ms={best:.3f} checksum={int(codes.long().sum())}", flush=True)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for sr, hop in ((24000, 320), (22050, 512)):
    S = sr * 3600
    wave = torch.randn(1, S, device="cuda") * 0.1
    mt = MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128).cuda()
    T = 1 + S // hop
    mel = torch.empty((1, 128, T), device="cuda")
    for name, fb in (("builtin fb", None), ("caller fb", mt.fb.contiguous().data_ptr())):
        ms = timeit(lambda: _lib.check(lib.nat_mel_power_f32(wave.data_ptr(), 1, S, sr, 2048, hop, 128, fb, mel.data_ptr(), None, st)))
        print(f"mel sr={sr} hop={hop} {name}: {ms:.3f} ms", flush=True)
    print(f"mel module call: {timeit(lambda: mt(wave)):.3f} ms", flush=True)
