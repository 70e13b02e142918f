"""Ad-hoc probe (not a test): one pass over every device path on small inputs, the command compute-sanitizer wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import MelSpectrogram, ResidualVectorQuantizer, align, spectral_stats, token_stats
torch.manual_seed(1)
D, K = 768, 1024
rvq = ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().cuda()
for N in (int(os.environ.get("PROBE_N", 19100)), 300, 1):
    x = torch.randn(1, D, N, device="cuda")
    q, codes, losses = rvq(x)                      # forward: quantised sum + losses
    enc = rvq.encode(x)                            # codes only (hot update form / small-input path)
    assert all(torch.equal(a, b) for a, b in zip(codes, enc))
    dec = rvq.decode(enc)
rvq2 = ResidualVectorQuantizer(96, 300, 3, use_stochastic=False).eval().cuda()      # ragged D / K, general update form
rvq2.encode(torch.randn(2, 96, 700, device="cuda"))
samp = ResidualVectorQuantizer(64, 128, 2).eval().cuda()
samp(torch.randn(1, 64, 50, device="cuda"))       # host-noise sampling
samp.sampling_mode = "philox"
samp.encode(torch.randn(1, 64, 500, device="cuda"))
mt = MelSpectrogram(sample_rate=24000, n_fft=2048, hop_length=320, n_mels=128).cuda()
mt(torch.randn(1, 24000, device="cuda"))
spectral_stats(torch.randn(24000, device="cuda"), 24000, 2048, 320)
align.interpolate_linear(torch.randn(1, 4, 131, device="cuda"), 128)
token_stats.pooled_counts([c for c in enc], K)
token_stats.mutual_information(enc[0], enc[1], K)
torch.cuda.synchronize()
print("all paths ran")
