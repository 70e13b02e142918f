"""Ad-hoc probe (not a test): device time of the mel / spectral front-end kernels on BASELINE config 4 geometry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import MelSpectrogram, spectral_stats

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for sr, hop, secs in ((24000, 320, 3600), (22050, 512, 3600)):
    S = sr * secs
    wave = torch.randn(1, S, device="cuda") * 0.1
    mt = MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128).cuda()
    ms = timeit(lambda: mt(wave))
    T = 1 + S // hop
    alg = T * (hop * 4 + 128 * 4)
    flop = T * 0.5 * 5 * 2048 * 11          # two real frames per complex 2048-point FFT
    print(f"mel sr={sr} hop={hop} frames={T} ms={ms:.3f} frames/s={T / ms * 1e3:.3e} alg_GB/s={alg / ms / 1e6:.1f} fft_TFLOP/s={flop / ms / 1e9:.2f}", flush=True)
    ms2 = timeit(lambda: spectral_stats(wave[0], sr, 2048, hop))
    Ts = 1 + (S - 2048) // hop
    print(f"spectral sr={sr} hop={hop} frames={Ts} ms={ms2:.3f} frames/s={Ts / ms2 * 1e3:.3e} alg_GB/s={Ts * (hop * 4 + 8) / ms2 / 1e6:.1f}", flush=True)
