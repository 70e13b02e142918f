"""Ad-hoc probe (not a test): a few launches of the mel kernel on the config-4 geometry, the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_audio_tokenizer_b200 import MelSpectrogram
sr, hop = 24000, 320
wave = torch.randn(1, sr * int(os.environ.get("PROBE_SECS", 3600)), device="cuda") * 0.1
mt = MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128).cuda()
for _ in range(3):
    mel = mt(wave)
torch.cuda.synchronize()
print("ok", float(mel.sum()))
