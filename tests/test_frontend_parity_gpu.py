"""Parity of the CUDA front-end (mel power, spectral centroid/bandwidth) against golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mel_oracle

pytestmark = pytest.mark.gpu

# Stated fp32 tolerances (see tests/test_oracle_golden.py for where they come from):
#   mel, our torch-built filterbank (bit-identical to torchaudio's): 2e-5 * max(mel) + 1e-7 per clip
#   mel, library's built-in double-precision filterbank:             1e-4 * max(mel) + 1e-6 per clip
#   centroid / bandwidth: rtol 2e-4, atol 1e-2 Hz
MEL_TOL_FB, MEL_TOL_BUILTIN = 2e-5, 1e-4


@pytest.mark.parametrize("name", ["mel_tone_22050_hop512", "mel_noise_24000_hop320"])
def test_mel_matches_torchaudio_golden(name):
    from neural_audio_tokenizer_b200 import MelSpectrogram, _lib
    g = load_golden(name)
    sr, hop = int(g["sr"]), int(g["hop"])
    wave = torch.from_numpy(g["wave"]).cuda()
    ref = g["mel"]                                                    # [1, 128, T]
    mt = MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128, normalized=True, log_mel=True).cuda()
    assert mt.sample_rate == sr
    np.testing.assert_array_equal(mt.fb.cpu().numpy(), g["fb"])       # same filterbank as the reference multiplies by
    mel = mt(wave[None])
    assert mel.shape == ref.shape
    err = np.abs(mel.cpu().numpy() - ref).max()
    assert err <= MEL_TOL_FB * ref.max() + 1e-7, err
    assert mt(wave).shape == ref.shape[1:]                            # 1-D input keeps torchaudio's shape rule
    np.testing.assert_allclose(mt.last_log_mel.cpu().numpy(), mel_oracle.log_mel_db(mel.cpu().numpy())[0],
                               rtol=0, atol=2e-4)
    # built-in filterbank of the C ABI (fb_dev = NULL)
    lib = _lib.load()
    out = torch.empty_like(mel)
    _lib.check(lib.nat_mel_power_f32(wave.data_ptr(), 1, wave.numel(), sr, 2048, hop, 128, None, out.data_ptr(), None,
                                     None))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - ref).max() <= MEL_TOL_BUILTIN * ref.max() + 1e-6
    # the same entry point with the caller's dense filterbank (prepared inside the call)
    out2 = torch.empty_like(mel)
    fbc = mt.fb.contiguous()
    _lib.check(lib.nat_mel_power_f32(wave.data_ptr(), 1, wave.numel(), sr, 2048, hop, 128, fbc.data_ptr(), out2.data_ptr(),
                                     None, None))
    torch.cuda.synchronize()
    assert torch.equal(out2, mel)                                     # same kernel, same layout as the module's call


def test_mel_batch_and_long_clip_vs_oracle():
    from neural_audio_tokenizer_b200 import MelSpectrogram
    rng = np.random.default_rng(9)
    wave = (rng.standard_normal((3, 24000 * 4 + 123)) * 0.1).astype(np.float32)
    mt = MelSpectrogram(sample_rate=24000, n_fft=2048, hop_length=320, n_mels=128).cuda()
    mel = mt(torch.from_numpy(wave).cuda()).cpu().numpy()
    fb = mt.fb.cpu().numpy().astype(np.float64)
    for b in range(3):
        ref = (mel_oracle.stft_power(wave[b], 2048, 320).T @ fb).T
        assert mel[b].shape == ref.shape == (128, 1 + wave.shape[1] // 320)
        assert np.abs(mel[b] - ref).max() <= MEL_TOL_FB * ref.max() + 1e-7


def _caller_filterbank(kind, n_mels, rng):
    """Filterbanks that exercise every projection layout of csrc/mel_fft.cuh (dual / single, staged / through L1)."""
    nb = 1025
    fb = np.zeros((nb, n_mels), dtype=np.float32)
    if kind == "wide_triangles":            # three bands overlap everywhere: not a two-band bank -> single layout
        centres = np.linspace(20, 1000, n_mels)
        half = 2.2 * (centres[1] - centres[0])
        k = np.arange(nb)[:, None]
        fb = np.maximum(0.0, 1.0 - np.abs(k - centres[None, :]) / half).astype(np.float32)
    elif kind == "dense":                   # every bin in every band: weights stay in global memory
        fb = rng.random((nb, n_mels), dtype=np.float32)
    elif kind == "empty_band":              # a band of zeros in the middle, and one single-bin band at Nyquist
        edges = np.linspace(0, 1024, n_mels + 1).astype(int)
        for m in range(n_mels):
            fb[edges[m]:edges[m + 1] + 1, m] = rng.random(edges[m + 1] + 1 - edges[m], dtype=np.float32)
        fb[:, n_mels // 2] = 0.0
        fb[:, n_mels - 1] = 0.0
        fb[1024, n_mels - 1] = 1.0
    return fb


@pytest.mark.parametrize("kind,n_mels", [("htk", 80), ("htk", 40), ("htk", 200), ("wide_triangles", 128),
                                         ("wide_triangles", 33), ("dense", 128), ("empty_band", 64), ("htk", 2)])
def test_mel_projection_layouts_vs_oracle(kind, n_mels):
    """The filterbank is a parameter of the transform (--n_mels, nat.py:5400; `fb` is a writable buffer): whatever its
    shape, the projection must equal power @ fb."""
    from neural_audio_tokenizer_b200 import MelSpectrogram
    rng = np.random.default_rng(n_mels)
    wave = (rng.standard_normal((2, 24000 + 77)) * 0.1).astype(np.float32)
    mt = MelSpectrogram(sample_rate=24000, n_fft=2048, hop_length=320, n_mels=n_mels).cuda()
    if kind != "htk":
        mt.fb.copy_(torch.from_numpy(_caller_filterbank(kind, n_mels, rng)))
    mel = mt(torch.from_numpy(wave).cuda()).cpu().numpy()
    fb = mt.fb.cpu().numpy().astype(np.float64)
    for b in range(2):
        ref = (mel_oracle.stft_power(wave[b], 2048, 320).T @ fb).T
        assert mel[b].shape == ref.shape
        assert np.abs(mel[b] - ref).max() <= MEL_TOL_FB * ref.max() + 1e-7, (kind, n_mels)


@pytest.mark.parametrize("name", ["spectral_tone_22050", "spectral_noise_24000", "spectral_short"])
def test_spectral_stats_match_reference_golden(name):
    from neural_audio_tokenizer_b200 import spectral_stats
    g = load_golden(name)
    st = spectral_stats(torch.from_numpy(g["wave"]).cuda()[None], int(g["sr"]))
    assert st.shape == g["stats"].shape
    # The stated fp32 tolerance (rtol 2e-4, atol 1e-2 Hz) is each fp32 implementation's distance from the exact
    # (fp64 oracle) value: the reference's own golden output sits up to 1.6e-4 from it on the pure tone, whose
    # bandwidth is a sum over leakage tails and amplifies FFT rounding noise. So: within the tolerance of the
    # oracle, and within twice the tolerance of the reference's fp32 output.
    exact = mel_oracle.spectral_stats(g["wave"], int(g["sr"]))
    np.testing.assert_allclose(st.cpu().numpy(), exact, rtol=2e-4, atol=1e-2)
    np.testing.assert_allclose(st.cpu().numpy(), g["stats"], rtol=4e-4, atol=2e-2)


def test_pipeline_fixture_mel():
    from neural_audio_tokenizer_b200 import MelSpectrogram
    g = load_golden("pipeline_tone_argmin")
    mt = MelSpectrogram(sample_rate=int(g["sr"]), n_fft=2048, hop_length=512, n_mels=128).cuda()
    mel = mt(torch.from_numpy(g["audio"]).cuda()[None]).cpu().numpy()
    assert np.abs(mel - g["mel"]).max() <= MEL_TOL_FB * g["mel"].max() + 1e-7


def test_frontend_errors():
    from neural_audio_tokenizer_b200 import MelSpectrogram
    with pytest.raises(ValueError):
        MelSpectrogram(sample_rate=22050, n_fft=1024)
    mt = MelSpectrogram(sample_rate=22050, n_fft=2048, hop_length=512).cuda()
    with pytest.raises(RuntimeError):
        mt(torch.zeros(1, 1000, device="cuda"))                       # shorter than the reflect pad
    with pytest.raises(RuntimeError):
        mt(torch.zeros(1, 4096))                                      # CPU tensor: no fallback
