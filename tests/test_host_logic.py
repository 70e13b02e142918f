"""Host-side mirror of the reference surface: constructor parity, error conventions, no-CPU-fallback. CPU only."""
import numpy as np
import pytest
import torch

from neural_audio_tokenizer_b200 import MelSpectrogram, ResidualVectorQuantizer, VectorQuantizer, spectral_stats
from neural_audio_tokenizer_b200.frontend import htk_filterbank
from neural_audio_tokenizer_b200.sharding import shard_clips, shard_range


def test_constructor_surface_matches_reference_contract():
    rvq = ResidualVectorQuantizer(64, 128, 4)
    assert (rvq.input_dim, rvq.codebook_size, rvq.num_quantizers, rvq.commitment_weight) == (64, 128, 4, 0.25)
    assert len(rvq.quantizers) == 4
    q = rvq.quantizers[0]
    assert isinstance(q, VectorQuantizer)
    names = dict(q.named_buffers())
    assert set(names) == {"codebook", "ema_count", "ema_weight"}          # the cache reads/writes these (nat.py:501-504)
    assert names["codebook"].shape == (128, 64) and names["ema_count"].shape == (128,)
    assert torch.equal(names["ema_weight"], names["codebook"])
    assert q.use_stochastic is True and q.temperature == 0.5              # the reference's defaults
    assert not list(rvq.parameters())                                     # buffers, not parameters
    sd = rvq.state_dict()
    assert "quantizers.3.codebook" in sd


def test_seeded_construction_draws_reference_codebooks():
    """Same global-RNG consumption as nat.py:2115: randn(K, D) per layer, in layer order."""
    torch.manual_seed(42)
    rvq = ResidualVectorQuantizer(16, 32, 3)
    torch.manual_seed(42)
    expect = [torch.randn(32, 16) for _ in range(3)]
    for q, e in zip(rvq.quantizers, expect):
        assert torch.equal(q.codebook, e)


def test_shape_errors_are_value_errors_before_any_device_work():
    rvq = ResidualVectorQuantizer(8, 16, 2, use_stochastic=False).eval()
    with pytest.raises(ValueError, match="2D or 3D"):
        rvq(torch.randn(8))
    with pytest.raises(ValueError, match="2D or 3D"):
        rvq(torch.randn(1, 1, 8, 4))
    with pytest.raises(ValueError, match="Expected 8 feature dimensions"):
        rvq(torch.randn(1, 9, 4))
    vq = rvq.quantizers[0]
    with pytest.raises(ValueError, match="expects 2D or 3D"):
        vq(torch.randn(8))
    with pytest.raises(ValueError, match="Expected 8 feature dimensions"):
        vq(torch.randn(7, 4))


def test_sampling_modes_never_fall_through_to_argmin():
    rvq = ResidualVectorQuantizer(8, 16, 2).eval()                        # default: use_stochastic=True
    with pytest.raises(RuntimeError, match="no CPU fallback"):            # native sampling needs the device too
        rvq(torch.randn(1, 8, 4))
    rvq.sampling_mode = "delegate"
    with pytest.raises(NotImplementedError, match="delegate"):
        rvq(torch.randn(1, 8, 4))
    with pytest.raises(NotImplementedError, match="delegate"):
        rvq.encode(torch.randn(1, 8, 4))
    rvq.sampling_mode = "host_noise"
    for q in rvq.quantizers:
        q.use_stochastic = False
    rvq.train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):            # training samples + EMA natively: device only
        rvq(torch.randn(1, 8, 4))
    calls = []

    class Delegate:
        def __call__(self, x):
            calls.append(tuple(x.shape))
            return "delegated"

    rvq.stochastic_delegate = Delegate()
    assert rvq(torch.randn(1, 8, 4)) == "delegated" and calls == [(1, 8, 4)]
    rvq.eval()
    rvq.stochastic_delegate = None
    assert rvq.training is False
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rvq(torch.randn(1, 8, 4), training_mode=True)
    assert rvq.training is False                                          # restored in finally, nat.py:1417-1420


def test_codebook_initializers_run_the_reference_code_and_invalidate():
    """The tokenizer calls `initialize_from_mert_model` & co. on its quantizers (nat.py:3075-3176); on the drop-in they
    resolve to the reference class's own functions, and their `.data` writes (which `_version` does not see)
    invalidate the uploaded codebook state."""
    class FakeReferenceRVQ:
        def initialize_from_mert_model(self, model_name="m", **kw):
            for i, q in enumerate(self.quantizers):
                q.codebook.data.copy_(torch.full_like(q.codebook, float(i + 1)))      # nat.py:1926 writes through .data
            return self._helper(model_name)

        def _helper(self, name):
            return f"{name}:{self.codebook_size}x{self.input_dim}"

        @staticmethod
        def _static(a):
            return a + 1

    rvq = ResidualVectorQuantizer(8, 16, 2, use_stochastic=False).eval()
    with pytest.raises(RuntimeError, match="codebook sourcing stays with the reference"):
        rvq.initialize_from_mert_model(model_name="x")
    with pytest.raises(AttributeError):
        rvq.no_such_thing
    rvq._reference_class = FakeReferenceRVQ
    rvq._pack.signature = ("uploaded",)
    for q in rvq.quantizers:
        q._pack.signature = ("uploaded",)
    v0 = rvq.quantizers[0].codebook._version
    assert rvq.initialize_from_mert_model(model_name="mert") == "mert:16x8"
    assert rvq.quantizers[0].codebook._version == v0                      # .data.copy_ leaves the version alone ...
    assert rvq._pack.signature is None and all(q._pack.signature is None for q in rvq.quantizers)   # ... hence this
    assert float(rvq.quantizers[1].codebook[0, 0]) == 2.0
    assert rvq._static(1) == 2
    rvq._pack.signature = ("uploaded",)
    rvq.invalidate_codebooks()
    assert rvq._pack.signature is None


def test_no_cpu_fallback():
    rvq = ResidualVectorQuantizer(8, 16, 2, use_stochastic=False).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rvq(torch.randn(1, 8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rvq.encode(torch.randn(1, 8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rvq.decode([torch.zeros(1, 4, dtype=torch.long)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MelSpectrogram(sample_rate=22050, n_fft=2048, hop_length=512)(torch.zeros(1, 4096))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        spectral_stats(torch.zeros(4096), 22050)
    assert rvq.decode([]).shape == (1, 8, 1)                              # nat.py:1430-1431


def test_mel_object_contract_and_filterbank():
    mt = MelSpectrogram(sample_rate=22050, n_fft=2048, hop_length=512, n_mels=128, normalized=True)
    assert (mt.sample_rate, mt.n_fft, mt.hop_length, mt.n_mels) == (22050, 2048, 512, 128)
    assert mt.to("cpu") is mt                                             # .to(device) works, nat.py:2287
    fb = htk_filterbank(22050, 2048, 128)
    assert fb.shape == (1025, 128) and (fb >= 0).all() and (fb.max(dim=0).values > 0).all()
    dens = (fb != 0).float().mean().item()
    assert 0.01 < dens < 0.02                                             # 1.5 % dense
    try:
        import torchaudio
        ref = torchaudio.functional.melscale_fbanks(1025, 0.0, 11025.0, 128, 22050)
        assert torch.equal(fb, ref)
    except ImportError:
        pass
    with pytest.raises(ValueError):
        MelSpectrogram(sample_rate=22050, n_fft=512)
    with pytest.raises(TypeError):
        MelSpectrogram(sample_rate=22050, n_fft=2048, power=1.0)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 270000, 270001):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) == -(-n // world)
    assert list(shard_clips(10, 4, 1)) == [1, 5, 9]


def test_peer_gatherer_needs_cuda_devices():
    """The copy-engine exchange moves device memory over NVLink: on a host device it refuses loudly (CodeGatherer is the
    form that runs on any torch.distributed backend), it does not fall back."""
    from neural_audio_tokenizer_b200.sharding import PeerCodeGatherer
    with pytest.raises(RuntimeError, match="CUDA"):
        PeerCodeGatherer(8, 100, 1, "cpu")
