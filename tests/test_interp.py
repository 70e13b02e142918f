"""Time-base alignment: F.interpolate(mode='linear', align_corners=False) of nat.py:3225-3236.
Oracle vs the golden outputs of the reference's call (CPU); device kernel vs the same goldens (GPU). Bar: bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import interp_oracle

G = load_golden("interp_cases")
CASES = sorted({tuple(int(v) for v in k.split("_")[1:]) for k in G.files if k.startswith("x_")})


@pytest.mark.parametrize("t_in,t_out", CASES)
def test_oracle_matches_reference_call(t_in, t_out):
    x, want = G[f"x_{t_in}_{t_out}"], G[f"y_{t_in}_{t_out}"]
    np.testing.assert_array_equal(interp_oracle.interpolate_linear(x, t_out), want)


def test_oracle_matches_torch_cpu_on_random_geometries():
    rng = np.random.default_rng(3)
    for t_in, t_out in rng.integers(1, 3000, (40, 2)):
        x = torch.randn(1, 2, int(t_in), generator=torch.Generator().manual_seed(int(t_in) * 7 + int(t_out)))
        want = torch.nn.functional.interpolate(x, size=int(t_out), mode="linear", align_corners=False).numpy()
        np.testing.assert_array_equal(interp_oracle.interpolate_linear(x.numpy(), int(t_out)), want)


def test_host_mirror_rejects_cpu_and_bad_rank():
    from neural_audio_tokenizer_b200 import align
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        align.interpolate_linear(torch.zeros(1, 2, 8), 4)
    with pytest.raises(ValueError):
        align.interpolate_linear(torch.zeros(2, 8), 4)


@pytest.mark.gpu
@pytest.mark.parametrize("t_in,t_out", CASES)
def test_device_kernel_is_bit_exact(t_in, t_out):
    from neural_audio_tokenizer_b200 import align
    x, want = G[f"x_{t_in}_{t_out}"], G[f"y_{t_in}_{t_out}"]
    got = align.interpolate_linear(torch.from_numpy(x).cuda(), t_out)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
def test_align_time_bases_and_namespace_shim():
    import types
    import torch.nn.functional as F
    from neural_audio_tokenizer_b200 import align
    sem = torch.randn(1, 16, 131, device="cuda")
    ac = torch.randn(1, 24, 128, device="cuda")
    s2, a2 = align.align_time_bases(sem, ac)
    assert s2.shape == (1, 16, 128) and a2 is ac
    np.testing.assert_array_equal(s2.cpu().numpy(), interp_oracle.interpolate_linear(sem.cpu().numpy(), 128))
    # long stream: one hour of 75 Hz frames shortened by one frame, against the oracle
    x = torch.randn(1, 3, 270001, device="cuda")
    np.testing.assert_array_equal(align.interpolate_linear(x, 270000).cpu().numpy(),
                                  interp_oracle.interpolate_linear(x.cpu().numpy(), 270000))
    fake = types.SimpleNamespace(F=F)
    align.install(fake)
    y = fake.F.interpolate(sem, size=100, mode="linear", align_corners=False)         # routed to the device kernel
    np.testing.assert_array_equal(y.cpu().numpy(), interp_oracle.interpolate_linear(sem.cpu().numpy(), 100))
    z = fake.F.interpolate(sem, size=100, mode="nearest")                               # anything else stays torch's
    assert torch.equal(z, F.interpolate(sem, size=100, mode="nearest"))
    assert fake.F.softmax is F.softmax and F.interpolate is not fake.F.interpolate
