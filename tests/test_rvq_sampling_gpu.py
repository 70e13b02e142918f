"""Sampling mode on the device: the reference's default selection, nat.py:2150-2154.

host_noise: the host draws the Exp(1) noise behind torch.multinomial from torch's CPU generator in the reference's
order, the device does everything else -> codes equal the reference's golden vectors except at near-ties of
probs / q (relative gap < 1e-4 in the reference's own values; counted and reported).
philox: device noise; checked in distribution against the oracle's probabilities."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer, VectorQuantizer
from oracle import rvq_oracle

pytestmark = pytest.mark.gpu

CASES = ["rvq_sampling_small", "rvq_sampling_mixed", "rvq_sampling_wide"]


def _module(g, dev="cuda"):
    cbs = torch.from_numpy(g["codebooks"])
    L, K, D = cbs.shape
    rvq = ResidualVectorQuantizer(D, K, L).eval()
    for q, cb, t in zip(rvq.quantizers, cbs, g["temperatures"]):
        q.codebook.copy_(cb)
        q.use_stochastic = bool(t > 0)
        if t > 0:
            q.temperature = float(t)
    return rvq.to(dev)


@pytest.mark.parametrize("name", CASES)
def test_host_noise_sampling_reproduces_reference_codes(name):
    g = load_golden(name)
    rvq = _module(g)
    x = torch.from_numpy(g["x"])
    torch.manual_seed(int(g["noise_seed"]))
    with torch.no_grad():
        quantized, codes, losses = rvq(x.cuda())
    got = np.stack([c.cpu().numpy().reshape(-1) for c in codes])
    ref = g["codes"].reshape(got.shape[0], -1)
    torch.manual_seed(int(g["noise_seed"]))
    _, _, _, aux = rvq_oracle.rvq_forward_sampling(x, [torch.from_numpy(c) for c in g["codebooks"]],
                                                   [float(t) for t in g["temperatures"]])
    rep = rvq_oracle.classify_sampling_mismatches(ref, got, aux)
    print(name, {k: rep[k] for k in ("frames", "exact_frames", "near_tie_flips", "real_mismatches")})
    assert rep["real_mismatches"] == 0, rep["flips"][:5]
    assert rep["near_tie_flips"] <= max(1, rep["frames"] // 100)
    if rep["mismatched_tokens"] == 0:
        np.testing.assert_array_equal(quantized.cpu().numpy(), g["quantized"])
        assert abs(losses["vq_loss"].item() - float(g["vq_loss"])) <= 2e-6 * abs(float(g["vq_loss"]))
    # encode() draws the same noise and returns the same codes
    torch.manual_seed(int(g["noise_seed"]))
    again = np.stack([c.cpu().numpy().reshape(-1) for c in rvq.encode(x.cuda())])
    np.testing.assert_array_equal(again, got)


def test_single_layer_sampling_matches_oracle():
    torch.manual_seed(3)
    vq = VectorQuantizer(32, 64, temperature=0.7).eval()
    cb = vq.codebook.clone()
    x = torch.randn(2, 32, 40, generator=torch.Generator().manual_seed(11))
    vq = vq.cuda()
    torch.manual_seed(21)
    quantized, codes, loss = vq(x.cuda())
    torch.manual_seed(21)
    ref_q, ref_codes, ref_loss, probs, q = rvq_oracle.vq_layer_sampling(x, cb, 0.7)
    rep = rvq_oracle.classify_sampling_mismatches(ref_codes.reshape(1, -1).numpy(), codes.cpu().reshape(1, -1).numpy(),
                                                  [(probs, q)])
    assert rep["real_mismatches"] == 0, rep["flips"][:3]
    assert codes.shape == (2, 40) and codes.dtype == torch.int64
    if rep["mismatched_tokens"] == 0:
        assert torch.equal(quantized.cpu(), ref_q)
        assert abs(loss.item() - float(ref_loss)) <= 2e-6 * float(ref_loss)


def test_sampling_does_not_return_argmin():
    """With a high temperature the draws must differ from the nearest code for most frames."""
    g = load_golden("rvq_sampling_small")
    rvq = _module(g)
    x = torch.from_numpy(g["x"]).cuda()
    torch.manual_seed(0)
    sampled = rvq.encode(x)[0].cpu()
    for q in rvq.quantizers:
        q.use_stochastic = False
    nearest = rvq.encode(x)[0].cpu()
    assert (sampled != nearest).float().mean().item() > 0.2


def test_philox_mode_matches_the_distribution():
    """One frame repeated N times: every repetition has its own Philox counter, so the histogram of first-layer
    codes must follow softmax(-d / T) (chi-square over the codes that carry mass; 6-sigma bound)."""
    torch.manual_seed(5)
    K, D, N, T = 16, 24, 40000, 2.0
    rvq = ResidualVectorQuantizer(D, K, 1, temperature=T).eval()
    cb = rvq.quantizers[0].codebook.clone()
    frame = torch.randn(D, generator=torch.Generator().manual_seed(2)) * 0.3
    x = frame[None, :, None].repeat(1, 1, N).contiguous()
    rvq = rvq.cuda()
    rvq.sampling_mode = "philox"
    codes = rvq.encode(x.cuda())[0].cpu().reshape(-1)
    p = torch.softmax(-torch.cdist(frame[None].double(), cb.double())[0] / T, dim=0).numpy()
    counts = np.bincount(codes.numpy(), minlength=K).astype(np.float64)
    keep = p * N >= 20
    chi2 = float((((counts - p * N) ** 2) / (p * N))[keep].sum())
    dof = int(keep.sum()) - 1
    assert chi2 < dof + 6.0 * np.sqrt(2.0 * dof), (chi2, dof)
    # a second call advances the draw counter: different codes, same distribution
    codes2 = rvq.encode(x.cuda())[0].cpu().reshape(-1)
    assert (codes2 != codes).float().mean().item() > 0.3


def test_host_noise_refuses_bulk_inputs():
    rvq = ResidualVectorQuantizer(8, 4096, 1).eval().cuda()
    with pytest.raises(RuntimeError, match="philox"):
        rvq(torch.zeros(1, 8, 100000, device="cuda"))


def test_philox_tensor_core_path_agrees_with_exact_scan():
    """Same Philox stream, distances from the tensor-core pass instead of the exact scan: the chosen codes may differ
    only where the two best values of -d/T - log q nearly tie. Layer by layer on identical residuals (L = 1 stacks)."""
    torch.manual_seed(12)
    D, K, N = 768, 1024, 20000
    x = torch.randn(1, D, N, device="cuda")
    for temperature in (0.5, 4.0):
        rvq = ResidualVectorQuantizer(D, K, 1, temperature=temperature).eval().cuda()
        rvq.sampling_mode = "philox_exact"
        exact = rvq.encode(x)[0]
        rvq.sampling_mode = "philox"
        rvq._draws = 0                                  # same draw counter as the first call
        fast = rvq.encode(x)[0]
        agree = (exact == fast).float().mean().item()
        print(f"T={temperature}: tensor-core Philox path agrees with the exact scan on {agree:.5f} of {N} frames")
        assert agree > 0.995
    # whole 4-layer stack: runs, in range, and is not the argmin
    rvq = ResidualVectorQuantizer(D, K, 4).eval().cuda()
    rvq.sampling_mode = "philox"
    q, codes, losses = rvq(x[:, :, :3000])
    assert len(codes) == 4 and all(int(c.min()) >= 0 and int(c.max()) < K for c in codes)
    assert torch.isfinite(q).all() and torch.isfinite(losses["vq_loss"])
