"""The oracle against the golden vectors minted from the reference (oracle/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_rvq_inputs, load_golden
from oracle import mel_oracle, rvq_oracle

RVQ_CASES = ["rvq_small", "rvq_ragged", "rvq_ties", "rvq_single_frame", "rvq_768x1024", "rvq_512x4096",
             "rvq_1024x1024"]


@pytest.mark.parametrize("name", RVQ_CASES)
def test_rvq_oracle_matches_reference_codes(name):
    g = load_golden(name)
    x, cbs = golden_rvq_inputs(g)
    quantized, codes, losses = rvq_oracle.rvq_forward(x, list(cbs))
    got = np.stack([c.numpy() for c in codes])
    ref = g["codes"]
    if not np.array_equal(got, ref):        # other BLAS build/threading: only near-ties may move
        x_nd = x.permute(0, 2, 1).reshape(-1, x.shape[1]).numpy()
        rep = rvq_oracle.classify_mismatches(x_nd, [c.numpy() for c in cbs], ref.reshape(ref.shape[0], -1),
                                             got.reshape(got.shape[0], -1))
        assert rep["real_mismatches"] == 0, rep["flips"][:5]
    assert abs(float(losses["vq_loss"]) - float(g["vq_loss"])) <= 2e-6 * abs(float(g["vq_loss"]))
    assert losses["num_layers"] == int(g["L"])
    np.testing.assert_array_equal(quantized[:, :8, :8].numpy(), g["quantized_head"])
    if "quantized" in g.files:
        np.testing.assert_array_equal(quantized.numpy(), g["quantized"])
        np.testing.assert_array_equal(rvq_oracle.rvq_decode(codes, list(cbs)).numpy(), g["decoded"])


def test_ties_pick_lowest_index():
    g = load_golden("rvq_ties")
    assert int(g["codes"][0, 0, 0]) == 3          # frame 0 equals code 3, which is duplicated at row 7


def test_exact_argmin_agrees_on_small_case():
    g = load_golden("rvq_small")
    x, cbs = golden_rvq_inputs(g)
    rows = x[0].T.numpy()
    np.testing.assert_array_equal(rvq_oracle.exact_argmin_f64(rows, cbs[0].numpy()), g["codes"][0, 0])


def test_classifier_counts_flip_and_cascade():
    g = load_golden("rvq_small")
    x, cbs = golden_rvq_inputs(g)
    ref = g["codes"].reshape(4, -1)
    test = ref.copy()
    test[1, 5] = (test[1, 5] + 1) % 128
    test[2, 5] = (test[2, 5] + 3) % 128
    rep = rvq_oracle.classify_mismatches(x[0].T.numpy(), [c.numpy() for c in cbs], ref, test)
    assert rep["real_mismatches"] == 1 and rep["cascade_tokens"] == 1 and rep["near_tie_flips"] == 0
    assert rep["flips"][0]["layer"] == 1 and rep["flips"][0]["frame"] == 5


def test_shape_errors_match_reference_conditions():
    cb = [torch.randn(16, 8)]
    with pytest.raises(ValueError):
        rvq_oracle.rvq_forward(torch.randn(8), cb)
    with pytest.raises(ValueError):
        rvq_oracle.rvq_forward(torch.randn(1, 9, 4), cb)


# Mel tolerance: the reference computes in fp32 (torch.stft on MKL, fp32 linspace/pow for the filterbank); the
# oracle is fp64 inside. The fp32 filterbank itself is only defined to ~2e-5 absolute (1-ulp differences in
# torch.linspace / powf move the narrow triangles), so the stated bound is
#   |oracle - torchaudio| <= 1e-4 * max(mel) + 1e-6 per clip        
# and 2e-5 * max(mel) when the reference's own filterbank is supplied.
MEL_RTOL_OF_MAX = 1e-4
MEL_RTOL_OF_MAX_GIVEN_FB = 2e-5


@pytest.mark.parametrize("name", ["mel_tone_22050_hop512", "mel_noise_24000_hop320"])
def test_mel_oracle_matches_torchaudio(name):
    g = load_golden(name)
    fb = mel_oracle.mel_filterbank(int(g["sr"]), 2048, 128)
    np.testing.assert_allclose(fb, g["fb"], rtol=0, atol=5e-5)
    assert np.array_equal(fb == 0, g["fb"] == 0) or np.abs(fb - g["fb"])[(fb == 0) != (g["fb"] == 0)].max() < 5e-5
    mel = mel_oracle.mel_power(g["wave"], int(g["sr"]), 2048, int(g["hop"]), 128)
    ref = g["mel"][0]
    assert mel.shape == ref.shape
    assert np.abs(mel - ref).max() <= MEL_RTOL_OF_MAX * ref.max() + 1e-6
    mel_fb = (mel_oracle.stft_power(g["wave"], 2048, int(g["hop"])).T @ g["fb"].astype(np.float64)).T
    assert np.abs(mel_fb - ref).max() <= MEL_RTOL_OF_MAX_GIVEN_FB * ref.max() + 1e-7


@pytest.mark.parametrize("name", ["spectral_tone_22050", "spectral_noise_24000", "spectral_short"])
def test_spectral_oracle_matches_reference(name):
    g = load_golden(name)
    st = mel_oracle.spectral_stats(g["wave"], int(g["sr"]))
    assert st.shape == g["stats"].shape
    np.testing.assert_allclose(st, g["stats"], rtol=2e-4, atol=1e-2)


def test_pipeline_fixture_is_consistent():
    """Config 1: codes in the NDJSON frames are what the RVQ oracle gives on the captured quantiser inputs."""
    g = load_golden("pipeline_tone_argmin")
    for feats, cbs, key in ((g["sem_in"], g["sem_codebooks"], "S"), (g["ac_in"], g["ac_codebooks"], "A")):
        codes = rvq_oracle.rvq_encode(torch.from_numpy(feats), [torch.from_numpy(c) for c in cbs])
        got = np.stack([c[0].numpy() for c in codes], axis=1)          # [T, L]
        np.testing.assert_array_equal(got, g[key])
    mel = mel_oracle.mel_power(g["audio"], int(g["sr"]), 2048, 512, 128)
    assert np.abs(mel - g["mel"][0]).max() <= MEL_RTOL_OF_MAX * g["mel"].max() + 1e-6


SAMPLING_CASES = ["rvq_sampling_small", "rvq_sampling_mixed", "rvq_sampling_wide"]


@pytest.mark.parametrize("name", SAMPLING_CASES)
def test_sampling_oracle_matches_reference_default_mode(name):
    """The reference's DEFAULT selection (use_stochastic=True, nat.py:2150-2154) with a seeded global generator:
    the oracle's spelled-out multinomial (exponential_ then argmax of probs / q) reproduces its codes bit for bit."""
    g = load_golden(name)
    x, cbs = torch.from_numpy(g["x"]), [torch.from_numpy(c) for c in g["codebooks"]]
    torch.manual_seed(int(g["noise_seed"]))
    quantized, codes, losses, aux = rvq_oracle.rvq_forward_sampling(x, cbs, [float(t) for t in g["temperatures"]])
    np.testing.assert_array_equal(np.stack([c.numpy() for c in codes]), g["codes"])
    np.testing.assert_array_equal(quantized.numpy(), g["quantized"])
    assert abs(float(losses["vq_loss"]) - float(g["vq_loss"])) <= 2e-6 * abs(float(g["vq_loss"]))
    assert [a is None for a in aux] == [float(t) <= 0 for t in g["temperatures"]]


def test_sampling_classifier_flags_real_flip():
    g = load_golden("rvq_sampling_small")
    x, cbs = torch.from_numpy(g["x"]), [torch.from_numpy(c) for c in g["codebooks"]]
    torch.manual_seed(int(g["noise_seed"]))
    _, codes, _, aux = rvq_oracle.rvq_forward_sampling(x, cbs, [float(t) for t in g["temperatures"]])
    ref = np.stack([c.numpy().reshape(-1) for c in codes])
    test = ref.copy()
    test[0, 3] = (test[0, 3] + 1) % 128
    rep = rvq_oracle.classify_sampling_mismatches(ref, test, aux)
    assert rep["real_mismatches"] == 1 and rep["near_tie_flips"] == 0 and rep["flips"][0]["frame"] == 3
