"""Token statistics: the oracle against figures minted from the reference's own
TokenizationEvaluator (CPU), and the device histograms + host mirror against both (GPU). Bar: equal floats."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import stats_oracle

with open(os.path.join(os.path.dirname(__file__), "golden", "token_stats.json")) as f:
    CASES = json.load(f)


def _streams(c, key, device="cpu", dtype=torch.int64):
    return [torch.tensor(s, dtype=dtype, device=device)[None] for s in c[key]]


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_evaluator(name):
    c = CASES[name]
    sem, ac = _streams(c, "semantic"), _streams(c, "acoustic")
    all_s, all_a = torch.cat([t.flatten() for t in sem]), torch.cat([t.flatten() for t in ac])
    assert stats_oracle.diversity(sem) == c["semantic_diversity"]
    assert stats_oracle.diversity(ac) == c["acoustic_diversity"]
    assert stats_oracle.token_entropy(all_s) == c["semantic_entropy"]
    assert stats_oracle.token_entropy(all_a) == c["acoustic_entropy"]
    assert stats_oracle.mutual_information(all_s, all_a) == c["mutual_information"]


def test_host_tail_of_mutual_information_matches_oracle():
    """mi_from_histogram is the reference's float64 tail applied to an integer histogram (no GPU needed)."""
    from neural_audio_tokenizer_b200.token_stats import diversity_from_counts, entropy_from_counts, mi_from_histogram
    c = CASES["sticky_1024"]
    a = np.concatenate(c["semantic"])
    b = np.concatenate(c["acoustic"])
    n = min(len(a), len(b))
    bins = min(64, max(len(np.unique(a[:n])), len(np.unique(b[:n])), 2))
    hist, _, _ = np.histogram2d(a[:n], b[:n], bins=bins)
    assert mi_from_histogram(hist) == c["mutual_information"]
    counts = np.bincount(a, minlength=c["vocab"])
    assert diversity_from_counts(counts) == c["semantic_diversity"]
    assert entropy_from_counts(counts) == c["semantic_entropy"]
    assert diversity_from_counts(np.zeros(8, dtype=np.int64)) == 0 and entropy_from_counts(np.zeros(8, dtype=np.int64)) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.int64, torch.int16])
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_statistics_equal_reference(name, dtype):
    from neural_audio_tokenizer_b200 import token_stats as ts
    c = CASES[name]
    sem, ac = _streams(c, "semantic", "cuda", dtype), _streams(c, "acoustic", "cuda", dtype)
    ds, da = ts.token_diversity(sem, ac, c["vocab"])
    assert ds == c["semantic_diversity"] and da == c["acoustic_diversity"]
    cs, ca = ts.pooled_counts(sem, c["vocab"]), ts.pooled_counts(ac, c["vocab"])
    np.testing.assert_array_equal(cs, np.bincount(np.concatenate(c["semantic"]), minlength=c["vocab"]))
    assert ts.entropy_from_counts(cs) == c["semantic_entropy"]
    assert ts.entropy_from_counts(ca) == c["acoustic_entropy"]
    all_s, all_a = torch.cat([t.flatten() for t in sem]), torch.cat([t.flatten() for t in ac])
    assert ts.mutual_information(all_s, all_a, c["vocab"]) == c["mutual_information"]


@pytest.mark.gpu
def test_device_histograms_large_and_edge_cases():
    from neural_audio_tokenizer_b200 import token_stats as ts
    g = torch.Generator().manual_seed(4)
    a = torch.randint(0, 4096, (8, 270_000), generator=g, dtype=torch.int64)
    b = (a // 3 + torch.randint(0, 5, a.shape, generator=g)) % 4096
    counts = ts.pooled_counts([a.cuda()], 4096)
    np.testing.assert_array_equal(counts, np.bincount(a.numpy().ravel(), minlength=4096))
    big = ts.pooled_counts([a.cuda().to(torch.int32)], 32768)                 # global-atomic path (vocab > 12 288 bins)
    np.testing.assert_array_equal(big[:4096], counts)
    assert big[4096:].sum() == 0
    assert ts.mutual_information(a.cuda(), b.cuda(), 4096) == stats_oracle.mutual_information(a, b)
    assert ts.mutual_information(a.cuda()[:0], b.cuda(), 4096) == 0.0
    with pytest.raises(ValueError, match="outside"):
        ts.pooled_counts([torch.tensor([5, 99], device="cuda")], 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ts.pooled_counts([torch.tensor([1])], 64)
