"""CPU restatement of the reference's token statistics.

TEST INFRASTRUCTURE ONLY (tests/ and nothing else may import this).

Follows the diversity expression of nat.py:4913-4917 / 3442-3447, `TokenizationEvaluator._calculate_entropy`
(nat.py:3577-3584) and `_calculate_mutual_information` (nat.py:3586-3637), with the host calls the reference makes
(`torch.unique`, `np.unique`, `np.histogram2d`, `scipy.stats.entropy`). Pinned by tests/golden/token_stats.json, minted
from the reference's own evaluator (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.stats import entropy


def diversity(streams) -> float:
    pooled = torch.cat([c.flatten().long().cpu() for c in streams]) if streams else torch.tensor([])
    return len(torch.unique(pooled)) / len(pooled) if len(pooled) > 0 else 0


def token_entropy(tokens: torch.Tensor) -> float:
    if len(tokens) == 0:
        return 0.0
    _, counts = torch.unique(tokens, return_counts=True)
    return float(entropy((counts.float() / len(tokens)).cpu().numpy()))


def mutual_information(tokens_a: torch.Tensor, tokens_b: torch.Tensor) -> float:
    if len(tokens_a) == 0 or len(tokens_b) == 0:
        return 0.0
    a = tokens_a.cpu().numpy().astype(np.int64).ravel()
    b = tokens_b.cpu().numpy().astype(np.int64).ravel()
    n = min(len(a), len(b))
    a, b = a[:n], b[:n]
    bins = min(64, max(len(np.unique(a)), len(np.unique(b)), 2))
    hist, _, _ = np.histogram2d(a, b, bins=bins)
    total = hist.sum()
    if total == 0:
        return 0.0
    pxy = hist / total
    px, py = pxy.sum(axis=1, keepdims=True), pxy.sum(axis=0, keepdims=True)
    mask = pxy > 1e-12
    if not mask.any():
        return 0.0
    val = np.sum(pxy[mask] * np.log2(pxy[mask] / (np.broadcast_to(px, pxy.shape)[mask] *
                                                  np.broadcast_to(py, pxy.shape)[mask] + 1e-12)))
    return float(val) if not np.isnan(val) else 0.0
