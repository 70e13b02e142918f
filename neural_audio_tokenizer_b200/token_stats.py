"""Token statistics over the index streams: host mirror of the reference's diversity,
entropy and mutual-information figures (nat.py:4913-4917, 3442-3447, 3577-3637).

The reference moves every stream to the host and runs `torch.unique` / `np.unique` / `np.histogram2d` there. Here
the integer work -- a pooled histogram per stack, and the 2-D histogram -- runs on the device over the streams where
they already are (`nat_token_histogram`, `nat_token_joint_histogram`); only the K counts (or bins x bins) cross PCIe.
The handful of floating-point operations on those counts are the reference's own numpy / scipy calls, applied to the
same integers in the same order, so the figures are equal to the reference's, not approximations of them.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib

_CODE_DTYPES = {torch.int64: _lib.CODES_I64, torch.int32: _lib.CODES_I32, torch.int16: _lib.CODES_I16}


def _flat_cuda(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"token stream is on {t.device}: the B200 statistics path has no CPU fallback")
    t = t.reshape(-1)
    if t.dtype not in _CODE_DTYPES:
        t = t.to(torch.int64)
    return t.contiguous()


def pooled_counts(streams: Sequence[torch.Tensor], vocab: int) -> np.ndarray:
    """int64 [vocab]: occurrences of every token over all the given streams (the reference pools a stack's layers
    with `torch.cat([codes.flatten() ...])`, nat.py:4913). Raises if a token lies outside [0, vocab)."""
    if not streams:
        return np.zeros(vocab, dtype=np.int64)
    lib = _lib.load()
    streams = [_flat_cuda(s) for s in streams]
    dev = streams[0].device
    with torch.cuda.device(dev):
        counts = torch.zeros(vocab + 1, dtype=torch.int64, device=dev)       # last slot: tokens outside the vocabulary
        st = torch.cuda.current_stream(dev).cuda_stream
        for f in streams:
            _lib.check(lib.nat_token_histogram(f.data_ptr(), _CODE_DTYPES[f.dtype], f.numel(), vocab, counts.data_ptr(),
                                               counts[vocab:].data_ptr(), st))
        host = counts.cpu().numpy()
    if host[vocab]:
        raise ValueError(f"{int(host[vocab])} tokens outside [0, {vocab})")
    return host[:vocab]


def diversity_from_counts(counts: np.ndarray) -> float:
    """len(torch.unique(all)) / len(all), 0 for an empty stream (nat.py:4916, 3445)."""
    total = int(counts.sum())
    return int(np.count_nonzero(counts)) / total if total > 0 else 0


def entropy_from_counts(counts: np.ndarray) -> float:
    """TokenizationEvaluator._calculate_entropy (nat.py:3577-3584): counts of the unique tokens in ascending token
    order, as float32 probabilities, through scipy.stats.entropy (natural log)."""
    from scipy.stats import entropy
    total = int(counts.sum())
    if total == 0:
        return 0.0
    nz = counts[counts > 0]
    probabilities = torch.from_numpy(nz.astype(np.int64)).float() / total
    return float(entropy(probabilities.numpy()))


def token_diversity(semantic_codes: Sequence[torch.Tensor], acoustic_codes: Sequence[torch.Tensor], vocab: int
                    ) -> Tuple[float, float]:
    """(semantic_diversity, acoustic_diversity) of nat.py:4913-4917 from device-side histograms."""
    return (diversity_from_counts(pooled_counts(semantic_codes, vocab)),
            diversity_from_counts(pooled_counts(acoustic_codes, vocab)))


def mutual_information(tokens_a: torch.Tensor, tokens_b: torch.Tensor, vocab: int) -> float:
    """TokenizationEvaluator._calculate_mutual_information (nat.py:3586-3637) with the histograms on the device.

    The bin count min(64, max(#unique a, #unique b, 2)) and the value ranges come from the 1-D histograms, the edges
    from numpy's own linspace (what histogram2d uses), the 2-D counts from the device, the rest is the reference's
    float64 arithmetic."""
    na, nb = tokens_a.numel(), tokens_b.numel()
    if na == 0 or nb == 0:
        return 0.0
    n = min(na, nb)
    a = _flat_cuda(tokens_a)[:n]
    b = _flat_cuda(tokens_b)[:n]
    if a.dtype != b.dtype:
        a, b = a.to(torch.int64), b.to(torch.int64)
    ca, cb = pooled_counts([a], vocab), pooled_counts([b], vocab)
    bins = min(64, max(int(np.count_nonzero(ca)), int(np.count_nonzero(cb)), 2))

    def edges(c):                                   # numpy/lib/_histograms_impl.py: _get_outer_edges + linspace
        nz = np.nonzero(c)[0]
        lo, hi = float(nz[0]), float(nz[-1])
        if lo == hi:
            lo, hi = lo - 0.5, hi + 0.5
        return np.linspace(lo, hi, bins + 1)

    ea, eb = edges(ca), edges(cb)
    lib = _lib.load()
    dev = a.device
    with torch.cuda.device(dev):
        hist = torch.zeros((bins, bins), dtype=torch.int64, device=dev)
        ea_d, eb_d = torch.from_numpy(ea).to(dev), torch.from_numpy(eb).to(dev)
        _lib.check(lib.nat_token_joint_histogram(a.data_ptr(), b.data_ptr(), _CODE_DTYPES[a.dtype], n, ea_d.data_ptr(),
                                                 eb_d.data_ptr(), bins, hist.data_ptr(),
                                                 torch.cuda.current_stream(dev).cuda_stream))
        hist_2d = hist.cpu().numpy().astype(np.float64)
    return mi_from_histogram(hist_2d)


def mi_from_histogram(hist_2d: np.ndarray) -> float:
    """The float64 tail of nat.py:3610-3634, verbatim in behaviour: probabilities, masked log2 ratio, NaN -> 0."""
    total_count = hist_2d.sum()
    if total_count == 0:
        return 0.0
    pxy = hist_2d / total_count
    px = pxy.sum(axis=1, keepdims=True)
    py = pxy.sum(axis=0, keepdims=True)
    mask = pxy > 1e-12
    pxy_nz = pxy[mask]
    if len(pxy_nz) == 0:
        return 0.0
    px_nz = np.broadcast_to(px, pxy.shape)[mask]
    py_nz = np.broadcast_to(py, pxy.shape)[mask]
    mi_val = np.sum(pxy_nz * np.log2(pxy_nz / (px_nz * py_nz + 1e-12)))
    return float(mi_val) if not np.isnan(mi_val) else 0.0
