"""Oracle for the residual vector quantiser (argmin contract, and the sampling mode at the end of the file).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with plain torch CPU ops and no nn.Module machinery, what the reference computes in
  * `VectorQuantizer.forward`          /root/reference/neural_audio_tokenizer.py:2119-2183 (argmin branch 2155-2157)
  * `VectorQuantizer.decode`           nat.py:2185-2203
  * `ResidualVectorQuantizer.forward`  nat.py:1358-1420
  * `ResidualVectorQuantizer.decode`   nat.py:1428-1446
The distance is `torch.cdist` on purpose: on CPU that is the same MKL sgemm the reference hits, so
on one machine this oracle is bit-identical to the reference, which `tests/test_oracle_vs_reference.py` asserts in
the authoring container and `tests/golden/rvq_*.npz` pin everywhere else.

`classify_mismatches` is the parity checker BASELINE.md section 3 asks for: per frame it finds the first layer whose
index differs, recomputes both candidates' distances in float64 on the oracle's residual and calls the flip a
near-tie when the relative gap is below 1e-6; later layers of that frame are cascade.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NEAR_TIE_REL_GAP = 1e-6


def _as_bct(x: torch.Tensor, input_dim: int) -> torch.Tensor:
    if x.dim() not in (2, 3):
        raise ValueError(f"Expected 2D or 3D input tensor, got {tuple(x.shape)}")
    if x.dim() == 2:
        x = x.unsqueeze(0)
    if x.shape[1] != input_dim:
        raise ValueError(f"Expected {input_dim} feature dimensions, got {x.shape[1]}")
    return x


def vq_layer(x_bct: torch.Tensor, codebook: torch.Tensor, commitment_weight: float = 0.25):
    """One VQ layer in eval/argmin mode. x_bct: [B, C, T] fp32. Returns (quantized [B,C,T], codes [B,T] int64, loss)."""
    B, C, T = x_bct.shape
    flat = x_bct.transpose(1, 2).contiguous().view(-1, C)          # nat.py:2141-2142
    dist = torch.cdist(flat, codebook)                              # nat.py:2146
    idx = torch.argmin(dist, dim=1)                                 # nat.py:2157 (first minimum wins)
    q = F.embedding(idx, codebook)                                  # nat.py:2159
    mse = F.mse_loss(q, flat)                                       # nat.py:2162-2163 (same value twice in no-grad)
    loss = mse + commitment_weight * mse                            # nat.py:2164
    q_ste = flat + (q - flat)                                       # nat.py:2167
    quantized = q_ste.view(B, T, C).transpose(1, 2).contiguous()    # nat.py:2170-2171
    return quantized, idx.view(B, T), loss


def rvq_forward(x: torch.Tensor, codebooks: Sequence[torch.Tensor], commitment_weight: float = 0.25
                ) -> Tuple[torch.Tensor, List[torch.Tensor], Dict[str, object]]:
    """The L-layer chain of nat.py:1393-1415 on CPU tensors, argmin selection."""
    x = _as_bct(x, codebooks[0].shape[1])
    with torch.no_grad():
        residual = x
        layers, codes = [], []
        total = 0
        for cb in codebooks:
            quantized, code, loss = vq_layer(residual, cb, commitment_weight)
            layers.append(quantized)
            codes.append(code)
            total = total + loss
            residual = residual - quantized                          # nat.py:1405
        final = sum(layers)                                          # nat.py:1408
    return final, codes, {"vq_loss": total, "num_layers": len(layers)}


def rvq_encode(x: torch.Tensor, codebooks: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    return rvq_forward(x, codebooks)[1]


def vq_decode(codes: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    squeeze = codes.dim() == 1
    if squeeze:
        codes = codes.unsqueeze(0)
    B, T = codes.shape
    q = F.embedding(codes.reshape(-1), codebook).view(B, T, codebook.shape[1]).transpose(1, 2).contiguous()
    return q.squeeze(0) if squeeze else q


def rvq_decode(codes: Sequence[torch.Tensor], codebooks: Sequence[torch.Tensor]) -> torch.Tensor:
    D = codebooks[0].shape[1]
    if not codes:
        return torch.zeros(1, D, 1)
    B, T = codes[0].shape
    out = torch.zeros(B, D, T, dtype=torch.float)
    for i, code in enumerate(codes):
        if i < len(codebooks):
            out += vq_decode(code, codebooks[i])
    return out


# ----------------------------------------------------------------------------------------------------------------
# Parity classification
# ----------------------------------------------------------------------------------------------------------------

def residual_before_layer(x_rows: np.ndarray, codebooks: Sequence[np.ndarray], codes: np.ndarray, layer: int
                          ) -> np.ndarray:
    """fp32 residual entering `layer` for the given rows, replaying the reference op order (2167 then 1405)."""
    r = x_rows.astype(np.float32).copy()
    for l in range(layer):
        q = codebooks[l][codes[l]].astype(np.float32)
        q_ste = r + (q - r)
        r = r - q_ste
    return r


def classify_mismatches(x_nd: np.ndarray, codebooks: Sequence[np.ndarray], codes_ref: np.ndarray,
                        codes_test: np.ndarray, rel_gap: float = NEAR_TIE_REL_GAP) -> Dict[str, object]:
    """x_nd: [N, D] fp32 rows. codes_*: [L, N] ints. Returns counts + the list of primary flips with their gaps."""
    codes_ref = np.asarray(codes_ref).astype(np.int64)
    codes_test = np.asarray(codes_test).astype(np.int64)
    L, N = codes_ref.shape
    assert codes_test.shape == (L, N), (codes_test.shape, (L, N))
    diff = codes_ref != codes_test
    out = {"frames": int(N), "layers": int(L), "exact_frames": int((~diff.any(axis=0)).sum()),
           "mismatched_tokens": int(diff.sum()), "near_tie_flips": 0, "real_mismatches": 0, "cascade_tokens": 0,
           "flips": []}
    bad_rows = np.nonzero(diff.any(axis=0))[0]
    for n in bad_rows:
        first = int(np.argmax(diff[:, n]))
        r = residual_before_layer(x_nd[n:n + 1], codebooks, codes_ref[:, n:n + 1], first)[0].astype(np.float64)
        cb = codebooks[first].astype(np.float64)
        da = float(np.sqrt(((r - cb[codes_ref[first, n]]) ** 2).sum()))
        db = float(np.sqrt(((r - cb[codes_test[first, n]]) ** 2).sum()))
        gap = abs(da - db) / max(da, db, 1e-300)
        kind = "near_tie" if gap < rel_gap else "real"
        out["near_tie_flips" if kind == "near_tie" else "real_mismatches"] += 1
        out["cascade_tokens"] += int(diff[first + 1:, n].sum())
        out["flips"].append({"frame": int(n), "layer": first, "ref": int(codes_ref[first, n]),
                             "test": int(codes_test[first, n]), "rel_gap": gap, "kind": kind})
    return out


def exact_argmin_f64(rows: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """True nearest code in float64 (first index on ties); the arbiter for tiny cases."""
    r = rows.astype(np.float64)
    c = codebook.astype(np.float64)
    d2 = (r * r).sum(1)[:, None] - 2.0 * r @ c.T + (c * c).sum(1)[None, :]
    return np.argmin(d2, axis=1)


# ----------------------------------------------------------------------------------------------------------------
# Sampling mode (the reference's default selection, nat.py:2150-2154)
# ----------------------------------------------------------------------------------------------------------------

SAMPLING_NEAR_TIE_REL_GAP = 1e-4


def vq_layer_sampling(x_bct: torch.Tensor, codebook: torch.Tensor, temperature: float, commitment_weight: float = 0.25,
                      generator: torch.Generator = None):
    """One VQ layer with `use_stochastic=True` in eval mode. `torch.multinomial(probs, 1)` (nat.py:2154) is spelled
    out as ATen evaluates it on CPU -- q = empty_like(probs).exponential_(1); argmax(probs / q) -- so that the draw the
    device path has to reproduce is explicit; tests/golden/rvq_sampling_*.npz (minted from the reference's own
    multinomial call) pin that the spelling is bit-identical. Also returns probs and q for the mismatch classifier."""
    B, C, T = x_bct.shape
    flat = x_bct.transpose(1, 2).contiguous().view(-1, C)
    dist = torch.cdist(flat, codebook)                              # nat.py:2146
    probs = F.softmax(-dist / temperature, dim=1)                   # nat.py:2153
    q = torch.empty_like(probs).exponential_(1, generator=generator)
    idx = torch.argmax(probs / q, dim=-1)                           # nat.py:2154
    quant = F.embedding(idx, codebook)
    mse = F.mse_loss(quant, flat)
    loss = mse + commitment_weight * mse
    q_ste = flat + (quant - flat)
    quantized = q_ste.view(B, T, C).transpose(1, 2).contiguous()
    return quantized, idx.view(B, T), loss, probs, q


def rvq_forward_sampling(x: torch.Tensor, codebooks: Sequence[torch.Tensor], temperatures: Sequence[float],
                         commitment_weight: float = 0.25, generator: torch.Generator = None):
    """The L-layer chain with per-layer selection: temperature > 0 samples, <= 0 takes the argmin (a layer whose
    `use_stochastic` is False). Draws from the global CPU generator when `generator` is None, like the reference.
    Returns (final, codes, losses, aux) where aux[l] = (probs, q) for sampling layers and None otherwise."""
    x = _as_bct(x, codebooks[0].shape[1])
    with torch.no_grad():
        residual = x
        layers, codes, aux = [], [], []
        total = 0
        for cb, t in zip(codebooks, temperatures):
            if t > 0:
                quantized, code, loss, probs, q = vq_layer_sampling(residual, cb, t, commitment_weight, generator)
                aux.append((probs, q))
            else:
                quantized, code, loss = vq_layer(residual, cb, commitment_weight)
                aux.append(None)
            layers.append(quantized)
            codes.append(code)
            total = total + loss
            residual = residual - quantized
        final = sum(layers)
    return final, codes, {"vq_loss": total, "num_layers": len(layers)}, aux


def classify_sampling_mismatches(codes_ref: np.ndarray, codes_test: np.ndarray, aux, rel_gap: float =
                                 SAMPLING_NEAR_TIE_REL_GAP) -> Dict[str, object]:
    """codes_*: [L, N]. A frame's first differing layer is a near-tie when the reference's own probs / q values of
    the two codes differ by less than `rel_gap` relatively (the device recomputes the distances exactly, the
    reference through an fp32 sgemm: their probabilities differ by ~1e-5 relative); later layers of
    that frame are cascade (the residual and hence every later draw's meaning changed)."""
    codes_ref = np.asarray(codes_ref).astype(np.int64)
    codes_test = np.asarray(codes_test).astype(np.int64)
    L, N = codes_ref.shape
    diff = codes_ref != codes_test
    out = {"frames": int(N), "layers": int(L), "exact_frames": int((~diff.any(axis=0)).sum()),
           "mismatched_tokens": int(diff.sum()), "near_tie_flips": 0, "real_mismatches": 0, "cascade_tokens": 0,
           "flips": []}
    for n in np.nonzero(diff.any(axis=0))[0]:
        first = int(np.argmax(diff[:, n]))
        if aux[first] is None:
            kind, gap = "real", float("nan")
        else:
            probs, q = aux[first]
            v = (probs[n].double() / q[n].double()).numpy()
            a, b = v[codes_ref[first, n]], v[codes_test[first, n]]
            gap = abs(a - b) / max(abs(a), abs(b), 1e-300)
            kind = "near_tie" if gap < rel_gap else "real"
        out["near_tie_flips" if kind == "near_tie" else "real_mismatches"] += 1
        out["cascade_tokens"] += int(diff[first + 1:, n].sum())
        out["flips"].append({"frame": int(n), "layer": first, "ref": int(codes_ref[first, n]),
                             "test": int(codes_test[first, n]), "rel_gap": float(gap), "kind": kind})
    return out


# ------------------------------------------------------------------------------------------------ training mode
def ema_update(codebook: torch.Tensor, ema_count: torch.Tensor, ema_weight: torch.Tensor, flat_input: torch.Tensor,
               codes_flat: torch.Tensor, ema_decay: float = 0.99) -> None:
    """`VectorQuantizer._update_ema` (nat.py:2205-2221), in place on the three buffers."""
    with torch.no_grad():
        onehot = F.one_hot(codes_flat, codebook.shape[0]).float()
        ema_count.mul_(ema_decay).add_(onehot.sum(dim=0), alpha=1 - ema_decay)
        ema_weight.mul_(ema_decay).add_(torch.matmul(onehot.t(), flat_input), alpha=1 - ema_decay)
        codebook.copy_(ema_weight / (ema_count + 1e-5).unsqueeze(1))


def rvq_forward_training(x: torch.Tensor, codebooks: Sequence[torch.Tensor], ema_counts: Sequence[torch.Tensor],
                         ema_weights: Sequence[torch.Tensor], temperature: float = 0.5, commitment_weight: float = 0.25,
                         ema_decay: float = 0.99, generator: torch.Generator = None):
    """The layer loop of nat.py:1393-1415 with every layer in training mode (nat.py:2150: a training layer samples
    whatever `use_stochastic` says; nat.py:2179-2181: then the EMA update). Mutates the buffers in place like the
    reference. Returns (final, codes, losses)."""
    x = _as_bct(x, codebooks[0].shape[1])
    with torch.no_grad():
        residual = x
        layers, codes = [], []
        total = 0
        for cb, cnt, wgt in zip(codebooks, ema_counts, ema_weights):
            quantized, code, loss, _, _ = vq_layer_sampling(residual, cb, temperature, commitment_weight, generator)
            flat = residual.transpose(1, 2).contiguous().view(-1, cb.shape[1])
            ema_update(cb, cnt, wgt, flat, code.reshape(-1), ema_decay)
            layers.append(quantized)
            codes.append(code)
            total = total + loss
            residual = residual - quantized
        final = sum(layers)
    return final, codes, {"vq_loss": total, "num_layers": len(layers)}
