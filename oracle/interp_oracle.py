"""CPU restatement of the time-base alignment `F.interpolate(x, size=T, mode='linear', align_corners=False)` that the
reference applies before quantisation (nat.py:3225-3236).

TEST INFRASTRUCTURE ONLY (tests/ and nothing else may import this).

The arithmetic is third-party (torch 2.11.0 ATen, UpSampleKernel.cpp: `area_pixel_compute_scale`,
`compute_source_index_and_lambda`, two-tap interpolation), restated in numpy with float32 steps; the two fused
multiply-adds are how this image's CPU kernel is compiled (found by bisection against torch itself, pinned by
tests/golden/interp_cases.npz which oracle/make_golden.py mints with the reference's own call).
"""
from __future__ import annotations

import numpy as np


def _fma32(a, b, c):
    """float32 fma(a, b, c): the product of two float32 is exact in float64 and the sum rounds once."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(np.float32)


def interpolate_linear(x: np.ndarray, t_out: int) -> np.ndarray:
    """x: [..., T_in] float32 -> [..., t_out] float32."""
    x = np.asarray(x, dtype=np.float32)
    t_in = x.shape[-1]
    scale = np.float32(t_in) / np.float32(t_out)
    i = np.arange(t_out, dtype=np.float32)
    real = np.maximum(_fma32(scale, i + np.float32(0.5), np.float32(-0.5)), np.float32(0))
    i0 = np.minimum(real.astype(np.int64), t_in - 1)
    i1 = i0 + (i0 < t_in - 1)
    l1 = np.clip(real - i0.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return _fma32(l0, x[..., i0], (l1 * x[..., i1]).astype(np.float32))
