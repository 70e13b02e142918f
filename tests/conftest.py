"""Shared pytest fixtures. `gpu` marks tests that need a B200; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)


def golden_rvq_inputs(g):
    """Rebuild (x [B,D,T], codebooks [L,K,D]) of an RVQ golden case; big cases are redrawn from their seeds."""
    D, K, L, B, T = (int(g[k]) for k in ("D", "K", "L", "B", "T"))
    if "codebooks" in g.files:
        cbs = torch.from_numpy(g["codebooks"])
        x = torch.from_numpy(g["x"])
    else:
        torch.manual_seed(int(g["seed"]))
        cbs = torch.stack([torch.randn(K, D) for _ in range(L)])      # construction order of nat.py:2115
        gen = torch.Generator().manual_seed(int(g["x_seed"]))
        x = torch.randn(B, D, T, generator=gen) * float(g["x_scale"])
        assert abs(float(cbs.double().sum()) - float(g["codebook_checksum"])) < 1e-6, "seeded codebooks drifted"
        assert abs(float(x.double().sum()) - float(g["x_checksum"])) < 1e-6, "seeded input drifted"
    return x, cbs


@pytest.fixture(scope="session")
def golden():
    return load_golden
