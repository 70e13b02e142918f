"""The 2048-point transform of csrc/mel_fft.cuh (three passes, 16 x 16 x 8, digit-reversed output) replayed on the host
against a direct DFT. nvcc compiles the host program without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_passes_match_direct_dft(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "fft_check"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "fft_check.cu")],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
