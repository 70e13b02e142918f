"""Build recipe for the sm_100a shared library `libnat_b200.so` (in-tree, next to this file).

`nvcc` cross-compiles without a GPU; the built `.so` is git-ignored but travels with `gpurun` snapshots.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")                 # kernels + their launcher: what bench.build_id hashes
CSRC_HOST = os.path.join(PKG_DIR, "csrc_host")       # host-only translation units (no kernel depends on them)
LIB_PATH = os.environ.get("NAT_B200_LIB_OUT") or os.path.join(PKG_DIR, "libnat_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def extra_flags():
    """A/B builds: NAT_B200_NVCC_FLAGS="-DNAT_REGS_UPD=88 ..." (with NAT_B200_LIB_OUT naming the output)."""
    return os.environ.get("NAT_B200_NVCC_FLAGS", "").split()


def sources():
    return [os.path.join(CSRC, "nat_b200.cu"), os.path.join(CSRC, "ndjson_emit.cpp"),
            os.path.join(CSRC_HOST, "peer_exchange.cpp")]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(CSRC_HOST, f) for f in os.listdir(CSRC_HOST)] + \
           [os.path.join(PKG_DIR, "..", "include", "nat_b200.h")]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libnat_b200.so; returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libnat_b200.so (there is no CPU fallback)")
    tmp = LIB_PATH + ".tmp"                      # built aside and renamed: a snapshot never sees a half-written library
    cmd = [nvcc, *NVCC_FLAGS, *extra_flags(), "-o", tmp, *sources()]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(f"nvcc failed with exit code {proc.returncode}")
    os.replace(tmp, LIB_PATH)
    with open(os.path.join(PKG_DIR, "libnat_b200.ptxas.log"), "w") as f:
        # registers / spills / shared memory per kernel; compile times vary from run to run and are left out
        f.write("".join(l for l in (proc.stdout + proc.stderr).splitlines(True) if "Compile time" not in l))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
