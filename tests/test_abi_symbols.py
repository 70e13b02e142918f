"""The C-ABI library loads and exports every symbol include/nat_b200.h declares (no compute calls: CPU only)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nat_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from neural_audio_tokenizer_b200 import _lib
    return _lib


def test_header_symbols_are_exported_and_bound(lib):
    declared = _declared_symbols()
    assert len(declared) >= 15
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in nat_b200.h but not exported by libnat_b200.so"
    assert sorted(lib.EXPORTED_SYMBOLS) == declared, "ctypes binding and header disagree"


def test_abi_version_and_pure_host_helpers(lib):
    handle = lib.load()
    assert handle.nat_abi_version() == 2
    assert handle.nat_mel_num_frames(22050, 512) == 44                 # 1 + S // hop (center=True)
    assert handle.nat_mel_num_frames(24000 * 3600, 320) == 270001
    assert handle.nat_spectral_num_frames(22050, 2048, 512) == 40      # nat.py:2400-2403
    assert handle.nat_spectral_num_frames(1000, 2048, 512) == 1
    assert handle.nat_rvq_workspace_bytes(None, 0) > 0


def test_sass_is_blackwell_native():
    """The built library carries tcgen05 MMA, TMEM loads and TMA loads (SASS names per B200_PROFILING.md)."""
    import shutil
    import subprocess
    from neural_audio_tokenizer_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass                                    # no legacy mma.sync path
