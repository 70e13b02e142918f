"""Import the UNMODIFIED reference (`/root/reference/neural_audio_tokenizer.py`) in the authoring container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (to mint golden vectors), by the cross-check tests, by
the config-1 GPU test (the reference pipeline with the drop-in installed) and by `bench.py --impl reference` / its
`cpu_baseline` leg (the unmodified reference timed on the host cores).  Nothing in the product package imports
this.  On the GPU box `/root/reference` does not exist: the git-ignored copy `baseline/_ref/` that
`__graft_entry__.build()` makes in the authoring container travels with the gpurun snapshot and is found here.

The reference imports `librosa` and `soundfile` unconditionally (nat.py:108-110); neither is installed
here.  The recipe is: touch the transformers symbols first, then register stub
modules that carry a ModuleSpec, then import the file by path.  The reference file itself is not edited.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import tempfile
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CARRIED_DIR = os.path.join(_ROOT, "baseline", "_ref")      # git-ignored copy that travels to the GPU box (see carry_reference)


def _find_reference_dir() -> str:
    """NAT_REFERENCE_DIR, then the mounted reference (authoring container), then the carried copy (GPU box)."""
    cands = [os.environ.get("NAT_REFERENCE_DIR"), "/root/reference", CARRIED_DIR]
    for d in cands:
        if d and os.path.isfile(os.path.join(d, "neural_audio_tokenizer.py")):
            return d
    return cands[0] or "/root/reference"


REFERENCE_DIR = _find_reference_dir()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "neural_audio_tokenizer.py"))


def carry_reference(src: str = "/root/reference") -> bool:
    """Copy the unmodified reference file and its test clips into the git-ignored `baseline/_ref/` so that the
    reference arm of bench.py and the config-1 GPU test can run on the GPU box, where /root/reference does not exist
    (SURVEY.md section 7). Nothing is copied into the tracked tree. Returns True when the carried copy is present."""
    import shutil
    names = ["neural_audio_tokenizer.py", "test_simple.wav", "test_simple2.wav", "test.wav", "LICENSE"]
    if os.path.isfile(os.path.join(src, names[0])) and os.path.abspath(src) != os.path.abspath(CARRIED_DIR):
        os.makedirs(CARRIED_DIR, exist_ok=True)
        for n in names:
            a, b = os.path.join(src, n), os.path.join(CARRIED_DIR, n)
            if os.path.isfile(a) and (not os.path.isfile(b) or os.path.getmtime(a) > os.path.getmtime(b)
                                      or os.path.getsize(a) != os.path.getsize(b)):
                shutil.copy2(a, b)
    return os.path.isfile(os.path.join(CARRIED_DIR, names[0]))


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    sys.modules[name] = mod
    return mod


def _soundfile_read(path, dtype="float64", always_2d=False):
    """Minimal `soundfile.read` built on scipy: float samples in [-1, 1) plus the rate."""
    import numpy as np
    from scipy.io import wavfile

    rate, data = wavfile.read(path)
    if data.dtype == np.int16:
        data = data.astype(np.float64) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float64) / 2147483648.0
    elif data.dtype == np.uint8:
        data = (data.astype(np.float64) - 128.0) / 128.0
    else:
        data = data.astype(np.float64)
    if always_2d and data.ndim == 1:
        data = data[:, None]
    return data.astype(dtype), rate


def load_reference():
    """Return the reference module object (cached in sys.modules)."""
    if "neural_audio_tokenizer" in sys.modules:
        return sys.modules["neural_audio_tokenizer"]
    if not reference_available():
        raise FileNotFoundError(f"reference not present under {REFERENCE_DIR}")

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    if not os.access(os.path.expanduser("~"), os.W_OK):
        os.environ["HOME"] = tempfile.mkdtemp(prefix="nat_home_")

    # transformers probes find_spec("librosa"); resolve its lazy symbols before stubbing.
    import transformers  # noqa: F401
    from transformers import AutoModel, Wav2Vec2Model, Wav2Vec2Processor  # noqa: F401

    if "librosa" not in sys.modules:
        librosa = _stub("librosa")
        librosa.display = _stub("librosa.display")

        def _no_librosa(*a, **k):
            raise RuntimeError("librosa is not installed (stub)")

        librosa.load = _no_librosa
    if "soundfile" not in sys.modules:
        sf = _stub("soundfile")
        sf.read = _soundfile_read

    sys.path.insert(0, REFERENCE_DIR)
    try:
        import neural_audio_tokenizer as nat  # type: ignore
    finally:
        sys.path.remove(REFERENCE_DIR)
    return nat
