"""Import the UNMODIFIED reference (`/root/reference/neural_audio_tokenizer.py`) in the authoring container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (to mint golden vectors) and by the
container-only cross-check tests (skipped when `/root/reference` is absent, i.e. on the GPU box).
Nothing in the product package, in `bench.py`, in `smoke()` or in the `-m gpu` tests imports this.

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

REFERENCE_DIR = os.environ.get("NAT_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "neural_audio_tokenizer.py"))


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
