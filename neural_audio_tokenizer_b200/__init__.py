"""B200-native drop-in for the RVQ and mel/spectral hot path of defcron/neural-audio-tokenizer.

Host-side mirror of the reference's Python surface for this path:
`ResidualVectorQuantizer`, `VectorQuantizer`, `MelSpectrogram`, `spectral_stats`, plus `install()` to graft them
into a live reference tokenizer, and `create_ndjson_stream` / `emit_frame_lines` (the NDJSON emission over the index
streams, native host code). All arithmetic happens in `libnat_b200.so` (hand-written sm_100a CUDA behind the
C ABI of include/nat_b200.h); there is no CPU or PyTorch-op fallback.
"""
from .frontend import MelSpectrogram, spectral_stats
from .install import install, patch_reference_module
from .ndjson import create_ndjson_stream, emit_frame_lines
from .quantizers import ResidualVectorQuantizer, VectorQuantizer
from .sharding import all_gather_codes, shard_range
from .stacks import HostContext, encode_stacks, encode_stacks_host
from . import token_stats
from .align import align_time_bases, interpolate_linear

__all__ = ["ResidualVectorQuantizer", "VectorQuantizer", "MelSpectrogram", "spectral_stats", "install",
           "patch_reference_module", "all_gather_codes", "shard_range", "create_ndjson_stream", "emit_frame_lines", "token_stats", "align_time_bases", "interpolate_linear", "encode_stacks", "encode_stacks_host", "HostContext"]
__version__ = "0.1.0"
