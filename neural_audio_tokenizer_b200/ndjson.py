"""NDJSON emission over the index streams: the host mirror of
`StreamingProtocol.create_ndjson_stream` (nat.py:4452-4520).

The header and end events stay with the reference's own `NDJSONStreamer.create_header` / `create_end_marker`
(json.dumps of small dicts, once per stream). Everything in between, which the reference produces with a Python loop
that reads the index streams one scalar at a time (nat.py:4482-4513, about 20 k frames/s), comes from
`nat_ndjson_emit_frames` in libnat_b200.so, byte for byte: dense `frame` events, RLE `tokens` events with the
reference's duration bookkeeping, keyframes, per-layer encodings, and the final flush.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import torch

from . import _lib

_CODE_DTYPES = {torch.int64: _lib.CODES_I64, torch.int32: _lib.CODES_I32, torch.int16: _lib.CODES_I16}


def _stack_host(codes: Sequence[torch.Tensor], num_frames: int) -> torch.Tensor:
    """[L, num_frames] contiguous host tensor from the reference's list of [B, T] (or [T]) code tensors; row 0 of the
    batch is what the reference emits (`codes[0, frame_idx]`, nat.py:4484-4487)."""
    rows = []
    for c in codes:
        c = c.detach()
        row = c[0] if c.dim() >= 2 else c
        rows.append(row.reshape(-1)[:num_frames])
    dt = rows[0].dtype
    if dt not in _CODE_DTYPES or any(r.dtype != dt for r in rows):
        rows = [r.to(torch.int64) for r in rows]
    return torch.stack(rows).cpu().contiguous()


def emit_frame_lines(semantic_codes: Sequence[torch.Tensor], acoustic_codes: Sequence[torch.Tensor], sample_rate: int,
                     hop_length: int, rle_mode: bool = False, per_layer_encoding: Optional[Dict[str, str]] = None,
                     keyframe_interval_seconds: float = 5.0) -> str:
    """Every line between the header and the end event, joined by newlines ('' when there is none)."""
    if not semantic_codes or not acoustic_codes:
        return ""                                            # nat.py:4465: no frame loop without both stacks
    num_frames = min(min(int(c.shape[-1]) for c in semantic_codes), min(int(c.shape[-1]) for c in acoustic_codes))
    n_sem, n_ac = len(semantic_codes), len(acoustic_codes)
    sem = _stack_host(semantic_codes, num_frames)
    ac = _stack_host(acoustic_codes, num_frames)
    if sem.dtype != ac.dtype:
        sem, ac = sem.to(torch.int64), ac.to(torch.int64)
    enc = per_layer_encoding or {}

    def is_rle(name: str) -> int:                            # NDJSONStreamer._should_use_rle_for_layer, nat.py:2707-2711
        return int(enc.get(name, "rle" if (rle_mode and name.startswith("S")) else "dense") == "rle")

    flags = bytes([is_rle(f"S{i}") for i in range(n_sem)] + [is_rle(f"A{i}") for i in range(n_ac)])
    lib = _lib.load()
    text, length = ctypes.c_void_p(), ctypes.c_size_t()
    _lib.check(lib.nat_ndjson_emit_frames(sem.data_ptr(), ac.data_ptr(), _CODE_DTYPES[sem.dtype], n_sem, n_ac,
                                          num_frames, num_frames, int(sample_rate), int(hop_length), int(bool(rle_mode)),
                                          flags, float(keyframe_interval_seconds), ctypes.byref(text),
                                          ctypes.byref(length)))
    try:
        return ctypes.string_at(text.value, length.value).decode("ascii")
    finally:
        lib.nat_free_host(text)


def create_ndjson_stream(protocol, tokens: Dict, metadata: Dict = None, processing_stats: Dict = None,
                         duration_seconds: float = None, include_legend: bool = True) -> str:
    """Drop-in for `StreamingProtocol.create_ndjson_stream` (same arguments, same text). `protocol` is the reference's
    StreamingProtocol instance: its NDJSONStreamer supplies the header and end lines and its configuration."""
    st = protocol.ndjson_streamer
    semantic_codes, acoustic_codes = tokens["semantic_codes"], tokens["acoustic_codes"]
    if semantic_codes and acoustic_codes and (len(semantic_codes) != st.num_semantic_layers or
                                              len(acoustic_codes) != st.num_acoustic_layers):
        # The reference pads / truncates every frame with a printed warning (nat.py:2731-2744) while its change
        # detection still sees the unpadded lists: a malformed-input path, served by the reference's own method so
        # that text and warnings stay what they were (no native form).
        original = getattr(type(protocol), "_nat_b200_reference_create_ndjson_stream", None)
        if original is None and getattr(type(protocol), "create_ndjson_stream", create_ndjson_stream) is not create_ndjson_stream:
            original = type(protocol).create_ndjson_stream
        if original is None:
            raise ValueError(f"layer count mismatch: streams {len(semantic_codes)}+{len(acoustic_codes)}, protocol "
                             f"{st.num_semantic_layers}+{st.num_acoustic_layers}, and the reference's own "
                             "create_ndjson_stream is not reachable from this protocol object")
        return original(protocol, tokens, metadata, processing_stats, duration_seconds, include_legend)
    lines = [st.create_header(duration_seconds, metadata, include_legend)]
    if semantic_codes and acoustic_codes:
        protocol.prev_semantic_tokens = None                 # same resets as nat.py:4476-4480
        protocol.prev_acoustic_tokens = None
        protocol.last_keyframe_time = 0.0
        st.buffered_event = None
        st.last_frame_index = -1
        body = emit_frame_lines(semantic_codes, acoustic_codes, st.sample_rate, st.hop_length, protocol.rle_mode,
                                st.per_layer_encoding, protocol.keyframe_interval_seconds)
        if body:
            lines.append(body)
    lines.append(st.create_end_marker(processing_stats))     # the native body already carries the final flush
    return "\n".join(lines)


def install(nat_module) -> None:
    """Rebind `StreamingProtocol.create_ndjson_stream` in the imported reference module."""
    cls = nat_module.StreamingProtocol
    if cls.create_ndjson_stream is not create_ndjson_stream:
        cls._nat_b200_reference_create_ndjson_stream = cls.create_ndjson_stream      # malformed-input path, see above
    cls.create_ndjson_stream = create_ndjson_stream
