"""NDJSON emission: the oracle restatement and the native emitter against the reference's own
output (tests/golden/ndjson_cases.json, minted by oracle/make_golden.py from StreamingProtocol.create_ndjson_stream),
and against each other on randomised streams. Host code only: runs without a GPU. Bar: byte-exact."""
import json
import os
import types

import numpy as np
import pytest
import torch

from neural_audio_tokenizer_b200 import ndjson as nd
from oracle import ndjson_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ndjson_cases.json")
with open(GOLDEN) as f:
    CASES = json.load(f)


def _tensors(streams, dtype=torch.int64):
    return [torch.tensor(s, dtype=dtype)[None] for s in streams]


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_output(name):
    c = CASES[name]
    got = ndjson_oracle.frame_lines(c["semantic"], c["acoustic"], c["sr"], c["hop"], c["rle"], c["per_layer_encoding"],
                                    c["keyframe_interval_seconds"])
    assert got == c["body"]


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.int16])
@pytest.mark.parametrize("name", sorted(CASES))
def test_native_matches_reference_output(name, dtype):
    c = CASES[name]
    text = nd.emit_frame_lines(_tensors(c["semantic"], dtype), _tensors(c["acoustic"], dtype), c["sr"], c["hop"],
                               c["rle"], c["per_layer_encoding"], c["keyframe_interval_seconds"])
    assert text == "\n".join(c["body"])


def _sticky(rng, n, vocab, stick):
    out = np.empty(n, dtype=np.int64)
    cur = 0
    for i in range(n):
        if i == 0 or rng.random() >= stick:
            cur = int(rng.integers(vocab))
        out[i] = cur
    return out


def test_native_matches_oracle_on_random_geometries():
    """Float formatting is the risk: sweep sample rates / hops (ts = round(t * frame_ms, 3), accumulated durations
    printed with repr) and layer counts, both modes."""
    rng = np.random.default_rng(2024)
    rates = [8000, 11025, 16000, 22050, 24000, 32000, 44100, 48000, 96000, 12345]
    hops = [64, 160, 256, 320, 441, 480, 512, 640, 1024, 333]
    for trial in range(60):
        sr, hop = int(rng.choice(rates)), int(rng.choice(hops))
        n_sem, n_ac = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        n = int(rng.integers(1, 700))
        rle = bool(trial % 2)
        enc = None
        if rle and trial % 4 == 3:
            enc = {f"{k}{i}": str(rng.choice(["rle", "dense"])) for k, m in (("S", n_sem), ("A", n_ac)) for i in range(m)}
        key_s = float(rng.choice([0.25, 1.0, 5.0, 1e9]))
        vocab = int(rng.choice([2, 1024, 32768]))
        sem = [_sticky(rng, n, vocab, float(rng.choice([0.0, 0.7, 0.97]))) for _ in range(n_sem)]
        ac = [_sticky(rng, n, vocab, float(rng.choice([0.0, 0.7, 0.97]))) for _ in range(n_ac)]
        want = "\n".join(ndjson_oracle.frame_lines(sem, ac, sr, hop, rle, enc, key_s))
        got = nd.emit_frame_lines([torch.from_numpy(s)[None] for s in sem], [torch.from_numpy(a)[None] for a in ac],
                                  sr, hop, rle, enc, key_s)
        assert got == want, (trial, sr, hop, rle, enc, key_s)


def test_long_stream_timestamps():
    """A ten-hour stream: timestamps reach 3.6e7 ms; every line must still equal the oracle's."""
    n = 2_700_000 // 40
    sem = [np.arange(n, dtype=np.int64) % 1024 for _ in range(4)]
    ac = [(np.arange(n, dtype=np.int64) * 7) % 1024 for _ in range(4)]
    # start far into the stream by emitting everything and comparing the tail only (the oracle is slow)
    got = nd.emit_frame_lines([torch.from_numpy(s)[None] for s in sem], [torch.from_numpy(a)[None] for a in ac],
                              24000, 320 * 40, False).split("\n")
    want = ndjson_oracle.frame_lines(sem, ac, 24000, 320 * 40, False)
    assert len(got) == n and got == want


def test_ragged_streams_use_the_shortest():
    sem = [torch.arange(10)[None], torch.arange(8)[None]]
    ac = [torch.arange(9)[None]]
    lines = nd.emit_frame_lines(sem, ac, 22050, 512).split("\n")
    assert len(lines) == 8 and json.loads(lines[-1])["S"] == [7, 7] and json.loads(lines[-1])["A"] == [7]


def test_empty_inputs():
    assert nd.emit_frame_lines([], [torch.arange(3)[None]], 22050, 512) == ""
    assert nd.emit_frame_lines([torch.zeros(1, 0, dtype=torch.int64)], [torch.zeros(1, 0, dtype=torch.int64)], 22050, 512) == ""


def test_create_ndjson_stream_drop_in_reproduces_whole_stream():
    """The drop-in keeps the reference's header / end lines (a stand-in streamer here) and its state resets."""
    c = CASES["rle_per_layer_24000"]

    class Streamer:
        sample_rate, hop_length = c["sr"], c["hop"]
        num_semantic_layers = num_acoustic_layers = 4
        per_layer_encoding = c["per_layer_encoding"]
        buffered_event, last_frame_index = {"stale": 1}, 99

        def create_header(self, duration_seconds, metadata, include_legend):
            return c["header"]

        def create_end_marker(self, stats):
            assert self.buffered_event is None
            return c["end"]

    proto = types.SimpleNamespace(ndjson_streamer=Streamer(), rle_mode=c["rle"],
                                  keyframe_interval_seconds=c["keyframe_interval_seconds"],
                                  prev_semantic_tokens=[1], prev_acoustic_tokens=[2], last_keyframe_time=3.0)
    text = nd.create_ndjson_stream(proto, {"semantic_codes": _tensors(c["semantic"]),
                                           "acoustic_codes": _tensors(c["acoustic"])})
    assert text == "\n".join([c["header"], *c["body"], c["end"]])
    with pytest.raises(ValueError):
        nd.create_ndjson_stream(proto, {"semantic_codes": _tensors(c["semantic"][:3]),
                                        "acoustic_codes": _tensors(c["acoustic"])})


def test_bad_arguments_are_reported():
    from neural_audio_tokenizer_b200 import _lib
    with pytest.raises(_lib.NatError):
        nd.emit_frame_lines([torch.arange(3)[None]], [torch.arange(3)[None]], 0, 512)
