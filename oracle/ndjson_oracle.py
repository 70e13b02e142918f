"""CPU restatement of the reference's NDJSON frame / token event emission.

TEST INFRASTRUCTURE ONLY (tests/ and nothing else may import this).

Follows `StreamingProtocol.create_ndjson_stream` (nat.py:4452-4520: frame loop, change detection 4413-4440,
keyframe clock 4442-4450) and `NDJSONStreamer.create_frame` / `_flush_buffered_event` / `create_end_marker`
(nat.py:2713-2853), restricted to the lines between the header event and the end event. Pinned against the
reference's own output by tests/golden/ndjson_*.json (minted by oracle/make_golden.py) and, in the authoring
container, against the live reference (tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence


def _dumps(event: dict) -> str:
    return json.dumps(event, separators=(",", ":"))          # nat.py:2718, 2767


def frame_lines(sem: Sequence[Sequence[int]], ac: Sequence[Sequence[int]], sample_rate: int, hop_length: int,
                rle_mode: bool = False, per_layer_encoding: Optional[Dict[str, str]] = None,
                keyframe_interval_seconds: float = 5.0) -> List[str]:
    """sem[l][t], ac[l][t] are the index streams. Returns the emitted lines in order (final flush included)."""
    enc = per_layer_encoding or {}
    n_frames = min(min(len(s) for s in sem), min(len(a) for a in ac)) if sem and ac else 0
    fps = sample_rate / hop_length                            # nat.py:2626-2627
    frame_ms = 1000.0 / fps

    def layer_is_rle(name: str) -> bool:                      # nat.py:2707-2711
        return enc.get(name, "rle" if (rle_mode and name.startswith("S")) else "dense") == "rle"

    out: List[str] = []
    pending: Optional[dict] = None                            # the buffered RLE event
    pending_at = -1
    prev = None
    last_key_s = 0.0
    for t in range(n_frames):
        s_now = [int(s[t]) for s in sem]
        a_now = [int(a[t]) for a in ac]
        ms = t * frame_ms
        keyframe = False
        if rle_mode and ms / 1000.0 - last_key_s >= keyframe_interval_seconds:      # nat.py:4447-4449
            last_key_s = ms / 1000.0
            keyframe = True
        if rle_mode:                                          # change detection runs for every frame in RLE mode
            if prev is None:
                changed = [f"S{i}" for i in range(len(s_now))] + [f"A{i}" for i in range(len(a_now))]
            else:
                changed = [f"S{i}" for i, (c, p) in enumerate(zip(s_now, prev[0])) if c != p] + \
                          [f"A{i}" for i, (c, p) in enumerate(zip(a_now, prev[1])) if c != p]
            prev = (s_now, a_now)
        if keyframe or not rle_mode:                          # nat.py:2746-2769: dense frame, buffered event first
            if pending is not None:
                out.append(_dumps(pending))
                pending = None
            ev = {"event": "frame", "fi": t, "ts": round(ms, 3), "dur": round(frame_ms, 3), "S": s_now, "A": a_now}
            if keyframe:
                ev["is_keyframe"] = True
            out.append(_dumps(ev))
        elif changed:                                         # nat.py:2772-2822
            if pending is not None:
                pending["dur"] += (t - pending_at) * frame_ms
                out.append(_dumps(pending))
            ev = {"event": "tokens", "fi": t, "ts": round(ms, 3), "dur": round(frame_ms, 3)}
            for name in changed:
                if layer_is_rle(name):
                    ev[name] = s_now[int(name[1:])] if name[0] == "S" else a_now[int(name[1:])]
            s_dense = [v for i, v in enumerate(s_now) if not layer_is_rle(f"S{i}")]
            a_dense = [v for i, v in enumerate(a_now) if not layer_is_rle(f"A{i}")]
            if s_dense:
                ev["S_dense"] = s_dense
            if a_dense:
                ev["A_dense"] = a_dense
            pending, pending_at = ev, t
        elif pending is not None:                             # nat.py:2823-2829
            pending["dur"] += (t - pending_at) * frame_ms
            pending_at = t
    if pending is not None:                                   # create_end_marker's flush, nat.py:2843-2845
        out.append(_dumps(pending))
    return out
