"""Development aid: wall-clock phase marks (with a device synchronisation each, so only for
probing -- enabled by ``BT_PHASE_TIMING=1``; a no-op otherwise)."""
from __future__ import annotations

import os
import time

ENABLED = bool(os.environ.get("BT_PHASE_TIMING"))
_marks: list[tuple[str, float]] = []


def mark(name: str) -> None:
    if not ENABLED:
        return
    import torch
    torch.cuda.synchronize()
    _marks.append((name, time.perf_counter()))


def report(reset: bool = True) -> dict[str, float]:
    """Milliseconds between consecutive marks, summed by the name of the LATER mark."""
    out: dict[str, float] = {}
    for (_, t0), (name, t1) in zip(_marks, _marks[1:]):
        if name != "start":
            out[name] = out.get(name, 0.0) + 1e3 * (t1 - t0)
    if reset:
        _marks.clear()
    return out
