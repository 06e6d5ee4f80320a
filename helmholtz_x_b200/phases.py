"""Per-phase wall-clock breakdown of one pass of the hot path (pattern / colouring / assembly /
flame vectors / multigrid set-up / per-shift set-up / Woodbury base solves / inner solves / ...).

Off by default: `phase(name)` is then a no-op.  bench.py switches it on for ONE extra, untimed step
(the device is synchronised at every phase boundary, which the timed steps must not pay).  Time is
attributed to the innermost active phase, so the numbers add up to the step."""
from __future__ import annotations

import time
from contextlib import contextmanager

_enabled = False
_stack = []
times = {}
counts = {}


def enable(on=True):
    global _enabled
    _enabled = on
    reset()


def reset():
    times.clear()
    counts.clear()
    del _stack[:]


def _sync():
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()


@contextmanager
def phase(name):
    if not _enabled:
        yield
        return
    _sync()
    now = time.perf_counter()
    if _stack:
        outer = _stack[-1]
        times[outer[0]] = times.get(outer[0], 0.0) + now - outer[1]
    _stack.append([name, now])
    try:
        yield
    finally:
        _sync()
        now = time.perf_counter()
        me = _stack.pop()
        times[name] = times.get(name, 0.0) + now - me[1]
        counts[name] = counts.get(name, 0) + 1
        if _stack:
            _stack[-1][1] = now


def report():
    return {k: round(v, 4) for k, v in sorted(times.items(), key=lambda kv: -kv[1])}
