"""Fork / join helpers for running independent pieces of the step on parallel CUDA streams
(capturable into a CUDA graph as parallel branches).  Used by trainer.py (three discriminators,
two generator calls, detached generator forward) and by miscc/utils.py (real / fake / wrong-pair
passes of one discriminator)."""
import torch

from . import nets as knets

ENABLED = True          # master switch: False = everything sequential on the current stream
HIGH_PRIORITY = -1      # CUDA: numerically lower = scheduled first
# priority of detached work.  Measured on B200: a lower priority (0) stretches the detached
# generator forward across the whole discriminator update, but the tensor-core GEMMs of both then
# interleave and the step is 0.3 ms slower than with equal priorities.
LOW_PRIORITY = -1

_POOLS = {}             # (kind, device, parent stream id) -> [streams]
_LENT = {}              # same key -> number of pool streams lent to forks that are still open
_DETACHED_OPEN = [0]    # detached branches issued and not yet joined
_FORK_DEPTH = [0]       # forks (concurrently) currently being issued: a nested join must not forget ordered-state events


def _borrow(kind, parent, n, priority=HIGH_PRIORITY):
    """n side streams of `parent` that no open fork of the same parent is using: the last piece of
    a fork runs on the parent itself and may fork again (nested), and must not be handed the
    streams its siblings are running on"""
    key = (kind, parent.device, parent.cuda_stream)
    streams = _POOLS.setdefault(key, [])
    first = _LENT.get(key, 0)
    while len(streams) < first + n:
        streams.append(torch.cuda.Stream(device=parent.device, priority=priority))
    _LENT[key] = first + n
    return key, streams[first:first + n]


def _give_back(key, n):
    _LENT[key] -= n


def step_stream(device=None):
    """a high-priority stream to run (or capture) train_step on"""
    return torch.cuda.Stream(device=device, priority=HIGH_PRIORITY)


def hold_state_order():
    """keep the ordered-state events (engine.StateOrder) alive across the next joins: a detached
    branch that touches the same module state is about to be issued"""
    _DETACHED_OPEN[0] += 1


def release_state_order():
    _DETACHED_OPEN[0] -= 1


def concurrently(*thunks, enabled=True):
    """Run the thunks on parallel streams: fork from the current stream, join back into it.  Issue
    order = list order (in-place module state shared by two pieces -- BatchNorm running
    statistics, spectral-norm vectors -- is updated in that order: engine.StateOrder); the last
    piece runs on the current stream.  Sequential on CPU or when switched off."""
    if not (enabled and ENABLED and torch.cuda.is_available() and len(thunks) > 1):
        return [t() for t in thunks]
    main = torch.cuda.current_stream()
    key, streams = _borrow("fork", main, len(thunks) - 1)
    _FORK_DEPTH[0] += 1
    try:
        results = [None] * len(thunks)
        for st in streams:
            st.wait_stream(main)
        for i, t in enumerate(thunks[:-1]):
            with torch.cuda.stream(streams[i]):
                results[i] = t()
        results[-1] = thunks[-1]()
        for st in streams:
            main.wait_stream(st)
    finally:
        _FORK_DEPTH[0] -= 1
        _give_back(key, len(streams))
    # the ordered-state events (engine.StateOrder) may only be forgotten once EVERYTHING has been joined: the join of
    # a fork nested inside one branch of an outer fork comes before the outer fork's later branches are even issued,
    # and those still have to wait for this branch's in-place updates (found by tests/streamcheck.py: the second
    # generator call of a phase did not wait for the first one's BatchNorm running-statistics updates)
    knets.sync_point(streams, reset_state_order=_DETACHED_OPEN[0] == 0 and _FORK_DEPTH[0] == 0)
    return results


class Detached:
    """Independent pieces issued on their own streams and joined LATER (not at the end of the
    issuing block): they overlap with everything the current stream does in between.  ``after``
    is an event of the current stream the pieces have to wait for (their inputs)."""

    def __init__(self, thunks, after=None):
        self.results = None
        self.streams = []
        if not (ENABLED and torch.cuda.is_available()):
            self.results = [t() for t in thunks]
            return
        main = torch.cuda.current_stream()
        self._key, self.streams = _borrow("detached", main, len(thunks), LOW_PRIORITY)
        self.results = []
        _DETACHED_OPEN[0] += 1
        for st, t in zip(self.streams, thunks):
            if after is not None:
                st.wait_event(after)
            else:
                st.wait_stream(main)
            with torch.cuda.stream(st):
                self.results.append(t())

    def join(self):
        if self.streams:
            main = torch.cuda.current_stream()
            for st in self.streams:
                main.wait_stream(st)
            _DETACHED_OPEN[0] -= 1
            _give_back(self._key, len(self.streams))
            knets.sync_point(self.streams, reset_state_order=_DETACHED_OPEN[0] == 0)
            self.streams = []
        return self.results
