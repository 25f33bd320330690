"""Happens-before checker for the step's multi-stream schedule, on the CPU.  TEST INFRASTRUCTURE.

The product issues its kernels on many CUDA streams and orders them with events / stream waits
(cpcsv_b200/streams.py, engine.StateOrder, engine.AuxBranch, the weight cache, optim.PackedAdam).  A missing wait is
a data race that only shows under the right timing on a GPU.  Here the same Python code runs on CPU tensors with

* ``torch.cuda`` stream / event API replaced by vector-clock fakes (every op ticks its stream's clock; an event
  carries the clock of its record; a wait joins clocks), ``Tensor.is_cuda`` forced True so that the product takes its
  multi-stream paths, the kernel contract emulated by tests/emulator.py;
* a ``TorchDispatchMode`` that sees EVERY aten op (the emulator's included) with its input / output tensors: reads of
  all tensor arguments, writes of mutated arguments and of outputs, each as a byte range of a storage;
* for every access, a check against the earlier conflicting accesses (overlapping range, at least one write, another
  stream): ordered iff the earlier op's clock entry is covered by the current stream's clock.

Limits: autograd runs backward nodes on the calling thread here, not on their forward streams -- ops executed inside
``Tensor.backward`` are neither logged nor checked (the forward losses are what deviated on the GPU); ranges are the
bounding boxes of strided views (interleaved views may be reported falsely: read the report).
"""
import contextlib
import sys
import weakref

import torch
from torch.utils._python_dispatch import TorchDispatchMode


class FakeStream:
    _next = [1]
    all = []

    def __init__(self, device=None, priority=0):
        self.id = FakeStream._next[0]
        FakeStream._next[0] += 1
        self.device, self.priority = device if device is not None else torch.device("cpu"), priority
        self.clock = {self.id: 0}
        FakeStream.all.append(self)

    cuda_stream = property(lambda self: self.id)

    def __eq__(self, other):
        return isinstance(other, FakeStream) and other.id == self.id

    def __hash__(self):
        return hash(self.id)

    def tick(self):
        self.clock[self.id] += 1
        return dict(self.clock)

    def _join(self, clock):
        for k, v in clock.items():
            if self.clock.get(k, 0) < v:
                self.clock[k] = v

    def wait_stream(self, other):
        self._join(other.clock)

    def wait_event(self, ev):
        if ev.clock is not None:
            self._join(ev.clock)

    def record_event(self, ev=None):
        ev = ev or FakeEvent()
        ev.record(self)
        return ev

    def synchronize(self):
        for s in FakeStream.all:
            self._join(s.clock)

    def __repr__(self):
        return "S%d" % self.id


class FakeEvent:
    def __init__(self, *a, **k):
        self.clock = None

    def record(self, stream=None):
        stream = stream or STATE["current"]
        self.clock = dict(stream.clock)

    def wait(self, stream=None):
        (stream or STATE["current"]).wait_event(self)

    def synchronize(self):
        pass

    def query(self):
        return True


STATE = {"current": None, "default": None, "in_backward": 0, "races": [], "log": {}, "keep": [], "enabled": False,
         "alloc": {}, "recorded": {}, "frees": [], "track_frees": False, "pending": {}, "custom_backward": 0}


@contextlib.contextmanager
def fake_stream_ctx(st):
    prev = STATE["current"]
    if st is not None:
        STATE["current"] = st
    try:
        yield
    finally:
        STATE["current"] = prev


def _global_sync(*a, **k):
    for s in FakeStream.all:
        for t in FakeStream.all:
            s._join(t.clock)


def _range(t):
    """(storage key, first byte, last byte + 1) of a strided tensor's bounding box"""
    if t.numel() == 0:
        return None
    st = t.untyped_storage()
    lo = t.storage_offset()
    hi = lo + sum((s - 1) * abs(d) for s, d in zip(t.shape, t.stride())) + 1
    es = t.element_size()
    return st.data_ptr(), lo * es, hi * es


def _where():
    """innermost product / emulator frames of the current call stack (cheap: no source lines are read)"""
    out, f, n = [], sys._getframe(2), 0
    while f is not None and n < 60 and len(out) < 6:
        fn = f.f_code.co_filename
        if "/cpcstoryvisualization-pytorch_b200/" in fn or fn.endswith("emulator.py"):
            out.append("%s:%d %s" % (fn.rsplit("/", 1)[-1], f.f_lineno, f.f_code.co_name))
        f, n = f.f_back, n + 1
    return " < ".join(out)


def _access(t, write, opname):
    r = _range(t)
    if r is None:
        return
    key, lo, hi = r
    cur = STATE["current"]
    ts = cur.tick()
    recs = STATE["log"].setdefault(key, [])
    for (plo, phi, pts, pstream, pwrite, pop, pwhere) in recs:
        if pstream == cur.id or not (pwrite or write) or phi <= lo or hi <= plo:
            continue
        if cur.clock.get(pstream, 0) >= pts[pstream]:
            continue        # ordered: the earlier access happens-before this one
        STATE["races"].append({"earlier": (pop, "S%d" % pstream, "W" if pwrite else "R", pwhere),
                               "later": (opname, "S%d" % cur.id, "W" if write else "R", _where()),
                               "bytes": (max(lo, plo), min(hi, phi)), "shape": tuple(t.shape)})
    recs.append((lo, hi, ts, cur.id, write, opname, _where()))
    if len(recs) > 64:          # keep the last writer(s) and recent readers
        del recs[:len(recs) - 64]
    if STATE["track_frees"]:
        # storage lifetimes: the log of a storage dies with it (the CPU allocator recycles addresses), and its death
        # is checked against the stream-ordered allocator's rule (see _freed)
        if len(recs) == 1:
            weakref.finalize(t.untyped_storage(), _freed, key)
    else:
        STATE["keep"].append(t)     # never let the CPU allocator recycle an address inside one check


def _freed(key):
    """A storage dies NOW (host order).  CUDA's caching allocator gives its block back to the pool of the stream it
    was allocated on (A): the next allocation on A may take it at once.  Safe only if every access by another stream X
    is already ordered before A's future work, i.e. covered by A's clock -- or was announced with record_stream."""
    recs = STATE["log"].pop(key, None)
    alloc = STATE["alloc"].pop(key, None)
    rec_streams = STATE["recorded"].pop(key, ())
    if not recs or alloc is None or not STATE["enabled"]:
        return
    a_stream, a_op, a_where, a_bytes = alloc
    pend = []
    for (lo, hi, ts, sid, write, op, where) in recs:
        if sid == a_stream.id or sid in rec_streams:
            continue
        if a_stream.clock.get(sid, 0) >= ts[sid]:
            continue
        pend.append((sid, ts[sid], op, where))
    if pend:
        # not a hazard yet: it becomes one when a LATER allocation on the same stream can take the block while those
        # accesses are still unordered (checked in _allocated)
        STATE["pending"].setdefault(a_stream.id, []).append(
            {"bytes": a_bytes, "alloc": (a_op, a_where), "uses": pend, "in_backward": STATE["in_backward"] > 0})


def _allocated(stream, nbytes, opname):
    """a fresh allocation on `stream`: may reuse any block freed earlier on this stream that is at least as large"""
    pend = STATE["pending"].get(stream.id)
    if not pend:
        return
    keep = []
    for p in pend:
        p["uses"] = [u for u in p["uses"] if stream.clock.get(u[0], 0) < u[1]]
        if not p["uses"]:
            continue            # the stream has caught up with every foreign access in the meantime
        if nbytes <= p["bytes"]:
            STATE["frees"].append({"alloc": ("%s (%d B)" % (p["alloc"][0], p["bytes"]), "S%d" % stream.id, p["alloc"][1]),
                                   "use": (p["uses"][0][2], "S%d" % p["uses"][0][0], p["uses"][0][3]),
                                   "reuser": (opname, nbytes, _where()), "in_backward": p["in_backward"]})
        else:
            keep.append(p)
    STATE["pending"][stream.id] = keep


class Checker(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        out = func(*args, **kwargs)
        if not STATE["enabled"] or (STATE["in_backward"] and not STATE["custom_backward"]):
            return out
        name = str(func)
        schema = getattr(func, "_schema", None)
        if name.startswith("aten.set_") or (schema is not None and any(
                r.alias_info is not None and not r.alias_info.is_write for r in schema.returns)):
            return out          # views / metadata only: no data is touched
        flat_in = []

        def walk(x, mut):
            if isinstance(x, torch.Tensor):
                flat_in.append((x, mut))
            elif isinstance(x, (list, tuple)):
                for y in x:
                    walk(y, mut)
        sargs = list(schema.arguments) if schema is not None else []
        for i, a in enumerate(args):
            mut = i < len(sargs) and sargs[i].alias_info is not None and sargs[i].alias_info.is_write
            walk(a, mut)
        for k, a in kwargs.items():
            mut = any(s.name == k and s.alias_info is not None and s.alias_info.is_write for s in sargs)
            walk(a, mut)
        seen = set()
        for t, mut in flat_in:
            _access(t, mut, name)
            seen.add(t.untyped_storage().data_ptr())
        outs = out if isinstance(out, (list, tuple)) else [out]
        for o in outs:
            if isinstance(o, torch.Tensor) and o.untyped_storage().data_ptr() not in seen and o.numel():
                k = o.untyped_storage().data_ptr()
                if STATE["track_frees"] and k not in STATE["log"]:
                    nb = o.untyped_storage().nbytes()
                    _allocated(STATE["current"], nb, name)
                    STATE["alloc"][k] = (STATE["current"], name, _where(), nb)
                _access(o, True, name)          # a fresh result: its first writer
        return out


@contextlib.contextmanager
def installed(monkeypatch, skip_backward=False, track_frees=False):
    """fake streams + forced is_cuda + dispatch logging; yields the STATE dict (``races`` after the block).
    ``skip_backward``: ``Tensor.backward`` does nothing (forward schedule only: half the run time; the optimiser
    steps then have no gradients to apply)"""
    FakeStream.all.clear()
    STATE.update(current=FakeStream(), races=[], log={}, keep=[], in_backward=0, enabled=True, alloc={}, recorded={},
                 frees=[], track_frees=track_frees, pending={}, custom_backward=0)
    STATE["default"] = STATE["current"]
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: STATE["current"])
    monkeypatch.setattr(torch.cuda, "default_stream", lambda device=None: STATE["default"])
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(torch.cuda, "synchronize", _global_sync)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    def record_stream(self, st):
        STATE["recorded"].setdefault(self.untyped_storage().data_ptr(), set()).add(st.id)
    monkeypatch.setattr(torch.Tensor, "record_stream", record_stream, raising=False)
    orig_backward = torch.Tensor.backward

    def backward(self, *a, **k):
        if skip_backward:
            return None
        STATE["in_backward"] += 1
        try:
            return orig_backward(self, *a, **k)
        finally:
            STATE["in_backward"] -= 1
    monkeypatch.setattr(torch.Tensor, "backward", backward)
    # Backward passes of the product's own autograd Function (nets.TapeFn: every kernel tape) run, as under the real
    # autograd engine, on the stream their forward pass ran on; the engine makes that stream wait for the producer of
    # the incoming gradients (approximated here by the stream that is current when the node runs: at least as much
    # ordering as the real engine gives) and, at the end of backward(), makes the calling stream wait for it.
    from cpcsv_b200 import nets as _nets
    fwd0, bwd0 = _nets.TapeFn.forward, _nets.TapeFn.backward

    def t_forward(ctx, runner, *tensors):
        ctx.fake_stream = STATE["current"]
        return fwd0(ctx, runner, *tensors)

    def t_backward(ctx, *grads):
        caller, st = STATE["current"], ctx.fake_stream
        st.wait_stream(caller)
        STATE["custom_backward"] += 1
        try:
            with fake_stream_ctx(st):
                return bwd0(ctx, *grads)
        finally:
            STATE["custom_backward"] -= 1
            caller.wait_stream(st)
    monkeypatch.setattr(_nets.TapeFn, "forward", staticmethod(t_forward))
    monkeypatch.setattr(_nets.TapeFn, "backward", staticmethod(t_backward))
    with Checker():
        try:
            yield STATE
        finally:
            STATE["enabled"] = False


def summarize(races, limit=12):
    seen, lines = set(), []
    for r in races:
        key = (r["earlier"][0], r["earlier"][3], r["later"][0], r["later"][3])
        if key in seen:
            continue
        seen.add(key)
        lines.append("%s %s %s  [%s]\n   vs later %s %s %s  [%s]  shape %s" % (
            r["earlier"][2], r["earlier"][1], r["earlier"][0], r["earlier"][3], r["later"][2], r["later"][1],
            r["later"][0], r["later"][3], r["shape"]))
        if len(lines) >= limit:
            break
    return "%d unordered conflicting accesses (%d distinct)\n" % (len(races), len(seen)) + "\n".join(lines)


def summarize_frees(frees, limit=20):
    seen, lines = {}, []
    for r in frees:
        key = (r["alloc"][0], r["alloc"][2], r["use"][0], r["use"][2], r["reuser"][0], r["reuser"][2])
        seen[key] = seen.get(key, 0) + 1
    for (aop, awhere, uop, uwhere, rop, rwhere), n in sorted(seen.items(), key=lambda kv: -kv[1])[:limit]:
        lines.append("%4d x block allocated by %s [%s]\n        still used on another stream by %s [%s]\n"
                     "        when the allocation stream allocates again: %s [%s]" % (n, aop, awhere, uop, uwhere, rop, rwhere))
    return "%d possible block reuses under a foreign reader / writer (%d distinct)\n" % (len(frees), len(seen)) + \
        "\n".join(lines)
