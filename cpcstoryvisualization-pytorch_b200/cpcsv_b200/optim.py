"""``PackedAdam``: the reference's ``optim.Adam(lr, betas=(0.5, 0.999))`` (trainer.py:212-220, stepped
at trainer.py:345-346 and 416) as hand-written multi-tensor kernels fused with the weight re-layout.

One optimiser step =
  * ``cpcsv_adam_tick``        step count += 1 and the two bias corrections (device side);
  * ``cpcsv_adam_pack_conv``   per tensor-core conv weight: Adam update + every persistent 16-bit
                               operand plane of that weight (forward fp16 / bf16 hi+lo, transposed
                               data-gradient plane, sub-pixel merged taps) in the same pass;
  * ``cpcsv_adam_pack_fc``     fc / fc_seg likewise (rows re-ordered to NHWC, transposed plane);
  * ``cpcsv_adam_multi``       everything else (BatchNorm affine, biases, GRUs, small Linears).
No pack kernel runs between an optimiser step and the next forward pass, and the planes are current
after a replayed CUDA graph as well (the step count and the learning rate are device tensors).

Arithmetic and state layout follow ``torch.optim.Adam`` (no weight decay / amsgrad / maximize):
``state[p] = {"step", "exp_avg", "exp_avg_sq"}``.  Difference: all parameters of one optimiser share
ONE step counter (``state[p]["step"]`` is the same device tensor for every p), i.e. a parameter that
had no gradient in some step still advances its bias correction -- on this path every parameter gets
a gradient in every step.
"""
import os

import torch

from . import ops

# issue the update kernels on a stream of the LOWEST priority (see PackedAdam.step); CPCSV_ADAM_LOW_PRIORITY=0: off
LOW_PRIORITY = os.environ.get("CPCSV_ADAM_LOW_PRIORITY", "1") != "0"


class PackedAdam(torch.optim.Optimizer):
    maintains_planes = True        # read by the weight cache's optimiser hook (nets.py)

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, cache=None):
        if weight_decay != 0:
            raise ValueError("PackedAdam: weight decay is not implemented (the reference uses none)")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("PackedAdam: betas %r" % (betas,))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0))
        if cache is None:
            from . import nets
            cache = nets.weight_cache()
        self.cache = cache
        self._dev = {}          # per group index: device-side step / bc / lr
        # overlap of the optimiser step with the tail of the backward pass (expect_backward)
        self._early_params, self._early_ids = [], set()
        self._armed, self._pending, self._done = False, set(), set()
        self._ticked = set()
        self._side, self._side_used, self._low = None, False, None
        self._buckets, self._bucket_of, self._expected, self._last_event = [], {}, {}, {}
        self._pre_update, self.early_fired = None, False

    # ---- device-side scalars ----------------------------------------------------------------
    def _group_state(self, gi, group, device):
        st = self._dev.get(gi)
        if st is None:
            st = self._dev[gi] = {"step": torch.zeros(1, device=device), "bc": torch.zeros(2, device=device),
                                  "lr": None, "lr_host": None}
        lr = group["lr"]
        if torch.is_tensor(lr):
            if lr.device != device or lr.dtype != torch.float32:
                raise ValueError("PackedAdam: a tensor learning rate must be an fp32 tensor on the parameters' device")
            st["lr"] = lr.reshape(1)
        elif st["lr"] is None or st["lr_host"] != lr:
            # a Python-float learning rate: mirrored into a device scalar (not capturable: inside a
            # CUDA graph use a tensor learning rate, trainer.build_capturable_optimizers)
            if st["lr"] is None or st["lr_host"] is None:
                st["lr"] = torch.zeros(1, device=device)
            st["lr"].fill_(float(lr))
            st["lr_host"] = lr
        return st

    # ---- optimiser step inside the backward pass ------------------------------------------------
    def overlap_with_backward(self, params, buckets=None):
        """Declare `params` (the generator trunk: 99.9 % of its bytes) as "early": once ``expect_backward``
        has armed the optimiser, their Adam + re-layout kernels are issued on a side stream the moment their
        gradients are complete -- while autograd still runs the rest of the backward pass -- instead of after it.
        ``step()`` then only updates the remaining parameters and joins the side stream.  Results are identical to
        a plain ``step()``.

        Two ways the gradients arrive.  (1) autograd: a post-accumulate hook per parameter; the update fires when
        the LAST early parameter has its gradient.  (2) ``buckets`` (lists of parameters, e.g. one per layer): the
        optimiser also registers as the kernel tapes' gradient sink (engine.set_grad_sink); a tape ``announce`` s the
        parameters it will produce gradients for at forward time and ``offer`` s each layer's gradients right after
        that layer's weight-gradient kernels, and a bucket's update fires as soon as every announced contribution
        to its parameters has been offered -- layer by layer, under the rest of the backward pass."""
        self._early_params = [p for p in params]
        self._early_ids = {id(p) for p in self._early_params}
        for p in self._early_params:
            p.register_post_accumulate_grad_hook(self._grad_ready)
        self._buckets, self._bucket_of = [], {}
        if buckets:
            for b in buckets:
                b = [p for p in b if id(p) in self._early_ids]
                if b:
                    for p in b:
                        self._bucket_of[id(p)] = len(self._buckets)
                    self._buckets.append(b)
            from . import engine
            engine.set_grad_sink(self)

    def expect_backward(self, pre_update=None):
        """arm the early step for the NEXT backward pass only (a backward pass that is not followed by
        ``step()`` -- gradient checks -- must not be armed).  ``pre_update(params)``: called on the side
        stream right before early parameters are updated -- the gradient exchange between ranks
        (trainer.GradSync) for exactly those parameters; ``early_fired`` tells the caller afterwards whether
        it ran (then ``updated_ids()`` are done and only the remaining parameters are left to exchange)."""
        self.early_fired = False
        self._pre_update = pre_update
        if self._early_params:
            self._armed = True
            self._pending = set(self._early_ids)

    # ---- gradient sink of the kernel tapes (engine.grad_sink) --------------------------------------
    def announce(self, params):
        """a tape whose backward pass will ``offer`` exactly one contribution (possibly None) for each of
        `params`; counted until the next ``step()``"""
        for p in params:
            if id(p) in self._bucket_of:
                self._expected[id(p)] = self._expected.get(id(p), 0) + 1

    def offer(self, p, grad):
        """one tape's gradient contribution for `p` (None: that tape has none), produced on the CURRENT stream.
        Returns True when the optimiser took it (the tape must then NOT hand it to autograd as well)."""
        pid = id(p)
        if not self._armed or pid not in self._bucket_of or self._expected.get(pid, 0) <= 0:
            return False
        cur = torch.cuda.current_stream() if p.is_cuda else None
        if grad is not None:
            if p.grad is None:
                p.grad = grad
            else:
                last = self._last_event.get(pid)
                if last is not None:
                    cur.wait_event(last)         # the earlier contribution was written on another stream
                p.grad.add_(grad)
            if cur is not None:
                ev = torch.cuda.Event()
                ev.record(cur)
                self._last_event[pid] = ev
        self._expected[pid] -= 1
        bucket = self._buckets[self._bucket_of[pid]]
        if any(self._expected.get(id(q), 0) > 0 for q in bucket):
            return True
        ready = [q for q in bucket if q.grad is not None and id(q) not in self._done]
        if not ready:
            return True
        if cur is None:                      # host-logic tests on the CPU emulator: no streams
            self._fire(ready)
            return True
        if self._side is None:
            self._side = torch.cuda.Stream(device=p.device, priority=0 if LOW_PRIORITY else cur.priority)
        self._side.wait_stream(cur)
        for q in bucket:
            ev = self._last_event.get(id(q))
            if ev is not None:
                self._side.wait_event(ev)
        self._side_used = True
        with torch.cuda.stream(self._side):
            self._fire(ready)
        return True

    def _fire(self, params):
        if self._pre_update is not None:
            self._pre_update(params)
        self.early_fired = True
        self._pending.difference_update(id(q) for q in params)
        self._update(params)

    def updated_ids(self):
        """ids of the parameters already updated in this step (by the early paths)"""
        return set(self._done)

    def early_parameters(self):
        return list(self._early_params)

    def disarm(self):
        self._armed = False

    def _grad_ready(self, p):
        if not self._armed:
            return
        if not p.is_cuda:                    # host-logic tests on the CPU emulator: no streams
            self._pending.discard(id(p))
            if not self._pending:
                self._armed = False
                self._early_step()
            return
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=p.device, priority=0 if LOW_PRIORITY else cur.priority)
        self._side.wait_stream(cur)          # this gradient is complete in stream order of `cur`
        self._side_used = True
        self._pending.discard(id(p))
        if not self._pending:
            self._armed = False
            with torch.cuda.stream(self._side):
                self._early_step()

    def _early_step(self):
        params = [q for q in self._early_params if q.grad is not None and id(q) not in self._done]
        if not params:
            return
        if self._pre_update is not None:
            self._pre_update(params)
        self.early_fired = True
        self._update(params)

    def _tick(self, gi, group, st):
        if gi not in self._ticked:
            beta1, beta2 = group["betas"]
            ops.adam_tick(st["step"], beta1, beta2, st["bc"])
            self._ticked.add(gi)

    @torch.no_grad()
    def _update(self, only=None):
        """Adam + re-layout kernels for the parameters of `only` (None: every parameter with a gradient
        that has not been updated yet in this step)"""
        only_ids = {id(p) for p in only} if only is not None else None
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None and id(p) not in self._done and
                      (only_ids is None or id(p) in only_ids)]
            if not params:
                continue
            device = params[0].device
            st = self._group_state(gi, group, device)
            beta1, beta2 = group["betas"]
            hyper = ops.AdamHyper(st["lr"], st["bc"], beta1, beta2, group["eps"])
            self._tick(gi, group, st)
            plain, touched = [], []
            for p in params:
                if p.grad.is_sparse:
                    raise RuntimeError("PackedAdam does not support sparse gradients")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = st["step"]
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                m, v = state["exp_avg"], state["exp_avg_sq"]
                ents = self.cache.maintained(p) if p.is_contiguous() else []
                conv = [(k, e) for k, e in ents if e.spec[0] == "conv"]
                fc = [(k, e) for k, e in ents if e.spec[0] == "fc"]
                if conv and p.dim() == 4 and len(conv) <= ops.MAX_PLANES:
                    planes = [(e.spec[1], e.spec[2], e.spec[3], e.spec[4], e.val[0], e.val[1]) for _, e in conv]
                    ops.adam_pack_conv(p, g, m, v, planes, hyper)
                    touched += [(e, p) for _, e in conv]
                elif fc and p.dim() == 2:
                    geo = {e.spec[2:] for _, e in fc}
                    assert len(geo) == 1, geo
                    C_, Cp, Kp = geo.pop()
                    by = {e.spec[1]: e for _, e in fc}
                    ops.adam_pack_fc(p, g, m, v, C_, p.shape[0] // C_, Cp, Kp,
                                     fwd16=by["fwd16"].val[0] if "fwd16" in by else None,
                                     fwd_hi=by["fwd"].val[0] if "fwd" in by else None,
                                     fwd_lo=by["fwd"].val[1] if "fwd" in by else None,
                                     bwd=by["bwd"].val if "bwd" in by else None, hyper=hyper)
                    touched += [(e, p) for _, e in fc]
                else:
                    plain.append((p, g, m, v))
                self._done.add(id(p))
            ops.adam_multi(plain, hyper)
            for ent, p in touched:
                self.cache.refreshed(ent, p)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._armed = False
        if self._side_used:
            # the early part runs on the side stream: same step count / bias corrections, other parameters
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_used = False
        if LOW_PRIORITY and torch.cuda.is_available() and any(p.is_cuda for g in self.param_groups for p in g["params"]):
            # the update kernels are large HBM-bound grids: on a stream of the lowest priority the block scheduler
            # hands free SM slots to the latency-bound chains of the other streams first (the conditioning-path
            # backward at the end of the generator update stalled ~0.5 ms behind them at equal priority)
            cur = torch.cuda.current_stream()
            if self._low is None:
                self._low = torch.cuda.Stream(device=cur.device, priority=0)
            self._low.wait_stream(cur)
            with torch.cuda.stream(self._low):
                self._update(None)
            cur.wait_stream(self._low)
        else:
            self._update(None)
        self._done.clear()
        self._ticked.clear()
        self._expected.clear()
        self._last_event.clear()
        return loss
