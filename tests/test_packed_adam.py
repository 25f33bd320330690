"""optim.PackedAdam (hand-written multi-tensor Adam fused with the weight re-layout) against
``torch.optim.Adam`` -- the optimiser the reference builds at trainer.py:212-220 -- and its persistent
operand planes against a fresh re-layout of the updated weights.

Every case runs with the CPU emulator of the kernel contract (host logic: grouping of parameters,
plane bookkeeping in the weight cache) and, marked ``gpu``, through libcpcsv.so.
"""
import pytest
import torch

import emulator
from cpcsv_b200 import engine, nets, ops
from cpcsv_b200.optim import PackedAdam


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "emu":
        emulator.install(monkeypatch)
        return torch.device("cpu")
    return torch.device("cuda")


def _params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = {"up": (20, 72, 3, 3), "s1": (24, 40, 3, 3), "s2": (12, 8, 4, 4), "fc": (16 * 10, 37),
              "bn_w": (33,), "gru": (45, 17), "bias": (1,), "big": (9000,)}
    return {k: torch.nn.Parameter((torch.randn(*s, generator=g) * 0.05).to(dev)) for k, s in shapes.items()}


class _FakeG:
    def __init__(self):
        self._cpcsv_maps = {}


def _register_planes(cache, P):
    """what the first forward / backward passes of a step would ask the cache for"""
    rup = engine.rup
    got = {}
    for name, geom in (("up", "up"), ("s1", "s1"), ("s2", "s2")):
        w = P[name]
        Co_pad, Ci_pad = rup(w.shape[0], 64), rup(w.shape[1], 64)
        _k, fkind, bkind, _nt, _uk = engine.CONV_GEOM[geom]
        got[name] = [engine.pack_conv(cache, w, geom, fkind, Co_pad, Ci_pad, 2, ops.BF16),
                     engine.pack_conv(cache, w, geom, bkind, Ci_pad, Co_pad, 1, ops.BF16)]
        if geom != "s2":
            got[name].append(engine.pack_conv(cache, w, geom, fkind, Co_pad, Ci_pad, 1, ops.FP16))

    class Lin:
        weight = P["fc"]
    G = _FakeG()
    Kp = rup(P["fc"].shape[1], 64)
    got["fc"] = [nets.pack_fc_fwd(cache, G, Lin, 10, Kp, 1, ops.FP16), nets.pack_fc_fwd(cache, G, Lin, 10, Kp, 2, ops.BF16),
                 nets.pack_fc_bwd(cache, G, Lin, 10, Kp)]
    return got


def _flat(planes):
    out = []
    for v in planes:
        for t in (v if isinstance(v, (list, tuple)) else [v]):
            if t is not None:
                out.append(t.detach().float().cpu().clone())
    return out


def test_packed_adam_matches_torch_adam_and_keeps_planes_current(dev):
    P = _params(dev)
    Q = {k: torch.nn.Parameter(v.detach().clone()) for k, v in P.items()}
    cache = engine.WeightCache()
    planes = _register_planes(cache, P)
    lr = torch.tensor(3e-3, device=dev)
    opt = PackedAdam(list(P.values()), lr=lr, betas=(0.5, 0.999), cache=cache)
    ref = torch.optim.Adam(list(Q.values()), lr=3e-3, betas=(0.5, 0.999))
    gen = torch.Generator().manual_seed(11)
    for step in range(3):
        for k in P:
            g = torch.randn(P[k].shape, generator=gen) * (0.3 if k != "bias" else 1e-3)
            P[k].grad = g.to(dev)
            Q[k].grad = g.to(dev).clone()
        if step == 2:
            lr.fill_(1e-3)                      # a schedule acting through the device tensor
            ref.param_groups[0]["lr"] = 1e-3
        opt.step()
        ref.step()
        for k in P:
            a, b = P[k].detach().double().cpu(), Q[k].detach().double().cpu()
            assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(b.abs().max())), (k, step)
        # state layout of torch.optim.Adam
        assert set(opt.state[P["up"]]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(opt.state[P["up"]]["step"]) == step + 1
        for k in ("up", "fc", "gru"):
            a, b = opt.state[P[k]]["exp_avg_sq"].double().cpu(), ref.state[Q[k]]["exp_avg_sq"].double().cpu()
            assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
    # the persistent planes hold the UPDATED weights: identical to a fresh re-layout of them
    kept = {k: _flat(v) for k, v in planes.items()}
    again = _register_planes(cache, P)
    for k in planes:      # cache hits: the very same buffers, nothing re-packed
        for a, b in zip(planes[k], again[k]):
            ta = a[0] if isinstance(a, (list, tuple)) else a
            tb = b[0] if isinstance(b, (list, tuple)) else b
            assert ta.data_ptr() == tb.data_ptr()
    fresh = _register_planes(engine.WeightCache(), P)
    for k in planes:
        for a, b in zip(kept[k], _flat(fresh[k])):
            assert torch.equal(a, b), k


def test_foreign_optimizer_and_invalidate_repack_in_place(dev):
    """another optimiser (the global post-step hook) or invalidate_weight_cache() marks the persistent
    planes stale; the next use re-packs them into the same buffers"""
    P = _params(dev, seed=3)
    cache = nets.weight_cache()      # the hook acts on the global cache
    planes = _register_planes(cache, P)
    before = _flat(planes["s1"])
    opt = torch.optim.Adam(list(P.values()), lr=1e-2)
    for v in P.values():
        v.grad = torch.ones_like(v)
    opt.step()
    again = _register_planes(cache, P)
    assert again["s1"][0][0].data_ptr() == planes["s1"][0][0].data_ptr()
    after = _flat(again["s1"])
    assert not torch.equal(before[0], after[0])
    fresh = _flat(_register_planes(engine.WeightCache(), P)["s1"])
    for a, b in zip(after, fresh):
        assert torch.equal(a, b)
    with torch.no_grad():
        P["s1"].data.add_(0.25)         # bypasses Tensor._version
    nets.invalidate_weight_cache()
    third = _flat(_register_planes(cache, P)["s1"])
    fresh = _flat(_register_planes(engine.WeightCache(), P)["s1"])
    for a, b in zip(third, fresh):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_adam_pack_kernels_match_the_pack_kernels_gpu():
    """the fused kernels' planes == cpcsv_pack_conv_weight / cpcsv_pack_matrix of the same weights"""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    for (Co, Ci, k), kinds in (((72, 136, 3), (0, 1, 2, 3)), ((124, 3, 4), (0, 1)), ((248, 124, 4), (0, 1))):
        w = (torch.randn(Co, Ci, k, k, generator=g) * 0.05).to(dev)
        for kind in kinds:
            tr = kind in (1, 3)
            rp, cp = (engine.rup(Ci, 64), engine.rup(Co, 64)) if tr else (engine.rup(Co, 64), engine.rup(Ci, 64))
            ntap = 16 if kind >= 2 else k * k
            for dtype, two in ((ops.BF16, True), (ops.FP16, False)):
                t16 = ops.TORCH16[dtype]
                a_hi = torch.full((ntap * rp, cp), 7.0, device=dev, dtype=t16)
                a_lo = torch.full((ntap * rp, cp), 7.0, device=dev, dtype=t16) if two else None
                b_hi, b_lo = torch.empty_like(a_hi), (torch.empty_like(a_hi) if two else None)
                ops.adam_pack_conv(w, None, None, None, [(kind, dtype, rp, cp, a_hi, a_lo)])
                ops.pack_conv_weight(w, kind, rp, cp, b_hi, b_lo, dtype)
                assert torch.equal(a_hi, b_hi), (Co, Ci, k, kind, dtype)
                if two:
                    assert torch.equal(a_lo, b_lo)


def test_early_step_inside_backward_gives_the_same_update(dev):
    """overlap_with_backward / expect_backward: part of the parameters is updated from a gradient hook in
    the middle of backward(), the rest in step(); together exactly one Adam step for every parameter"""
    P = _params(dev, seed=7)
    Q = {k: torch.nn.Parameter(v.detach().clone()) for k, v in P.items()}
    cache = engine.WeightCache()
    _register_planes(cache, P)
    opt = PackedAdam(list(P.values()), lr=torch.tensor(2e-3, device=dev), betas=(0.5, 0.999), cache=cache)
    ref = torch.optim.Adam(list(Q.values()), lr=2e-3, betas=(0.5, 0.999))
    early = ["up", "s1", "fc", "bn_w"]
    opt.overlap_with_backward([P[k] for k in early])

    def loss(W):
        return sum(((w * (i + 1)) ** 2).sum() + w.sum() for i, w in enumerate(W.values()))
    for step in range(2):
        for W, o in ((P, opt), (Q, ref)):
            o.zero_grad(set_to_none=True)
        before = {k: P[k].detach().clone() for k in P}
        opt.expect_backward()
        loss(P).backward()
        # the early parameters have already moved, the others not yet
        if dev.type == "cuda":
            torch.cuda.synchronize()
        assert all(not torch.equal(before[k], P[k].detach()) for k in early)
        assert all(torch.equal(before[k], P[k].detach()) for k in P if k not in early)
        opt.step()
        loss(Q).backward()
        ref.step()
        for k in P:
            a, b = P[k].detach().double().cpu(), Q[k].detach().double().cpu()
            assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max())), (k, step)
    # not armed: backward alone must not touch the weights
    before = {k: P[k].detach().clone() for k in P}
    opt.zero_grad(set_to_none=True)
    loss(P).backward()
    assert all(torch.equal(before[k], P[k].detach()) for k in P)


def test_generator_update_inside_backward_matches_plain_adam(dev, monkeypatch):
    """one whole train step with the generator trunk's Adam fired from inside the backward pass -- (a) layer by
    layer through the tapes' gradient sink, (b) once, from autograd hooks -- against (c) plain torch.optim.Adam
    after the backward pass: same accumulated gradients (every tape's contribution counted exactly once) and the
    same weights.  The first Adam step is sign(g)-like, so weights are compared in the mean, in units of lr."""
    import harness
    import trainer
    from oracle import presets
    # discriminator learning rate 0: their (sign-like) first Adam step would turn the GPU's summation-order noise
    # into 2 * lr weight differences and those into percent-level differences of the generator gradients
    p = presets.get("tiny", DISCRIMINATOR_LR=0.0)
    lr = p["GENERATOR_LR"]
    # GPU: fp32 summation order varies run to run; one flipped bf16 rounding is 4e-3 of an element (measured 5e-4
    # on the GRU weights' gradient); a lost or doubled contribution would be O(1)
    tol = 1e-5 if dev.type == "cpu" else 1e-2
    monkeypatch.setattr(trainer, "LAYERWISE_G_ADAM", True)
    nets_a, _o, grads_a = harness.run_product_step(p, dev, fused=True)
    sink = engine.grad_sink()
    assert sink is not None and len(sink._buckets) >= 10
    monkeypatch.setattr(trainer, "LAYERWISE_G_ADAM", False)
    engine.set_grad_sink(None)
    nets_b, _o, grads_b = harness.run_product_step(p, dev, fused=True)
    nets_c, _o, grads_c = harness.run_product_step(p, dev, fused=False)
    trunk = set(nets.TrunkRunner.parameter_names())
    cat = {k: [] for k in "abc"}
    for n in grads_a["G"]:
        if n in harness.ZERO_GRAD or n not in trunk:
            continue        # only the trunk's gradients travel through the sink
        ga, gb, gc = (g["G"][n].double().cpu().flatten() for g in (grads_a, grads_b, grads_c))
        for k, g in zip("abc", (ga, gb, gc)):
            cat[k].append(g)
        if dev.type == "cpu":       # deterministic: every tensor, tightly (a lost or doubled contribution is O(1))
            assert float((ga - gb).norm()) <= tol * float(gb.norm()), n
            assert float((ga - gc).norm()) <= tol * float(gc.norm()), n
    # GPU: the 'tiny' preset has 2-element BatchNorm tensors whose gradient moves by percents with the summation
    # order of a run; the trunk as a whole does not
    ga, gb, gc = (torch.cat(cat[k]) for k in "abc")
    assert float((ga - gb).norm()) <= tol * float(gb.norm())
    assert float((ga - gc).norm()) <= tol * float(gc.norm())
    for (n, a), (_n, b), (_m, c) in zip(nets_a["G"].named_parameters(), nets_b["G"].named_parameters(),
                                         nets_c["G"].named_parameters()):
        if n in harness.ZERO_GRAD or (dev.type != "cpu" and a.numel() < 256):
            continue        # (GPU: one flipped sign in a 2-element tensor is a mean difference of lr)
        a, b, c = (t.detach().double().cpu() for t in (a, b, c))
        assert float((a - b).abs().mean()) <= (0.02 if dev.type == "cpu" else 0.10) * lr, n
        assert float((a - c).abs().mean()) <= 0.10 * lr, (n, n in trunk)


def test_plain_cache_entries_are_rebuilt_after_a_graph_replay():
    """weights updated inside a replayed CUDA graph leave no Python-side trace; the owner of the
    graph calls WeightCache.note_replay() (GraphedStep.step does), after which eagerly packed plain entries are
    rebuilt at their next use while maintained planes (rewritten by the graph's own optimiser kernels) stay"""
    cache = engine.WeightCache()
    w = torch.nn.Parameter(torch.randn(8, 4))
    built = {"plain": 0, "kept": 0}

    def build_plain():
        built["plain"] += 1
        return [w.detach().clone(), None]

    def build_kept():
        built["kept"] += 1
        return [w.detach().clone(), None]
    cache.get((id(w), "plain"), w, build_plain)
    cache.get((id(w), "kept"), w, build_kept, spec=("conv", 0), fill=lambda val: None)
    cache.get((id(w), "plain"), w, build_plain)
    assert built == {"plain": 1, "kept": 1}
    cache.note_replay()
    cache.get((id(w), "plain"), w, build_plain)
    cache.get((id(w), "kept"), w, build_kept, spec=("conv", 0), fill=lambda val: None)
    assert built == {"plain": 2, "kept": 1}
