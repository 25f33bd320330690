"""trainer.GradSync (the only collective on the path: gradient averaging across ranks) on a
world-size-2 gloo group on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "cpcstoryvisualization-pytorch_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import trainer
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)),
              torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7.0) * (rank + 1)
    # params[2] has no gradient on any rank: must be skipped consistently
    sync = trainer.GradSync()
    assert sync.enabled and sync.world == world
    sync(params)
    out[rank] = [p.grad.clone() if p.grad is not None else None for p in params]
    # in-place variant (replayed CUDA-graph optimiser steps read the original gradient memory)
    q = torch.nn.Parameter(torch.zeros(4))
    q.grad = torch.full((4,), float(rank + 1))
    ptr = q.grad.data_ptr()
    sync([q], inplace=True)
    assert q.grad.data_ptr() == ptr and torch.allclose(q.grad, torch.full((4,), 1.5))
    # every rank starts from rank 0's parameters and buffers
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    net[1].running_mean.fill_(float(rank))
    trainer.broadcast_initial_state({"net": net})
    torch.manual_seed(100)
    ref = torch.nn.Linear(4, 3)
    assert torch.equal(net[0].weight.data, ref.weight.data) and float(net[1].running_mean.sum()) == 0.0
    dist.destroy_process_group()


def test_grad_sync_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        g0, g1, g2 = out[rank]
        assert torch.allclose(g0, torch.full((5, 3), 1.5))
        assert torch.allclose(g1, torch.arange(7.0) * 1.5)
        assert g2 is None


class _Patch:
    """minimal stand-in for pytest's monkeypatch inside a spawned worker"""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _dp_worker(rank, world, port, out, fused=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p_ in (root, os.path.join(root, "cpcstoryvisualization-pytorch_b200"), os.path.join(root, "tests")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emulator
    import harness
    import trainer
    from oracle import params, presets, synth
    emulator.install(_Patch)
    p = presets.get("tiny")
    dev = torch.device("cpu")
    N, B = p["IM_BATCH"], p["ST_BATCH"]
    labels = (torch.ones(N), torch.zeros(N), torch.ones(B), torch.zeros(B))
    x = harness.product_inputs(synth.make_batch(p, 10 + rank))          # this rank's shard of the job
    noise = synth.make_noise(p, 20 + rank)
    grads = {}
    for mode, sync in (("local", None), ("synced", trainer.GradSync())):
        nets = harness.build_product(p, params.init_all(p, 0), dev)     # identical weights on every rank
        harness.inject_noise(nets["G"], synth.NoiseFeed(noise))
        # fused: cpcsv_b200.optim.PackedAdam on the emulator -- the generator's trunk gradients are then
        # exchanged and applied from inside the backward pass, the rest after it
        opts = trainer.build_optimizers(nets, fused=fused)
        trainer.train_step(nets, opts, x, labels, 1.0, grad_sync=sync, apply_optim=(mode == "synced"))
        if fused and mode == "synced":
            assert opts["G"].early_fired
        grads[mode] = torch.cat([q.grad.flatten() for k in ("D_im", "G") for q in nets[k].parameters()
                                 if q.grad is not None])
        if mode == "synced":
            weights = torch.cat([q.detach().flatten() for n in nets.values() for q in n.parameters()])
    # D gradients do not depend on the optimiser coupling: synced == mean over ranks of the local ones
    n_d = sum(q.numel() for q in nets["D_im"].parameters())
    local = [torch.zeros(n_d) for _ in range(world)]
    dist.all_gather(local, grads["local"][:n_d].contiguous())
    out[rank] = (float((grads["synced"][:n_d] - sum(local) / world).abs().max()),
                 float(grads["synced"][:n_d].abs().max()), weights)
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("fused", [False, True])
def test_data_parallel_step_world2(fused):
    """SURVEY.md section 8e on two gloo ranks (kernel-contract emulator): each rank runs train_step on
    its own shard; after the one exchange per optimiser step the discriminator gradients are the
    mean of the per-rank gradients and every rank ends the step with identical weights.  ``fused``: with
    PackedAdam (per-discriminator exchange inside the discriminator stage, trunk exchange + update inside
    the generator's backward pass)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, _free_port(), out, fused), nprocs=world, join=True)
    for rank in range(world):
        err, scale, _ = out[rank]
        assert err <= 1e-6 * max(scale, 1.0), (rank, err, scale)
    assert torch.equal(out[0][2], out[1][2])
