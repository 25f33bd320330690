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
