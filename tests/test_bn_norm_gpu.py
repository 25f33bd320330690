"""BatchNorm forward / backward as the product issues it -- statistics accumulated with fp64 atomics
(cpcsv_bn_stats), constants derived inside the apply kernel (cpcsv_bn_norm_act_pack), backward sums
accumulated likewise -- against the contract emulated on the CPU (reference semantics: model.py:31-33,
SURVEY.md Appendix A).  GPU only."""
import pytest
import torch

import emulator as emu
from cpcsv_b200 import ops

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def close(a, b, tol):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)) < tol


@pytest.mark.parametrize("rows,C", [(1440, 1024), (90, 4096), (23040, 256), (7, 64), (368640, 128)])
@pytest.mark.parametrize("act,use_mod", [(1, False), (2, False), (1, True)])
def test_bn_norm_forward_backward(rows, C, act, use_mod):
    if rows * C > 3e7 and use_mod:
        pytest.skip("one large case is enough")
    x = rnd(rows, C, seed=1) * 1.7 + 0.3
    gamma, beta = 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    mod = rnd(rows, C, seed=4, scale=0.3) if use_mod else None
    dy = rnd(rows, C, seed=5)
    res = {}
    for dev, m in (("cuda", ops), ("cpu", emu)):
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        vec = torch.empty(4, C, device=dev)
        y = torch.empty(rows, C, device=dev)
        hi = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
        lo = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
        ws = m.bn_workspace(rows, C, dev)
        xd, modd = x.to(dev), (mod.to(dev) if use_mod else None)
        m.bn_stats(xd, ws)
        m.bn_norm_act_pack(xd, ws, gamma.to(dev), beta.to(dev), rm, rv, None, C - 4, vec, act, modd, y, hi, lo, 1)
        dx16 = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
        dmod16 = torch.empty(rows, C, device=dev, dtype=torch.bfloat16) if use_mod else None
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        ws2 = m.bn_workspace(rows, C, dev)
        m.bn_bwd_reduce(xd, dy.to(dev), vec[2], vec[3], vec[0], vec[1], act, modd, ws2)
        m.bn_bwd_apply(xd, dy.to(dev), vec[2], vec[3], vec[0], vec[1], None, C - 4, act, modd, ws2, True,
                       dx16=dx16, dmod16=dmod16, dgamma=dg, dbeta=db)
        if dev == "cuda":
            torch.cuda.synchronize()
        res[dev] = dict(stats=ws.cpu(), vec=vec.cpu(), rm=rm.cpu(), rv=rv.cpu(), y=y.cpu(),
                        hl=(hi.float() + lo.float()).cpu(), dx=dx16.float().cpu(),
                        dmod=dmod16.float().cpu() if use_mod else None, dg=dg.cpu(), db=db.cpu())
    g, c = res["cuda"], res["cpu"]
    assert close(g["stats"], c["stats"], 1e-9)
    assert torch.allclose(g["vec"], c["vec"], rtol=2e-5, atol=2e-6)
    assert torch.allclose(g["rm"], c["rm"], rtol=1e-5, atol=1e-6) and torch.allclose(g["rv"], c["rv"], rtol=1e-5, atol=1e-6)
    assert close(g["y"], c["y"], 1e-5) and close(g["hl"], c["y"], 3e-5)
    assert close(g["dx"], c["dx"], 4e-3)
    if use_mod:
        assert close(g["dmod"], c["dmod"], 4e-3)
    assert close(g["dg"], c["dg"], 1e-4) and close(g["db"], c["db"], 1e-4)
    # padding channels (>= C_valid) must come out exactly zero
    assert float(g["y"][:, C - 4:].abs().max()) == 0.0
