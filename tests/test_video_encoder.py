"""the order-consistency critic ``VideoEncoder`` (reference model.py:99-210, used
by STAGE1_D_STY_V2 when cfg.USE_SEQ_CONSISTENCY, losses miscc/utils.py:110-122,155-169).

* the oracle restatement (oracle/video_encoder.py) against outputs of the REAL reference class
  (tests/golden/video_encoder.pt, written by oracle/make_golden_video.py);
* the product's VideoEncoder (model.py, kernels of libcpcsv.so: im2col'd 7x7 stem, pointwise / 3x3-stride-2 /
  temporal 3-tap stride-2 convolutions on the tcgen05 GEMM, BatchNorm3d through the same BatchNorm kernels)
  against the fp64 oracle: on the CPU emulator of the kernel contract and, ``-m gpu``, on the kernels.
"""
import os

import pytest
import torch

import emulator
from oracle import video_encoder as VE

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _inputs(gold):
    g = torch.Generator().manual_seed(gold["seed"] + 1)
    story = torch.rand(gold["B"], 3, gold["T"], 64, 64, generator=g) * 2 - 1
    labels = (torch.rand(gold["B"], generator=g) < 0.5).float()
    assert torch.equal(labels, gold["labels"])
    return story, labels


def _oracle(sd, story, labels, dtype):
    sd = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    leaves = {k: v.requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and k.rsplit(".", 1)[-1] in ("weight_orig", "weight", "bias")}
    x = story.to(dtype).requires_grad_(True)
    loss, logits = VE.order_loss_d(sd, x, labels.to(dtype))
    loss.backward()
    return loss.detach(), logits.detach(), {k: v.grad for k, v in leaves.items()}, x.grad, sd


def test_oracle_matches_the_real_reference():
    gold = torch.load(os.path.join(GOLD, "video_encoder.pt"))
    story, labels = _inputs(gold)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    loss, logits, grads, dstory, sd = _oracle(VE.init_state(gold["seed"]), story, labels, torch.float32)
    assert abs(float(loss) - gold["loss"]) <= 1e-5 * abs(gold["loss"])
    assert torch.allclose(logits, gold["logits"], atol=1e-5)
    assert set(grads) == set(gold["grad_norms"])
    for n, v in gold["grad_norms"].items():
        assert abs(float(grads[n].norm()) - v) <= 1e-3 * v + 1e-9, n
        assert torch.allclose(grads[n].flatten()[:16], gold["grad_heads"][n], rtol=1e-3, atol=1e-6 + 1e-3 * v / max(1.0, grads[n].numel() ** 0.5)), n
    assert abs(float(dstory.norm()) - gold["dstory_norm"]) <= 1e-3 * gold["dstory_norm"]
    for n, t in gold["buffers"].items():
        assert torch.allclose(sd[n].flatten()[:16].float(), t, atol=1e-5, rtol=1e-4), n


# ------------------------------------------------------------------------------------------ product
@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "emu":
        emulator.install(monkeypatch)
        return torch.device("cpu")
    return torch.device("cuda")


def _product(sd, dev):
    import model
    net = model.VideoEncoder()
    net.load_state_dict(sd, strict=True)
    return net.to(dev).train()


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("B,T", [(4, 5), (3, 4)])
def test_product_video_encoder_against_fp64_oracle(dev, B, T):
    """order logits, BCE order loss (miscc/utils.py:110-120), every parameter gradient, the gradient w.r.t. the
    story and the updated BatchNorm / spectral-norm buffers.  Tolerances: forward GEMMs use bf16 hi/lo split
    operands (relative 2^-16), backward GEMMs single bf16 -> gradient cosine >= 0.999 per tensor."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g = torch.Generator().manual_seed(7 + T)
    story = torch.rand(B, 3, T, 64, 64, generator=g) * 2 - 1
    labels = (torch.arange(B) % 2).float()
    sd0 = VE.init_state(3)
    loss64, logits64, grads64, dstory64, sd64 = _oracle(sd0, story, labels, torch.float64)
    net = _product(sd0, dev)
    x = story.to(dev).requires_grad_(True)
    logits = net(x)
    assert tuple(logits.shape) == (B, 1)
    loss = torch.nn.BCEWithLogitsLoss()(logits, labels.to(dev).unsqueeze(-1))
    loss.backward()
    assert float((logits.detach().cpu().double() - logits64).abs().max()) <= 2e-3 * max(1.0, float(logits64.abs().max()))
    assert abs(float(loss.detach()) - float(loss64)) <= 1e-3 * abs(float(loss64))
    worst = (1.0, None)
    for n, p in net.named_parameters():
        ref = grads64[n]
        if n.endswith("detector.0.bias"):
            # a bias directly in front of a batch-statistics BatchNorm: identically zero gradient
            assert float(p.grad.norm()) <= 1e-4 * float(grads64["detector.0.weight_orig"].norm())
            continue
        c = _cos(p.grad, ref)
        if c < worst[0]:
            worst = (c, n)
        assert abs(float(p.grad.norm()) / float(ref.norm()) - 1) <= 2e-2, n
    assert worst[0] >= 0.999, worst
    assert _cos(x.grad, dstory64) >= 0.999
    # state: running statistics and power-iteration vectors moved exactly once
    new = net.state_dict()
    for n, t in sd64.items():
        if n.rsplit(".", 1)[-1] in ("running_mean", "running_var", "weight_u", "weight_v"):
            assert torch.allclose(new[n].cpu().double(), t, atol=2e-4, rtol=2e-3), n
        if n.endswith("num_batches_tracked"):
            assert int(new[n]) == 1, n


@pytest.mark.timeout(900)
def test_no_grad_call_and_generator_side_loss(dev):
    """the generator-side term (miscc/utils.py:155-169): MSE(logits(fake), logits(real).detach()) -- gradient
    reaches the fake story only; the torch.no_grad() call of the real stories keeps the hi/lo split operands"""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g = torch.Generator().manual_seed(21)
    real = torch.rand(4, 3, 5, 64, 64, generator=g) * 2 - 1
    fake = torch.rand(4, 3, 5, 64, 64, generator=g) * 2 - 1
    sd0 = VE.init_state(5)
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    f64 = fake.double().requires_grad_(True)
    ref = VE.order_loss_g(sd64, real.double(), f64)
    ref.backward()
    net = _product(sd0, dev)
    for p in net.parameters():
        p.requires_grad_(False)
    xf = fake.to(dev).requires_grad_(True)
    with torch.no_grad():
        real_logits = net(real.to(dev))
    loss = torch.nn.MSELoss()(net(xf), real_logits.detach())
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 2e-3 * abs(float(ref))
    assert _cos(xf.grad, f64.grad) >= 0.999
    assert all(p.grad is None for p in net.parameters())
