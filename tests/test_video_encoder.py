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
