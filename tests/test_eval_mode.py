"""Eval-mode generator (``netG.eval()`` under ``torch.no_grad()``: reference inference.py:88-89 and
the FID / SSIM loops of trainer.py:161-182): BatchNorm normalises with its running statistics.

Golden: tests/golden/eval_tiny*.pt, produced by the REAL reference (``python -m oracle.make_golden
--eval tiny``): one train-mode no-grad pass of sample_videos + sample_images (moves the running
statistics), then both calls with ``seg=True`` in eval mode.  Checked: the oracle restatement, the
product's host logic on the kernel-contract emulator and (``-m gpu``, tests/test_zz_cascade.py) the
product on the real kernels."""
import os

import pytest
import torch

import emulator
import harness
from oracle import functional as Fn
from oracle import params, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("st_fake", "st_seg", "im_fake", "im_seg")


def _inputs(p, batch):
    T = p["TEXT_DIM"]
    return (torch.cat((batch["st_desc"][:, :, :T], batch["st_labels"]), 2), batch["st_desc"][:, :, :T],
            torch.cat((batch["im_desc"][:, :T], batch["im_labels"]), 1), batch["im_content"][:, :, :T])


@pytest.mark.parametrize("name", ["tiny", "tiny_cascade"])
def test_oracle_eval_mode_vs_reference(name):
    gold = torch.load(os.path.join(GOLD, "eval_%s.pt" % name))
    p = gold["preset"]
    G = Fn.OracleModel(params.init_all(p, 0), p, with_optim=False).nets["G"]
    feed = synth.NoiseFeed(synth.make_noise(p, 2))
    st_m, st_c, im_m, im_c = _inputs(p, synth.make_batch(p, 1))
    with torch.no_grad():
        Fn.sample_videos(G, st_m, st_c, feed)
        Fn.sample_images(G, im_m, im_c, feed, seg=True)
        with Fn.eval_mode():
            _, st_fake, _, _, _, _, st_seg = Fn.sample_videos(G, st_m, st_c, feed, seg=True)
            _, im_fake, _, _, _, _, im_seg = Fn.sample_images(G, im_m, im_c, feed, seg=True)
    assert feed.pos == len(feed.tensors)
    for k, v in zip(KEYS, (st_fake, st_seg, im_fake, im_seg)):
        assert torch.allclose(v, gold[k], atol=5e-6), k


def run_product_eval(name, device):
    """shared with the -m gpu case: product generator, train-mode no-grad pass then eval-mode pass"""
    gold = torch.load(os.path.join(GOLD, "eval_%s.pt" % name))
    p = gold["preset"]
    netG = harness.build_product(p, params.init_all(p, 0), device)["G"]
    harness.inject_noise(netG, synth.NoiseFeed(synth.make_noise(p, 2, device=device)))
    st_m, st_c, im_m, im_c = _inputs(p, synth.make_batch(p, 1, device=device))
    with torch.no_grad():
        netG.train()
        netG.sample_videos(st_m, st_c)
        netG.sample_images(im_m, im_c, seg=True)
        before = {k: v.clone() for k, v in netG.state_dict().items() if "running" in k or "tracked" in k}
        netG.eval()
        _, st_fake, _, _, _, _, st_seg = netG.sample_videos(st_m, st_c, seg=True)
        _, im_fake, _, _, _, _, im_seg = netG.sample_images(im_m, im_c, seg=True)
    for k, v in netG.state_dict().items():          # eval mode leaves every BatchNorm buffer alone
        if k in before:
            assert torch.equal(v, before[k]), k
    for k, v in zip(KEYS, (st_fake, st_seg, im_fake, im_seg)):
        r = harness.rel_l2(v.float().cpu(), gold[k])
        print("eval %s %s relL2 %.2e" % (name, k, r))
        assert r <= 2e-2, (k, r)
    # eval mode with gradients is refused loudly, not computed with batch statistics
    harness.inject_noise(netG, synth.NoiseFeed(synth.make_noise(p, 3, device=device, calls=("images",))))
    with pytest.raises((RuntimeError, NotImplementedError), match="eval-mode"):
        netG.sample_images(im_m, im_c, seg=True)
    return netG


@pytest.mark.parametrize("name", ["tiny", "tiny_cascade"])
def test_product_eval_mode_emulated(monkeypatch, name):
    emulator.install(monkeypatch)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    run_product_eval(name, torch.device("cpu"))
