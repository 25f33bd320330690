"""The oracle restatement (oracle/functional.py) against outputs of the REAL reference
(tests/golden/step_*.pt, written by oracle/make_golden.py in the build container)."""
import os

import pytest
import torch

from oracle import functional as Fn
from oracle import params, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
# Linear biases directly followed by BatchNorm have an identically-zero true gradient;
# autograd returns rounding noise for them.
ZERO_GRAD = ("filter_net.0.bias", "image_net.0.bias", "m_net.0.bias", "c_net.0.bias",
             # cascade downBlocks: conv bias directly followed by BatchNorm (cascade_model.py:36-41)
             "downsample1_seg.0.bias", "downsample2_seg.0.bias", "downsample3_seg.0.bias",
             "downsample4_seg.0.bias")


def _run(name):
    gold = torch.load(os.path.join(GOLD, "step_%s.pt" % name))
    p = gold["preset"]
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    model = Fn.OracleModel(params.init_all(p, 0), p)
    feed = synth.NoiseFeed(synth.make_noise(p, 2))
    out = Fn.train_step(model, synth.make_batch(p, 1), feed)
    assert feed.pos == len(feed.tensors)
    return gold, model, out


def _grads(out):
    return dict(out["D_grads"], G=out["G_grads"])


@pytest.mark.parametrize("name", ["tiny", "tiny_cascade"])
def test_tiny_full_tensors(name):
    """every tensor of one step of the real reference; 'tiny_cascade' = cascade_model.py with the
    latent-MSE / reconstruction losses of trainer.py:369-384"""
    gold, model, out = _run(name)
    for k, v in gold["losses"].items():
        assert abs(float(out[k]) - v) <= 1e-5 * abs(v) + 1e-7, k
    for k in ("p1_st_fake", "p1_im_fake", "p1_se_fake", "p3_st_fake", "p3_im_fake", "p3_se_fake"):
        assert torch.allclose(out[k], gold[k], atol=5e-6), k
    mine = _grads(out)
    for net, gd in gold["grads"].items():
        assert set(gd) == set(mine[net])
        for n, g in gd.items():
            if n in ZERO_GRAD:
                # rounding noise only (the conv-bias ones see the O(1) latent-MSE gradients)
                assert mine[net][n].norm() < (1e-3 if "downsample" in n else 1e-4)
                continue
            rel = (g - mine[net][n]).norm() / g.norm().clamp_min(1e-12)
            assert rel < 1e-4, (net, n, float(rel))
    for net, bd in gold["post_buffers"].items():
        for n, t in bd.items():
            assert torch.allclose(t.float(), model.nets[net][n].float(), atol=5e-4, rtol=1e-4), (net, n)


@pytest.mark.parametrize("name", ["small", "clevr", "small_cascade", "pororo"])
def test_summary_presets(name):
    gold, model, out = _run(name)
    for k, v in gold["losses"].items():
        assert abs(float(out[k]) - v) <= 5e-4 * abs(v) + 1e-7, (k, float(out[k]), v)
    mine = _grads(out)
    for net, gn in gold["grad_norms"].items():
        for n, v in gn.items():
            if n in ZERO_GRAD:
                continue
            assert abs(float(mine[net][n].norm()) - v) <= 5e-2 * v + 1e-9, (net, n)
            head = gold["grad_heads"][net][n]
            got = mine[net][n].flatten()[:16]
            rms = v / max(1.0, mine[net][n].numel() ** 0.5)
            assert (head - got).norm() <= 0.1 * head.norm() + 0.1 * rms + 1e-9, (net, n)
    for k, (mean, std, head) in gold["image_stats"].items():
        assert abs(float(out[k].mean()) - mean) < 1e-4
        assert torch.allclose(out[k].flatten()[:64], head, atol=1e-4), k


def test_inventory_counts_match_reference():
    """Parameter counts published in SURVEY.md Appendix B (measured on the reference)."""
    from oracle import presets
    p = presets.get("pororo")
    def count(inv):
        n = 0
        for k, spec in inv.items():
            if params.is_parameter(k):
                m = 1
                for s in spec[1]:
                    m *= s
                n += m
        return n
    assert count(params.generator_inventory(p)) == 86995977
    assert count(params.discriminator_inventory(p, "img")) == 23725169
    assert count(params.discriminator_inventory(p, "sty")) == 23582321
    assert count(params.discriminator_inventory(p, "seg")) == 23721201
