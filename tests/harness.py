"""Shared helpers of the parity tests: build the product networks for a preset, load the
oracle-initialised weights, inject the oracle's noise list, run one step, compare."""
import random

import numpy as np
import torch

from oracle import functional as Fn
from oracle import params, presets, synth

ZERO_GRAD = ("filter_net.0.bias", "image_net.0.bias", "m_net.0.bias", "c_net.0.bias",
             # cascade downBlocks: conv bias directly followed by BatchNorm (cascade_model.py:36-41)
             "downsample1_seg.0.bias", "downsample2_seg.0.bias", "downsample3_seg.0.bias",
             "downsample4_seg.0.bias",
             # order-consistency critic: Linear bias in front of BatchNorm1d (model.py:192-194)
             "seq_consisten_model.detector.0.bias")
LOSS_KEYS = ("se_errD", "im_errD", "st_errD", "se_errG", "im_errG", "st_errG", "im_kl", "st_kl")
CASCADE_LOSS_KEYS = ("video_latent_loss", "image_latent_loss", "reconstruct_loss")
IMG_KEYS = ("p1_st_fake", "p1_im_fake", "p1_se_fake", "p3_st_fake", "p3_im_fake", "p3_se_fake")


def build_product(p, states, device):
    from miscc.config import cfg
    presets.apply_to_cfg(cfg, p)
    cfg.CUDA = device.type == "cuda"
    if p.get("CASCADE_MODEL"):
        import cascade_model as model      # reference trainer.py:83-84
    else:
        import model
    nets = {"G": model.StoryGAN(p["VIDEO_LEN"]), "D_im": model.STAGE1_D_IMG(),
            "D_st": model.STAGE1_D_STY_V2(), "D_se": model.STAGE1_D_SEG()}
    for k, net in nets.items():
        net.load_state_dict(states[k], strict=True)
        net.to(device).train()
    return nets


def inject_noise(netG, feed):
    netG.ca_net.draw_eps = lambda like: feed.pop(tuple(like.shape))
    netG.get_gru_initial_state = lambda n: feed.pop((n, netG.motion_dim))
    netG.get_iteration_input = lambda m: torch.cat((feed.pop((m.shape[0], netG.noise_dim)), m), dim=1)


def product_inputs(batch):
    import trainer
    st = {"images": batch["st_real"], "description": batch["st_desc"], "labels": batch["st_labels"]}
    im = {"images": batch["im_real"], "description": batch["im_desc"], "content": batch["im_content"],
          "labels": batch["im_labels"], "images_seg": batch["se_real"]}
    return trainer.prepare_inputs(st, im)


def _seed_host_rngs(seed):
    """create_random_shuffle (cfg.USE_SEQ_CONSISTENCY) draws from the two global host generators"""
    random.seed(seed)
    np.random.seed(seed)


def run_product_step(p, device, seed_w=0, seed_b=1, seed_n=2, fused=False):
    import trainer
    _seed_host_rngs(seed_b)
    states = params.init_all(p, seed_w)
    nets = build_product(p, states, device)
    feed = synth.NoiseFeed(synth.make_noise(p, seed_n, device=device))
    inject_noise(nets["G"], feed)
    batch = synth.make_batch(p, seed_b, device=device)
    x = product_inputs(batch)
    N, B = p["IM_BATCH"], p["ST_BATCH"]
    labels = (torch.ones(N, device=device), torch.zeros(N, device=device),
              torch.ones(B, device=device), torch.zeros(B, device=device))
    opts = trainer.build_optimizers(nets, fused=fused)
    out = trainer.train_step(nets, opts, x, labels, ratio=1.0)
    assert feed.pos == len(feed.tensors)
    grads = {k: {n: (q.grad.detach().clone() if q.grad is not None else torch.zeros_like(q))
                 for n, q in nets[k].named_parameters()} for k in nets}
    return nets, out, grads


def run_oracle_step(p, device, dtype=torch.float32, seed_w=0, seed_b=1, seed_n=2):
    _seed_host_rngs(seed_b)
    states = params.init_all(p, seed_w)
    if dtype != torch.float32:
        states = {k: {n: (t.to(dtype) if t.is_floating_point() else t) for n, t in sd.items()}
                  for k, sd in states.items()}
    model = Fn.OracleModel(states, p, device=device)
    feed = synth.NoiseFeed([t.to(dtype) for t in synth.make_noise(p, seed_n, device=device)])
    batch = {k: v.to(dtype) for k, v in synth.make_batch(p, seed_b, device=device).items()}
    out = Fn.train_step(model, batch, feed)
    grads = dict(out["D_grads"], G=out["G_grads"])
    return model, out, grads


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def compare(out, grads, ref_out, ref_grads, verbose=False):
    """returns dict of worst-case metrics against the reference results"""
    res = {"loss_rel": 0.0, "img_rel": 0.0, "cos_min": 1.0, "cos_min_name": None, "cos_net": {},
           "zero_grad_ratio": 0.0}
    for k in LOSS_KEYS + tuple(k for k in CASCADE_LOSS_KEYS if k in ref_out):
        r = abs(float(out[k]) - float(ref_out[k])) / abs(float(ref_out[k]))
        res["loss_rel"] = max(res["loss_rel"], r)
        if verbose:
            print("  loss %-8s %.6f ref %.6f rel %.2e" % (k, float(out[k]), float(ref_out[k]), r))
    for k in IMG_KEYS:
        r = rel_l2(out[k].float().cpu(), ref_out[k].float().cpu())
        res["img_rel"] = max(res["img_rel"], r)
        if verbose:
            print("  %-11s relL2 %.2e" % (k, r))
    for net, gd in ref_grads.items():
        cat_a, cat_b = [], []
        for n, g in gd.items():
            mine = grads[net][n].cpu()
            g = g.cpu()
            cat_a.append(mine.flatten().double())
            cat_b.append(g.flatten().double())
            if n in ZERO_GRAD:
                # true gradient identically zero (bias in front of a batch-statistics BatchNorm): checked by
                # norm, relative to the gradient of the same layer's weight
                wref = gd.get(n[:-len("bias")] + "weight", gd.get(n[:-len("bias")] + "weight_orig"))
                scale = float(wref.norm()) if wref is not None else 1.0
                ratio = float(mine.double().norm()) / max(scale, 1e-30)
                if verbose:
                    print("  zero-grad %s.%s |g| %.3e (|ref| %.3e) vs |dW| %.3e" % (
                        net, n, float(mine.norm()), float(g.norm()), scale))
                res["zero_grad_ratio"] = max(res["zero_grad_ratio"], ratio)
                continue
            c = cosine(mine, g)
            if verbose and c < 0.9995:
                print("  grad %s.%s cos %.6f |ref| %.3e" % (net, n, c, float(g.norm())))
            if c < res["cos_min"]:
                res["cos_min"], res["cos_min_name"] = c, net + "." + n
        res["cos_net"][net] = cosine(torch.cat(cat_a), torch.cat(cat_b))
    return res
