"""Generate ``tests/golden/*.pt`` by running the REAL reference (build container only).

    python -m oracle.make_golden [tiny small clevr tiny_cascade small_cascade]
    python -m oracle.make_golden --eval [tiny tiny_cascade]      (eval-mode generator, see run_eval)

For every preset: oracle-initialised weights (``oracle/params.py``, seed 0) are loaded into
the reference's own modules (strict), the seeded synthetic batch (``oracle/synth.py``,
seed 1) and noise list (seed 2) are fed through ``reference_step`` (Adam applied), and the
results are saved.  'tiny' keeps every tensor (images, masks, losses, all gradients,
post-step buffers); 'small' and 'clevr' (full-width model, 150 M parameters) keep losses, per-tensor
gradient norms, the leading 16 entries of every gradient and image statistics.
"""
import os
import sys
import time
import torch

from . import presets, synth, params
from .ref_import import load_reference, build_reference_nets, inject_noise, reference_step

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
BUFFER_TAILS = ("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v")


def run(name):
    torch.manual_seed(1234)
    torch.set_num_threads(os.cpu_count() or 1)
    p = presets.get(name)
    ref_model, ref_utils, cfg = load_reference(p)
    states = params.init_all(p, seed=0)
    nets = build_reference_nets(ref_model, p, states)
    feed = synth.NoiseFeed(synth.make_noise(p, seed=2))
    inject_noise(nets["G"], feed)
    batch = synth.make_batch(p, seed=1)
    opts = {k: torch.optim.Adam([q for q in nets[k].parameters() if q.requires_grad],
                                lr=(p["GENERATOR_LR"] if k == "G" else p["DISCRIMINATOR_LR"]),
                                betas=(0.5, 0.999)) for k in nets}
    t0 = time.time()
    out = reference_step(nets, ref_utils, cfg, batch, ratio=1.0, opts=opts)
    dt = time.time() - t0
    assert feed.pos == len(feed.tensors)
    losses = {k: float(v) for k, v in out.items() if torch.is_tensor(v) and v.dim() == 0}
    post_buffers = {k: {n: t.detach().clone() for n, t in nets[k].state_dict().items()
                        if n.rsplit(".", 1)[-1] in BUFFER_TAILS} for k in nets}
    gold = {"preset": p, "losses": losses, "seconds": dt, "torch": str(torch.__version__),
            "threads": torch.get_num_threads()}
    grads = dict(out["D_grads"], G=out["G_grads"])
    if name.startswith("tiny"):
        for k in ("p1_st_fake", "p1_im_fake", "p1_se_fake", "p3_st_fake", "p3_im_fake", "p3_se_fake"):
            gold[k] = out[k].contiguous().clone()
        gold["grads"] = grads
        gold["post_buffers"] = post_buffers
        gold["post_params_sample"] = {k: {n: q.detach().flatten()[:32].clone()
                                          for n, q in nets[k].named_parameters()} for k in nets}
    else:
        gold["grad_norms"] = {k: {n: float(g.norm()) for n, g in v.items()} for k, v in grads.items()}
        gold["grad_heads"] = {k: {n: g.flatten()[:16].clone() for n, g in v.items()} for k, v in grads.items()}
        gold["image_stats"] = {k: (float(out[k].mean()), float(out[k].std()), out[k].flatten()[:64].clone())
                               for k in ("p1_st_fake", "p1_im_fake", "p1_se_fake",
                                         "p3_st_fake", "p3_im_fake", "p3_se_fake")}
        gold["post_buffers_sample"] = {k: {n: t.flatten()[:16].clone() for n, t in v.items()}
                                       for k, v in post_buffers.items()}
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "step_%s.pt" % name)
    torch.save(gold, path)
    print(name, "%.1fs" % dt, {k: round(v, 6) for k, v in losses.items()},
          "%.2f MB" % (os.path.getsize(path) / 1e6))


def run_eval(name):
    """eval-mode generator (reference inference.py:88-89 / trainer.py:161-162: ``netG.eval()`` under
    no_grad): one train-mode no-grad pass first, so the running statistics are no longer their
    initial 0 / 1, then ``sample_videos`` / ``sample_images`` with ``seg=True`` in eval mode.
    Writes tests/golden/eval_<name>.pt."""
    torch.manual_seed(1234)
    p = presets.get(name)
    ref_model, _ref_utils, cfg = load_reference(p)
    states = params.init_all(p, seed=0)
    netG = build_reference_nets(ref_model, p, states)["G"]
    feed = synth.NoiseFeed(synth.make_noise(p, seed=2))
    inject_noise(netG, feed)
    batch = synth.make_batch(p, seed=1)
    T = cfg.TEXT.DIMENSION
    im_motion = torch.cat((batch["im_desc"][:, :T], batch["im_labels"]), 1)
    im_content = batch["im_content"][:, :, :T]
    st_motion = torch.cat((batch["st_desc"][:, :, :T], batch["st_labels"]), 2)
    st_content = batch["st_desc"][:, :, :T]
    with torch.no_grad():
        netG.train()
        netG.sample_videos(st_motion, st_content)
        netG.sample_images(im_motion, im_content, seg=True)
        netG.eval()
        _, st_fake, _, _, _, _, st_seg = netG.sample_videos(st_motion, st_content, seg=True)
        _, im_fake, _, _, _, _, im_seg = netG.sample_images(im_motion, im_content, seg=True)
    assert feed.pos == len(feed.tensors)
    gold = {"preset": p, "torch": str(torch.__version__),
            "st_fake": st_fake.contiguous().clone(), "st_seg": st_seg.contiguous().clone(),
            "im_fake": im_fake.contiguous().clone(), "im_seg": im_seg.contiguous().clone()}
    path = os.path.join(OUT, "eval_%s.pt" % name)
    torch.save(gold, path)
    print("eval", name, "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--eval":
        run_eval(sys.argv[2])
        sys.exit(0)
    # one preset per process: the reference reads cfg at import/construction time
    names = sys.argv[1:] or ["tiny"]
    assert len(names) == 1, "run one preset per process"
    run(names[0])
