"""tests/golden/video_encoder.pt from the REAL reference class (build container only):

    python -m oracle.make_golden_video

Loads the seeded state dict of oracle/video_encoder.py into the reference's own ``VideoEncoder``
(model.py:151-210, ``load_state_dict(strict=True)`` pins the inventory), runs the discriminator-side
order loss of miscc/utils.py:110-120 on a seeded story batch with given labels, and stores the logits,
the loss, per-tensor gradient norms / leading entries, the input gradient statistics and the post-forward
BatchNorm / spectral-norm buffers (leading entries)."""
import os

import torch

from . import presets
from .ref_import import load_reference
from .video_encoder import init_state

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run(B=4, T=5, seed=0):
    torch.manual_seed(1234)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_model, _utils, _cfg = load_reference(presets.get("tiny"))
    net = ref_model.VideoEncoder()
    net.load_state_dict(init_state(seed), strict=True)
    net.train()
    g = torch.Generator().manual_seed(seed + 1)
    story = (torch.rand(B, 3, T, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    labels = (torch.rand(B, generator=g) < 0.5).float()
    logits = net(story)
    loss = torch.nn.BCEWithLogitsLoss()(logits, labels.unsqueeze(-1))
    loss.backward()
    grads = {n: p.grad.detach() for n, p in net.named_parameters()}
    gold = {"B": B, "T": T, "seed": seed, "labels": labels, "logits": logits.detach().clone(), "loss": float(loss),
            "grad_norms": {n: float(v.norm()) for n, v in grads.items()},
            "grad_heads": {n: v.flatten()[:16].clone() for n, v in grads.items()},
            "dstory_norm": float(story.grad.norm()), "dstory_head": story.grad.flatten()[:64].clone(),
            "buffers": {n: t.detach().flatten()[:16].clone() for n, t in net.state_dict().items()
                        if n.rsplit(".", 1)[-1] in ("running_mean", "running_var", "weight_u", "weight_v")},
            "torch": str(torch.__version__)}
    path = os.path.join(OUT, "video_encoder.pt")
    torch.save(gold, path)
    print("video_encoder", float(loss), logits.flatten().tolist(), "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    run()
