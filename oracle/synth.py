"""Seeded synthetic inputs for the CP-CSV training step.

Shapes follow the batch dicts the reference's loaders emit (reference
``trainer.py:252-288``, ``datasets/pororo.py:150-151,239-243``); value ranges follow
``Normalize(0.5, 0.5)`` (``main_pororo.py:76,82``).  Everything is drawn on the CPU from
explicit generators so the same tensors can be reproduced on any box.
"""
import torch


def make_batch(p, seed=1, device="cpu"):
    """One training batch: dict of float32 tensors."""
    g = torch.Generator().manual_seed(seed)
    B, N, V = p["ST_BATCH"], p["IM_BATCH"], p["VIDEO_LEN"]
    T, L = p["TEXT_DIM"], p["LABEL_NUM"]

    def unif(*s):
        return torch.rand(*s, generator=g) * 2.0 - 1.0

    def normal(*s):
        return torch.randn(*s, generator=g)

    def bern(*s):
        return (torch.rand(*s, generator=g) < 0.3).float()

    batch = {}
    batch["st_real"] = unif(B, 3, V, 64, 64)
    batch["im_real"] = unif(N, 3, 64, 64)
    batch["se_real"] = unif(N, 1, 64, 64)
    batch["st_desc"] = normal(B, V, T)
    batch["st_labels"] = bern(B, V, L)
    batch["im_desc"] = normal(N, T)
    batch["im_content"] = normal(N, V, T)
    batch["im_labels"] = bern(N, L)
    return {k: v.to(device) for k, v in batch.items()}


def g_call_noise_shapes(p, kind):
    """Noise tensors one generator call consumes, in draw order.

    ``sample_videos`` (reference model.py:348-368): eps (B,C) -> GRU h0 noise (B,M) ->
    V x per-frame noise (B,Z).  ``sample_images`` (model.py:426-434): eps (N,C) ->
    (N,M) -> one (N,Z).
    """
    C, Z = p["CONDITION_DIM"], p["Z_DIM"]
    M = p["TEXT_DIM"] + p["LABEL_NUM"]
    if kind == "videos":
        B, V = p["ST_BATCH"], p["VIDEO_LEN"]
        return [(B, C), (B, M)] + [(B, Z)] * V
    N = p["IM_BATCH"]
    return [(N, C), (N, M), (N, Z)]


def make_noise(p, seed=2, device="cpu", calls=("videos", "images", "videos", "images")):
    """Flat list of N(0,1) tensors covering the four generator calls of one step
    (reference trainer.py:295-300 and 365-368)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for kind in calls:
        for s in g_call_noise_shapes(p, kind):
            out.append(torch.randn(*s, generator=g).to(device))
    return out


class NoiseFeed:
    """Pops pre-generated noise tensors in order; used to make two implementations
    consume identical noise."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.pos = 0

    def pop(self, shape):
        t = self.tensors[self.pos]
        assert tuple(t.shape) == tuple(shape), (self.pos, tuple(t.shape), tuple(shape))
        self.pos += 1
        return t
