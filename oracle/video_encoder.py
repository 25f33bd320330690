"""Oracle restatement of the reference's ``VideoEncoder`` (model.py:99-210), the order-consistency critic
inside ``STAGE1_D_STY_V2`` when ``cfg.USE_SEQ_CONSISTENCY`` (model.py:602-608), and of the two loss terms
that use it (miscc/utils.py:110-122, 155-169).  TEST INFRASTRUCTURE: imported by tests / make_golden only.

Layer list = ``VideoEncoder.story_encoder`` (model.py:155-190) with the ``R2Plus1dStem`` of model.py:99-113 in
front and the ``detector`` of model.py:192-197 behind; every Conv3d / Linear carries the legacy spectral-norm
hook, no conv has a bias, BatchNorm3d / 1d in train mode.
"""
import random
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .functional import batch_norm, spectral_weight

# (state-dict index inside story_encoder, Cout, Cin, kernel (t, h, w), stride, padding)
ENCODER_CONVS = (
    (1, 128, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (4, 128, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    (7, 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (10, 256, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    (13, 256, 256, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (16, 512, 256, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    (19, 512, 512, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (22, 512, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
)


def inventory():
    """key -> (kind, shape[, fan]) in the reference's state_dict order (pinned by load_state_dict(strict=True)
    into the real class in make_golden.py)"""
    d = OrderedDict()

    def sn(prefix, shape):
        cout = shape[0]
        cols = 1
        for s in shape[1:]:
            cols *= s
        d[prefix + ".weight_orig"] = ("w", shape)
        d[prefix + ".weight_u"] = ("unit", (cout,))
        d[prefix + ".weight_v"] = ("unit", (cols,))

    def bn(prefix, c):
        d[prefix + ".weight"] = ("bn_w", (c,))
        d[prefix + ".bias"] = ("zero", (c,))
        d[prefix + ".running_mean"] = ("zero", (c,))
        d[prefix + ".running_var"] = ("one", (c,))
        d[prefix + ".num_batches_tracked"] = ("count", ())

    sn("story_encoder.0.0", (45, 3, 1, 7, 7))
    bn("story_encoder.0.1", 45)
    sn("story_encoder.0.3", (64, 45, 1, 1, 1))
    bn("story_encoder.0.4", 64)
    for idx, co, ci, k, _s, _p in ENCODER_CONVS:
        sn("story_encoder.%d" % idx, (co, ci) + k)
        bn("story_encoder.%d" % (idx + 1), co)
    d["detector.0.bias"] = ("zero", (128,))
    sn("detector.0", (128, 512))
    bn("detector.1", 128)
    d["detector.3.bias"] = ("zero", (1,))
    sn("detector.3", (1, 128))
    return d


def init_state(seed=0):
    """seeded state dict with the distributions of weights_init (miscc/utils.py:191-201; Conv3d / Linear
    weights N(0, 0.02), BatchNorm N(1, 0.02) / 0, zero biases), spectral-norm u / v unit normals"""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, spec in inventory().items():
        kind, shape = spec[0], spec[1]
        if kind == "w":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "bn_w":
            t = 1.0 + torch.randn(shape, generator=g) * 0.02
        elif kind == "zero":
            t = torch.zeros(shape)
        elif kind == "one":
            t = torch.ones(shape)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.long)
        elif kind == "unit":
            t = F.normalize(torch.randn(shape, generator=g), dim=0, eps=1e-12)
        else:
            raise KeyError(kind)
        sd[k] = t
    return sd


def video_encoder(sd, story):
    """VideoEncoder.forward (model.py:199-210): story (B, 3, T, H, W) -> order logits (B, 1)."""
    B = story.shape[0]
    # R2Plus1dStem (model.py:99-113)
    h = F.conv3d(story, spectral_weight(sd, "story_encoder.0.0"), None, (1, 2, 2), (0, 3, 3))
    h = F.relu(batch_norm(sd, "story_encoder.0.1", h))
    h = F.conv3d(h, spectral_weight(sd, "story_encoder.0.3"), None, (1, 1, 1), (1, 0, 0))
    h = F.relu(batch_norm(sd, "story_encoder.0.4", h))
    for idx, _co, _ci, _k, stride, pad in ENCODER_CONVS:
        h = F.conv3d(h, spectral_weight(sd, "story_encoder.%d" % idx), None, stride, pad)
        h = F.leaky_relu(batch_norm(sd, "story_encoder.%d" % (idx + 1), h), 0.2)
    lat = F.adaptive_avg_pool3d(h, 1).view(B, -1)
    x = F.linear(lat, spectral_weight(sd, "detector.0"), sd["detector.0.bias"])
    x = F.relu(batch_norm(sd, "detector.1", x))
    return F.linear(x, spectral_weight(sd, "detector.3"), sd["detector.3.bias"])


def order_loss_d(sd, shuffled_stories, order_labels):
    """the discriminator-side term (miscc/utils.py:110-120): BCEWithLogits(order_logits, labels)"""
    logits = video_encoder(sd, shuffled_stories)
    return F.binary_cross_entropy_with_logits(logits, order_labels.unsqueeze(-1)), logits


def order_loss_g(sd, real_stories, fake_stories):
    """the generator-side term (miscc/utils.py:155-169): MSE(logits(fake), logits(real).detach())"""
    real_logits = video_encoder(sd, real_stories)
    fake_logits = video_encoder(sd, fake_stories)
    return F.mse_loss(fake_logits, real_logits.detach())


def create_random_shuffle(stories, random_rate=0.5):
    """miscc/utils.py:14-44: per story, with probability ``random_rate`` (numpy global RNG) permute the frames
    into a non-sorted order (python ``random``; re-drawn with numpy while still sorted) and, when another story
    is drawn, replace one frame by that story's frame of the same position; label 1 = shuffled.  Consumes the
    two global generators in the reference's order, so equal seeds give equal batches."""
    host = stories.detach().cpu()
    n, T = host.shape[0], host.shape[2]
    out, labels = [], []
    for i in range(n):
        label = 1 if random_rate > np.random.random() else 0
        labels.append(label)
        if label == 0:
            out.append(host[i].clone())
            continue
        seq = random.sample(range(T), T)
        while all(seq[j] <= seq[j + 1] for j in range(T - 1)):
            np.random.shuffle(seq)
        story = host[i][:, list(seq)].clone()
        j = random.randint(0, n - 1)
        if j != i:
            mix = random.sample(range(T), 1)
            story[:, mix] = host[j][:, mix].clone()
        out.append(story)
    return torch.stack(out, 0).to(stories.device), torch.tensor(labels, dtype=stories.dtype, device=stories.device)


def sub_state(sd, prefix="seq_consisten_model."):
    """the critic's tensors inside a story-discriminator state dict (same tensor objects: in-place buffer
    updates and .grad land in the parent's dict)"""
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
