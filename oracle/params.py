"""Parameter / buffer inventory of the four CP-CSV networks and a seeded initialiser.

The inventory restates the module constructors of the reference (``model.py:214-311``
generator, ``model.py:487-618`` discriminators, ``model.py:68-97`` logits head; key list
in SURVEY.md Appendix B).  ``make_golden.py`` loads these dicts into the REAL reference
modules with ``load_state_dict(strict=True)``, which pins the inventory.

The initialiser draws every tensor from one explicit CPU generator in inventory order, with
the distributions ``weights_init`` uses (reference ``miscc/utils.py:191-201``): N(0, 0.02)
for Conv/Linear weights, N(1, 0.02) / 0 for BatchNorm, 0 for Linear biases; GRUCell tensors
keep torch's U(-1/sqrt(H), 1/sqrt(H)); spectral-norm ``u``/``v`` are unit-normalised normals.
It does NOT reproduce the reference's global-RNG stream -- parity tests load the same dicts
into both implementations instead.
"""
from collections import OrderedDict
import math
import torch


def _bn(d, prefix, c):
    d[prefix + ".weight"] = ("bn_w", (c,))
    d[prefix + ".bias"] = ("zero", (c,))
    d[prefix + ".running_mean"] = ("zero", (c,))
    d[prefix + ".running_var"] = ("one", (c,))
    d[prefix + ".num_batches_tracked"] = ("count", ())


def _sn_conv(d, prefix, cout, cin, k, bias=False):
    d[prefix + ".weight_orig"] = ("w", (cout, cin, k, k))
    if bias:
        d[prefix + ".bias"] = ("conv_b", (cout,), cin * k * k)
    d[prefix + ".weight_u"] = ("unit", (cout,))
    d[prefix + ".weight_v"] = ("unit", (cin * k * k,))


def generator_inventory(p):
    """reference model.py:214-311 (StoryGAN.__init__ / define_module); with
    ``p["CASCADE_MODEL"]`` also the mask re-encoder of cascade_model.py:36-41, 312-320."""
    V, T, L = p["VIDEO_LEN"], p["TEXT_DIM"], p["LABEL_NUM"]
    C, Z = p["CONDITION_DIM"], p["Z_DIM"]
    M = T + L
    ngf = p["GF_DIM"] * 8
    nseg = p["GF_SEG_DIM"]
    ninput = M + C + 124
    d = OrderedDict()
    for name, i, h in (("recurrent", Z + M, M), ("mocornn", M, C)):
        d[name + ".weight_ih"] = ("gru", (3 * h, i), h)
        d[name + ".weight_hh"] = ("gru", (3 * h, h), h)
        d[name + ".bias_ih"] = ("gru", (3 * h,), h)
        d[name + ".bias_hh"] = ("gru", (3 * h,), h)
    d["ca_net.fc.weight"] = ("w", (2 * C, T * V))
    d["ca_net.fc.bias"] = ("zero", (2 * C,))
    d["filter_net.0.weight"] = ("w", (63, C))
    d["filter_net.0.bias"] = ("zero", (63,))
    _bn(d, "filter_net.1", 63)
    d["image_net.0.weight"] = ("w", (372, M))
    d["image_net.0.bias"] = ("zero", (372,))
    _bn(d, "image_net.1", 372)
    d["fc.0.weight"] = ("w", (ngf * 16, ninput))
    _bn(d, "fc.1", ngf * 16)
    c = ngf
    for i in range(1, 5):
        d["upsample%d.1.weight" % i] = ("w", (c // 2, c, 3, 3))
        _bn(d, "upsample%d.2" % i, c // 2)
        c //= 2
    d["img.0.weight"] = ("w", (3, ngf // 16, 3, 3))
    d["seg_c.weight"] = ("w", (ngf, nseg, 3, 3))
    d["seg_c1.weight"] = ("w", (ngf // 2, nseg // 2, 3, 3))
    d["fc_seg.0.weight"] = ("w", (nseg * 16, ninput))
    _bn(d, "fc_seg.1", nseg * 16)
    c = nseg
    for i in range(1, 5):
        d["upsample%d_seg.1.weight" % i] = ("w", (c // 2, c, 3, 3))
        _bn(d, "upsample%d_seg.2" % i, c // 2)
        c //= 2
    d["img_seg.0.weight"] = ("w", (1, nseg // 16, 3, 3))
    d["m_net.0.weight"] = ("w", (M, M))
    d["m_net.0.bias"] = ("zero", (M,))
    _bn(d, "m_net.1", M)
    d["c_net.0.weight"] = ("w", (C, C))
    d["c_net.0.bias"] = ("zero", (C,))
    _bn(d, "c_net.1", C)
    if p.get("CASCADE_MODEL"):
        # cascade_model.py:312-320 (appended last so the non-cascade draws are unchanged)
        d["presample.0.weight"] = ("w", (nseg // 16, 1, 3, 3))
        _bn(d, "presample.1", nseg // 16)
        c = nseg // 16
        for i in range(1, 5):
            d["downsample%d_seg.0.weight" % i] = ("w", (c * 2, c, 3, 3))
            d["downsample%d_seg.0.bias" % i] = ("conv_b", (c * 2,), c * 9)
            _bn(d, "downsample%d_seg.1" % i, c * 2)
            c *= 2
    return d


def discriminator_inventory(p, kind):
    """kind in {'img','seg','sty'}: reference model.py:487-527 / 529-569 / 571-608 and the
    D_GET_LOGITS head model.py:68-84."""
    ndf = p["DF_DIM"]
    nef = p["CONDITION_DIM"] + p["TEXT_DIM"] + p["LABEL_NUM"]
    cin0 = 1 if kind == "seg" else 3
    d = OrderedDict()
    if kind == "sty":
        _sn_conv(d, "encode_img.0", ndf, cin0, 4)
    else:
        d["encode_img.0.weight"] = ("w", (ndf, cin0, 4, 4))
    c = ndf
    for idx in (2, 5, 8):
        _sn_conv(d, "encode_img.%d" % idx, c * 2, c, 4)
        _bn(d, "encode_img.%d" % (idx + 1), c * 2)
        c *= 2
    _sn_conv(d, "get_cond_logits.outlogits.0", ndf * 8, ndf * 8 + nef, 3)
    _bn(d, "get_cond_logits.outlogits.1", ndf * 8)
    _sn_conv(d, "get_cond_logits.outlogits.3", 1, ndf * 8, 4, bias=True)
    if kind != "sty":
        d["cate_classify.weight"] = ("w", (p["LABEL_NUM"], ndf * 8, 4, 4))
    elif p.get("USE_SEQ_CONSISTENCY"):
        # model.py:599-601: the VideoEncoder critic is a child module of the story discriminator
        from .video_encoder import inventory as video_inventory
        for k, spec in video_inventory().items():
            d["seq_consisten_model." + k] = spec
    return d


def _draw(spec, g):
    kind, shape = spec[0], spec[1]
    if kind == "w":
        return torch.randn(*shape, generator=g) * 0.02
    if kind == "bn_w":
        return 1.0 + torch.randn(*shape, generator=g) * 0.02
    if kind == "zero":
        return torch.zeros(*shape)
    if kind == "one":
        return torch.ones(*shape)
    if kind == "count":
        return torch.zeros((), dtype=torch.long)
    if kind == "gru":
        k = 1.0 / math.sqrt(spec[2])
        return (torch.rand(*shape, generator=g) * 2.0 - 1.0) * k
    if kind == "conv_b":
        k = 1.0 / math.sqrt(spec[2])
        return (torch.rand(*shape, generator=g) * 2.0 - 1.0) * k
    if kind == "unit":
        v = torch.randn(*shape, generator=g)
        return v / v.norm().clamp_min(1e-12)
    raise KeyError(kind)


def init_state(inventory, seed):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((k, _draw(spec, g)) for k, spec in inventory.items())


def init_all(p, seed=0):
    """State dicts of (G, D_im, D_st, D_se) in the trainer's construction order
    (reference trainer.py:87-97)."""
    return OrderedDict(
        G=init_state(generator_inventory(p), seed * 4 + 0),
        D_im=init_state(discriminator_inventory(p, "img"), seed * 4 + 1),
        D_st=init_state(discriminator_inventory(p, "sty"), seed * 4 + 2),
        D_se=init_state(discriminator_inventory(p, "seg"), seed * 4 + 3),
    )


def is_parameter(key):
    """True for trainable tensors, False for buffers."""
    tail = key.rsplit(".", 1)[-1]
    return tail not in ("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v")
