"""Every non-GEMM libcpcsv.so kernel against the CPU emulator of its contract (and thereby
against the torch formulas the emulator is written in).  GPU only."""
import pytest
import torch

import emulator as emu
from cpcsv_b200 import ops

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def both(fn_name, inputs, outputs, **kw):
    """run ops.<fn> on cuda copies and emulator.<fn> on cpu copies; returns (gpu_outs, cpu_outs)"""
    def to(x, dev):
        return x.to(dev).clone() if torch.is_tensor(x) else x
    res = []
    for dev, mod in (("cuda", ops), ("cpu", emu)):
        args = [to(a, dev) for a in inputs]
        outs = {k: (to(v, dev) if v is not None else None) for k, v in outputs.items()}
        getattr(mod, fn_name)(*args, **outs, **kw)
        if dev == "cuda":
            torch.cuda.synchronize()
        res.append(({k: (v.cpu() if v is not None else None) for k, v in outs.items()},
                    [a.cpu() if torch.is_tensor(a) else a for a in args]))
    return res


def close(a, b, tol=1e-5):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)) < tol


def test_bn_forward_chain():
    rows, C = 1000, 192
    x = rnd(rows, C, seed=1) * 2 + 0.5
    stats_g = ops.bn_workspace(rows, C, "cuda")
    ops.bn_stats(x.cuda(), stats_g)
    stats_c = torch.zeros(2 * C, dtype=torch.float64)
    emu.bn_stats(x, stats_c)
    assert close(stats_g[:2 * C].cpu(), stats_c, 1e-10)
    gamma, beta = rnd(256, seed=2), rnd(256, seed=3)
    cmap = torch.randperm(256, generator=torch.Generator().manual_seed(4))[:C].int()
    outs = {}
    for dev, mod in (("cuda", ops), ("cpu", emu)):
        rm, rv = torch.zeros(256, device=dev), torch.ones(256, device=dev)
        o = [torch.empty(C, device=dev) for _ in range(4)]
        mod.bn_finalize(stats_c.to(dev), rows, gamma.to(dev), beta.to(dev), rm, rv, cmap.to(dev), C - 8, *o)
        outs[dev] = [t.cpu() for t in o] + [rm.cpu(), rv.cpu()]
    for a, b in zip(outs["cuda"], outs["cpu"]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    mean, invstd, scale, shift = outs["cpu"][:4]
    mod_t = rnd(rows, C, seed=5, scale=0.3)
    for act in (0, 1, 2):
        for use_mod in (False, True):
            (g, _), (c, _) = both("bn_act_pack", [x, scale, shift, act],
                                  dict(mod=mod_t if use_mod else None, y=torch.empty(rows, C),
                                       hi=torch.empty(rows, C, dtype=torch.bfloat16),
                                       lo=torch.empty(rows, C, dtype=torch.bfloat16)))
            assert close(g["y"], c["y"], 1e-6)
            assert close(g["hi"].float() + g["lo"].float(), c["y"], 2e-5)
            assert torch.equal(g["hi"], c["hi"]) or close(g["hi"].float(), c["hi"].float(), 1e-3)
    # backward
    dy = rnd(rows, C, seed=6)
    for act, has_bn, use_mod in ((1, True, True), (2, True, False), (2, False, False), (0, True, False)):
        sums = {}
        for dev, mod in (("cuda", ops), ("cpu", emu)):
            s = mod.bn_workspace(rows, C, dev)
            mod.bn_bwd_reduce(x.to(dev), dy.to(dev), scale.to(dev) if has_bn else None,
                              shift.to(dev) if has_bn else None, mean.to(dev), invstd.to(dev), act,
                              mod_t.to(dev) if use_mod else None, s)
            sums[dev] = s[:2 * C].cpu()
        assert close(sums["cuda"], sums["cpu"], 1e-6)
        (g, _), (c, _) = both(
            "bn_bwd_apply",
            [x, dy, scale if has_bn else None, shift if has_bn else None, mean, invstd, cmap, C - 8, act,
             mod_t if use_mod else None, sums["cpu"], has_bn],
            dict(dx=torch.empty(rows, C), dx16=torch.empty(rows, C, dtype=torch.bfloat16),
                 dmod=torch.empty(rows, C) if use_mod else None,
                 dmod16=torch.empty(rows, C, dtype=torch.bfloat16) if use_mod else None,
                 dgamma=torch.zeros(256), dbeta=torch.zeros(256)))
        for k in g:
            if g[k] is not None:
                assert close(g[k].float(), c[k].float(), 2e-3 if "16" in k else 1e-5), (k, act, has_bn)


def test_layout_kernels():
    x = rnd(5, 7, 4, 4, seed=7).permute(0, 1, 3, 2)       # non-contiguous NCHW
    b = rnd(5, 9, seed=8)
    (g, _), (c, _) = both("pack_nchw", [x, b],
                          dict(hi=torch.empty(5 * 16, 64, dtype=torch.bfloat16),
                               lo=torch.empty(5 * 16, 64, dtype=torch.bfloat16)), cpad=64)
    assert torch.equal(g["hi"], c["hi"]) and torch.equal(g["lo"], c["lo"])
    img = rnd(3, 3, 16, 16, seed=9)
    (g, _), (c, _) = both("im2col_small", [img, 4, 2, 1],
                          dict(hi=torch.empty(3 * 64, 64, dtype=torch.bfloat16),
                               lo=torch.empty(3 * 64, 64, dtype=torch.bfloat16)), ldp=64)
    assert torch.equal(g["hi"], c["hi"]) and torch.equal(g["lo"], c["lo"])
    dcol = rnd(3 * 64, 64, seed=10)
    (g, _), (c, _) = both("col2im_small", [dcol, 3, 3, 16, 16, 4, 2, 1], dict(dx=torch.empty(3, 3, 16, 16)))
    assert close(g["dx"], c["dx"])
    yy = torch.tanh(rnd(2, 3, 8, 8, seed=11))
    dy = rnd(2, 8, 3, 8, seed=12).permute(0, 2, 1, 3)     # strided
    (g, _), (c, _) = both("tanh_bwd_im2col", [dy, yy], dict(col=torch.empty(128, 64, dtype=torch.bfloat16)))
    assert close(g["col"].float(), c["col"].float(), 2e-3)


def test_weight_layout_kernels():
    w = rnd(20, 12, 3, 3, seed=13, scale=0.05)
    for kind in (0, 1, 2, 3):
        rp, cp = (64, 64)
        nt = 16 if kind >= 2 else 9
        (g, _), (c, _) = both("pack_conv_weight", [w, kind, rp, cp],
                              dict(hi=torch.empty(nt * rp, cp, dtype=torch.bfloat16),
                                   lo=torch.empty(nt * rp, cp, dtype=torch.bfloat16)))
        assert close(g["hi"].float() + g["lo"].float(), c["hi"].float() + c["lo"].float(), 1e-5), kind
        dwt = rnd(nt, rp, cp, seed=14)
        alpha = torch.tensor([0.7])
        (g, _), (c, _) = both("unpack_conv_wgrad", [dwt, rp * cp, cp, kind, alpha], dict(dw=torch.empty(20, 12, 3, 3)))
        assert close(g["dw"], c["dw"], 1e-6), kind
    m = rnd(30, 17, seed=15)
    rmap = torch.randint(0, 30, (40,), generator=torch.Generator().manual_seed(16)).int()
    (g, _), (c, _) = both("pack_matrix", [m, 40, 64, 17, 17, 1, rmap],
                          dict(hi=torch.empty(40, 64, dtype=torch.bfloat16),
                               lo=torch.empty(40, 64, dtype=torch.bfloat16)))
    assert torch.equal(g["hi"], c["hi"]) and torch.equal(g["lo"], c["lo"])


def test_small_fp32_ops():
    x, w, bias = rnd(37, 150, seed=17), rnd(70, 150, seed=18), rnd(70, seed=19)
    (g, _), (c, _) = both("linear_f32", [x, w, bias], dict(y=torch.zeros(37, 70)))
    assert close(g["y"], c["y"], 1e-5)
    (g, _), (c, _) = both("linear_tn_f32", [rnd(37, 70, seed=20), x], dict(y=torch.ones(70, 150)), accumulate=True)
    assert close(g["y"], c["y"], 1e-5)
    (g, _), (c, _) = both("linear_nn_f32", [rnd(37, 70, seed=21), w], dict(y=torch.zeros(37, 150)))
    assert close(g["y"], c["y"], 1e-5)
    B, H = 9, 50
    gi, gh, h = rnd(B, 3 * H, seed=22), rnd(B, 3 * H, seed=23), rnd(B, H, seed=24)
    (g, _), (c, _) = both("gru_gates_fwd", [gi, gh, h], dict(hnew=torch.empty(B, H), save=torch.empty(B, 4 * H)))
    assert close(g["hnew"], c["hnew"], 1e-5) and close(g["save"], c["save"], 1e-5)
    (g2, _), (c2, _) = both("gru_gates_bwd", [rnd(B, H, seed=25), h, c["save"]],
                            dict(dgi=torch.empty(B, 3 * H), dgh=torch.empty(B, 3 * H), dh=torch.empty(B, H)))
    for k in g2:
        assert close(g2[k], c2[k], 1e-5)
    pre, eps = rnd(B, 24, seed=26), rnd(B, 12, seed=27)
    (g, _), (c, _) = both("ca_fwd", [pre, eps], dict(mu=torch.empty(B, 12), logvar=torch.empty(B, 12), code=torch.empty(B, 12)))
    for k in g:
        assert close(g[k], c[k], 1e-5)
    (g, _), (c, _) = both("ca_bwd", [pre, eps, rnd(B, 12, seed=28), rnd(B, 12, seed=29), rnd(B, 12, seed=30)],
                          dict(dpre=torch.empty(B, 24)))
    assert close(g["dpre"], c["dpre"], 1e-5)
    img, filt = rnd(6, 3, 124, seed=31), rnd(6, 1, 3, 21, seed=32)
    (g, _), (c, _) = both("dfn1d_fwd", [img, filt], dict(out=torch.empty(6, 1, 124)))
    assert close(g["out"], c["out"], 1e-5)
    (g, _), (c, _) = both("dfn1d_bwd", [img, filt, rnd(6, 1, 124, seed=33)],
                          dict(dimg=torch.empty(6, 3, 124), dfilt=torch.empty(6, 1, 3, 21)))
    assert close(g["dimg"], c["dimg"], 1e-5) and close(g["dfilt"], c["dfilt"], 1e-5)


def test_spectral_norm_kernels():
    R, Cc = 50, 333
    w = rnd(R, Cc, seed=34, scale=0.02)
    u = torch.nn.functional.normalize(rnd(R, seed=35), dim=0)
    v = torch.nn.functional.normalize(rnd(Cc, seed=36), dim=0)
    res = {}
    for dev, mod in (("cuda", ops), ("cpu", emu)):
        uu, vv = u.to(dev).clone(), v.to(dev).clone()
        s, si = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        mod.spectral_sigma(w.to(dev), uu, vv, True, s, si, torch.empty(R + Cc, device=dev))
        g = rnd(R, Cc, seed=37).to(dev)
        dw = torch.empty(R, Cc, device=dev)
        mod.spectral_bwd(g, w.to(dev), uu, vv, s, dw, torch.zeros(4, device=dev))
        res[dev] = [t.cpu() for t in (uu, vv, s, si, dw)]
    for a, b in zip(res["cuda"], res["cpu"]):
        assert close(a, b, 1e-5)


def test_skinny_linear_and_big_unpack():
    x, w, bias = rnd(90, 5000, seed=40), rnd(9, 5000, seed=41), rnd(9, seed=42)
    (g, _), (c, _) = both("linear_f32", [x, w, bias], dict(y=torch.zeros(90, 9)))
    assert close(g["y"], c["y"], 1e-5)
    (g, _), (c, _) = both("linear_f32", [x, w[:1], None], dict(y=torch.ones(90, 1)), accumulate=True)
    assert close(g["y"], c["y"], 1e-5)
    for kind, k in ((0, 4), (0, 3), (2, 3)):
        nt = 16 if kind == 2 else k * k
        dwt = rnd(nt, 192, 320, seed=43)
        (g, _), (c, _) = both("unpack_conv_wgrad", [dwt, 192 * 320, 320, kind, None],
                              dict(dw=torch.empty(150, 300, k, k)))
        assert close(g["dw"], c["dw"], 1e-6), kind


def test_weight_cache_sees_fused_optimizer_updates():
    """Adam(fused=True) does not bump Tensor._version; the packed-operand cache must still be
    refreshed (global optimizer post-step hook in cpcsv_b200.nets)."""
    import harness
    from oracle import functional as Fn
    from oracle import params, presets, synth
    p = presets.get("small")
    dev = torch.device("cuda")
    nets = harness.build_product(p, params.init_all(p, 0), dev)
    G = nets["G"]
    batch = synth.make_batch(p, 1, device=dev)
    x = harness.product_inputs(batch)

    def run():
        feed = synth.NoiseFeed(synth.make_noise(p, 2, device=dev, calls=("images",)))
        harness.inject_noise(G, feed)
        with torch.no_grad():
            return G.sample_images(x["im_motion"], x["im_content"], seg=True)[1]
    img0 = run()
    opt = torch.optim.Adam(G.parameters(), lr=2e-2, betas=(0.5, 0.999), fused=True)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for q in G.parameters():
        q.grad = torch.randn(q.shape, device=dev, generator=gen)
    opt.step()
    img1 = run()
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    # the product forward above already advanced the BN buffers once more than the oracle will
    feed = synth.NoiseFeed(synth.make_noise(p, 2, device=dev, calls=("images",)))
    with torch.no_grad():
        ref = Fn.sample_images(sd, x["im_motion"], x["im_content"], feed, seg=True)[1]
    # no-grad generator calls run single-pass fp16 GEMMs: ~1.5e-3 relative on the images
    assert harness.rel_l2(img1, ref) < 5e-3, harness.rel_l2(img1, ref)
    assert harness.rel_l2(img0, ref) > 1e-2      # the update really changed the output


def test_sgemm_split_k_paths():
    """the conditioning-path SGEMM splits K over the grid (atomic partial sums into a kernel-zeroed
    output, bias from split 0): CA_NET-sized K, all three operand layouts, with and without
    accumulation into an existing output"""
    x, w, bias = rnd(18, 1780, seed=31), rnd(248, 1780, seed=32, scale=0.05), rnd(248, seed=33)
    (g, _), (c, _) = both("linear_f32", [x, w, bias], dict(y=torch.full((18, 248), 7.0)))
    assert close(g["y"], c["y"], 2e-6)
    (g, _), (c, _) = both("linear_f32", [x, w, bias], dict(y=torch.full((18, 248), 7.0)), accumulate=True)
    assert close(g["y"], c["y"], 2e-6)
    a, b = rnd(450, 124, seed=34), rnd(450, 372, seed=35)          # K = rows (5 steps x 90)
    (g, _), (c, _) = both("linear_tn_f32", [a, b], dict(y=torch.zeros(124, 372)))
    assert close(g["y"], c["y"], 2e-6)
    dy, w2 = rnd(90, 1095, seed=36), rnd(1095, 465, seed=37, scale=0.05)
    (g, _), (c, _) = both("linear_nn_f32", [dy, w2], dict(y=torch.zeros(90, 465)))
    assert close(g["y"], c["y"], 2e-6)


def test_gru_sequence_matches_torch_grucell():
    """functions.GRUSeqFn (hoisted input projections, per-step recurrent GEMM + gate kernel) vs a
    loop of torch.nn.GRUCell in fp64: outputs and every gradient (reference model.py:313-346)"""
    from cpcsv_b200 import functions as Fx
    T, B, I, H = 5, 18, 465, 365
    cell = torch.nn.GRUCell(I, H).cuda()
    x = rnd(T, B, I, seed=41).cuda().requires_grad_(True)
    h0 = rnd(B, H, seed=42).cuda().requires_grad_(True)
    gout = rnd(T, B, H, seed=43).cuda()
    out = Fx.gru_sequence(x, h0, cell)
    out.backward(gout)
    got = [out.detach(), x.grad, h0.grad] + [p.grad for p in cell.parameters()]
    ref_cell = torch.nn.GRUCell(I, H).double().cuda()
    ref_cell.load_state_dict({k: v.double() for k, v in cell.state_dict().items()})
    xr = x.detach().double().requires_grad_(True)
    hr = h0.detach().double().requires_grad_(True)
    h, outs = hr, []
    for t in range(T):
        h = ref_cell(xr[t], h)
        outs.append(h)
    ref = torch.stack(outs, 0)
    ref.backward(gout.double())
    want = [ref.detach(), xr.grad, hr.grad] + [p.grad for p in ref_cell.parameters()]
    for a, b in zip(got, want):
        assert close(a.cpu(), b.cpu(), 2e-5)


def test_nested_forks_borrow_distinct_streams():
    """streams.concurrently: the last piece of a fork runs on the parent stream and may fork again;
    the nested fork must not be handed the side streams its siblings are running on"""
    from cpcsv_b200 import streams
    seen = {}

    def leaf(tag):
        seen[tag] = torch.cuda.current_stream().cuda_stream
        return tag

    def inner():
        return streams.concurrently(lambda: leaf("in0"), lambda: leaf("in1"))

    streams.concurrently(lambda: leaf("a"), lambda: leaf("b"), inner)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream().cuda_stream
    assert seen["in1"] == main                       # last piece of the nested fork: the parent
    assert len({seen["a"], seen["b"], seen["in0"], main}) == 4
    assert not any(streams._LENT.values())           # everything given back


def test_planes_stay_current_across_graph_replays_that_update_the_weights():
    """ADVICE r1 (stale packed weights after CUDA-graph replays): an eager generator call, replays of a
    captured graph whose optimiser step changes the weights, an eager call again -- the second eager call
    must see the updated conv / fc weights (PackedAdam rewrites the persistent planes inside the graph)."""
    import harness
    from cpcsv_b200.optim import PackedAdam
    from oracle import functional as Fn
    from oracle import params, presets, synth
    p = presets.get("small")
    dev = torch.device("cuda")
    nets = harness.build_product(p, params.init_all(p, 0), dev)
    G = nets["G"]
    batch = synth.make_batch(p, 1, device=dev)
    x = harness.product_inputs(batch)

    def run():
        feed = synth.NoiseFeed(synth.make_noise(p, 2, device=dev, calls=("images",)))
        harness.inject_noise(G, feed)
        with torch.no_grad():
            return G.sample_images(x["im_motion"], x["im_content"], seg=True)[1]
    img0 = run()
    opt = PackedAdam(list(G.parameters()), lr=torch.tensor(5e-3, device=dev), betas=(0.5, 0.999))
    gen = torch.Generator(device="cuda").manual_seed(5)
    for q in G.parameters():
        q.grad = torch.randn(q.shape, device=dev, generator=gen)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()                                   # eager: allocates the optimiser state
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert float(opt.state[next(iter(G.parameters()))]["step"]) == 4     # 1 eager step + 3 replays
    img1 = run()
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    feed = synth.NoiseFeed(synth.make_noise(p, 2, device=dev, calls=("images",)))
    with torch.no_grad():
        ref = Fn.sample_images(sd, x["im_motion"], x["im_content"], feed, seg=True)[1]
    # no-grad generator calls run single-pass fp16 GEMMs: ~1.5e-3 relative on the images
    assert harness.rel_l2(img1, ref) < 5e-3, harness.rel_l2(img1, ref)
    assert harness.rel_l2(img0, ref) > 1e-2          # the updates really changed the output
    # ... and again AFTER that eager call has re-packed the small plain entries (head weights, ...): replays move
    # the weights without any Python-side trace, the owner of the graph tells the cache (GraphedStep.step does)
    from cpcsv_b200 import nets as knets
    for _ in range(3):
        knets.weight_cache().note_replay()
        graph.replay()
    torch.cuda.synchronize()
    img2 = run()
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    feed = synth.NoiseFeed(synth.make_noise(p, 2, device=dev, calls=("images",)))
    with torch.no_grad():
        ref2 = Fn.sample_images(sd, x["im_motion"], x["im_content"], feed, seg=True)[1]
    assert harness.rel_l2(img2, ref2) < 5e-3, harness.rel_l2(img2, ref2)
    assert harness.rel_l2(img1, ref2) > 1e-2


@pytest.mark.parametrize("Co", [1, 3])
def test_head_gather_tanh(Co):
    """second half of the img / img_seg heads (reference model.py:272-274): 3x3 gather of the per-pixel
    tap products + tanh"""
    N, H, W = 3, 16, 64
    ld = 16 if Co == 1 else 32
    z = rnd(N * H * W, ld, seed=31)
    (g, _), (c, _) = both("head_gather_tanh", [z, N, H, W, Co], dict(y=torch.empty(N, Co, H, W)))
    assert close(g["y"], c["y"], 1e-6)


def test_unpack_wgrad_dot_and_spectral_apply():
    """the fused form of the spectral-norm backward: sum(dW .* W) out of the un-packing pass, then one
    in-place pass -- same result as cpcsv_unpack_conv_wgrad + cpcsv_spectral_bwd"""
    Co, Ci, k = 24, 200, 4
    rp, cp = 64, 256
    dwt = rnd(k * k * rp, cp, seed=61)
    w = rnd(Co, Ci, k, k, seed=62)
    u, v = rnd(Co, seed=63), rnd(Ci * k * k, seed=64)
    sigma = torch.tensor([1.7])
    res = {}
    for dev, mod in (("cuda", ops), ("cpu", emu)):
        g = torch.empty(Co, Ci, k, k, device=dev)
        dot = torch.zeros(1, device=dev)
        mod.unpack_conv_wgrad_dot(dwt.to(dev), rp * cp, cp, 0, None, g, w.to(dev), dot)
        g_ref = torch.empty(Co, Ci, k, k, device=dev)
        mod.unpack_conv_wgrad(dwt.to(dev), rp * cp, cp, 0, None, g_ref)
        assert torch.equal(g, g_ref)
        dw_ref = torch.empty(Co, Ci * k * k, device=dev)
        mod.spectral_bwd(g_ref.view(Co, -1), w.to(dev).view(Co, -1), u.to(dev), v.to(dev), sigma.to(dev), dw_ref,
                         torch.zeros(4, device=dev))
        mod.spectral_bwd_apply(g.view(Co, -1), u.to(dev), v.to(dev), sigma.to(dev), dot, g.view(Co, -1))
        res[dev] = (g.cpu().view(Co, -1), dw_ref.cpu(), float(dot))
    assert close(res["cuda"][0], res["cuda"][1], 1e-5)
    assert close(res["cuda"][0], res["cpu"][0], 1e-5)
    assert abs(res["cuda"][2] - res["cpu"][2]) <= 1e-4 * abs(res["cpu"][2]) + 1e-4


def test_images_to_u8_matches_the_host_formula():
    """device-side image conversion == reference miscc/utils.py:230-235 on the host"""
    import numpy as np
    x = (rnd(3, 70, 130, seed=71) * 0.8).permute(0, 2, 1).contiguous().permute(0, 2, 1)   # strided
    x[0, 0, :4] = torch.tensor([-1.5, -1.0, 1.0, 1.5])
    ref = ((np.clip(x.numpy().transpose(1, 2, 0), -1.0, 1.0) + 1.0) / 2.0 * 255.0).astype("uint8")
    out = torch.empty(70, 130, 3, dtype=torch.uint8, device="cuda")
    ops.images_to_u8(x.cuda(), out)
    assert np.array_equal(out.cpu().numpy(), ref)
    from miscc.outputs import images_to_numpy
    assert np.array_equal(images_to_numpy(x.cuda()), ref)


@pytest.mark.parametrize("C,Co,ldp,n,strided", [(3, 124, 128, 5, False), (1, 124, 128, 3, False), (3, 31, 64, 4, True),
                                                  (1, 8, 64, 2, False)])
def test_first_discriminator_layer_direct_kernel(C, Co, ldp, n, strided):
    """cpcsv_enc0_lrelu_fwd (conv4x4 s2 p1 + 1/sigma + LeakyReLU + hi/lo planes in one launch) against
    torch.nn.functional.conv2d in fp64, and cpcsv_lrelu_bwd16 against the activation's derivative"""
    import torch.nn.functional as F
    dev = torch.device("cuda")
    x = rnd(n, C, 64, 64, seed=1).to(dev)
    if strided:     # frames of a story: [B, C, T, H, W] -> view of frame t
        big = rnd(n, C, 3, 64, 64, seed=2).to(dev)
        x = big[:, :, 1]
    w = (rnd(Co, C, 4, 4, seed=3) * 0.1).to(dev)
    alpha = torch.tensor([0.37], device=dev)
    for dtype, t16 in ((ops.BF16, torch.bfloat16), (ops.FP16, torch.float16)):
        hi = torch.full((n, 32, 32, ldp), 7.0, device=dev, dtype=t16)
        lo = torch.full((n, 32, 32, ldp), 7.0, device=dev, dtype=t16)
        ops.enc0_lrelu_fwd(x, w, alpha, 0.2, hi, lo, ldp, dtype)
        ref = F.leaky_relu(F.conv2d(x.double(), w.double(), stride=2, padding=1) * 0.37, 0.2).permute(0, 2, 3, 1)
        got = hi.double() + lo.double()
        assert float((got[..., :Co] - ref).abs().max()) <= 3e-5 * float(ref.abs().max())
        assert float(got[..., Co:].abs().max()) == 0.0
        # hi plane = the rounded value itself (fp32 vs fp64 accumulation may break a rounding tie differently)
        assert float((hi[..., :Co] != ref.float().to(t16)).float().mean()) < 5e-3
        only_hi = torch.empty_like(hi)
        ops.enc0_lrelu_fwd(x, w, None, 0.2, only_hi, None, ldp, dtype)
        ref1 = F.leaky_relu(F.conv2d(x.double(), w.double(), stride=2, padding=1), 0.2).permute(0, 2, 3, 1)
        assert float((only_hi.double()[..., :Co] - ref1).abs().max()) <= (2 ** -8) * float(ref1.abs().max())
        dy = rnd(n, 32, 32, ldp, seed=5).to(dev)
        dz = torch.empty(n, 32, 32, ldp, device=dev, dtype=torch.bfloat16)
        ops.lrelu_bwd16(dy, hi, 0.2, dz)
        want = (dy * torch.where(hi.float() > 0, 1.0, 0.2)).to(torch.bfloat16)
        assert torch.equal(dz, want)
