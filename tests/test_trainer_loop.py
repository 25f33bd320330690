"""Host logic of the product's training loop (trainer.GANTrainer.train, trainer.GraphedStep;
reference trainer.py:187-485) on the kernel-contract emulator:
loader plumbing, static-buffer loading, the three-segment step used for multi-GPU runs, the
learning-rate halving and the checkpoint files.  The CUDA-graph capture itself is covered by the
``-m gpu`` case in tests/test_zz_cascade.py and by bench.py."""
import copy
import os
import types

import torch

import emulator
import harness
from oracle import params, presets, synth


def _setup(monkeypatch, name="tiny"):
    emulator.install(monkeypatch)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    p = presets.get(name)
    dev = torch.device("cpu")
    nets = harness.build_product(p, params.init_all(p, 0), dev)
    batch = synth.make_batch(p, 1)
    st = {"images": batch["st_real"], "description": batch["st_desc"], "labels": batch["st_labels"]}
    im = {"images": batch["im_real"], "description": batch["im_desc"], "content": batch["im_content"],
          "labels": batch["im_labels"], "images_seg": batch["se_real"]}
    N, B = p["IM_BATCH"], p["ST_BATCH"]
    labels = (torch.ones(N), torch.zeros(N), torch.ones(B), torch.zeros(B))
    return p, nets, st, im, labels


def test_graphed_step_segments_match_whole_step(monkeypatch):
    """the 3-segment step (discriminator stage | D Adam + generator stage | G Adam) that multi-GPU
    runs replay around the NCCL exchanges computes what the one-piece step computes"""
    import trainer
    p, nets, st, im, labels = _setup(monkeypatch)
    nets2 = copy.deepcopy(nets)
    noise = synth.make_noise(p, 2)
    results = []
    for which, nn_ in (("whole", nets), ("segments", nets2)):
        harness.inject_noise(nn_["G"], synth.NoiseFeed(noise))
        opts = trainer.build_optimizers(nn_, fused=False)
        gs = trainer.GraphedStep(nn_, opts, labels, {k: v.clone() for k, v in st.items()},
                                 {k: v.clone() for k, v in im.items()}, grad_sync=None, use_graph=False)
        assert gs.fits(st, im) and not gs.fits({k: v[:1] for k, v in st.items()}, im)
        gs.load(st, im)
        if which == "whole":
            gs.step()
        else:
            gs._seg_d()
            gs._seg_g()
            gs._seg_opt()
        results.append((gs.losses(), {k: [q.detach().clone() for q in n.parameters()] for k, n in nn_.items()}))
    (l1, w1), (l2, w2) = results
    assert set(l1) == set(trainer.LOSS_KEYS)
    for k in l1:
        assert abs(l1[k] - l2[k]) <= 1e-5 * abs(l1[k]) + 1e-7, (k, l1[k], l2[k])
    for k in w1:
        for a, b in zip(w1[k], w2[k]):
            # Adam's first step is lr * sign(g)-like: identical unless a gradient element flips sign
            assert float((a - b).abs().max()) <= 2.5 * 4e-4, k


def test_gan_trainer_loop(monkeypatch, tmp_path):
    """GANTrainer.train with list 'loaders': two epochs of one iteration, LR halving after epoch 1,
    checkpoints written with the reference's file names (miscc/utils.py:323-338)"""
    import trainer
    from miscc.config import cfg
    p, nets, st, im, labels = _setup(monkeypatch)
    cfg.TRAIN.MAX_EPOCH, cfg.TRAIN.SNAPSHOT_INTERVAL, cfg.TRAIN.LR_DECAY_EPOCH = 2, 1, 1
    cfg.NET_G = ""
    t = trainer.GANTrainer.__new__(trainer.GANTrainer)      # __init__ binds a CUDA device
    t.model_dir = str(tmp_path)
    t.image_dir = str(tmp_path / "Image")
    os.makedirs(t.image_dir)
    t.video_len, t.max_epoch, t.snapshot_interval = p["VIDEO_LEN"], 2, 1
    t.imbatch_size, t.stbatch_size, t.ratio = p["IM_BATCH"], p["ST_BATCH"], 1.0
    t.con_ckpt, t.device = None, torch.device("cpu")
    logged = []
    sheets = []
    t._logger = types.SimpleNamespace(add_scalar=lambda k, v, s: logged.append((k, v, s)),
                                      add_image=lambda tag, img, ep: sheets.append((tag, img.shape, ep)))
    captured = {}
    orig = trainer.build_optimizers

    def spy(nets_, fused=None):
        captured["opts"] = orig(nets_, fused=False)
        return captured["opts"]
    monkeypatch.setattr(trainer, "build_optimizers", spy)
    B, V = p["ST_BATCH"], p["VIDEO_LEN"]
    st_loader = [dict(st, text=[["frame %d of story %d" % (f, b) for b in range(B)] for f in range(V)])]
    im_loader = [dict(im, text=["an image"])]
    out_nets = t.train(im_loader, st_loader, None)
    assert set(out_nets) == {"G", "D_im", "D_st", "D_se"}
    # halved once (after epoch 1) for G, D_st, D_im; the reference never halves D_se (trainer.py:452-455)
    lr = {k: o.param_groups[0]["lr"] for k, o in captured["opts"].items()}
    assert abs(lr["G"] - 0.5 * p["GENERATOR_LR"]) < 1e-12
    assert abs(lr["D_im"] - 0.5 * p["DISCRIMINATOR_LR"]) < 1e-12 and abs(lr["D_st"] - lr["D_im"]) < 1e-12
    assert abs(lr["D_se"] - p["DISCRIMINATOR_LR"]) < 1e-12
    files = set(os.listdir(tmp_path))
    assert {"netG_epoch_0.pth", "netG_epoch_1.pth", "netG_epoch_2.pth", "netD_im_epoch_last.pth",
            "netD_st_epoch_last.pth", "netD_se_epoch_last.pth"} <= files, files
    assert {k for k, _, _ in logged} == set(trainer.LOSS_KEYS) and len(logged) == 2 * len(trainer.LOSS_KEYS)
    # end-of-epoch sample sheets (reference trainer.py:437-444): stories next to the ground truth, masks
    assert [tag for tag, _, _ in sheets] == ["pororo", "segment"] * 2
    h, w = B * (68 + 2) + 2, (V * 66 + 2) + 4      # make_grid of per-story rows, 2-pixel gutters twice
    assert sheets[0][1] == (3, h, 2 * w) and sheets[1][1] == (3, h, w)
    assert os.path.exists(os.path.join(t.image_dir, "fake_samples_1.txt"))
    sd = torch.load(os.path.join(tmp_path, "netG_epoch_2.pth"))
    assert set(sd) == set(out_nets["G"].state_dict())


class _FakeGraph:
    """stand-in for torch.cuda.CUDAGraph on a box without a GPU: 'capturing' runs the body once and
    keeps it, 'replay' runs it again -- enough to drive the host logic of GraphedStep.capture / step"""

    def __init__(self):
        self.body = None

    def pool(self):
        return "pool"

    def replay(self):
        self.body()


class _FakeCapture:
    def __init__(self, graph, pool=None, stream=None, **kw):
        self.graph = graph

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def test_graphed_step_capture_paths_host_logic(monkeypatch, tmp_path):
    """capture() + step() of the one-graph and of the three-graph (gradient exchange) variants with
    a fake CUDAGraph: every attribute / call the real capture goes through exists and the segmented
    replay issues the two exchanges between the segments"""
    import torch.distributed as dist
    import trainer
    p, nets, st, im, labels = _setup(monkeypatch)
    dist.init_process_group("gloo", init_method="file://%s" % (tmp_path / "pg"), rank=0, world_size=1)
    try:
        _capture_paths(monkeypatch, trainer, p, nets, st, im, labels)
    finally:
        dist.destroy_process_group()


def _capture_paths(monkeypatch, trainer, p, nets, st, im, labels):
    monkeypatch.setattr(trainer, "step_stream", lambda device=None: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", _FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", _FakeCapture)
    noise = synth.make_noise(p, 2)
    for exchange, segmented in ((False, False), (True, False), (True, True)):
        nn_ = copy.deepcopy(nets)
        opts = trainer.build_optimizers(nn_, fused=False)
        calls = []
        sync = trainer.GradSync(enabled=exchange)
        orig_call = sync.__call__

        class Spy:
            enabled = exchange

            def __call__(self, params, inplace=False):
                calls.append((len(list(params)), inplace))
                return orig_call(params, inplace=inplace)
        gs = trainer.GraphedStep(nn_, opts, labels, {k: v.clone() for k, v in st.items()},
                                 {k: v.clone() for k, v in im.items()}, grad_sync=Spy() if exchange else None,
                                 segmented=segmented)
        assert gs.segmented == segmented
        bodies = []

        def run_capture():
            # the fake context cannot intercept the body: record it by wrapping the segment functions
            if gs.segmented:
                for name in ("_seg_d", "_seg_g", "_seg_opt"):
                    fn = getattr(gs, name)
                    bodies.append(fn)
            else:
                bodies.append(gs._step_body)
            harness.inject_noise(nn_["G"], synth.NoiseFeed(noise))
            gs.capture()
        run_capture()
        if gs.segmented:
            assert len(gs.graphs) == 3
            for g, fn in zip(gs.graphs, bodies):
                g.body = fn
        else:
            gs.graph.body = bodies[0]
        harness.inject_noise(nn_["G"], synth.NoiseFeed(noise))
        calls.clear()
        gs.load(st, im)
        gs.step()
        losses = gs.losses()
        assert all(v == v for v in losses.values())          # finite, not NaN
        if exchange:
            # D_se, D_im, D_st (each between its backward pass and its Adam step), G after the generator stage
            assert len(calls) == 4, calls
            # segmented: replayed optimiser steps read the captured gradient memory, the exchange between the
            # graphs must be in place; whole graph: the exchange is part of the captured step
            assert all(inplace == segmented for _, inplace in calls), calls
        gs.close()
        assert gs.graph is None


def test_snapshot_model_copy_is_importable_under_another_name(tmp_path):
    """reference trainer.py:55-61 + inference.py:61-68: the run directory gets model.py (the cascade
    file when cfg.CASCADE_MODEL) and a copy imported under another module name still builds the generator"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "cpcstoryvisualization-pytorch_b200")
    code = r'''
import importlib, os, shutil, sys
sys.path[:0] = [%r, %r]
from miscc.config import cfg
from oracle import presets
import trainer
for cascade in (False, True):
    presets.apply_to_cfg(cfg, presets.get("tiny_cascade" if cascade else "tiny"))
    out = os.path.join(%r, "run%%d" %% cascade)
    os.makedirs(out)
    trainer.snapshot_sources(out)
    assert sorted(os.listdir(out)) == ["model.py", "trainer.py"]
    name = "model_saved_%%d" %% cascade
    shutil.copyfile(os.path.join(out, "model.py"), os.path.join(%r, name + ".py"))
    sys.path.insert(0, %r)
    mod = importlib.import_module(name)
    G = mod.StoryGAN(cfg.VIDEO_LEN)
    assert hasattr(G, "presample") == cascade and hasattr(G, "sample_videos")
print("ok")
''' % (pkg, root, str(tmp_path), str(tmp_path), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-3000:]
