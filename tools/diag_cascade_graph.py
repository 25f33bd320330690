"""diagnostic: losses of three graph-replayed steps (zero learning rates, injected noise, three batches) from
independent engines -- serial loads twice, staged load_async once -- for a preset; prints the per-step losses"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200"))
import harness  # noqa: E402
from oracle import params, presets, synth  # noqa: E402


def main(name, eager_flag):
    import trainer
    for kv in filter(None, os.environ.get("DIAG_FLAGS", "").split(",")):      # e.g. EARLY_D_REAL=0,CONCURRENT_D=0
        k, v = kv.split("=")
        setattr(trainer, k, bool(int(v)))
    import miscc.utils as mu
    for kv in filter(None, os.environ.get("DIAG_UTILS_FLAGS", "").split(",")):      # e.g. PARALLEL_PASSES=0
        k, v = kv.split("=")
        setattr(mu, k, bool(int(v)))
    p = presets.get(name)
    dev = torch.device("cuda")
    base = harness.build_product(p, params.init_all(p, 0), dev)
    noise = synth.make_noise(p, 2, device=dev)

    def host_batch(seed):
        b = synth.make_batch(p, seed)
        st = {"images": b["st_real"], "description": b["st_desc"], "labels": b["st_labels"]}
        im = {"images": b["im_real"], "description": b["im_desc"], "content": b["im_content"],
              "labels": b["im_labels"], "images_seg": b["se_real"]}
        return ({k: v.pin_memory() for k, v in st.items()}, {k: v.pin_memory() for k, v in im.items()})
    batches = [host_batch(s) for s in (11, 12, 13, 14)]
    N, B = p["IM_BATCH"], p["ST_BATCH"]
    labels = (torch.ones(N, device=dev), torch.zeros(N, device=dev), torch.ones(B, device=dev), torch.zeros(B, device=dev))
    for mode in ("serial", "serial", "staged", "eager" if eager_flag else "serial"):
        nets = copy.deepcopy(base)
        opts = trainer.build_capturable_optimizers(nets, dev)
        for o in opts.values():
            trainer.set_lr(o, 0.0)
        st0, im0 = batches[0]
        gs = trainer.GraphedStep(nets, opts, labels, {k: v.to(dev) for k, v in st0.items()},
                                 {k: v.to(dev) for k, v in im0.items()}, grad_sync=None)
        for _ in range(2):
            harness.inject_noise(nets["G"], synth.NoiseFeed(noise))
            gs.step()
        harness.inject_noise(nets["G"], synth.NoiseFeed(noise))
        if mode != "eager":
            gs.capture()
        out = []
        if mode in ("serial", "eager"):
            for st, im in batches[1:]:
                if mode == "eager":
                    harness.inject_noise(nets["G"], synth.NoiseFeed(noise))
                gs.load(st, im)
                gs.step()
                out.append(gs.losses())
        else:
            gs.load(*batches[1])
            for i in range(1, 4):
                gs.step()
                if i + 1 < 4:
                    gs.load_async(*batches[i + 1])
                out.append(gs.losses())
        torch.cuda.synchronize()
        for i, o in enumerate(out):
            print(name, mode, i, " ".join("%s=%.5f" % (k, v) for k, v in o.items()), flush=True)


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
