"""The two large-batch BASELINE.json configurations that are not the contract bench line:
  configs[3]  inference.py-style story generation (generator + segmentation branch, no_grad,
              train-mode BN as the reference does), batch sweep 16..1024 stories
  configs[4]  large-batch stress: 512 stories + 2560 images per step on one GPU (train step)
Prints one JSON line per measurement.  Development / evidence tool (profiles/), not bench.py."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

PEAK = 1368.4


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def inference_sweep(batches=(16, 64, 256, 1024)):
    from miscc.config import cfg
    p = bench.preset_dict()
    bench.apply_cfg(cfg, p)
    import trainer
    dev = torch.device("cuda", 0)
    nets = trainer.build_networks(5)
    G = nets["G"].to(dev).train()
    for B in batches:
        motion = torch.randn(B, 5, 365, device=dev)
        content = torch.randn(B, 5, 356, device=dev)

        def run():
            with torch.no_grad():
                G.sample_videos(motion, content, seg=True)
        g = torch.cuda.CUDAGraph()
        run(); run()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            run()
        ms = timed(g.replay, 3, 10)
        gf = 66.9 * B      # GFLOP per 5-frame story, SURVEY.md section 8(d)
        print(json.dumps({"config": "inference", "stories": B, "ms": ms, "stories_per_s": B / ms * 1e3,
                          "nominal_tflops": gf / ms, "frac_of_peak": gf / ms / PEAK,
                          "precision": "single-pass fp16 (no-grad path)"}), flush=True)
        del g
    del G, nets


def stress(st_batch=512, im_batch=2560):
    from cpcsv_b200 import _lib
    p = bench.preset_dict(st_batch, im_batch)
    dev = torch.device("cuda", 0)
    eng = bench.StepEngine(p, dev, use_graph=False, grad_sync=None)
    t0 = time.time()
    ms = timed(eng.step, 2, 3)
    gf = 659.7 * st_batch
    print(json.dumps({"config": "stress", "stories": st_batch, "images": im_batch, "ms_per_step": ms,
                      "stories_per_s": st_batch / ms * 1e3, "nominal_tflops": gf / ms,
                      "frac_of_peak_nominal": gf / ms / PEAK,
                      "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "wall_s": time.time() - t0, "launches": _lib.launch_count()}), flush=True)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("inference", "all"):
        inference_sweep()
    if what in ("stress", "all"):
        torch.cuda.empty_cache()
        stress(int(sys.argv[2]) if len(sys.argv) > 2 else 512, int(sys.argv[3]) if len(sys.argv) > 3 else 2560)
