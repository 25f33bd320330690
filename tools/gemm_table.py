"""Per-launch table of the tcgen05 GEMM jobs of one train step: the step runs eagerly on ONE
stream behind a long device-side sleep (so the host stays ahead and every event pair brackets
its kernel alone), each launch is timed with CUDA events and grouped by job signature.

    python tools/gemm_table.py [out.txt]
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def signature(job):
    n, h, w = job.grid
    kind = "fprop/dgrad" if job.mode == 0 else "wgrad"
    return "%-11s pl=%d grid=%dx%dx%d groups=%d taps=%d kb=%d nv=%d bn=%d mv=%d splits=%d acc=%d" % (
        kind, job.planes, n, h, w, job.groups, job.taps_per_group, job.k_blocks, job.n_valid, job.block_n,
        job.m_valid, job.splits, int(job.accumulate))


def tiles(job):
    n, h, w = job.grid
    if job.mode == 0:
        m = 1
        for g, t in zip(job.grid, job.tile):
            m *= -(-g // t)
        return job.groups * job.splits * job.n_tiles * m
    return job.groups * job.splits * job.n_tiles * (-(-job.m_valid // 128))


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else None
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    eng = bench.StepEngine(bench.preset_dict(), dev, use_graph=False, grad_sync=None)
    tr = eng.trainer
    tr.CONCURRENT_D = tr.CONCURRENT_G = tr.EARLY_G = tr.streams.ENABLED = False
    from cpcsv_b200 import engine, ops
    engine.WGRAD_ON_AUX_STREAM = False
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    orig = ops.conv_gemm
    recs = []

    def timed(job):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if job.splits > 1 and not job.accumulate:
            job.out.zero_()
            job.accumulate, restore = True, True     # keep the zero fill out of the bracket
        else:
            restore = False
        e0.record()
        orig(job)
        e1.record()
        if restore:
            job.accumulate = False
        recs.append((signature(job), bench.job_flops(job), tiles(job), e0, e1))

    ops.conv_gemm = timed
    torch.cuda._sleep(int(0.4 * 1.9e9))
    eng.step()
    torch.cuda.synchronize()
    ops.conv_gemm = orig
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0])
    for sig, fl, nt, e0, e1 in recs:
        a = agg[sig]
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += fl
        a[3] = nt
    tot_ms = sum(a[1] for a in agg.values())
    tot_fl = sum(a[2] for a in agg.values())
    lines = ["%d launches, %.3f ms, %.1f GFLOP executed, %.1f TFLOP/s average" % (
        len(recs), tot_ms, tot_fl / 1e9, tot_fl / tot_ms / 1e9),
        "%6s %8s %8s %7s %6s  %s" % ("count", "ms_tot", "us_each", "TF/s", "tiles", "signature")]
    for sig, (c, ms, fl, nt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%6d %8.3f %8.1f %7.1f %6d  %s" % (c, ms, 1e3 * ms / c, fl / ms / 1e9, nt, sig))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")
