"""Split-K sweep of the tcgen05 GEMM jobs at the cfg/final.yml shapes: for every layer/direction
time the job with splits in {1, 2, 3, 4, 6, 8, 12, 16} (zero fill of the output included when
splits > 1, as in the step) and mark what conv.pick_splits chooses.  Development tool."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200"))
from cpcsv_b200 import conv, ops  # noqa: E402

dev = torch.device("cuda")
SWEEP = (1, 2, 3, 4, 6, 8, 12, 16)


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


def timeit(job, reps=7):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops.conv_gemm(job)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        ops.conv_gemm(job)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def sweep(name, job):
    chosen = job.splits
    iters = job.taps_per_group * job.k_blocks if job.mode == 0 else None
    if job.mode == 1:
        n, h, w = job.grid
        iters = 1
        for g, t in zip(job.grid, job.tile):
            iters *= -(-g // t)
    res = []
    for s in SWEEP:
        if s > iters:
            break
        job.splits = s
        res.append((s, timeit(job)))
    best = min(res, key=lambda r: r[1])
    cur = dict(res).get(chosen)
    print("%-26s iters=%5d chosen=%3d (%s us) best=%3d (%.1f us) | %s" % (
        name, iters, chosen, ("%.1f" % cur) if cur else "n/a", best[0], best[1],
        " ".join("%d:%.1f" % r for r in res)), flush=True)


def up_layer(N, H, Ci, Co, tag):
    x2 = [bf(N, H, H, Ci), bf(N, H, H, Ci)]
    w2 = [bf(16 * Co, Ci), bf(16 * Co, Ci)]
    out = torch.empty(N, 2 * H, 2 * H, Co, device=dev)
    sweep(tag + " fwd 2pl", conv.upconv_fwd(x2, w2, out))
    sweep(tag + " fwd 1pl", conv.upconv_fwd([x2[0], None], [w2[0], None], out))
    dz = bf(N, 2 * H, 2 * H, Co)
    dx = torch.empty(N, H, H, Ci, device=dev)
    sweep(tag + " dgrad", conv.upconv_dgrad(dz, bf(16 * Ci, Co), dx))
    sweep(tag + " wgrad", conv.upconv_wgrad(dz, x2[0], torch.empty(16, Co, Ci, device=dev)))


def s2_layer(N, H, Ci, Co, tag):
    x2 = [bf(N, H, H, Ci), bf(N, H, H, Ci)]
    w2 = [bf(16 * Co, Ci), bf(16 * Co, Ci)]
    out = torch.empty(N, H // 2, H // 2, Co, device=dev)
    sweep(tag + " fwd 2pl", conv.conv_s2_fwd(x2, w2, out))
    dy = bf(N, H // 2, H // 2, Co)
    dx = torch.empty(N, H, H, Ci, device=dev)
    sweep(tag + " dgrad", conv.conv_s2_dgrad(dy, bf(16 * Ci, Co), dx))
    sweep(tag + " wgrad", conv.conv_s2_wgrad(dy, x2[0], torch.empty(16, Co, Ci, device=dev)))


def s1_layer(N, H, Ci, Co, tag):
    x2 = [bf(N, H, H, Ci), bf(N, H, H, Ci)]
    w2 = [bf(9 * Co, Ci), bf(9 * Co, Ci)]
    out = torch.empty(N, H, H, Co, device=dev)
    sweep(tag + " fwd 2pl", conv.conv_s1_fwd(x2, w2, out))
    sweep(tag + " fwd 1pl", conv.conv_s1_fwd([x2[0], None], [w2[0], None], out))
    dy = bf(N, H, H, Co)
    dx = torch.empty(N, H, H, Ci, device=dev)
    sweep(tag + " dgrad", conv.conv_s1_dgrad(dy, bf(9 * Ci, Co), dx))
    sweep(tag + " wgrad", conv.conv_s1_wgrad(dy, x2[0], torch.empty(9, Co, Ci, device=dev)))


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 90
    print(torch.cuda.get_device_name(0), "N =", N)
    up_layer(N, 4, 2048, 1024, "up1")
    up_layer(N, 8, 1024, 512, "up2")
    up_layer(N, 16, 512, 256, "up3")
    up_layer(N, 32, 256, 128, "up4")
    up_layer(N, 4, 1024, 512, "up1_seg")
    up_layer(N, 8, 512, 256, "up2_seg")
    up_layer(N, 16, 256, 128, "up3_seg")
    up_layer(N, 32, 128, 64, "up4_seg")
    s1_layer(N, 4, 1024, 2048, "seg_c")
    s1_layer(N, 8, 512, 1024, "seg_c1")
    s2_layer(N, 32, 128, 256, "D1")
    s2_layer(N, 16, 256, 512, "D2")
    s2_layer(N, 8, 512, 1024, "D3")
    s1_layer(N, 4, 1536, 1024, "logits")
    s1_layer(18, 4, 1536, 1024, "logits18")
