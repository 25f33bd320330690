"""cta_group::2 pairs vs single CTAs for every fprop / dgrad job shape of the step (median of 9
runs each, L2 flushed, zero fill of split outputs included).  Development tool behind the
`conv.use_pair` rule."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cpcsv_b200 import conv, ops  # noqa: E402
from sweep_splits import bf, timeit  # noqa: E402

dev = torch.device("cuda")


def ab(name, job):
    res = {}
    for pair in (False, True):
        job.pair = pair
        res[pair] = timeit(job, reps=9)
    n, h, w = job.grid
    mt = -(-n * h * w // 128)
    iters = job.taps_per_group * job.k_blocks // job.splits
    flops = 2.0 * n * h * w * job.groups * job.taps_per_group * job.k_blocks * 64 * job.n_valid * (
        3 if job.planes == 2 else 1)
    print("%-22s pl=%d groups=%d taps=%2d kb=%2d nv=%4d bn=%3d splits=%2d mtiles=%4d iters=%3d gflop=%6.1f | "
          "single %6.1f us  pair %6.1f us  ratio %.2f" % (
              name, job.planes, job.groups, job.taps_per_group, job.k_blocks, job.n_valid, job.block_n,
              job.splits, mt, iters, flops / 1e9, res[False], res[True], res[True] / res[False]), flush=True)


def up_layer(N, H, Ci, Co, tag):
    x2 = [bf(N, H, H, Ci), bf(N, H, H, Ci)]
    w2 = [bf(16 * Co, Ci), bf(16 * Co, Ci)]
    out = torch.empty(N, 2 * H, 2 * H, Co, device=dev)
    ab(tag + " fwd 2pl", conv.upconv_fwd(x2, w2, out))
    ab(tag + " fwd 1pl", conv.upconv_fwd([x2[0], None], [w2[0], None], out))
    ab(tag + " dgrad", conv.upconv_dgrad(bf(N, 2 * H, 2 * H, Co), bf(16 * Ci, Co), torch.empty(N, H, H, Ci, device=dev)))


def s2_layer(N, H, Ci, Co, tag):
    x2 = [bf(N, H, H, Ci), bf(N, H, H, Ci)]
    w2 = [bf(16 * Co, Ci), bf(16 * Co, Ci)]
    ab(tag + " fwd 2pl", conv.conv_s2_fwd(x2, w2, torch.empty(N, H // 2, H // 2, Co, device=dev)))
    ab(tag + " dgrad", conv.conv_s2_dgrad(bf(N, H // 2, H // 2, Co), bf(16 * Ci, Co), torch.empty(N, H, H, Ci, device=dev)))


def s1_layer(N, H, Ci, Co, tag):
    x2 = [bf(N, H, H, Ci), bf(N, H, H, Ci)]
    w2 = [bf(9 * Co, Ci), bf(9 * Co, Ci)]
    out = torch.empty(N, H, H, Co, device=dev)
    ab(tag + " fwd 2pl", conv.conv_s1_fwd(x2, w2, out))
    ab(tag + " fwd 1pl", conv.conv_s1_fwd([x2[0], None], [w2[0], None], out))
    ab(tag + " dgrad", conv.conv_s1_dgrad(bf(N, H, H, Co), bf(9 * Ci, Co), torch.empty(N, H, H, Ci, device=dev)))


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
