"""A handful of representative launches for `ncu --set full` (keep it short: ncu replays every
captured launch ~40 times): the direct head kernel and tcgen05 GEMM jobs of the main classes at
the cfg/final.yml shapes.

    ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|head_conv' \
        -o gpurun_out/prof python tools/ncu_targets.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200"))
from cpcsv_b200 import conv, ops  # noqa: E402

dev = torch.device("cuda")
N = 90


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "head"):
        hi, lo = bf(N * 64 * 64, 128), bf(N * 64 * 64, 128)
        wp = [bf(32, 128), bf(32, 128)]
        z = torch.empty(N * 64 * 64, 32, device=dev)
        y = torch.empty(N, 3, 64, 64, device=dev)
        ops.conv_gemm(conv.gemm_nt([hi, lo], wp, z))                 # launch 0: img head GEMM, 2 planes
        ops.head_gather_tanh(z, N, 64, 64, 3, y)                     # launch 0b: gather + tanh
    if which in ("all", "gemm"):
        # launch 1: upsample3 forward, hi/lo split operands (3 MMAs / k-step)
        x2 = [bf(N, 16, 16, 512), bf(N, 16, 16, 512)]
        w2 = [bf(16 * 256, 512), bf(16 * 256, 512)]
        out = torch.empty(N, 32, 32, 256, device=dev)
        ops.conv_gemm(conv.upconv_fwd(x2, w2, out))
        # launch 2: the same layer single-plane (no-grad pass)
        ops.conv_gemm(conv.upconv_fwd([x2[0], None], [w2[0], None], out))
        # launch 3: D encoder layer 2 data gradient (4 output parities x 4 taps)
        dy = bf(N, 8, 8, 512)
        dx = torch.empty(N, 16, 16, 256, device=dev)
        ops.conv_gemm(conv.conv_s2_dgrad(dy, bf(16 * 256, 512), dx))
        # launch 4: upsample4 weight gradient (K = 92160 pixels, split-K)
        dz = bf(N, 64, 64, 128)
        ops.conv_gemm(conv.upconv_wgrad(dz, bf(N, 32, 32, 256), torch.empty(16, 128, 256, device=dev)))
        # launch 5: upsample2 weight gradient as cta_group::2 pairs (Co = 512: four 128-channel M tiles)
        dz2 = bf(N, 16, 16, 512)
        ops.conv_gemm(conv.upconv_wgrad(dz2, bf(N, 8, 8, 1024), torch.empty(16, 512, 1024, device=dev)))
    if which in ("all", "enc0"):
        # first discriminator layer: direct conv4x4 s2 + LeakyReLU + hi/lo planes (csrc/enc0.cu)
        x = torch.randn(N, 3, 64, 64, device=dev)
        w = torch.randn(124, 3, 4, 4, device=dev) * 0.05
        hi = torch.empty(N, 32, 32, 128, device=dev, dtype=torch.bfloat16)
        ops.enc0_lrelu_fwd(x, w, None, 0.2, hi, torch.empty_like(hi), 128)
    torch.cuda.synchronize()
