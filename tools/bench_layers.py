"""Per-layer timing of the tcgen05 GEMM jobs at the cfg/final.yml shapes (N = 90 frames).
Prints executed TFLOP/s per job; development tool, not the contract benchmark."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200"))
from cpcsv_b200 import conv, ops  # noqa: E402

dev = torch.device("cuda")


def planes(shape, n=2):
    hi = torch.randn(*shape, device=dev).to(torch.bfloat16)
    return [hi, hi.clone() if n == 2 else None]


def timeit(job, reps=5):
    ops.conv_gemm(job)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
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
    return ts[len(ts) // 2]


def report(name, job, macs, mmas):
    ms = timeit(job)
    print("%-28s %8.3f ms  %7.1f TF/s executed (%dx MMA)  tiles=%d splits=%d bn=%d" % (
        name, ms, 2 * macs * mmas / ms / 1e9, mmas,
        job.groups * job.n_tiles * (job.splits), job.splits, job.block_n), flush=True)


def up_layer(N, H, Ci, Co, tag):
    x = planes((N, H, H, Ci))
    w = planes((16 * Co, Ci))
    out = torch.empty(N, 2 * H, 2 * H, Co, device=dev)
    macs = N * H * H * 16 * Ci * Co
    report("%s up fwd %dx%d %d->%d" % (tag, H, H, Ci, Co), conv.upconv_fwd(x, w, out), macs, 3)
    x1 = [x[0], None]
    w1 = [w[0], None]
    report("%s up fwd 1-plane" % tag, conv.upconv_fwd(x1, w1, out), macs, 1)
    dz = torch.randn(N, 2 * H, 2 * H, Co, device=dev).to(torch.bfloat16)
    wt = torch.randn(16 * Ci, Co, device=dev).to(torch.bfloat16)
    dx = torch.empty(N, H, H, Ci, device=dev)
    report("%s up dgrad" % tag, conv.upconv_dgrad(dz, wt, dx), macs, 1)
    dwt = torch.empty(16, Co, Ci, device=dev)
    report("%s up wgrad" % tag, conv.upconv_wgrad(dz, x[0], dwt), macs, 1)


def s2_layer(N, H, Ci, Co, tag):
    x = planes((N, H, H, Ci))
    w = planes((16 * Co, Ci))
    out = torch.empty(N, H // 2, H // 2, Co, device=dev)
    macs = N * (H // 2) ** 2 * 16 * Ci * Co
    report("%s s2 fwd %dx%d %d->%d" % (tag, H, H, Ci, Co), conv.conv_s2_fwd(x, w, out), macs, 3)
    dy = torch.randn(N, H // 2, H // 2, Co, device=dev).to(torch.bfloat16)
    wt = torch.randn(16 * Ci, Co, device=dev).to(torch.bfloat16)
    dx = torch.empty(N, H, H, Ci, device=dev)
    report("%s s2 dgrad" % tag, conv.conv_s2_dgrad(dy, wt, dx), macs, 1)
    dwt = torch.empty(16, Co, Ci, device=dev)
    report("%s s2 wgrad" % tag, conv.conv_s2_wgrad(dy, x[0], dwt), macs, 1)


def s1_layer(N, H, Ci, Co, tag):
    x = planes((N, H, H, Ci))
    w = planes((9 * Co, Ci))
    out = torch.empty(N, H, H, Co, device=dev)
    macs = N * H * H * 9 * Ci * Co
    report("%s 3x3 fwd %dx%d %d->%d" % (tag, H, H, Ci, Co), conv.conv_s1_fwd(x, w, out), macs, 3)
    dy = torch.randn(N, H, H, Co, device=dev).to(torch.bfloat16)
    wt = torch.randn(9 * Ci, Co, device=dev).to(torch.bfloat16)
    dx = torch.empty(N, H, H, Ci, device=dev)
    report("%s 3x3 dgrad" % tag, conv.conv_s1_dgrad(dy, wt, dx), macs, 1)
    dwt = torch.empty(9, Co, Ci, device=dev)
    report("%s 3x3 wgrad" % tag, conv.conv_s1_wgrad(dy, x[0], dwt), macs, 1)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 90
    print(torch.cuda.get_device_name(0), "N =", N)
    up_layer(N, 4, 2048, 1024, "up1")
    up_layer(N, 8, 1024, 512, "up2")
    up_layer(N, 16, 512, 256, "up3")
    up_layer(N, 32, 256, 128, "up4")
    up_layer(N, 32, 128, 64, "up4_seg")
    s1_layer(N, 4, 1024, 2048, "seg_c")
    s1_layer(N, 8, 512, 1024, "seg_c1")
    s2_layer(N, 32, 128, 256, "D1")
    s2_layer(N, 16, 256, 512, "D2")
    s2_layer(N, 8, 512, 1024, "D3")
    s1_layer(N, 4, 1536, 1024, "logits")
    a = planes((N, 640))
    b = planes((32768, 640))
    out = torch.empty(N, 32768, device=dev)
    report("fc fwd", conv.gemm_nt(a, b, out), N * 640 * 32768, 3)
    # heads: pixel-major GEMM (N = 9*Co padded) + gather-tanh (alone on the GPU)
    for C, Co, two in ((128, 3, True), (128, 3, False), (64, 1, True), (64, 1, False)):
        hi = torch.randn(N * 64 * 64, C, device=dev).to(torch.bfloat16)
        lo = hi.clone() if two else None
        Np = 32 if Co == 3 else 16
        wp = [torch.randn(Np, C, device=dev).to(torch.bfloat16), None]
        if two:
            wp[1] = wp[0].clone()
        z = torch.empty(N * 64 * 64, Np, device=dev)
        y = torch.empty(N, Co, 64, 64, device=dev)

        def run():
            ops.conv_gemm(conv.gemm_nt([hi, lo], wp, z))
            ops.head_gather_tanh(z, N, 64, 64, Co, y)
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        gb = hi.numel() * 2 * (2 if two else 1) / 1e9
        print("head C=%d Co=%d planes=%d   %.1f us  (%.0f GB/s of operand planes)" % (
            C, Co, 2 if two else 1, ts[2] * 1e3, gb / (ts[2] * 1e-3)))

