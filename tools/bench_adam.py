"""Time the fused Adam + re-layout kernels (csrc/optim.cu) alone on the step's real weight shapes and
print achieved HBM GB/s against the algorithmic bytes (p, g, m, v read; p, m, v and the planes written).

    python tools/bench_adam.py            # CUDA events, L2 flushed between launches
    ncu --set full -k regex:adam_pack -c 8 python tools/bench_adam.py --once
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200"))
from cpcsv_b200 import engine, ops  # noqa: E402

dev = torch.device("cuda")
SHAPES = [  # (name, Cout, Cin, k, geometry)
    ("G.upsample1", 1024, 2048, 3, "up"), ("G.seg_c", 2048, 1024, 3, "s1"), ("G.upsample2", 512, 1024, 3, "up"),
    ("G.upsample4", 128, 256, 3, "up"), ("D.logits", 992, 1481, 3, "s1"), ("D.enc8", 992, 496, 4, "s2"),
    ("D.enc5", 496, 248, 4, "s2"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lr = torch.tensor(1e-4, device=dev)
    bc = torch.ones(2, device=dev)
    hyper = ops.AdamHyper(lr, bc, 0.5, 0.999, 1e-8)
    for name, Co, Ci, k, geom in SHAPES:
        w = torch.randn(Co, Ci, k, k, device=dev) * 0.02
        g, m, v = torch.randn_like(w), torch.zeros_like(w), torch.zeros_like(w)
        _k, fkind, bkind, ntap, _u = engine.CONV_GEOM[geom]
        Cop, Cip = engine.rup(Co, 64), engine.rup(Ci, 64)
        planes = [(fkind, ops.BF16, Cop, Cip, torch.empty(ntap * Cop, Cip, device=dev, dtype=torch.bfloat16),
                   torch.empty(ntap * Cop, Cip, device=dev, dtype=torch.bfloat16)),
                  (bkind, ops.BF16, Cip, Cop, torch.empty(ntap * Cip, Cop, device=dev, dtype=torch.bfloat16), None)]
        if name.startswith("G."):
            planes.append((fkind, ops.FP16, Cop, Cip, torch.empty(ntap * Cop, Cip, device=dev, dtype=torch.float16), None))
        nbytes = w.numel() * 28 + sum(p[4].numel() * 2 * (2 if p[5] is not None else 1) for p in planes)
        reps = 1 if args.once else 5
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.adam_pack_conv(w, g, m, v, planes, hyper)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print("%-14s %9.1f MB  %7.1f us  %6.0f GB/s" % (name, nbytes / 1e6, t * 1e3, nbytes / t / 1e6), flush=True)
    # fc
    C_, K = 2048, 613
    w = torch.randn(C_ * 16, K, device=dev) * 0.02
    g, m, v = torch.randn_like(w), torch.zeros_like(w), torch.zeros_like(w)
    Cp, Kp = engine.rup(C_, 64), engine.rup(K, 64)
    f16 = torch.empty(16 * Cp, Kp, device=dev, dtype=torch.float16)
    hi, lo = (torch.empty(16 * Cp, Kp, device=dev, dtype=torch.bfloat16) for _ in range(2))
    bw = torch.empty(Kp, 16 * Cp, device=dev, dtype=torch.bfloat16)
    nbytes = w.numel() * 28 + 4 * f16.numel() * 2
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.adam_pack_fc(w, g, m, v, C_, 16, Cp, Kp, f16, hi, lo, bw, hyper)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print("%-14s %9.1f MB  %7.1f us  %6.0f GB/s" % ("G.fc", nbytes / 1e6, t * 1e3, nbytes / t / 1e6))


if __name__ == "__main__":
    main()
