"""Per-shape lower bounds for the GEMM launches of a step, from a table written by tools/gemm_table.py
(no GPU needed): for every job signature the time it would take if it were bound only by

  * the tensor pipe       : executed FLOPs / measured sustained bf16 peak (MEASURED_PEAKS.json),
  * the L2 -> SM operand feed: operand bytes every CTA pulls through L2 per k-step, at the LTS cap
                               ncu showed the kernel saturating (12.4 TB/s, profiles/r01_ncu_full_v2_*),
  * wave quantisation     : tiles / (ceil(tiles / 148) * 148) on the persistent grid,

and the gap between that bound and the measured time, ranked.  Usage:
    python tools/gemm_bounds.py profiles/r01_gemm_table_c_pair_vs_single.txt ["single CTAs"]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS = 148
LTS_CAP = 12.4e12          # B/s, L2 -> SM (ncu lts__t_bytes cap observed on B200)


def parse(path, section):
    rows, on = [], section is None
    for line in open(path):
        if line.startswith("#"):
            on = section is None or section in line
            continue
        m = re.match(r"\s*(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)\s+(\S+)\s+pl=(\d) grid=(\d+)x(\d+)x(\d+) "
                     r"groups=(\d+) taps=(\d+) kb=(\d+) nv=(\d+) bn=(\d+) mv=(\d+) splits=(\d+)", line)
        if on and m:
            g = m.groups()
            rows.append(dict(count=int(g[0]), ms_tot=float(g[1]), us=float(g[2]), tfs=float(g[3]), tiles=int(g[4]),
                             kind=g[5], pl=int(g[6]), grid=(int(g[7]), int(g[8]), int(g[9])), groups=int(g[10]),
                             taps=int(g[11]), kb=int(g[12]), nv=int(g[13]), bn=int(g[14]), mv=int(g[15]),
                             splits=int(g[16])))
    return rows


def bound_us(r, peak_tfs):
    flops = r["tfs"] * 1e12 * r["us"] * 1e-6                       # executed FLOPs of one launch
    t_tensor = flops / (peak_tfs * 1e12)
    if r["kind"] == "wgrad":
        # per k-step of 64 pixels a CTA loads [64 x 128] of A and [64 x bn] of B (bf16)
        bytes_per_mac = (128 + r["bn"]) * 64 * 2 / (128.0 * r["bn"] * 64)
    else:
        # per k-step of 64 channels: planes x ([128 x 64] of A + [bn x 64] of B); planes=2 issues 3 MMAs
        macs = 128.0 * r["bn"] * 64 * (3 if r["pl"] == 2 else 1)
        bytes_per_mac = r["pl"] * (128 + r["bn"]) * 64 * 2 / macs
    t_feed = (flops / 2.0) * bytes_per_mac / LTS_CAP
    wave_eff = r["tiles"] / float(-(-r["tiles"] // SMS) * SMS)
    return max(t_tensor, t_feed) / wave_eff * 1e6, t_tensor * 1e6, t_feed * 1e6, wave_eff


def main():
    path = sys.argv[1]
    section = sys.argv[2] if len(sys.argv) > 2 else None
    peak = 1368.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    rows = parse(path, section)
    tot_meas = tot_bound = 0.0
    out = []
    for r in rows:
        b, tt, tf, we = bound_us(r, peak)
        gap = (r["us"] - b) * r["count"] * 1e-3
        tot_meas += r["ms_tot"]
        tot_bound += b * r["count"] * 1e-3
        out.append((gap, r, b, tt, tf, we))
    out.sort(key=lambda x: -x[0])
    print("# %s%s: %d signatures, measured %.2f ms, sum of per-shape bounds %.2f ms (peak %.1f TFLOP/s, LTS cap %.1f TB/s)"
          % (os.path.basename(path), " [%s]" % section if section else "", len(rows), tot_meas, tot_bound, peak,
             LTS_CAP / 1e12))
    print("# gap_ms count  us_meas us_bound (tensor  feed  wave_eff)  signature")
    for gap, r, b, tt, tf, we in out:
        print("%7.3f %5d %8.1f %8.1f (%6.1f %6.1f %5.2f)  %s pl=%d grid=%dx%dx%d groups=%d taps=%d kb=%d nv=%d bn=%d splits=%d"
              % (gap, r["count"], r["us"], b, tt, tf, we, r["kind"], r["pl"], r["grid"][0], r["grid"][1], r["grid"][2],
                 r["groups"], r["taps"], r["kb"], r["nv"], r["bn"], r["splits"]))


if __name__ == "__main__":
    main()
