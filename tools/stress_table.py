"""Per-layer tensor-pipe table of the stress-512 step (BASELINE.json configs[4]): joins the ncu launch list of
the conv_gemm launches (gpu__time_duration, sm__pipe_tensor_cycles_active) with bench.py's --job-log (one line
per launch, same order) and groups by job signature.

    python tools/stress_table.py ncu.csv jobs.log SKIP [out.txt]
"""
import csv
import sys
from collections import OrderedDict


def read_ncu(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Metric Name" in r and "Metric Value" in r)
    H = rows[hdr]
    iid, iname, ival = H.index("ID"), H.index("Metric Name"), H.index("Metric Value")
    iunit = H.index("Metric Unit")
    per = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= ival:
            continue
        v = float(r[ival].replace(",", ""))
        if r[iname] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iunit], 1e-3)      # -> us
        per.setdefault(int(r[iid]), {})[r[iname]] = v
    return list(per.values())


def main(ncu_csv, job_log, skip, out=None):
    launches = read_ncu(ncu_csv)
    jobs = [ln.rstrip("\n").split("\t") for ln in open(job_log)][int(skip):int(skip) + len(launches)]
    assert len(jobs) == len(launches), (len(jobs), len(launches))
    agg = OrderedDict()
    for (sig, fl), m in zip(jobs, launches):
        a = agg.setdefault(sig, [0, 0.0, 0.0, 0.0])
        us = m["gpu__time_duration.sum"]
        a[0] += 1
        a[1] += us
        a[2] += float(fl)
        a[3] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * us
    tot_us = sum(a[1] for a in agg.values())
    tot_fl = sum(a[2] for a in agg.values())
    lines = ["# stress-512 step (512 stories + 2560 images), conv_gemm launches of one step under ncu (--clock-control none;",
             "# per-launch times are serialised and cold-cache).  tensor%% = sm__pipe_tensor_cycles_active (time-weighted).",
             "%d launches, %.2f ms, %.1f TFLOP executed, %.0f TFLOP/s, tensor pipe %.1f %% active" % (
                 len(launches), tot_us / 1e3, tot_fl / 1e12, tot_fl / tot_us / 1e6,
                 sum(a[3] for a in agg.values()) / tot_us),
             "%5s %9s %8s %7s %8s  %s" % ("count", "ms_total", "us_each", "TF/s", "tensor%", "signature")]
    for sig, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%5d %9.3f %8.1f %7.0f %8.1f  %s" % (a[0], a[1] / 1e3, a[1] / a[0], a[2] / a[1] / 1e6, a[3] / a[1], sig))
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:])
