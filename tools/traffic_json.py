"""profiles/<round>_gemm_traffic.json (bench.py quotes it as roofline.traffic) from the ncu CSV of
tools/r02_profile.sh step 2: per-launch dram__bytes_read / dram__bytes_write of every conv_gemm launch.

    python tools/traffic_json.py ncu.csv out.json
"""
import csv
import json
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    h = next(i for i, r in enumerate(rows) if "Metric Name" in r and "Metric Value" in r)
    H = rows[h]
    iid, iname, iunit, ival = H.index("ID"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value")
    per = {}
    for r in rows[h + 1:]:
        if len(r) > ival:
            per.setdefault(int(r[iid]), {})[r[iname]] = float(r[ival].replace(",", "")) * SCALE.get(r[iunit], 1.0)
    n = len(per)
    rd = sum(v["dram__bytes_read.sum"] for v in per.values()) / n
    wr = sum(v["dram__bytes_write.sum"] for v in per.values()) / n
    us = sum(v["gpu__time_duration.sum"] for v in per.values()) / n
    d = {"avg_bytes_per_launch": rd + wr, "avg_read_bytes": rd, "avg_write_bytes": wr, "launches": n,
         "avg_duration_us_under_ncu": us,
         "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 "-k regex:conv_gemm over the %d conv_gemm launches of two eager steps (bench.py --no-graph; "
                 "tools/r02_profile.sh step 2; tools/traffic_json.py; durations are cold-cache / serialised).  "
                 "Algorithmic bytes per launch (operand planes + fp32 output once, from the job descriptors) are "
                 "reported next to it by bench.py as roofline.algorithmic_bytes_per_launch." % n}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
