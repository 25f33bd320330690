"""Aggregate an ncu `--csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"<.*", "", name)
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<.*", "", name)
        agg[name][0] += 1
        agg[name][1] += ns
        total += ns
    print("total %.3f ms over %d launches" % (total / 1e6, sum(a[0] for a in agg.values())))
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%8.3f ms %5.1f%% %6d  %s" % (ns / 1e6, 100 * ns / total, cnt, name[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
