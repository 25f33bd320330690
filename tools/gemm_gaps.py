"""List the intervals of a step timeline (tools/timeline_graph.py CSV) in which no tensor-core GEMM is
in flight, with the kernels that run in them: what the step is waiting for when the tensor cores idle.
    python tools/gemm_gaps.py gpurun_out/timeline.csv [min_gap_us]
"""
import csv
import sys


def main(path, min_gap=100.0):
    rows = []
    for r in csv.DictReader(open(path)):
        rows.append((float(r["start_us"]), float(r["dur_us"]), r["stream"], r["kernel"]))
    gem = sorted((s, s + d) for s, d, st, k in rows if "conv_gemm" in k)
    m = []
    for a, b in gem:
        if m and a <= m[-1][1]:
            m[-1][1] = max(m[-1][1], b)
        else:
            m.append([a, b])
    gaps, prev = [], 0.0
    for a, b in m:
        if a - prev > 30:
            gaps.append((prev, a))
        prev = b
    end = max(s + d for s, d, _, _ in rows)
    gaps.append((prev, end))
    print("span %.2f ms; GEMM in flight %.2f ms; gaps > 30 us: n=%d total=%.2f ms" % (
        end / 1e3, sum(b - a for a, b in m) / 1e3, len(gaps), sum(b - a for a, b in gaps) / 1e3))
    for a, b in gaps:
        if b - a < min_gap:
            continue
        ks = {}
        for s, d, st, k in rows:
            o = min(b, s + d) - max(a, s)
            if o > 0:
                ks[k] = ks.get(k, 0) + o
        top = sorted(ks.items(), key=lambda x: -x[1])[:5]
        print("%6.2f-%6.2f ms (%4.0f us): " % (a / 1e3, b / 1e3, b - a) + ", ".join(
            "%s %.0f" % (k.split("::")[-1][-26:], v) for k, v in top))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 100.0)
