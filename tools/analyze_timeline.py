"""Digest of tools/timeline_graph.py's CSV: per-kernel totals, per-stream busy time, how much of
the step has a tensor-core GEMM in flight, and the gaps where nothing runs."""
import csv
import sys
from collections import defaultdict


def union(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    merged = []
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
                merged.append((cur_s, cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
        merged.append((cur_s, cur_e))
    return tot, merged


def main(path, nbins=30):
    rows = []
    for r in csv.DictReader(open(path)):
        rows.append((float(r["start_us"]), float(r["dur_us"]), r["stream"], r["kernel"]))
    span = max(s + d for s, d, _, _ in rows)
    print("records %d  span %.3f ms" % (len(rows), span / 1e3))
    agg = defaultdict(lambda: [0, 0.0])
    for s, d, st, k in rows:
        agg[k][0] += 1
        agg[k][1] += d
    total = sum(v[1] for v in agg.values())
    print("sum of kernel time %.3f ms" % (total / 1e3))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print("%9.3f ms %5.1f%% %6d  %s" % (us / 1e3, 100 * us / total, c, k))
    busy, merged = union([(s, s + d) for s, d, _, _ in rows])
    gemm, gm = union([(s, s + d) for s, d, _, k in rows if "conv_gemm" in k])
    print("any-kernel busy %.3f ms (idle %.3f ms); GEMM in flight %.3f ms; no-GEMM %.3f ms" % (
        busy / 1e3, (span - busy) / 1e3, gemm / 1e3, (span - gemm) / 1e3))
    by_stream = defaultdict(list)
    for s, d, st, k in rows:
        by_stream[st].append((s, s + d))
    for st, iv in sorted(by_stream.items(), key=lambda kv: -len(kv[1])):
        b, _ = union(iv)
        print("  stream %-6s launches %5d busy %.3f ms  [%.2f .. %.2f ms]" % (
            st, len(iv), b / 1e3, min(x[0] for x in iv) / 1e3, max(x[1] for x in iv) / 1e3))
    # timeline in bins: fraction of each bin with a GEMM in flight / with anything in flight,
    # and the top non-GEMM kernel of the bin
    w = span / nbins
    print("bin(ms)   gemm%  busy%  conc  top non-GEMM kernels")
    for b in range(nbins):
        lo, hi = b * w, (b + 1) * w
        def clip(iv):
            return sum(max(0.0, min(e, hi) - max(s, lo)) for s, e in iv)
        g = clip(gm) / w
        a = clip(merged) / w
        tot_k = sum(max(0.0, min(s + d, hi) - max(s, lo)) for s, d, _, _ in rows) / w
        top = defaultdict(float)
        for s, d, _, k in rows:
            if "conv_gemm" in k:
                continue
            o = max(0.0, min(s + d, hi) - max(s, lo))
            if o > 0:
                top[k] += o
        tops = ", ".join("%s %.0f%%" % (k.split("::")[-1][:28], 100 * v / w)
                         for k, v in sorted(top.items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f-%4.1f  %4.0f  %4.0f  %4.2f  %s" % (lo / 1e3, hi / 1e3, 100 * g, 100 * a, tot_k, tops))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
