"""Per-launch (start, duration, stream, kernel) records of ONE replay of the whole-step CUDA
graph, taken with torch.profiler (CUPTI concurrent-kernel activity records keep the
cross-stream overlap).  Writes a CSV that tools/analyze_timeline.py digests offline.

    python tools/timeline_graph.py gpurun_out/timeline_graph.csv
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def short(name):
    key = name.replace("(anonymous namespace)::", "").replace("void ", "")
    return key.split("(")[0].split("<")[0][-60:]


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_graph.csv"
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    eng = bench.StepEngine(bench.preset_dict(), dev, use_graph=True, grad_sync=None)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            eng.step()
    torch.cuda.synchronize()
    eng.capture()
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start,
                     getattr(ev, "device_resource_id", -1), short(ev.name)))
    rows.sort()
    t0 = rows[0][0]
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        f.write("start_us,dur_us,stream,kernel\n")
        for s, d, st, n in rows:
            f.write("%.3f,%.3f,%s,%s\n" % (s - t0, d, st, n))
    print("wrote %d records, span %.3f ms" % (len(rows), (rows[-1][0] + rows[-1][1] - t0) / 1e3))
