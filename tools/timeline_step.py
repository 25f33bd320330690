"""Per-kernel GPU time of one eager CP-CSV train step measured with torch.profiler (CUPTI
activity records: warm caches, kernels back to back -- unlike ncu, which serialises and flushes).
A long device-side sleep is queued first so the CPU launch path never starves the GPU."""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    eng = bench.StepEngine(bench.preset_dict(), dev, use_graph=False, grad_sync=None)
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        torch.cuda._sleep(int(0.15 * 1.9e9))
        eng.step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    t_first, t_last = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name
        if "sleep" in name.lower() or "spin" in name.lower():
            continue
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        key = name.replace("(anonymous namespace)::", "").replace("void ", "")
        key = key.split("(")[0].split("<")[0][-70:]
        agg[key][0] += 1
        agg[key][1] += dur
        st = ev.time_range.start
        en = ev.time_range.end
        t_first = st if t_first is None else min(t_first, st)
        t_last = en if t_last is None else max(t_last, en)
    total = sum(v[1] for v in agg.values())
    print("kernel time %.3f ms over %d launches; span first->last %.3f ms" % (
        total / 1e3, sum(v[0] for v in agg.values()), (t_last - t_first) / 1e3))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%9.3f ms %5.1f%% %6d  %s" % (us / 1e3, 100 * us / total, c, k))
