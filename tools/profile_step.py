"""One eager CP-CSV train step between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum ...` launch lists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    p = bench.preset_dict()
    eng = bench.StepEngine(p, dev, use_graph=False, grad_sync=None)
    for _ in range(2):
        eng.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one step")
