"""Peak device memory of the train step per stage, for batch-scaling analysis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

if __name__ == "__main__":
    st, im = int(sys.argv[1]), int(sys.argv[2])
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    eng = bench.StepEngine(bench.preset_dict(st, im), dev, use_graph=False, grad_sync=None)
    tr = eng.trainer
    def gb(): return torch.cuda.max_memory_allocated() / 2**30, torch.cuda.memory_allocated() / 2**30
    torch.cuda.synchronize(); print("after build: peak %.1f live %.1f GB" % gb())
    for it in range(2):
        torch.cuda.reset_peak_memory_stats()
        x = tr.prepare_inputs(eng.dev_st, eng.dev_im)
        with torch.no_grad():
            a = eng.nets["G"].sample_videos(x["st_motion"], x["st_content"])
            torch.cuda.synchronize(); print("  p1 videos: peak %.1f live %.1f" % gb())
            b = eng.nets["G"].sample_images(x["im_motion"], x["im_content"], seg=True)
            torch.cuda.synchronize(); print("  p1 images: peak %.1f live %.1f" % gb())
        del a, b
        out = tr.stage_discriminators(eng.nets, x, eng.labels)
        torch.cuda.synchronize(); print("  stage D: peak %.1f live %.1f" % gb())
        for k in tr.D_NETS: eng.opts[k].step()
        out.update(tr.stage_generator(eng.nets, x, eng.labels, 1.0))
        torch.cuda.synchronize(); print("  stage G: peak %.1f live %.1f" % gb())
        eng.opts["G"].step()
        del out
