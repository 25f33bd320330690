"""One line per launch of OUR kernels from an .ncu-rep (ncu --set full): duration, DRAM traffic, DRAM / L2 / SM
throughput, tensor-pipe activity, grid, registers, dynamic shared memory.

    python tools/ncu_compact.py rep.ncu-rep [out.txt] [label ...]
"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "dur us", 1e-3), ("dram__bytes_read.sum", "rd MB", None),
        ("dram__bytes_write.sum", "wr MB", None), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1), ("launch__grid_size", "grid", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__shared_mem_per_block_dynamic", "dsmem KB", None)]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(rep, out=None, *labels):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    iname = hdr.index("Kernel Name")
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# %s: ncu --set full --clock-control none of tools/ncu_targets.py, the repo's own kernels only" % rep.split("/")[-1],
             "%-34s " % "kernel" + " ".join("%9s" % c[1] for c in COLS) + "  what"]
    k = 0
    for r in data:
        name = r[iname]
        if "cpcsv" not in name:
            continue
        short = name.split("cpcsv::", 1)[1].replace("<unnamed>::", "").split("(")[0][:34]
        vals = []
        for key, _t, scale in COLS:
            i = idx.get(key)
            v = r[i] if i is not None else ""
            try:
                f = float(v.replace(",", ""))
                if "bytes" in key:
                    f *= UNIT.get(units[i], 1.0)
                elif key.endswith("dynamic"):
                    f *= {"byte": 1 / 1024, "Kbyte": 1.0}.get(units[i], 1.0)
                elif key == "gpu__time_duration.sum":
                    f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1e-3)
                vals.append("%9.1f" % f)
            except ValueError:
                vals.append("%9s" % "-")
        lines.append("%-34s " % short + " ".join(vals) + "  " + (labels[k] if k < len(labels) else ""))
        k += 1
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:])
