"""Compact per-launch table of the metrics the roofline discussion uses, from an .ncu-rep."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    sel = [i for i, h in enumerate(hdr) if h in KEYS or ("pipe_tensor" in h and "realtime" in h and "hmma" in h)
           or h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")]
    lines = ["%-86s %-16s %s" % ("metric", "unit", " | ".join("launch %d" % i for i in range(len(data))))]
    for i in sel:
        lines.append("%-86s %-16s %s" % (hdr[i][:86], units[i], " | ".join(r[i][:16] for r in data)))
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:])
