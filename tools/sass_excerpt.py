"""profiles/<round>_sass_conv_gemm.txt: the tcgen05 / TMEM / TMA instructions of conv_gemm_kernel in the built
libcpcsv.so (cuobjdump -sass), with per-function instruction counts.

    python tools/sass_excerpt.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cpcstoryvisualization-pytorch_b200", "cpcsv_b200", "libcpcsv.so")
KEEP = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS", "UTMACCTL", "SYNCS", "UCGABAR", "ATOMS", "BAR.SYNC",
        "UTMAPF", "REDG", "RED.")


def main(out=None):
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    lines = ["# cuobjdump -sass excerpt of libcpcsv.so (sm_100a), conv_gemm_kernel<false> / <true>: the tcgen05 / TMEM / "
             "TMA", "# instructions (UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld, UTCBAR = "
             "tcgen05.commit,", "# UTCATOMSWS = tcgen05.alloc / dealloc / relinquish, SYNCS = mbarrier ops, UCGABAR = "
             "cluster barrier).", "# Regenerate: python tools/sass_excerpt.py profiles/<name>.txt", ""]
    fn, body = None, []

    def flush():
        if fn and "conv_gemm_kernel" in fn:
            ops = collections.Counter()
            kept = []
            for ln in body:
                m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
                if not m:
                    continue
                ins = m.group(1).strip()
                op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
                if any(op.startswith(k) for k in KEEP):
                    ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("BAR") else "")] += 1
                    kept.append("    " + ins)
            lines.append("Function : " + fn)
            lines.append("  instruction counts: " + ", ".join("%s x%d" % kv for kv in sorted(ops.items())))
            seen = set()
            for k in kept:          # each distinct instruction form once, in first-occurrence order
                key = re.sub(r"\b(R|UR|P|UP)\d+\b", r"\1n", k)
                key = re.sub(r"0x[0-9a-f]+", "0x..", key)
                if key not in seen:
                    seen.add(key)
                    lines.append(k)
            lines.append("")

    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            fn, body = m.group(1), []
        else:
            body.append(ln)
    flush()
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text + "\n")
    print(text[:3000])


if __name__ == "__main__":
    main(*sys.argv[1:])
