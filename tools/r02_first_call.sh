#!/bin/bash
# First GPU call of round 2 (run through gpurun from the repo root, ~6-8 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02_first_call.sh'
# Everything the last session of round 1 could not measure, each step under
# its own timeout, logs in gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 secs=$2; shift 2; echo "== $name" ; timeout "$secs" "$@" > "$O/r02_$name.log" 2>&1; echo "$name rc=$?" | tee -a "$O/r02_summary.log"; }

run gpu_tests      240 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider
run bench          90  python bench.py --steps 20 --warmup 5
run bench_overlap  90  python bench.py --steps 20 --warmup 5 --overlap-io --no-cpu-baseline
run stock_cuda     240 python bench.py --impl stock-cuda --steps 3 --warmup 2
run smoke          120 python -c "import __graft_entry__ as g; g.smoke()"
run timeline       120 python tools/timeline_graph.py gpurun_out/r02_timeline_a.csv
# pair-mode dead-lock: bisect the mix (each run exits 3 on a hang; the watchdog prints the configuration)
for mix in pairs pairs+single pairs+wgrad pairs+small all; do
  CPCSV_PAIR=1 run "pair_${mix//+/_}" 90 python tools/repro_pair_hang.py --mix "$mix" --replays 300 --heat 10
done
tail -n 3 "$O"/r02_gpu_tests.log
grep -h '"metric"' "$O"/r02_bench.log "$O"/r02_bench_overlap.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%.2f ms/step  %.1f stories/s  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
"
grep -h '"impl": "stock-cuda"' "$O"/r02_stock_cuda.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('stock PyTorch fp32: TF32 off %.1f ms/step, TF32 on %.1f ms/step' % (d['tf32_off']['ms_per_step'], d['tf32_on']['ms_per_step']))
"
cat "$O/r02_summary.log"
