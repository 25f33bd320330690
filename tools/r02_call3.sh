#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-c3}
run() { local name=$1 secs=$2; shift 2; echo "== $name" ; timeout "$secs" "$@" > "$O/r02_${T}_$name.log" 2>&1; echo "$name rc=$?" | tee -a "$O/r02_${T}_summary.log"; }
run gpu_tests 400 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider -x
run bench     120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run timeline  120 python tools/timeline_graph.py gpurun_out/r02_${T}_timeline.csv
run adam      60  python tools/bench_adam.py
run adam_ncu  180 ncu --set full --clock-control none -k regex:adam_pack -c 8 -o gpurun_out/r02_${T}_adam python tools/bench_adam.py --once
tail -n 6 "$O"/r02_${T}_gpu_tests.log
cat "$O"/r02_${T}_adam.log
grep -h '"metric"' "$O"/r02_${T}_bench.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%.2f ms/step  %.1f stories/s  e2e %.1f  launches/step %s  gemm frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches_per_step'], d['roofline']['frac']))
"
cat "$O/r02_${T}_summary.log"
