#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-c6}
run() { local name=$1 secs=$2; shift 2; echo "== $name" ; timeout "$secs" "$@" > "$O/r02_${T}_$name.log" 2>&1; echo "$name rc=$?" | tee -a "$O/r02_${T}_summary.log"; }
run gpu_tests 600 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider
run bench_a   120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
CPCSV_HEAD_GEMM=0 run bench_headdirect 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
run bench_b   120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
CPCSV_PAIR=0 run bench_single 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
run timeline  120 python tools/timeline_graph.py gpurun_out/r02_${T}_timeline.csv
run inference 200 python bench.py --config inference
grep -E "passed|failed" "$O"/r02_${T}_gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR" "$O"/r02_${T}_gpu_tests.log | head
for f in bench_a bench_headdirect bench_b bench_single inference; do grep -h '"metric"' "$O"/r02_${T}_$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$f: %.2f ms/step  %.1f %s  e2e %s  frac %s' % (d['ms_per_step'], d['value'], d['unit'], d.get('e2e',{}).get('value'), d.get('config',{}).get('nominal_frac_of_peak')))
"; done
cat "$O/r02_${T}_summary.log"
