#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-c8}
run() { local name=$1 secs=$2; shift 2; echo "== $name" ; timeout "$secs" "$@" > "$O/r02_${T}_$name.log" 2>&1; echo "$name rc=$?" | tee -a "$O/r02_${T}_summary.log"; }
run gpu_tests 600 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider -x
run bench_a   150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
CPCSV_PIPELINE_G=0 run bench_nopipe 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
run bench_b   150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
CPCSV_NOGRAD_SPLIT=1 run bench_split 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline
run timeline  120 python tools/timeline_graph.py gpurun_out/r02_${T}_timeline.csv
grep -E "passed|failed" "$O"/r02_${T}_gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR" "$O"/r02_${T}_gpu_tests.log | head
for f in bench_a bench_nopipe bench_b bench_split; do grep -h '"metric"' "$O"/r02_${T}_$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$f: %.2f ms/step  %.1f %s  e2e %s  gemm frac %.3f  serial gemm ms %.2f' % (d['ms_per_step'], d['value'], d['unit'], d.get('e2e',{}).get('value'), d['roofline']['frac'], d['roofline']['gemm_ms_per_step_serial_events']))
"; done
cat "$O/r02_${T}_summary.log"
