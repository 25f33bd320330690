#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-c5}
run() { local name=$1 secs=$2; shift 2; echo "== $name" ; timeout "$secs" "$@" > "$O/r02_${T}_$name.log" 2>&1; echo "$name rc=$?" | tee -a "$O/r02_${T}_summary.log"; }
run gpu_tests 600 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider
run smoke     120 python -c "import __graft_entry__ as g; g.smoke()"
run bench     300 python bench.py --steps 20 --warmup 5
run timeline  120 python tools/timeline_graph.py gpurun_out/r02_${T}_timeline.csv
run inference 200 python bench.py --config inference
run stress    300 python bench.py --config stress
run reference 400 python bench.py --impl reference --steps 2 --warmup 1
run traffic   400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_gemm --launch-skip 657 -c 438 --csv --log-file gpurun_out/r02_${T}_gemm_traffic.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
grep -E "passed|failed" "$O"/r02_${T}_gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR" "$O"/r02_${T}_gpu_tests.log | head
tail -n 2 "$O"/r02_${T}_smoke.log
for f in bench inference stress reference; do grep -h '"metric"' "$O"/r02_${T}_$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$f: %.2f ms/step  %.1f %s  e2e %s  frac %s' % (d['ms_per_step'], d['value'], d['unit'], d.get('e2e',{}).get('value'), d.get('config',{}).get('nominal_frac_of_peak')))
"; done
cat "$O/r02_${T}_summary.log"
