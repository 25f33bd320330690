#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_packed_adam.py tests/test_ops_gpu.py tests/test_trainer_loop.py -q -m gpu -x > $O/c14_tests.log 2>&1; echo "tests rc=$?" >> $O/c14_tests.log
tail -n 3 $O/c14_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c14_bench_low.log 2>&1; echo "rc=$?" >> $O/c14_bench_low.log
CPCSV_ADAM_LOW_PRIORITY=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c14_bench_eq.log 2>&1; echo "rc=$?" >> $O/c14_bench_eq.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c14_bench_low2.log 2>&1; echo "rc=$?" >> $O/c14_bench_low2.log
CPCSV_ADAM_LOW_PRIORITY=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c14_bench_eq2.log 2>&1; echo "rc=$?" >> $O/c14_bench_eq2.log
for f in c14_bench_low c14_bench_eq c14_bench_low2 c14_bench_eq2; do echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $O/$f.log | head -2; tail -n 1 $O/$f.log; done
timeout 600 python tools/timeline_graph.py $O/c14_timeline.csv > $O/c14_timeline.log 2>&1; echo "timeline rc=$?"
