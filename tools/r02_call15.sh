#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_packed_adam.py tests/test_trainer_loop.py tests/test_grad_sync.py -q -m gpu -x > $O/c15_tests.log 2>&1; echo "tests rc=$?" >> $O/c15_tests.log
tail -n 3 $O/c15_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c15_bench_layer.log 2>&1; echo "rc=$?" >> $O/c15_bench_layer.log
CPCSV_LAYERWISE_ADAM=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c15_bench_once.log 2>&1; echo "rc=$?" >> $O/c15_bench_once.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c15_bench_layer2.log 2>&1; echo "rc=$?" >> $O/c15_bench_layer2.log
CPCSV_LAYERWISE_ADAM=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c15_bench_once2.log 2>&1; echo "rc=$?" >> $O/c15_bench_once2.log
for f in c15_bench_layer c15_bench_once c15_bench_layer2 c15_bench_once2; do echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $O/$f.log | head -2; tail -n 1 $O/$f.log; done
timeout 600 python tools/timeline_graph.py $O/c15_timeline.csv > $O/c15_timeline.log 2>&1; echo "timeline rc=$?"
timeout 900 python -m pytest tests/test_step_parity.py -q -m gpu -x -k "coupled or pororo_step_gpu" > $O/c15_parity.log 2>&1; echo "parity rc=$?" >> $O/c15_parity.log
tail -n 3 $O/c15_parity.log
