#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 180 python -m pytest tests/test_conv_jobs.py -q -m gpu -x -k "cta_pairs" > $O/c13_wpair.log 2>&1; echo "wpair rc=$?" >> $O/c13_wpair.log
tail -n 3 $O/c13_wpair.log
if ! grep -q "wpair rc=0" $O/c13_wpair.log; then echo "pair wgrad test failed: stop"; grep -n "^E" $O/c13_wpair.log | head -20; exit 0; fi
timeout 1500 python -m pytest tests -q -m gpu -x > $O/c13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c13_pytest.log
tail -n 3 $O/c13_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c13_bench_pair.log 2>&1; echo "rc=$?" >> $O/c13_bench_pair.log
CPCSV_WGRAD_PAIR=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c13_bench_nopair.log 2>&1; echo "rc=$?" >> $O/c13_bench_nopair.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c13_bench_pair2.log 2>&1; echo "rc=$?" >> $O/c13_bench_pair2.log
for f in c13_bench_pair c13_bench_nopair c13_bench_pair2; do echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $O/$f.log | head -2; grep -o '"gemm_ms_per_step_serial_events": [0-9.]*' $O/$f.log; tail -n 1 $O/$f.log; done
: > $O/c13_soak.log
for i in 1 2 3 4 5 6; do timeout 120 python tools/soak_replay.py --replays 150 --heat 2 >> $O/c13_soak.log 2>&1; echo "soak run $i rc=$?" >> $O/c13_soak.log; done
grep "soak run" $O/c13_soak.log
