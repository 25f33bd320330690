#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "first_discriminator" > $O/c12_enc0.log 2>&1; echo "enc0 rc=$?" >> $O/c12_enc0.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c12_bench_direct.log 2>&1; echo "rc=$?" >> $O/c12_bench_direct.log
CPCSV_DIRECT_ENC0=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c12_bench_im2col.log 2>&1; echo "rc=$?" >> $O/c12_bench_im2col.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c12_bench_direct2.log 2>&1; echo "rc=$?" >> $O/c12_bench_direct2.log
timeout 1500 python -m pytest tests -q -m gpu -x > $O/c12_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c12_pytest.log
timeout 700 ncu --replay-mode application --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:conv_gemm --launch-skip 454 -c 227 --csv --log-file $O/c12_stress_gemm.csv \
    python bench.py --config stress --steps 2 --warmup 2 --job-log $O/c12_stress_jobs.log > $O/c12_stress.log 2>&1; echo "stress rc=$?" >> $O/c12_stress.log
for f in c12_enc0 c12_pytest c12_stress; do echo "== $f"; tail -n 4 $O/$f.log; done
for f in c12_bench_direct c12_bench_im2col c12_bench_direct2; do echo "== $f"; grep -o '"ms_per_step": [0-9.]*' $O/$f.log | head -2; tail -n 1 $O/$f.log; done
wc -l $O/c12_stress_jobs.log $O/c12_stress_gemm.csv
