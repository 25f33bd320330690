#!/bin/bash
# Evidence captures for profiles/ (one GPU): ncu launch list of the bench command, per-launch DRAM traffic of
# the dominant kernel, ncu --set full of the GEMM classes, the stress-512 per-layer tensor-pipe table.
set -u
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 secs=$2; shift 2; echo "== $name"; timeout "$secs" "$@" > "$O/r02_prof_$name.log" 2>&1; echo "$name rc=$?" | tee -a "$O/r02_prof_summary.log"; }
: > $O/r02_prof_summary.log
# 1. every launch of the bench command (graph replays are profiled kernel node by kernel node)
run launches 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 3600 --csv \
    --log-file $O/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline
# 2. DRAM bytes of every conv_gemm launch of two eager steps
run traffic 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:conv_gemm --launch-skip 654 -c 436 --csv --log-file $O/r02_gemm_traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
# 3. full sets of the GEMM classes (pair mode) and of the head GEMM + gather
run full 600 ncu --set full --clock-control none --import-source on -o $O/r02_ncu_targets python tools/ncu_targets.py all
# 4. stress-512: tensor-pipe activity of every conv_gemm launch of one step
# (application replay: kernel replay would save / restore the 65 GB working set around every launch)
run stress 600 ncu --replay-mode application --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:conv_gemm --launch-skip 436 -c 218 --csv --log-file $O/r02_stress_gemm.csv \
    python bench.py --config stress --steps 2 --warmup 2 --job-log $O/r02_stress_jobs.log
cat $O/r02_prof_summary.log
