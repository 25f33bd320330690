#!/bin/bash
# final evidence of the round on one box: GPU tests, smoke, both bench arms, a timeline of one graph replay
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r02_final_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02_final_gpu_tests.log
tail -n 3 $O/r02_final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_final_smoke.log
tail -n 2 $O/r02_final_smoke.log
timeout 900 python bench.py > $O/r02_final_bench.log 2>&1; echo "bench rc=$?" >> $O/r02_final_bench.log
grep -o '"ms_per_step": [0-9.]*' $O/r02_final_bench.log | head -2; tail -n 1 $O/r02_final_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_final_bench_reference.log 2>&1; echo "ref rc=$?" >> $O/r02_final_bench_reference.log
tail -c 600 $O/r02_final_bench_reference.log
timeout 600 python tools/timeline_graph.py $O/r02_final_timeline.csv > $O/r02_final_timeline.log 2>&1; echo "timeline rc=$?"
timeout 600 python bench.py --config inference --steps 10 --warmup 3 > $O/r02_final_inference.log 2>&1; echo "inference rc=$?"
grep -o '"stories": [0-9]*\|"value": [0-9.]*' $O/r02_final_inference.log | head -16
