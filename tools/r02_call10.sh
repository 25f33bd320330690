#!/bin/bash
# f4 (VideoEncoder) on the GPU: the new tests first, then the whole GPU suite, smoke and a bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_video_encoder.py tests/test_conv_jobs.py -q -m gpu -x -k "video or temporal" > gpurun_out/c10_video.log 2>&1
echo "video rc=$?" >> gpurun_out/c10_video.log
timeout 300 python -m pytest tests/test_step_parity.py -q -m gpu -x -k "clevr_seq" -s > gpurun_out/c10_seq_step.log 2>&1
echo "seq rc=$?" >> gpurun_out/c10_seq_step.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c10_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c10_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/c10_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/c10_bench.log
tail -3 gpurun_out/c10_video.log gpurun_out/c10_seq_step.log gpurun_out/c10_pytest.log gpurun_out/c10_smoke.log
tail -c 1500 gpurun_out/c10_bench.log
