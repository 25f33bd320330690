#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_video_encoder.py tests/test_conv_jobs.py -q -m gpu -k "video or temporal" > gpurun_out/c11_video.log 2>&1
echo "video rc=$?" >> gpurun_out/c11_video.log
timeout 300 python -m pytest tests/test_step_parity.py -q -m gpu -x -k "clevr_seq" -s > gpurun_out/c11_seq_step.log 2>&1
echo "seq rc=$?" >> gpurun_out/c11_seq_step.log
tail -n 5 gpurun_out/c11_video.log; tail -n 5 gpurun_out/c11_seq_step.log
