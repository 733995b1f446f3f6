#!/bin/bash
# scratch runner for gpurun calls: edit, run as `gpurun -- 'bash tools/_run.sh'`; outputs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
( time python bench.py > gpurun_out/bench_default_timed.json ) 2> gpurun_out/bench_default_time.txt
tail -3 gpurun_out/bench_default_time.txt
bash tools/final_capture_r02.sh s3
