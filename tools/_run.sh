#!/bin/bash
# scratch runner for gpurun calls: edit, run as `gpurun -- 'bash tools/_run.sh'`; outputs under gpurun_out/
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_s6_bench_n1.json 2> gpurun_out/bench_err.log
tail -1 gpurun_out/r02_s6_bench_n1.json | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_s6_bench_reference_arm.json 2>> gpurun_out/bench_err.log
tail -1 gpurun_out/r02_s6_bench_reference_arm.json | cut -c1-200
