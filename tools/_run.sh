#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
python tools/sweep.py restir > gpurun_out/restir_frame3.json 2>&1
cat gpurun_out/restir_frame3.json | cut -c1-700
