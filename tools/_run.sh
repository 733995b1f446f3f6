#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests.log
O=gpurun_out/sweep_bsdf_at_sample.txt; : > $O
python tools/stage_times.py --opt PASS_PARTS=1 --tag "bsdf at sample time" >> $O 2>&1
python tools/pass_time.py --passes 30 --tag "bsdf at sample time" >> $O 2>&1
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag C3 >> $O 2>&1
python tools/pass_time.py --passes 30 --flags 32 --tag "fast math" >> $O 2>&1
cat gpurun_out/gpu_tests.log; cat $O
