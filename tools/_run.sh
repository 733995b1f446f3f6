#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
O=gpurun_out/sweep_pipeline.txt; : > $O
python tools/pass_time.py --passes 40 --tag "C2 pipelined (default)" >> $O 2>&1
python tools/pass_time.py --passes 40 --opt PASS_PIPELINE=0 --tag "C2 pipeline off" >> $O 2>&1
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --tag "C3 pipelined (default)" >> $O 2>&1
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --opt PASS_PIPELINE=0 --tag "C3 pipeline off" >> $O 2>&1
python tools/pass_time.py --passes 30 --flags 32 --tag "C2 fast pipelined" >> $O 2>&1
python tools/pass_time.py --passes 30 --flags 32 --opt PASS_PIPELINE=0 --tag "C2 fast pipeline off" >> $O 2>&1
cat $O
