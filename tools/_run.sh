#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
O=gpurun_out/sweep_interleave3.txt; : > $O
python tools/pass_time.py --passes 40 --tag "C2 default (auto rows)" >> $O 2>&1
python tools/pass_time.py --passes 40 --opt PART_ROWS=0 --tag "C2 contiguous" >> $O 2>&1
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --tag "C3 default (auto rows)" >> $O 2>&1
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --opt PART_ROWS=0 --tag "C3 contiguous" >> $O 2>&1
python tools/sweep.py restir >> $O 2>&1
cat $O | cut -c1-400
