#!/bin/bash
# scratch runner for gpurun calls: edit, run as `gpurun -- 'bash tools/_run.sh'`; outputs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests.log
O=gpurun_out/sweep_side.txt; : > $O
for r in 1 2; do
for so in 0 1; do
python tools/pass_time.py --passes 30 --opt SHADOW_OVERLAP=$so --tag "C2 parts=2 side=$so" >> $O 2>&1
python tools/pass_time.py --passes 30 --opt PASS_PARTS=1 --opt SHADOW_OVERLAP=$so --tag "C2 parts=1 side=$so" >> $O 2>&1
python tools/pass_time.py --passes 30 --opt PASS_PARTS=3 --opt SHADOW_OVERLAP=$so --tag "C2 parts=3 side=$so" >> $O 2>&1
done; done
for so in 0 1; do
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --opt SHADOW_OVERLAP=$so --tag "C3 parts=2 side=$so" >> $O 2>&1
done
cat gpurun_out/gpu_tests.log; cat $O
