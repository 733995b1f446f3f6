#!/bin/bash
# A/B of traversal variants on the bench workload (run under gpurun): build/variants/*.so against the default build
mkdir -p gpurun_out
out=gpurun_out/ab_trace.txt
rm -f $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
python tools/stage_times.py --tag default >> $out 2>&1
for v in build/variants/*.so; do RTX_B200_LIB=$v python tools/stage_times.py >> $out 2>&1; done
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag default_C3 >> $out 2>&1
for v in build/variants/*.so; do RTX_B200_LIB=$v python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag C3_$v >> $out 2>&1; done
cat $out
