#!/bin/bash
mkdir -p gpurun_out
python tools/stage_times.py --tag "C2 default" 2>&1 | cut -c1-220
RTX_B200_LIB=build/variants/tea.so python tools/stage_times.py --tag "C2 tea noinline" 2>&1 | cut -c1-220
for l in 0 1; do
RTX_B200_LIB=build/variants/timeline.so python tools/pass_timeline.py $l > gpurun_out/tl_$l.txt 2>&1
python tools/pass_timeline.py --analyse gpurun_out/tl_$l.txt > gpurun_out/r02_s1_trace_timeline_lpt$l.txt
done
wc -l gpurun_out/r02_s1_trace_timeline_lpt*.txt
