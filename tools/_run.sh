#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/sweep3.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> $out
tail -6 gpurun_out/pytest_gpu.txt >> $out
C3="--scene inst --width 3840 --height 2160 --bounces 3 --passes 4"
for o in "TRACE_FETCH_TH=20" "TRACE_FETCH_TH=24" "TRACE_FETCH_TH=28" "TRACE_FETCH_TH=24 TRACE_SCHED=0x080808" "TRACE_FETCH_TH=24 TRACE_SCHED=0x0c0808" "TRACE_FETCH_TH=24 TRACE_SCHED=0x060c0c" "TRACE_FETCH_TH=24 TRACE_SCHED=0x0a0a0a"; do
  args=""; for kv in $o; do args="$args --opt $kv"; done
  python tools/stage_times.py $args --tag "C2 $o" >> $out 2>&1
  python tools/stage_times.py $C3 $args --tag "C3 $o" >> $out 2>&1
done
cut -c1-230 $out
