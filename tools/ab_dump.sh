#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ab_dump.txt; rm -f $out
for cfg in "1 1 0" "2 1 0" "2 2 0" "3 3 0" "4 4 0" "4 2 0" "2 2 12" "3 2 0"; do
  set -- $cfg
  RTX_PARTS=$1 RTX_TRACE_GRID_DIV=$2 RTX_DUMP_TH=$3 python tools/pass_time.py --tag C2_parts$1_div$2_dump$3 >> $out 2>&1
  RTX_PARTS=$1 RTX_TRACE_GRID_DIV=$2 RTX_DUMP_TH=$3 python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag C3_parts$1_div$2_dump$3 >> $out 2>&1
done
cat $out
