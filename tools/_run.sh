#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_spminb.txt; : > $O
for v in default spminb6 spminb5 spminb4; do
L=build/variants/$v.so; [ $v = default ] && L=royaltracer-dx_b200/librtx_b200.so
RTX_B200_LIB=$L python tools/stage_times.py --opt PASS_PARTS=1 --tag "C2 $v" >> $O 2>&1
RTX_B200_LIB=$L python tools/pass_time.py --passes 30 --tag "C2 $v" >> $O 2>&1
RTX_B200_LIB=$L python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag "C3 $v" >> $O 2>&1
done
cat $O
