#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_gistage2.txt; : > $O
for r in 1 2 3; do
for v in default gistage; do
L=build/variants/$v.so; [ $v = default ] && L=royaltracer-dx_b200/librtx_b200.so
RTX_B200_LIB=$L python tools/pass_time.py --passes 40 --tag "C2 $v" >> $O 2>&1
done; done
for v in default gistage; do
L=build/variants/$v.so; [ $v = default ] && L=royaltracer-dx_b200/librtx_b200.so
RTX_B200_LIB=$L python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --tag "C3 $v" >> $O 2>&1
RTX_B200_LIB=$L python tools/pass_time.py --passes 30 --flags 32 --tag "C2 fast $v" >> $O 2>&1
done
cat $O
