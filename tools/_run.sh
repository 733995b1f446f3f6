#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/tlas_nodes.txt; : > $O
RTX_B200_LIB=build/variants/tlasstat.so python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 2 --tag "C3 (tri column = TLAS node visits)" >> $O 2>&1
RTX_B200_LIB=build/variants/tlasstat.so python tools/stage_times.py --passes 2 --tag "C2 (tri column = TLAS node visits)" >> $O 2>&1
cat $O
