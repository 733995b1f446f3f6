#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out; T=s5
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"trace_kernel|k_gi_step|k_shade_primary|k_di_finish" -s 52 -c 13 -o $O/r02_${T}_pass -f python tools/prof_pass.py --passes 3 --opt PASS_PARTS=1 --opt PASS_PIPELINE=0 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
for l in 0 1; do
RTX_B200_LIB=build/variants/timeline.so python tools/pass_timeline.py $l > $O/tl_$l.txt 2>&1
python tools/pass_timeline.py --analyse $O/tl_$l.txt > $O/r02_${T}_trace_timeline_lpt$l.txt
done
head -7 $O/r02_${T}_trace_timeline_lpt1.txt
