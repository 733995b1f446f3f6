#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_ploc.txt; : > $O
for v in default ploc8 ploc32 ploc64; do
L=build/variants/$v.so; [ $v = default ] && L=royaltracer-dx_b200/librtx_b200.so
RTX_B200_LIB=$L python tools/stage_times.py --opt PASS_PARTS=1 --tag "C2 $v" >> $O 2>&1
RTX_B200_LIB=$L python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag "C3 $v" >> $O 2>&1
RTX_B200_LIB=$L python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5', '$v', round(d['value']), d['roofline']['frac'], [ (round(b['build_ms'],1), round(b.get('build_ms_warm') or 0,1)) for b in d.get('blas',[])])" >> $O 2>&1
done
cat $O
