#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_cprim2.txt; : > $O
for v in cprim09 cprim14 cprim25; do
L=build/variants/$v.so
RTX_B200_LIB=$L python tools/stage_times.py --opt PASS_PARTS=1 --tag "C2 $v" >> $O 2>&1
RTX_B200_LIB=$L python tools/pass_time.py --passes 30 --tag "C2 $v" >> $O 2>&1
RTX_B200_LIB=$L python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag "C3 $v" >> $O 2>&1
RTX_B200_LIB=$L python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5', '$v', round(d['value']), d['config'].get('workload','')[:60], d['roofline']['frac'])" >> $O 2>&1
done
L=royaltracer-dx_b200/librtx_b200.so
RTX_B200_LIB=$L python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5', 'default', round(d['value']), d['config'].get('workload','')[:60], d['roofline']['frac'])" >> $O 2>&1
cat $O
