#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_th_final2.txt; : > $O
for th in 18 20 22 24; do
python tools/pass_time.py --passes 40 --opt TRACE_FETCH_TH=$th --tag "C2 fetch_th=$th" >> $O 2>&1
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --opt TRACE_FETCH_TH=$th --tag "C3 fetch_th=$th" >> $O 2>&1
done
python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 default', round(d['value']), d['roofline']['frac'])" >> $O 2>&1
cat $O
