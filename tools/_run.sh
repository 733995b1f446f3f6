#!/bin/bash
# scratch runner for gpurun calls: edit, run as `gpurun -- 'bash tools/_run.sh'`; outputs under gpurun_out/
mkdir -p gpurun_out
O=gpurun_out/sweep_push2.txt; : > $O
for r in 1 2; do
python tools/stage_times.py --opt PASS_PARTS=1 --tag "warp atomics" >> $O 2>&1
python tools/pass_time.py --passes 30 --tag "warp atomics" >> $O 2>&1
RTX_B200_LIB=build/variants/pushcta.so python tools/stage_times.py --opt PASS_PARTS=1 --tag "cta atomics" >> $O 2>&1
RTX_B200_LIB=build/variants/pushcta.so python tools/pass_time.py --passes 30 --tag "cta atomics" >> $O 2>&1
done
RTX_B200_LIB=build/variants/pushcta.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render or concurrent or graph" 2>&1 | tail -3 >> $O
cat $O
