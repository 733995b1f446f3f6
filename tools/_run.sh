#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.txt | tail -3
python tools/stage_times.py --opt TRACE_COOP=0 --tag "C2 coop off" 2>&1 | cut -c1-200
python tools/stage_times.py --opt TRACE_COOP=1 --tag "C2 coop on" 2>&1 | cut -c1-200
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --opt TRACE_COOP=0 --tag "C3 coop off" 2>&1 | cut -c1-200
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --opt TRACE_COOP=1 --tag "C3 coop on" 2>&1 | cut -c1-200
python tools/pass_time.py --tag "C2 2parts" 2>&1
