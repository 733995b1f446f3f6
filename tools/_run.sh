#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
tail -6 gpurun_out/pytest_gpu.txt
python tools/stage_times.py --tag "C2" 2>&1 | cut -c1-220
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag "C3" 2>&1 | cut -c1-220
python tools/pass_time.py --tag "C2 2parts" 2>&1
