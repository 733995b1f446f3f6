#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
tail -6 gpurun_out/pytest_gpu.txt
python tools/pass_time.py --opt PASS_GRAPH=0 --tag "C2 2parts nograph" 2>&1
python tools/pass_time.py --opt PASS_GRAPH=1 --tag "C2 2parts graph" 2>&1
python tools/pass_time.py --opt PASS_GRAPH=1 --opt PASS_PARTS=1 --tag "C2 1part graph" 2>&1
python tools/pass_time.py --opt PASS_GRAPH=0 --opt PASS_PARTS=1 --tag "C2 1part nograph" 2>&1
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
