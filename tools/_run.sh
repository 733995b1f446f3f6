#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
python bench.py > gpurun_out/r02_s7_bench_n1.json 2> gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_s7_bench_reference_arm.json 2>> gpurun_out/bench_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r02_s7_bench_n1.json').read().strip().splitlines()[-1]); print('full   value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'fast', round(d['fast_math']['value']), 'C3', round(d['configs']['C3']['value']))"
