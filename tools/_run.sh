#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_s2_bench_n1.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s2_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], "frac", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["parity"]["float_mismatches"], d["parity"]["hit_id_mismatches"])
print(d.get("fast_math"))
for k,v in d["configs"].items(): print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"], (v.get("parity") or {}).get("float_mismatches"), (v.get("parity") or {}).get("hit_id_mismatches"), [b["build_ms"] for b in v.get("blas") or []])
print(d["c4_strong"], d["blas"])
PY
python tools/stage_times.py --tag C2 > gpurun_out/r02_s2_stage_times.txt 2>&1
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag C3 >> gpurun_out/r02_s2_stage_times.txt 2>&1
python tools/stage_times.py --scene cornell --width 256 --height 256 --bounces 2 --flags 3 --tag C1 >> gpurun_out/r02_s2_stage_times.txt 2>&1
python tools/stage_times.py --flags 32 --tag "C2 fast math" >> gpurun_out/r02_s2_stage_times.txt 2>&1
cut -c1-230 gpurun_out/r02_s2_stage_times.txt
