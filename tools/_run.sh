#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "fast_math" 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo "bench rc $?"; tail -3 gpurun_out/bench_r02_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["parity"]["float_mismatches"], d.get("fast_math"))
for k,v in d["configs"].items(): print(k, v["value"], v["roofline"]["frac"], (v.get("parity") or {}).get("float_mismatches"), (v.get("parity") or {}).get("hit_id_mismatches"))
print(d["c4_strong"])
PY
