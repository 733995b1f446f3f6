#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_s1_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
echo "bench n$N rc $?"; tail -3 gpurun_out/r02_scale_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_s1_scale_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}, d["e2e"], d["c4_strong"])
PY
