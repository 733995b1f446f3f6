#!/bin/bash
# sweeps the traversal knobs on the bench workload (run under gpurun)
for cfg in "24 0 2" "24 4 2" "24 8 2" "24 12 2" "16 0 2" "28 0 2" "20 8 2" "24 8 1" "24 8 4"; do
  set -- $cfg
  RTX_FETCH_TH=$1 RTX_POSTPONE_TH=$2 RTX_TRACE_WAVES=$3 python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fetch/postpone/waves $cfg  Mrays/s %.0f  ms/step %.2f  trace avg ms %.3f  frac %.3f' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac']))"
done
