#!/bin/bash
# scratch runner for gpurun calls: edit, run as `gpurun -- 'bash tools/_run.sh'`; outputs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests.log
O=gpurun_out/sweep_gi_parts.txt; : > $O
for p in 1 2 3 4; do python tools/pass_time.py --opt PASS_PARTS=$p --passes 30 --tag "parts=$p" >> $O 2>&1; done
for p in 1 2 3 4; do python tools/pass_time.py --scene inst --width 3840 --height 2160 --opt PASS_PARTS=$p --passes 10 --tag "C3 parts=$p" >> $O 2>&1; done
python tools/stage_times.py --opt PASS_PARTS=1 --tag "default gi384x2" >> $O 2>&1
for v in gi256x3 gi256x4 gi128x6 gi128x8 gi512x1; do
  RTX_B200_LIB=build/variants/$v.so python tools/stage_times.py --opt PASS_PARTS=1 --tag "$v" >> $O 2>&1
  RTX_B200_LIB=build/variants/$v.so python tools/pass_time.py --passes 30 --tag "$v" >> $O 2>&1
done
S=gpurun_out/sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $S 2>&1; echo "memcheck rc=$?" >> $S
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $S 2>&1; echo "racecheck rc=$?" >> $S
timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $S 2>&1; echo "initcheck rc=$?" >> $S
tail -5 $O
