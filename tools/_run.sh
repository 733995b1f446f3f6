#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/sanitizer4.txt; : > $S
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or async_readback or instance_list_changes or concurrent_pass_parts or graph_replay or engine_side_reduce or restir_frames_in" >> $S 2>&1; echo "memcheck tests rc=$?" >> $S
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $S 2>&1; echo "memcheck smoke rc=$?" >> $S
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $S 2>&1; echo "racecheck smoke rc=$?" >> $S
timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> $S 2>&1; echo "initcheck smoke rc=$?" >> $S
grep -E "SUMMARY|rc=|passed|failed|Error" $S
