#!/bin/bash
# scratch runner for gpurun calls: edit, run as `gpurun -- 'bash tools/_run.sh'`; outputs under gpurun_out/
mkdir -p gpurun_out
bash tools/final_capture_r02.sh s4
