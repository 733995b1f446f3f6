#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_pipeline_share.txt; : > $O
python tools/pass_time.py --passes 60 --tag "C2 pipelined, half grid per lane" >> $O 2>&1
python tools/pass_time.py --passes 60 --tag "C2 pipelined, half grid per lane" >> $O 2>&1
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 12 --tag "C3 pipelined, half grid per lane" >> $O 2>&1
python tools/pass_time.py --passes 40 --flags 32 --tag "C2 fast pipelined, half grid per lane" >> $O 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or graph or concurrent" 2>&1 | tail -2 >> $O
cat $O
