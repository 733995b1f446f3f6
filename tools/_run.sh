#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_pipeline_ctas.txt; : > $O
for ct in 0 4 5 6 7; do
python tools/pass_time.py --passes 60 --opt TRACE_CTAS=$ct --tag "C2 pipelined, TRACE_CTAS=$ct" >> $O 2>&1
done
for th in 20 24; do
python tools/pass_time.py --passes 60 --opt TRACE_FETCH_TH=$th --tag "C2 pipelined, fetch_th=$th" >> $O 2>&1
done
python tools/pass_time.py --passes 60 --opt SHADOW_OVERLAP=0 --tag "C2 pipelined, shadow rays in sequence" >> $O 2>&1
cat $O
