mkdir -p gpurun_out
python tools/stage_times.py > gpurun_out/stages.txt 2>&1
cat gpurun_out/stages.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_kernel" -s 1 -c 4 -o gpurun_out/prof_trace2 -f python tools/prof_pass.py --passes 1 > gpurun_out/ncu_trace2.log 2>&1
tail -2 gpurun_out/ncu_trace2.log
