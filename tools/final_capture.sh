#!/bin/bash
# End-of-session captures (run under gpurun, one GPU): bench line, reference arm, ncu launch list of the bench command, ncu --set full of the
# traversal + shading kernels of one C2 pass, per-stage times, C1/C3/C5/ReSTIR sweep.
T=${1:-s4}
O=gpurun_out
mkdir -p $O
python bench.py > $O/r01_${T}_bench_n1.json 2> $O/bench_err.log
tail -1 $O/r01_${T}_bench_n1.json | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > $O/r01_${T}_bench_reference_arm.json 2>> $O/bench_err.log
tail -1 $O/r01_${T}_bench_reference_arm.json | cut -c1-300
python tools/stage_times.py --tag C2 > $O/r01_${T}_stage_times.txt 2>&1
python tools/stage_times.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag C3 >> $O/r01_${T}_stage_times.txt 2>&1
cat $O/r01_${T}_stage_times.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r01_${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"trace_kernel|k_gi_step|k_shade_primary|k_di_finish" -s 52 -c 8 -o $O/r01_${T}_pass -f python tools/prof_pass.py --passes 3 > $O/ncu_full.log 2>&1
python tools/sweep.py c1 > $O/r01_${T}_sweep.jsonl 2>> $O/bench_err.log
python tools/sweep.py c3 >> $O/r01_${T}_sweep.jsonl 2>> $O/bench_err.log
python tools/sweep.py c5 10000 1000000 10000000 50000000 >> $O/r01_${T}_sweep.jsonl 2>> $O/bench_err.log
python tools/sweep.py restir >> $O/r01_${T}_sweep.jsonl 2>> $O/bench_err.log
cut -c1-260 $O/r01_${T}_sweep.jsonl
tail -5 $O/bench_err.log
ls -la $O
