mkdir -p gpurun_out
rm -f gpurun_out/stages.txt
RTX_B200_LIB=build/variants/tps2.so python tools/stage_times.py >> gpurun_out/stages.txt 2>&1
cat gpurun_out/stages.txt
(RTX_B200_LIB=build/variants/tps2.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
