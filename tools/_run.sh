mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/gputests.log
python tools/stage_times.py > gpurun_out/stages.txt 2>&1
RTX_TRACE_WAVES=2 python tools/stage_times.py --tag waves2 >> gpurun_out/stages.txt 2>&1
RTX_FETCH_TH=16 python tools/stage_times.py --tag fetch16 >> gpurun_out/stages.txt 2>&1
RTX_FETCH_TH=28 python tools/stage_times.py --tag fetch28 >> gpurun_out/stages.txt 2>&1
for v in conv0 minb5 minb6 chunk32; do RTX_B200_LIB=build/variants/$v.so python tools/stage_times.py >> gpurun_out/stages.txt 2>&1; done
cat gpurun_out/gputests.log gpurun_out/stages.txt
