mkdir -p gpurun_out
rm -f gpurun_out/stages.txt
python tools/stage_times.py >> gpurun_out/stages.txt 2>&1
for v in build/variants/*.so; do RTX_B200_LIB=$v python tools/stage_times.py >> gpurun_out/stages.txt 2>&1; done
cat gpurun_out/stages.txt
