#!/bin/bash
# whole-pass time against RTX_OPT_PASS_PARTS (concurrent path ranges) and the traversal CTA size — events around 8 passes, no per-stage timing
mkdir -p gpurun_out
out=gpurun_out/ab_parts.txt; rm -f $out
for lib in default build/variants/b64.so build/variants/b32.so; do
for p in 1 2 3; do
  L=""; [ $lib != default ] && L="RTX_B200_LIB=$lib"
  env $L python tools/pass_time.py --opt PASS_PARTS=$p --tag C2_${lib##*/}_parts$p >> $out 2>&1
  env $L python tools/pass_time.py --opt PASS_PARTS=$p --scene inst --width 3840 --height 2160 --bounces 3 --passes 4 --tag C3_${lib##*/}_parts$p >> $out 2>&1
done
done
cat $out
