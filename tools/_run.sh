#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/sweep_stagger.txt; : > $O
for r in 1 2; do
for st in 0 1; do
python tools/pass_time.py --passes 30 --opt PART_STAGGER=$st --tag "C2 parts=2 stagger=$st" >> $O 2>&1
python tools/pass_time.py --passes 30 --opt PASS_PARTS=3 --opt PART_STAGGER=$st --tag "C2 parts=3 stagger=$st" >> $O 2>&1
python tools/pass_time.py --passes 30 --opt PASS_PARTS=4 --opt PART_STAGGER=$st --tag "C2 parts=4 stagger=$st" >> $O 2>&1
done; done
for st in 0 1; do
python tools/pass_time.py --scene inst --width 3840 --height 2160 --bounces 3 --passes 10 --opt PART_STAGGER=$st --tag "C3 parts=2 stagger=$st" >> $O 2>&1
done
cat $O
