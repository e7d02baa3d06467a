#!/bin/bash
# GPU call 8: variance and A/B of lane settings in one process; where the multi-buffer hashers' time goes with 4 hardware threads.
out=gpurun_out/r02b_probe8.txt
: > $out
echo "## 16 cpus, A/B in one process (3 repetitions each, interleaved)" >> $out
timeout 600 python scripts/pipeline_ab.py 3 "16:lanes=4,lane_wait=0" "16:lanes=6,lane_wait=2" "16:lanes=8,lane_wait=2" "16:" "64:" "64:lanes=8" "64:lanes=4,lane_wait=0" 2>&1 | grep -v Warn >> $out
echo "## 4 cpus" >> $out
taskset -c 0-3 timeout 600 python scripts/pipeline_ab.py 2 "192:hash_trace=1" "192:hash_trace=1,hash_nice=0" "64:" "64:lanes=8" "16:" "16:lanes=4,lane_wait=1" 2>gpurun_out/r02b_probe8_trace.err | grep -v Warn >> $out
grep "\[hash\]" gpurun_out/r02b_probe8_trace.err | head -60 >> $out
cat $out
