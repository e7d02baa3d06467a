#!/bin/bash
# GPU call 9: one rank's host share of the 8-GPU box, emulated faithfully: 4 hardware threads (taskset) AND LOCAL_WORLD_SIZE=4 on this
# 16-vCPU box, so the library computes the same share (hardware_concurrency() ignores the affinity mask) -- multi-buffer hashing traced.
out=gpurun_out/r02b_probe9.txt
: > $out
python -c "import os; print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))" >> $out
echo "## 4 cpus, LOCAL_WORLD_SIZE=4" >> $out
LOCAL_WORLD_SIZE=4 taskset -c 0-3 timeout 900 python scripts/pipeline_ab.py 2 "192:hash_trace=1" "192:hash_mb=0" "64:hash_trace=1" "64:hash_mb=0" "64:hash_trace=1,hash_threads=2" "192:hash_trace=1,hash_threads=4" "16:" "16:lanes=4,lane_wait=1" 2>gpurun_out/r02b_probe9_trace.err | grep -v Warn >> $out
grep "hash\]" gpurun_out/r02b_probe9_trace.err | awk '{w+=$9; c+=$12; n++} END {print "groups", n, "mean wall ms", w/n, "mean cpu ms", c/n}' >> $out
grep "hash\]" gpurun_out/r02b_probe9_trace.err | head -40 >> $out
echo "## accumulate with an L2 prefetch of the next point (isolated 2^19 MSM, then pipeline)" >> $out
for o in "acc_prefetch=0" "acc_prefetch=1" "acc_prefetch=0" "acc_prefetch=1"; do timeout 120 python scripts/msm_opt_probe.py 19 $o 2>&1 | tail -1 >> $out; done
timeout 600 python scripts/pipeline_ab.py 3 "64:acc_prefetch=0" "64:acc_prefetch=1" 2>&1 | grep -v Warn >> $out
cat $out
