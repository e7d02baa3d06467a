#!/bin/bash
# GPU call 4: the blob pipeline with one rank's host share of the 8-GPU box (taskset to 4 hardware threads): hash pool priority,
# pool size, lanes.
out=gpurun_out/r02b_probe4.txt
: > $out
run() {  # label, cpu list or "-", B, steps, opts...
  label=$1; cpus=$2; B=$3; steps=$4; shift 4
  opts=""; for kv in "$@"; do opts="$opts --option $kv"; done
  pre=""; [ "$cpus" != "-" ] && pre="taskset -c $cpus"
  echo -n "$label B=$B $* : " >> $out
  $pre timeout -s KILL 400 python bench.py --skip-cpu-baseline --skip-msm-leg --steps $steps --warmup 3 --blobs-per-step $B $opts 2>>gpurun_out/r02b_probe4.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))" >> $out 2>&1
}
echo "## bench: value e2e ms_per_step" >> $out
run 4cpu 0-3 192 3 lane_wait=1
run 4cpu 0-3 192 3 lane_wait=1 hash_nice=0
run 4cpu 0-3 192 3 lane_wait=1 hash_threads=3
run 4cpu 0-3 192 3 lane_wait=1 hash_threads=2
run 4cpu 0-3 192 3 lane_wait=1 lanes=8
run 4cpu 0-3 192 3 lane_wait=0 lanes=4 hash_threads=2
run 4cpu 0-3 64 4 lane_wait=1
run 4cpu 0-3 64 4 lane_wait=1 hash_threads=3
run 4cpu 0-3 16 8 lane_wait=1
run 4cpu 0-3 16 8 lane_wait=1 hash_mb=1
run 2cpu 0-1 192 3 lane_wait=1
run full - 16 8
cat $out
