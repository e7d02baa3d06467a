#!/bin/bash
# GPU call 6: lane wait modes with one rank's host share of the 8-GPU box (4 hardware threads); two-tier chunk sweep.
out=gpurun_out/r02b_probe6.txt
: > $out
run() {  # label, cpu list or "-", B, steps, opts...
  label=$1; cpus=$2; B=$3; steps=$4; shift 4
  opts=""; for kv in "$@"; do opts="$opts --option $kv"; done
  pre=""; [ "$cpus" != "-" ] && pre="taskset -c $cpus"
  echo -n "$label B=$B $* : " >> $out
  $pre timeout -s KILL 400 python bench.py --skip-cpu-baseline --skip-msm-leg --steps $steps --warmup 3 --blobs-per-step $B $opts 2>>gpurun_out/r02b_probe6.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))" >> $out 2>&1
}
echo "## bench: value e2e ms_per_step" >> $out
run 4cpu 0-3 192 3 lane_wait=2 lanes=6
run 4cpu 0-3 192 3 lane_wait=2 lanes=8
run 4cpu 0-3 192 3 lane_wait=2 lanes=8 hash_mb=0 hash_threads=16
run 4cpu 0-3 192 3 lane_wait=1 lanes=8 hash_mb=0 hash_threads=16
run 4cpu 0-3 64 4 lane_wait=2 lanes=8
run 4cpu 0-3 64 4 lane_wait=2 lanes=8 hash_threads=3
run 4cpu 0-3 16 8 lane_wait=2 lanes=8
run 4cpu 0-3 16 8 lane_wait=2 lanes=6 hash_threads=16
run full - 16 8 lane_wait=2 lanes=6
run full - 64 4 lane_wait=2 lanes=8
echo "## isolated 2^19 MSM" >> $out
for o in "acc_long_pct=0" "acc_long_pct=75 acc_waves=8" "acc_long_pct=80 acc_waves=8" "acc_long_pct=70 acc_waves=6" "acc_long_pct=65 acc_waves=8" "acc_long_pct=80 acc_waves=6" "acc_long_pct=0 acc_waves=6"; do
  timeout 120 python scripts/msm_opt_probe.py 19 $o 2>&1 | tail -1 >> $out
done
cat $out
