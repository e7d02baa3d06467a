#!/bin/bash
# Round 2, session 2, first GPU call: box topology, the GPU test suite on the committed tree, and the blob pipeline at
# deeper batches / with the host share of one rank of the 8-GPU box (4 hardware threads, emulated with taskset).
out=gpurun_out/r02b_probe.txt
: > $out
{ echo "## topology"; nproc; lscpu | grep -E "Model name|Thread|Core|Socket|^CPU\(s\)|MHz|L3|Flags" | cut -c1-400; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu -e | head -40; } >> $out 2>&1
echo "## pytest -m gpu" >> $out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) >> $out 2>&1
run() {  # label, cpu list or "-", B, steps, opts...
  label=$1; cpus=$2; B=$3; steps=$4; shift 4
  opts=""; for kv in "$@"; do opts="$opts --option $kv"; done
  pre=""; [ "$cpus" != "-" ] && pre="taskset -c $cpus"
  echo -n "$label B=$B $* : " >> $out
  $pre timeout -s KILL 400 python bench.py --skip-cpu-baseline --skip-msm-leg --steps $steps --warmup 3 --blobs-per-step $B $opts 2>>gpurun_out/r02b_probe.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))" >> $out 2>&1
}
echo "## bench: value e2e ms_per_step" >> $out
run full - 16 8
run full - 64 4
run full - 192 3
run full - 192 3 device_hash=0
run 4cpu 0-3 16 8 hash_threads=16 lane_wait=1
run 4cpu 0-3 192 3 hash_threads=16 lane_wait=1 device_hash=0
run 4cpu 0-3 192 3 hash_threads=16 lane_wait=1 device_hash=48
run 4cpu 0-3 192 3 hash_threads=4 lane_wait=1 device_hash=48
cat $out
