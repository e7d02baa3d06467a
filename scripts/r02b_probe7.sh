#!/bin/bash
# GPU call 7: library defaults (blocking lanes for deep batches, multi-buffer hashing by plan) with 4 and 16 hardware threads.
out=gpurun_out/r02b_probe7.txt
: > $out
run() {  # label, cpu list or "-", B, steps, opts...
  label=$1; cpus=$2; B=$3; steps=$4; shift 4
  opts=""; for kv in "$@"; do opts="$opts --option $kv"; done
  pre=""; [ "$cpus" != "-" ] && pre="taskset -c $cpus"
  echo -n "$label B=$B $* : " >> $out
  $pre timeout -s KILL 400 python bench.py --skip-cpu-baseline --skip-msm-leg --steps $steps --warmup 3 --blobs-per-step $B $opts 2>>gpurun_out/r02b_probe7.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['single_blob_latency_ms']['value'])" >> $out 2>&1
}
echo "## pytest" >> $out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $out
echo "## bench: value e2e ms_per_step single_blob_ms" >> $out
run 4cpu 0-3 16 8
run 4cpu 0-3 64 4
run 4cpu 0-3 192 3
run 4cpu 0-3 32 6
run full - 16 10
run full - 64 4
run full - 192 3
run full - 32 6
cat $out
