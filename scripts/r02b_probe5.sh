#!/bin/bash
# GPU call 5: two-tier accumulate chunks (parity, isolated MSM, pipeline) and which vCPUs of the box share a core.
out=gpurun_out/r02b_probe5.txt
: > $out
echo "## pytest -m gpu" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $out
echo "## isolated 2^19 MSM" >> $out
for o in "acc_long_pct=0" "acc_long_pct=75" "acc_long_pct=85" "acc_long_pct=60" "acc_long_pct=75 acc_waves=6" "acc_long_pct=85 acc_waves=8" "acc_long_pct=50 acc_waves=3"; do
  timeout 120 python scripts/msm_opt_probe.py 19 $o 2>&1 | tail -1 >> $out
done
run() {  # label, cpu list or "-", B, steps, opts...
  label=$1; cpus=$2; B=$3; steps=$4; shift 4
  opts=""; for kv in "$@"; do opts="$opts --option $kv"; done
  pre=""; [ "$cpus" != "-" ] && pre="taskset -c $cpus"
  echo -n "$label B=$B $* : " >> $out
  $pre timeout -s KILL 400 python bench.py --skip-cpu-baseline --skip-msm-leg --steps $steps --warmup 3 --blobs-per-step $B $opts 2>>gpurun_out/r02b_probe5.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), round(d['roofline']['launch_ms_isolated'],3), round(d['roofline']['msm_total_ms_isolated'],3))" >> $out 2>&1
}
echo "## bench: value e2e ms_per_step acc_ms msm_ms" >> $out
run full - 16 10 acc_long_pct=0
run full - 16 10 acc_long_pct=75
run full - 16 10 acc_long_pct=85
run full - 64 4 acc_long_pct=75
echo "## host SHA-256 by cpu set (which vCPUs share a core)" >> $out
g++ -O3 -std=c++17 -pthread -o /tmp/hb scripts/host_sha_mb_bench.cpp rust-kzg-bn254_b200/csrc/sha256.cpp
for set in 0,1 0,2 0,8 0-3 0,2,4,6; do
  n=$(echo $set | python -c "import sys; s=sys.stdin.read().strip(); print(4 if s in ('0-3','0,2,4,6') else 2)")
  echo "# taskset -c $set, $n threads" >> $out
  taskset -c $set /tmp/hb 16 $n 2>&1 | grep "T=" >> $out
done
cat $out
