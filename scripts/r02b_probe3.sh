#!/bin/bash
# GPU call 3: AVX-512 multi-buffer transcript hashing (parity, host throughput on this box, the blob pipeline with one rank's
# host share of the 8-GPU box) and the accumulate kernel at 1 / 2 / 4 waves under ncu.
out=gpurun_out/r02b_probe3.txt
: > $out
echo "## pytest -m gpu" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $out
echo "## host SHA-256 on this box" >> $out
g++ -O3 -std=c++17 -pthread -o /tmp/host_sha_mb_bench scripts/host_sha_mb_bench.cpp rust-kzg-bn254_b200/csrc/sha256.cpp && /tmp/host_sha_mb_bench 16 1 4 8 16 >> $out 2>&1
run() {  # label, cpu list or "-", B, steps, opts...
  label=$1; cpus=$2; B=$3; steps=$4; shift 4
  opts=""; for kv in "$@"; do opts="$opts --option $kv"; done
  pre=""; [ "$cpus" != "-" ] && pre="taskset -c $cpus"
  echo -n "$label B=$B $* : " >> $out
  $pre timeout -s KILL 400 python bench.py --skip-cpu-baseline --skip-msm-leg --steps $steps --warmup 3 --blobs-per-step $B $opts 2>>gpurun_out/r02b_probe3.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))" >> $out 2>&1
}
echo "## bench: value e2e ms_per_step" >> $out
run 4cpu 0-3 192 3 lane_wait=1
run 4cpu 0-3 192 3 lane_wait=1 hash_mb=0 hash_threads=16
run 4cpu 0-3 64 4 lane_wait=1
run 4cpu 0-3 64 4 lane_wait=1 hash_mb=0 hash_threads=16
run 4cpu 0-3 16 8 lane_wait=1 hash_threads=16
run 4cpu 0-3 384 2 lane_wait=1
run full - 192 3 hash_mb=1
run full - 192 3
echo "## ncu accumulate waves 1 / 2 / 4" >> $out
M=gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__warps_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio
for w in 1 2 4; do
  timeout 300 ncu --metrics $M --clock-control none -k k_accumulate -s 3 -c 1 --csv --log-file gpurun_out/r02b_acc_waves$w.csv python scripts/msm_opt_probe.py 19 acc_waves=$w >> $out 2>&1
done
python - >> $out <<'P'
import csv
for w in (1,2,4):
    rows=[r for r in csv.reader(l for l in open(f'gpurun_out/r02b_acc_waves{w}.csv') if l.startswith('"'))]
    hdr=rows[0]
    for r in rows[1:]:
        d=dict(zip(hdr,r)); print(w, d.get('Metric Name'), d.get('Metric Value'))
P
cat $out
