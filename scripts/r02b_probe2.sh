#!/bin/bash
# GPU call 2: lane-per-transcript device hashing (parity, then throughput with the host share of one rank of the 8-GPU box)
# and why fewer accumulate waves are slower (ncu of the isolated MSM at 1 and 4 waves).
out=gpurun_out/r02b_probe2.txt
: > $out
{ echo "## flags"; grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -E "avx|sha|bmi|vaes|gfni" | tr '\n' ' '; echo; } >> $out
echo "## parity" >> $out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "device_hashed or batch" 2>&1 | tail -3 >> $out
echo "## 4 cpus (one rank's share of the 8-GPU box): B=192" >> $out
taskset -c 0-3 timeout 600 python scripts/device_hash_bench.py 192 hash_threads=16 lane_wait=1 0 32 64 96 128 48:0 2>&1 | grep -v Warn >> $out
echo "## 16 cpus: B=192" >> $out
timeout 300 python scripts/device_hash_bench.py 192 0 64 128 192 2>&1 | grep resident >> $out
echo "## ncu accumulate waves 1 / 4" >> $out
M=gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__warps_active.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio
for w in 1 2 4; do
  cat > /tmp/w.py <<P
import sys; sys.argv=['x','19']
import os; sys.path.insert(0, os.getcwd())
from __graft_entry__ import load_package
pkg = load_package(); pkg.lib.kzgb_set_option(b"acc_waves", $w)
exec(open('scripts/msm_sweep.py').read().split('pkg = load_package()')[1])
P
  timeout 300 ncu --metrics $M --clock-control none -k k_accumulate -s 3 -c 1 --csv --log-file gpurun_out/r02b_acc_waves$w.csv python /tmp/w.py >> $out 2>&1
done
python - >> $out <<'P'
import csv
for w in (1,2,4):
    rows=[r for r in csv.reader(l for l in open(f'gpurun_out/r02b_acc_waves{w}.csv') if l.startswith('"'))]
    hdr=rows[0]; 
    for r in rows[1:]:
        d=dict(zip(hdr,r)); print(w, d.get('Metric Name'), d.get('Metric Value'))
P
cat $out
