#!/bin/bash
# 1-GPU call: GPU test suite, the default bench line, the ncu launch list of the same command (short), smoke.
out=gpurun_out/r02b_final1.txt
: > $out
echo "## pytest -m gpu" >> $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $out
echo "## smoke" >> $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $out
echo "## bench.py (default)" >> $out
timeout 600 python bench.py > gpurun_out/r02b_bench_n1_final.json 2> gpurun_out/r02b_bench_n1_final.err
python -c "import sys,json; d=json.loads(open('gpurun_out/r02b_bench_n1_final.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['extra']['batch16'], d['host_cpu_ms_per_blob'], d['clocks'], d['roofline']['frac'], d['roofline']['step']['frac'], d['cpu_baseline']['value'], d['extra']['msm_mpts']['value'], d['extra']['msm_mpts']['e2e']['value'])" >> $out 2>&1
echo "## ncu launch list" >> $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_launches_final.csv python bench.py --steps 2 --warmup 1 --blobs-per-step 16 --skip-msm-leg --skip-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02b_launches_final.csv 2>&1 | head -30 >> $out
cat $out
