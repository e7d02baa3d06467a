#!/bin/bash
# Round 2, first GPU call: which integer accumulate variant wins at 2^19, and how many waves.
#   0 = production (k_accumulate_relaxed), 28 = identity case peeled, 34 = 28 + predicated subtractions, 31 = 28 + PP squaring,
#   33 = 31 + 34, 35 = 33 + lean loop head, 29/30 = both squarings at 4/3 blocks per SM.
# Usage:  gpurun --timeout 900 -- 'bash scripts/r02_variant_sweep.sh'
mkdir -p gpurun_out
out=gpurun_out/r02_variant_sweep.txt
: > $out
echo "## parity" >> $out
timeout 200 python scripts/variant_check.py 0 28 34 31 33 35 29 30 >> $out 2>&1
echo "## isolated 2^19 MSM: total and accumulate kernel (scripts/msm_sweep.py)" >> $out
for w in 4 2 1; do
  for v in 0 28 34 31 33 35 29 30; do
    echo "# waves=$w variant=$v" >> $out
    KZGB_ACC_WAVES=$w KZGB_ACC_VARIANT=$v timeout 120 python scripts/msm_sweep.py 19 2>&1 | tail -1 >> $out
  done
done
echo "## headline pipeline (bench.py --skip-cpu-baseline): blobs/s resident, e2e" >> $out
for cfg in "4 0" "4 28" "4 33" "4 35" "4 29" "2 0" "2 35" "1 0" "1 35"; do
  set -- $cfg
  echo "# waves=$1 variant=$2" >> $out
  KZGB_ACC_WAVES=$1 KZGB_ACC_VARIANT=$2 timeout 200 python bench.py --skip-cpu-baseline --steps 8 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['launch_ms_isolated'],3))" >> $out 2>&1
done
cat $out
