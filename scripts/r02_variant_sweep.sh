#!/bin/bash
# First GPU call of the next round: parity of the opt-in accumulate variants through the whole MSM, then what they are worth.
#   23 = every block on the FP64 pipe, 25 = half of the blocks, 28 = integer kernel with the identity case peeled,
#   34 = 28 + predicated subtractions, 31 = 28 + PP squaring, 33 = both, 35 = 33 + lean loop head (modelled -4.2 %), 29/30 = both squarings
#   (modelled -1.5 %, profiles/r01_accumulate_sass_census.txt); KZGB_ACC_WAVES 1/2/4 = accumulate threads (fewer waves =
#   fewer chunk partials for k_bucket_fix to stitch: 303 k x 14 Fq-mul at 4 waves, a quarter of that at 1).
# Usage (about 10 GPU-minutes):  gpurun --timeout 900 -- 'bash scripts/r02_variant_sweep.sh'
mkdir -p gpurun_out
out=gpurun_out/r02_variant_sweep.txt
: > $out
echo "## parity (xfail-marked tests report XPASS when the variants are exact)" >> $out
timeout 120 python scripts/variant_check.py 0 28 34 31 33 35 32 29 30 23 25 >> $out 2>&1
timeout 600 python -m pytest tests/test_zz_gpu_dfma.py -q -m gpu -rxX 2>&1 | tail -12 >> $out
echo "## isolated 2^19 MSM: total and accumulate kernel (scripts/msm_sweep.py)" >> $out
for w in 4 2 1; do
  for v in 0 28 34 31 33 35 29 30 23 25; do
    echo "# waves=$w variant=$v" >> $out
    KZGB_ACC_WAVES=$w KZGB_ACC_VARIANT=$v timeout 120 python scripts/msm_sweep.py 19 2>&1 | tail -1 >> $out
  done
done
echo "## headline pipeline (bench.py --skip-cpu-baseline): blobs/s resident, e2e" >> $out
for cfg in "4 0" "4 28" "4 31" "4 33" "4 35" "4 29" "2 0" "2 28" "1 0" "1 28" "1 35"; do
  set -- $cfg
  echo "# waves=$1 variant=$2" >> $out
  KZGB_ACC_WAVES=$1 KZGB_ACC_VARIANT=$2 timeout 200 python bench.py --skip-cpu-baseline --steps 8 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['launch_ms_isolated'],3))" >> $out 2>&1
done
cat $out
