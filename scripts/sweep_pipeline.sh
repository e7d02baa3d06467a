#!/bin/bash
# Sweep of the blob-batch pipeline knobs (accumulate waves, lanes) on the headline config; one JSON line per setting.
out=gpurun_out/sweep_pipeline.jsonl
: > $out
run() { echo "## $*" >> $out; env "$@" timeout 120 python bench.py --skip-cpu-baseline --steps 6 --warmup 3 < /dev/null >> $out 2>> gpurun_out/sweep_pipeline.err; }
run KZGB_ACC_WAVES=2
run KZGB_ACC_WAVES=3
run KZGB_ACC_WAVES=6
run KZGB_ACC_WAVES=8
run KZGB_LANES=3
run KZGB_LANES=5
run KZGB_LANES=6
run KZGB_LANES=6 KZGB_ACC_WAVES=8
python - <<'PY'
import json
lab=None
for l in open('gpurun_out/sweep_pipeline.jsonl'):
    l=l.strip()
    if l.startswith('##'): lab=l
    elif l.startswith('{'):
        d=json.loads(l); print(lab, round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['launch_ms_isolated'],3), round(d['roofline']['msm_total_ms_isolated'],3))
PY
