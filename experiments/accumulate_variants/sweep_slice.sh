#!/bin/bash
# bucket-reduce slice size on the headline config (KZGB_SLICE), one JSON line per setting
out=gpurun_out/sweep_slice.jsonl
: > $out
run() { echo "## $*" >> $out; env "$@" timeout 120 python bench.py --skip-cpu-baseline --steps 8 --warmup 3 < /dev/null >> $out 2>> gpurun_out/sweep_slice.err; }
run KZGB_SLICE=4
run KZGB_SLICE=8
run KZGB_SLICE=16
run KZGB_SLICE=32
python - <<'PY'
import json
lab=None
for l in open('gpurun_out/sweep_slice.jsonl'):
    l=l.strip()
    if l.startswith('##'): lab=l
    elif l.startswith('{'):
        d=json.loads(l); print(lab, round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['launch_ms_isolated'],3), round(d['roofline']['msm_total_ms_isolated'],3), round(d['single_blob_latency_ms']['value'],2))
PY
