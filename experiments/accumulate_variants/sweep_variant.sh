#!/bin/bash
# accumulate-kernel register caps on the headline config (KZGB_ACC_VARIANT), one JSON line per setting
out=gpurun_out/sweep_variant.jsonl
: > $out
run() { echo "## $*" >> $out; env "$@" timeout 120 python bench.py --skip-cpu-baseline --steps 8 --warmup 3 < /dev/null >> $out 2>> gpurun_out/sweep_variant.err; }
for v in "$@"; do run KZGB_ACC_VARIANT=$v; done
python - <<'PY'
import json
lab=None
for l in open('gpurun_out/sweep_variant.jsonl'):
    l=l.strip()
    if l.startswith('##'): lab=l
    elif l.startswith('{'):
        d=json.loads(l); print(lab, round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['launch_ms_isolated'],3), round(d['roofline']['msm_total_ms_isolated'],3), round(d['single_blob_latency_ms']['value'],2))
PY
