#!/bin/bash
# Second 8-GPU call: rank-0-only in-process clock sampling, host CPU accounting; default line, 192-blob step, 192 with a device share.
out=gpurun_out/r02b_scale8b.txt
: > $out
run() {  # name, extra args
  name=$1; shift
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 100)) bench.py --gpus 8 "$@" \
    > gpurun_out/r02b_bench_n8b_$name.json 2> gpurun_out/r02b_bench_n8b_$name.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02b_bench_n8b_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), (d['extra'].get('batch16') or {}).get('value'), d['host_cpu_ms_per_blob']['resident'], d['clocks'])" >> $out 2>&1
  grep "hash\]" gpurun_out/r02b_bench_n8b_$name.err | awk '{w+=$9; c+=$12; n++} END {if (n) print "  rank-interleaved multi-buffer groups:", n, "mean wall ms", w/n, "mean cpu ms", c/n}' >> $out
}
echo "## bench --gpus 8: value e2e ms_per_step batch16 host_cpu_ms_per_blob clocks" >> $out
run b64 --steps 8 --warmup 3 --skip-msm-leg --skip-cpu-baseline --option hash_trace=1
run b192 --steps 4 --warmup 3 --blobs-per-step 192 --skip-msm-leg --option hash_trace=1
run b192dev64 --steps 4 --warmup 3 --blobs-per-step 192 --skip-msm-leg --option device_hash=64 --option device_hash_lanes=0
run b64mb4 --steps 8 --warmup 3 --skip-msm-leg --option hash_threads=4 --option hash_mb=1
cat $out
