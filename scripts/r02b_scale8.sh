#!/bin/bash
# 8-GPU call: the box's topology and host SHA-256 rates, then the default bench line (64 blobs per step, config-4 leg included) and
# the 192-blob step.
out=gpurun_out/r02b_scale8.txt
: > $out
{ echo "## topology"; nproc; lscpu | grep -E "Model name|Thread|Core|Socket|^CPU\(s\)|NUMA"; } >> $out 2>&1
g++ -O3 -std=c++17 -pthread -o /tmp/hb scripts/host_sha_mb_bench.cpp rust-kzg-bn254_b200/csrc/sha256.cpp && /tmp/hb 16 4 16 32 >> $out 2>&1
run() {  # name, extra args
  name=$1; shift
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 100)) bench.py --gpus 8 "$@" \
    > gpurun_out/r02b_bench_n8_$name.json 2> gpurun_out/r02b_bench_n8_$name.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02b_bench_n8_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), (d['extra'].get('batch16') or {}).get('value'), ((d['extra'].get('msm_mpts') or {}).get('value'), ((d['extra'].get('msm_mpts') or {}).get('e2e') or {}).get('value')))" >> $out 2>&1
}
echo "## bench --gpus 8: value e2e ms_per_step batch16 (msm_mpts, e2e)" >> $out
run b64 --steps 8 --warmup 3
run b192 --steps 4 --warmup 3 --blobs-per-step 192 --skip-msm-leg
run b128 --steps 5 --warmup 3 --blobs-per-step 128 --skip-msm-leg
cat $out
