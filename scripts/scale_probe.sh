#!/bin/bash
# 8-GPU probes of the blob-batch leg: bash scripts/scale_probe.sh "B opt=val opt=val" ...   -> gpurun_out/scale_probe.txt
out=gpurun_out/scale_probe.txt
: > $out
port=29600
for cfg in "$@"; do
  set -- $cfg
  B=$1; shift
  opts=""
  for kv in "$@"; do opts="$opts --option $kv"; done
  port=$((port+1))
  echo -n "B=$B $* : " >> $out
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 4 --warmup 2 --blobs-per-step $B --skip-msm-leg $opts 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))" >> $out 2>&1
done
cat $out
