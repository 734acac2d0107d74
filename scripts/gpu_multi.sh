#!/bin/bash
# usage: r2_multi.sh N [steps]
N=$1; STEPS=${2:-5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 3 --no-traversal > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('weak', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('strong', d.get('strong'))
print({k:(v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d.get('configs',{}).items()})
PY
