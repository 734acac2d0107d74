#!/bin/bash
# 4-GPU weak-scaling bench line (torchrun, NCCL film reduce), short form
mkdir -p gpurun_out
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline --no-traversal --no-volpath > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1]); print('N=4', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['n_gpus'], d['ms_per_step'])"; tail -2 gpurun_out/bench_n4.err
