#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_volpath.py tests/test_gpu_xml.py -m gpu -q --timeout 180 2>&1 | tail -4
echo "== lockstep+partition"; SMOKE_ORACLE=0 timeout 300 python scripts/bench_scenes.py smoke 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['scene'], d['msamples_s'], d['ms'], d['iters'])"
