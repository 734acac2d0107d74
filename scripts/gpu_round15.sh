#!/bin/bash
mkdir -p gpurun_out
B2_VOL_LOCKSTEP=1 timeout 600 python -m pytest tests/test_gpu_volpath.py tests/test_gpu_xml.py -m gpu -q --timeout 180 2>&1 | tail -8
for ls in 0 1; do
echo "== lockstep $ls"; B2_VOL_LOCKSTEP=$ls SMOKE_ORACLE=0 timeout 300 python scripts/bench_scenes.py smoke 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['scene'], d['msamples_s'], d['ms'])"
done
