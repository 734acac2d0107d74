#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for r in 8 12 16 24 32; do
  echo "=== B2_REFILL=$r"
  B2_REFILL=$r python scripts/bench_scenes.py ball stress 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'scene' in d: print(d['scene'], d['msamples_s'], d['kernel_ms'])
    else: print(d['trace'], d['mode'], d['mrays_s'], d['node_visits_per_ray'], d['frac_of_hbm'])
"
done
