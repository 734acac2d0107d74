#!/bin/bash
# A/B: paired flat-leaf records (Cornell) and the leaf-vote traversal loop (BVH scenes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B2_VERBOSE=1 python bench.py --steps 2 --warmup 3 --spp 256 --no-cpu-baseline --no-traversal 2> gpurun_out/q.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('quads', d['value'], d['kernel_ms'] if 'kernel_ms' in d else d.get('roofline'))"
grep -m1 "b2mts" gpurun_out/q.err
B2_NO_QUADS=1 python bench.py --steps 2 --warmup 3 --spp 256 --no-cpu-baseline --no-traversal 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('noquads', d['value'])"
for lv in 1 4 8 12 16 24; do
  echo "== leafVote $lv refill 16"; B2_LEAFVOTE=$lv python scripts/bench_scenes.py ball stress 2>&1 | grep -v unsorted | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if 'scene' in d: print(d['scene'], d['msamples_s'], d['kernel_ms'])
    else: print(d['trace'], d['mode'], d['mrays_s'], d['frac_of_hbm'])
"
done
for rf in 8 24; do
  echo "== leafVote 8 refill $rf"; B2_REFILL=$rf B2_LEAFVOTE=8 python scripts/bench_scenes.py stress 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if 'scene' in d: print(d['scene'], d['msamples_s'], d['kernel_ms'])
    else: print(d['trace'], d['mode'], d['mrays_s'], d['frac_of_hbm'])
"
done
