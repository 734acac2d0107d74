#!/bin/bash
for v in "" _pf; do
  export B2MTS_LIB=$PWD/mitsuba_b200/libb2mts$v.so
  python bench.py --steps 3 --warmup 3 --spp 512 --no-cpu-baseline --no-traversal --no-volpath 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v', round(d['value'],1), {k: round(v,1) for k,v in d['roofline']['kernel_ms'].items()})"
done
unset B2MTS_LIB
python scripts/bench_scenes.py cornell 2>&1 | head -3 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['scene'], d['rfilter'], d.get('sampler',''), d['msamples_s'], d['kernel_ms'])"
