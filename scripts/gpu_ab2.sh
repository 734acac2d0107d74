#!/bin/bash
for v in "" _sp; do
  export B2MTS_LIB=$PWD/mitsuba_b200/libb2mts$v.so
  python bench.py --steps 3 --warmup 3 --spp 512 --no-cpu-baseline --no-traversal --no-volpath 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v', round(d['value'],1), {k: round(v,1) for k,v in d['roofline']['kernel_ms'].items()})"
done
B2MTS_LIB=$PWD/mitsuba_b200/libb2mts_sp.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 180 -k "cornell or image or material" 2>&1 | tail -3
