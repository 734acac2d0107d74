#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/f3_tests.log 2>&1; tail -6 gpurun_out/f3_tests.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traversal --no-volpath 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cornell', d['value'], d['roofline']['kernel_ms'])"
