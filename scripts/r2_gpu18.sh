#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/trace_bench.py 100 22 2>&1 | tail -1 | cut -c1-330
timeout 300 python scripts/trace_bench.py 10 22 2>&1 | tail -1 | cut -c1-330
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu18_pytest.log
tail -4 gpurun_out/r2_gpu18_pytest.log
for s in c3 stress; do echo "$s: $(python scripts/render_once.py $s 64 2>&1 | tail -1)"; done
