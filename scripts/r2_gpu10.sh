#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/trace_bench.py 100 22 > gpurun_out/r2_trace_10m.json 2>&1; tail -1 gpurun_out/r2_trace_10m.json
timeout 300 python scripts/trace_bench.py 10 22 > gpurun_out/r2_trace_1m.json 2>&1; tail -1 gpurun_out/r2_trace_1m.json
for f in box gaussian; do for r in 0 64 16 4; do
  echo "filter $f round $r: $(B2_RFILTER=$f B2_ROUND_SPP=$r python scripts/render_once.py cornell 256 2>&1 | tail -1)"
done; done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu10_pytest.log
tail -4 gpurun_out/r2_gpu10_pytest.log
