#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/trace_bench.py 100 22 > gpurun_out/r2_trace_10m_v2.json 2>&1; tail -1 gpurun_out/r2_trace_10m_v2.json
timeout 300 python scripts/trace_bench.py 10 22 > gpurun_out/r2_trace_1m_v2.json 2>&1; tail -1 gpurun_out/r2_trace_1m_v2.json
for lv in 4 12 16 24; do echo "leafVote $lv: $(B2_LEAFVOTE=$lv timeout 300 python scripts/trace_bench.py 100 22 2>&1 | tail -1 | cut -c1-330)"; done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu13_pytest.log
tail -4 gpurun_out/r2_gpu13_pytest.log
