#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/trace_bench.py 100 22 > gpurun_out/r2_trace_10m_bin.json 2>&1; tail -1 gpurun_out/r2_trace_10m_bin.json
timeout 300 python scripts/trace_bench.py 10 22 > gpurun_out/r2_trace_1m_bin.json 2>&1; tail -1 gpurun_out/r2_trace_1m_bin.json
B2_NO_BIN=1 timeout 600 python scripts/trace_bench.py 100 22 > gpurun_out/r2_trace_10m_nobin.json 2>&1; tail -1 gpurun_out/r2_trace_10m_nobin.json
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu12_pytest.log
tail -4 gpurun_out/r2_gpu12_pytest.log
