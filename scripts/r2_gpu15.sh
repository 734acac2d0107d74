#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gpu15_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu15_pytest.log
tail -6 gpurun_out/r2_gpu15_pytest.log
timeout 300 python scripts/trace_bench.py 100 22 2>&1 | tail -1 | cut -c1-300
