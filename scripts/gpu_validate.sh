#!/bin/bash
# Round-2 validation at HEAD: whole GPU suite, then the default bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_validate_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_validate_pytest.log
tail -8 gpurun_out/gpu_validate_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/gpu_validate_bench.json 2> gpurun_out/gpu_validate_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/gpu_validate_bench.json
