#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu11_pytest.log
tail -4 gpurun_out/r2_gpu11_pytest.log
for f in box gaussian; do echo "filter $f: $(B2_RFILTER=$f python scripts/render_once.py cornell 1024 2>&1 | tail -1)"; done
