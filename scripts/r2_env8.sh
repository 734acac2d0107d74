#!/bin/bash
# textured plastic + parallel commit phases: whole GPU suite, commit timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_env8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_env8_pytest.log
tail -30 gpurun_out/r2_env8_pytest.log
B2_COMMIT_TIMING=1 B2_NINST=100 python scripts/render_once.py stress 1 256 2>&1 | grep -E "b2 commit|Msamples"
