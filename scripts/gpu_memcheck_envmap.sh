#!/bin/bash
# memory checker over the environment-map kernels (probes, images, scene-file route) + the new instanced / thin-lens test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_envmap.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_env7_memcheck.log python -m pytest tests/test_gpu_envmap.py -q -p no:cacheprovider -x 2>&1 | tail -4
echo "memcheck rc=$?"
tail -5 gpurun_out/r2_env7_memcheck.log
