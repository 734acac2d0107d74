#!/bin/bash
# host build threads on the leased box (quota 16 of 128 hardware threads) + the EXR route of the scene-file envmap test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_xml.py -q -p no:cacheprovider -k envmap 2>&1 | tail -3
for t in 16 32 64 128; do
  echo "threads $t"; B2_BUILD_THREADS=$t B2_COMMIT_TIMING=1 B2_NINST=100 python scripts/render_once.py stress 1 256 2>&1 | grep -E "binned SAH|8-wide|relayout|BVH \(world\)|upload"
done
