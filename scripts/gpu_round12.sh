#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/vol_tests.log 2>&1; tail -4 gpurun_out/vol_tests.log
timeout 600 python scripts/bench_scenes.py smoke 2>&1 | tail -6
