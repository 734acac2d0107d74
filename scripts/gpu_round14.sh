#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/inst_tests.log 2>&1; tail -6 gpurun_out/inst_tests.log
