#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench_full.json; tail -3 gpurun_out/bench.err
