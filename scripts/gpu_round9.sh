#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench_full.json; tail -3 gpurun_out/bench.err
