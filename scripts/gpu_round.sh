#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list + full capture of the four wavefront kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1200 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --res 512 --spp 64 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(extend|shade|occluded|generate)' -s 80 -c 8 -o gpurun_out/prof_wave \
    python bench.py --steps 1 --warmup 3 --res 512 --spp 64 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
