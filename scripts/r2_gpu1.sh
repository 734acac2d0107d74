#!/bin/bash
# round 2, GPU call 1: where does HEAD stand?  Full -m gpu suite (no -x), parity probe, baseline bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu1_smi.txt 2>&1
nproc > gpurun_out/r2_gpu1_nproc.txt; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count()); print(open('/sys/fs/cgroup/cpu.max').read() if os.path.exists('/sys/fs/cgroup/cpu.max') else 'no cpu.max')" >> gpurun_out/r2_gpu1_nproc.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gpu1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu1_pytest.log
timeout 600 python scripts/parity_probe.py > gpurun_out/r2_gpu1_parity_probe.json 2> gpurun_out/r2_gpu1_parity_probe.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-volpath > gpurun_out/r2_gpu1_bench.json 2> gpurun_out/r2_gpu1_bench.err
tail -5 gpurun_out/r2_gpu1_pytest.log
