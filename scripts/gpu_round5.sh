#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/bench_scenes.py cornell ball stress > gpurun_out/scenes.jsonl 2> gpurun_out/scenes.err
cat gpurun_out/scenes.jsonl; tail -5 gpurun_out/scenes.err
