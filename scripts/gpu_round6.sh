#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python scripts/bench_scenes.py ball stress > gpurun_out/scenes2.jsonl 2> gpurun_out/scenes.err
cat gpurun_out/scenes2.jsonl | cut -c1-420; tail -5 gpurun_out/scenes.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2>> gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cur.json"))
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", {k:round(v,1) for k,v in r["kernel_ms"].items()}, "ms/step", round(d["ms_per_step"],1), r["kernel"], round(r["achieved"],1), round(r["frac"],3))
PY
