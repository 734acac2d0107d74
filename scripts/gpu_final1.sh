#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('N=1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'trav', round(d['traversal']['frac'],3), 'vol', round(d['volpath']['value'],1), 'cpu', round(d['cpu_baseline']['value'],2), d['clocks'])"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | cut -c1-300
