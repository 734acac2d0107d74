#!/bin/bash
# texture feature validation + full regression in one box session
mkdir -p gpurun_out; rm -f gpurun_out/tex_diag.jsonl
B2_TEST_DIAG=gpurun_out/tex_diag.jsonl timeout 400 python -m pytest tests/test_gpu_texture.py -q --timeout 180 2>&1 | grep -v "^E  *+\|^E  *where\|^E  *and " > gpurun_out/tex_tests.log; tail -6 gpurun_out/tex_tests.log
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('N=1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'trav', round(d['traversal']['frac'],3), 'vol', round(d['volpath']['value'],1), 'tex', round(d['textured'].get('value', -1), 1), 'cpu', round(d['cpu_baseline']['value'],2), d['clocks'])"; tail -3 gpurun_out/bench.err
