#!/usr/bin/env python3
"""The isolated traversal measurement of bench.py (10 M flattened triangles, incoherent rays) as a stand-alone program: ncu target.
usage: trace_bench.py [n_instances=100] [n_rays_log2=22]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mitsuba_b200 import api
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 22
ctx = api.Context(0)
print(json.dumps(bench.traversal_metric(ctx, bench.measured_peaks()[0].get("hbm_gbs", 6650.0), n_inst=n_inst, n_rays=1 << lg)))
