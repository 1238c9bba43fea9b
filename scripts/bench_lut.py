"""LUT side bench alone (bench.py's lut_side_bench): python scripts/bench_lut.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mkids_sdr_b200 import _lib
out = bench.lut_side_bench(_lib.default_context(0))
for k, v in out.items():
    print(k, json.dumps(v))
