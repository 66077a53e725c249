"""developer tool (GPU box): the compute_source_cfg4 leg of bench.py alone (6.55 M points x NLM 256)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from at3d_b200 import backend as B
peak, src = bench.measured_peak()
print(json.dumps(bench.compute_source_cfg4_leg(B, 5, 2, peak)))
