"""developer tool (GPU box): the cfg4-size COMPUTE_SOURCE leg of bench.py alone."""
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import bench
from at3d_b200 import backend as B
peak, _ = bench.measured_peak()
dims = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else [256, 256, 100]
print(json.dumps(bench.compute_source_cfg4_leg(B, 5, 2, peak, *dims)))
