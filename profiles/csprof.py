"""developer tool: two COMPUTE_SOURCE calls on a bench workload (for ncu -k regex:cs_)"""
import sys, os, types, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from at3d_b200 import backend as B
a = types.SimpleNamespace(workload=sys.argv[1] if len(sys.argv) > 1 else 'cfg4s', pixels=8)
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
print(json.dumps(bench.compute_source_leg(B, sc.state, 2, 1)))
