import sys, json, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from at3d_b200 import backend as B
class A: pass
a = A(); a.workload = sys.argv[1]; a.pixels = 8
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
print(json.dumps(dict(tr=bench.transform_leg(B, sc.state, 3, 1), cs=bench.compute_source_leg(B, sc.state, 3, 1))))
