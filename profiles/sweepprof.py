import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import bench
from at3d_b200 import backend as B
class A: pass
a = A(); a.workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'; a.pixels = 8
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
print(json.dumps(bench.sweep3d_leg(sc.state, 2, 1)), flush=True)
