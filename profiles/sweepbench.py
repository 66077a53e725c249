import sys, os, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import bench
from at3d_b200 import backend as B
class A: pass
a = A(); a.workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'; a.pixels = 8
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
for env in [dict(AT3D_SWEEP_LEVELS='1'), dict(AT3D_SWEEP_LEVELS='0'), dict(AT3D_SWEEP_LEVELS='0', AT3D_SWEEP_GROUP='177'),
            dict(AT3D_SWEEP_LEVELS='0', AT3D_SWEEP_GROUP='4'), dict(AT3D_SWEEP_LEVELS='1', AT3D_SWEEP_GROUP='30')]:
    for k in ('AT3D_SWEEP_LEVELS', 'AT3D_SWEEP_GROUP'):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(env, json.dumps(bench.sweep3d_leg(sc.state, 3, 1)), flush=True)
