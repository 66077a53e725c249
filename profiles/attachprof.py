"""developer tool (GPU box): where the host time of DeviceState(...) + attach_gradient goes (cfg2)."""
import os, sys, time, argparse, cProfile, pstats, io
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from at3d_b200 import backend as B, gradsetup
from at3d_b200.device import DeviceState
args = argparse.Namespace(workload='cfg2', pixels=0)
sc, rays, cfg = bench.build_scene(args)
B.finalize_scene(sc)
st = sc.state
gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
B.memory_reuse(True)
for rep in range(3):
    t0 = time.perf_counter(); dev = DeviceState(st); t1 = time.perf_counter(); dev.attach_gradient(gi); t2 = time.perf_counter(); dev.close()
    print('create %.1f ms attach %.1f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
dev = DeviceState(st)
pr = cProfile.Profile(); pr.enable(); dev.attach_gradient(gi); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(14); print(s.getvalue()[:3000])
os.environ['AT3D_B200_ATTACH_TIMING'] = '1'
dev.attach_gradient(gi)
dev.close()
