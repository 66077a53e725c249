"""developer tool (GPU box): the cfg2 gradient step a few times, for `ncu -k regex:gwalk|apply|weights` captures."""
import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import bench
from at3d_b200 import backend as B, gradsetup
from at3d_b200.device import DeviceState
class A: pass
a = A(); a.workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'; a.pixels = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
dev = DeviceState(sc.state)
gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
dev.attach_gradient(gi)
rad = dev.render(rays)
pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=1)
import torch
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()      # ncu --profile-from-start off: the last call only
    res = dev.gradient(rays, pix, timing=True)
    if i == n - 1:
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
    print(json.dumps(dict(ms=res[-1])), flush=True)
