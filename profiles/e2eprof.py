"""developer tool (GPU box): wall clock of the host-buffer gradient call (bench.py's e2e leg), call by call."""
import os, sys, time, argparse
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from at3d_b200 import backend as B, gradsetup
from at3d_b200.device import DeviceState
import torch
args = argparse.Namespace(workload='cfg2', pixels=0)
sc, rays, cfg = bench.build_scene(args)
B.finalize_scene(sc)
st = sc.state
gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
dev = DeviceState(st)
dev.attach_gradient(gi)
rad = dev.render(rays)
pix = gradsetup.make_pixels(st.nstokes, rays.nrays, rad, seed=1)
for k in ('camx', 'camy', 'camz', 'cammu', 'camphi'):
    setattr(rays, k, torch.from_numpy(np.ascontiguousarray(getattr(rays, k))).pin_memory().numpy())
ts = []
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = dev.gradient(rays, pix, timing=True)
    torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    kms = out[-1]
print('OMP', os.environ.get('OMP_NUM_THREADS'), 'cores', len(os.sched_getaffinity(0)), 'wall ms', ' '.join('%.1f' % t for t in ts), '| kernels %.1f' % kms[3])
dev.close()
