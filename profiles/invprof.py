"""Where the host time of one inversion evaluation goes (cfg2): state upload, derivative tables, first / later gradient calls."""
import os, sys, time, argparse
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from at3d_b200 import backend as B, gradsetup, solver
from at3d_b200.device import DeviceState
import torch
args = argparse.Namespace(workload='cfg2', pixels=0)
sc, rays, cfg = bench.build_scene(args)
B.finalize_scene(sc)
st = sc.state
gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
dev = DeviceState(st); rad = dev.render(rays); dev.close()
pix = gradsetup.make_pixels(st.nstokes, rays.nrays, rad, seed=1)
wtmu = (st.wtdo[:, 0] / (np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32))).astype(np.float32)
import sys
B.memory_reuse(len(sys.argv) < 2 or sys.argv[1] != 'plain')
sv = solver.SweepSolver(st.copy().normalize(), wtmu)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
sol = None
for rep in range(3):
    t0 = T(); sol, iters, solcrit, tm = sv.solve(solacc=1e-4, maxiter=60, initial=(sol if rep and 'warm' in sys.argv else None))
    t1 = T(); dev = DeviceState(sol)
    t2 = T(); dev.attach_gradient(gi)
    t3 = T(); dev.gradient(rays, pix)
    t4 = T(); dev.gradient(rays, pix)
    t5 = T(); dev.gradient(rays, pix)
    t6 = T(); dev.close()
    t7 = T()
    print('rep %d: iters %d solve %.1f (loop %.1f)  state_create %.1f  attach %.1f  grad#1 %.1f  grad#2 %.1f  grad#3 %.1f  close %.1f ms'
          % (rep, iters, 1e3*(t1-t0), tm['loop_ms'], 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(t6-t5), 1e3*(t7-t6)))
sv.close()
