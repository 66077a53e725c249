"""developer tool (GPU box): phase times of the cfg2 gradient step with device-resident inputs (as bench.py's value leg)."""
import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import bench
from at3d_b200 import backend as B, gradsetup
from at3d_b200.device import DeviceState
class A: pass
a = A(); a.workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'; a.pixels = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
st = sc.state
dev = DeviceState(st)
gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
dev.attach_gradient(gi)
rad = dev.render(rays)
pix = gradsetup.make_pixels(st.nstokes, rays.nrays, rad, seed=1)
class Bag: pass
dr, dp = Bag(), Bag()
for k in ('camx', 'camy', 'camz', 'cammu', 'camphi'):
    setattr(dr, k, torch.from_numpy(getattr(rays, k)).cuda())
dp.measurements = torch.from_numpy(np.ascontiguousarray(pix.measurements.T)).cuda()
dp.uncertainties = torch.from_numpy(np.ascontiguousarray(pix.uncertainties.transpose(2, 1, 0))).cuda()
dp.rays_per_pixel = torch.from_numpy(pix.rays_per_pixel).cuda()
dp.ray_weights = torch.from_numpy(pix.ray_weights).cuda()
dp.stokes_weights = torch.from_numpy(np.ascontiguousarray(pix.stokes_weights.T)).cuda()
gout = torch.zeros((gi.numder, gi.maxpg), dtype=torch.float64, device='cuda')
sout = torch.zeros((pix.npix, st.nstokes), dtype=torch.float32, device='cuda')
cout = torch.zeros(1, dtype=torch.float64, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
ms = []
for i in range(n + 2):
    out = dev.gradient(dr, dp, gradout=gout, stokesout=sout, cost=cout, stream=stream, timing=True)
    if i >= 2: ms.append(out[-1])
if os.environ.get('GRADBENCH_EACH'):
    print('forward per call:', ' '.join('%.1f' % x[0] for x in ms), flush=True)
m = np.mean(np.array(ms), axis=0)
print(os.environ.get('AT3D_B200_LIB', 'default').split('/')[-1], 'forward %.2f deriv %.2f (weights %.2f pairs %.2f) beam %.2f total %.2f' % (m[0], m[1], m[4], m[5], m[2], m[3]), flush=True)
