"""developer tool (GPU box): SH_TO_DO FP32-FMA vs tensor-core (3xTF32 tcgen05) variants on a workload's SOURCE."""
import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import bench
from at3d_b200 import backend as B
class A: pass
a = A(); a.workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'; a.pixels = 8
sc, rays, cfg = bench.build_scene(a)
B.finalize_scene(sc)
st = sc.state
delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
wtmu = (st.wtdo[:, 0] / delphi).astype(np.float32)
res = {}
outs = {}
for v in ('fp32', 'tc'):
    os.environ['AT3D_B200_TRANSFORM'] = v
    ms = []
    for i in range(6):
        do, t = B.sh_to_do(st, wtmu, st.shptr, st.source, timing=True)
        if i >= 2: ms.append(t)
    res[v] = float(np.mean(ms)); outs[v] = do
nang = int(st.nphi0.sum())
resb, outb = {}, {}
for v in ('fp32', 'tc'):
    os.environ['AT3D_B200_TRANSFORM'] = v
    ms = []
    for i in range(6):
        sh, t = B.do_to_sh(st, wtmu, st.rshptr, outs['fp32'], timing=True)
        if i >= 2: ms.append(t)
    resb[v] = float(np.mean(ms)); outb[v] = sh
errb = float(np.abs(outb['tc'] - outb['fp32']).max() / np.abs(outb['fp32']).max())
print(json.dumps(dict(do_to_sh_ms=resb, max_rel_diff=errb)))
err = float(np.abs(outs['tc'] - outs['fp32']).max() / np.abs(outs['fp32']).max())
dense = 2.0 * st.npts * st.nlm * nang
print(json.dumps(dict(npts=int(st.npts), nang=nang, ms=res, max_rel_diff=err,
                      tc_tensor_tflops_3x=3 * dense / (res['tc'] * 1e-3) / 1e12)))
