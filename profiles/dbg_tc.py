import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, scenes, oracle_lib as O
from at3d_b200 import backend as B
st = scenes.make('scalar_periodic_split', O).state
delphi = np.float32(2.0*np.pi)/st.nphi0.astype(np.float32)
w = (st.wtdo[:,0]/delphi).astype(np.float32)
fp = B.sh_to_do(st, w, st.shptr, st.source)
os.environ['AT3D_B200_TRANSFORM']='tc'
tc = B.sh_to_do(st, w, st.shptr, st.source)
print(fp.shape, st.npts, int(st.nphi0.sum()), st.nlm)
bad = np.abs(tc-fp) > 1e-4*np.abs(fp).max()
idx = np.argwhere(bad)
print(len(idx))
for ax in range(idx.shape[1]):
    u = np.unique(idx[:,ax]); print('axis',ax,'n',len(u), u[:20], u[-5:])
