"""Shared synthetic scenes and ray sets for the parity tests (seeded, oracle finishes in seconds)."""
import numpy as np
from at3d_b200 import synthetic as S


def ray_set(scene, n_persp=10, res=0.021):
    m = scene.meta
    cx, cy = 0.5 * m['xmax'], 0.5 * m['ymax']
    r = []
    r.append(S.orthographic_rays(scene, 0.0, 0.0, res)[0])                 # nadir (cx=cy=0 exactly)
    r.append(S.orthographic_rays(scene, 35.0, 25.0, res)[0])               # oblique
    r.append(S.orthographic_rays(scene, 60.0, 200.0, 1.7 * res)[0])        # oblique, other octant
    r.append(S.perspective_rays((cx, cy, 3.0), (cx, cy * 1.1, 0.1), 12.0, n_persp, n_persp)[0])
    r.append(S.perspective_rays((cx - 0.9, cy + 0.4, 1.2), (cx, cy, 0.15), 25.0, n_persp, n_persp)[0])
    r.append(S.perspective_rays((cx, cy, 0.0), (cx * 1.1, cy * 0.9, 0.3), 50.0, 6, 6)[0])   # up-looking
    r.append(S.perspective_rays((cx, cy, 0.5 * m['zmax']), (cx + 0.1, cy + 0.2, 0.55 * m['zmax']), 60.0, 5, 5)[0])  # in-cloud
    return S.concat_rays(r)


SCENE_CASES = {
    'scalar_periodic': dict(nx=8, ny=7, nz=9, nstokes=1, bc='periodic', nsplits=0, seed=1),
    'scalar_periodic_split': dict(nx=8, ny=7, nz=9, nstokes=1, bc='periodic', nsplits=12, seed=2),
    'scalar_open_split': dict(nx=7, ny=8, nz=10, nstokes=1, bc='open', nsplits=10, seed=3),
    'scalar_nmu16': dict(nx=6, ny=6, nz=8, nmu=16, nphi=32, nstokes=1, bc='periodic', nsplits=4, seed=4),
    'polarized_periodic_split': dict(nx=7, ny=7, nz=8, nstokes=3, bc='periodic', nsplits=8, seed=5),
    'polarized_open': dict(nx=6, ny=7, nz=8, nstokes=3, bc='open', nsplits=0, seed=6),
    'rayleigh_two_species': dict(nx=7, ny=6, nz=9, nstokes=1, bc='periodic', rayleigh=True, nsplits=5, seed=7),
    'polarized_rayleigh_varsfc': dict(nx=6, ny=6, nz=8, nstokes=3, bc='open', rayleigh=True, nsplits=4,
                                      variable_sfc=True, seed=8),
    'scalar_no_deltam': dict(nx=7, ny=7, nz=8, nstokes=1, bc='periodic', deltam=False, nsplits=4, seed=10),
    'polarized_rayleigh_no_deltam': dict(nx=6, ny=6, nz=7, nstokes=3, bc='open', deltam=False, rayleigh=True,
                                         nsplits=3, seed=12),
    'thick_transcut': dict(nx=8, ny=8, nz=10, nstokes=1, bc='periodic', ext_max=120.0, cloud='slab',
                           nsplits=3, seed=9),
}


def make(case, oracle, **overrides):
    sc = S.make_scene(**dict(SCENE_CASES[case], **overrides))
    oracle.finalize_scene(sc)
    return sc
