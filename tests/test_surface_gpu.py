"""GPU RENDER over general BRDF surfaces and with thermal sources, through the C ABI.

The solved states come from the oracle's fixed-grid solver (test infrastructure); the GPU renders them and is
compared (a) with the oracle's RENDER at the north-star tolerance (rtol 1e-4, abs 1e-6 on near-zero Q/U) and
(b) directly with SHDOM's own verification outputs (tests/golden/brdf_*1r.out) at the reference's tolerance."""
import os
import numpy as np
import pytest
import oracle_lib as O
import shdom_verification as V
import scenes
from at3d_b200 import synthetic as S
from at3d_b200.device import DeviceState

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_cache = {}


def solved(kind):
    if kind not in _cache:
        st, pg, wtmu = V.make_state(O, kind)
        st.nscatangle = 721
        st.phasetab = O.precompute_phase_check(pg.legenp, 721, st.nstokes, st.ml, True)
        _cache[kind] = O.solve_fixed_grid(st, wtmu, solacc=1e-5)[0]
    return _cache[kind]


def close(a, b, rtol=1e-4, atol=1e-6):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('kind', ['L', 'O', 'R', 'W', 'D'])
def test_render_brdf_surface_vs_oracle_and_shdom(kind):
    sol = solved(kind)
    rays = V.sensor_rays()
    ref = O.render(sol, rays)
    dev = DeviceState(sol)
    out = dev.render(rays)
    dev.close()
    close(out, ref)
    gold = V.parse_shdom_output(os.path.join(GOLD, 'brdf_%s1r.out' % kind))
    np.testing.assert_allclose(out[0], gold[:, 2], rtol=0, atol=9e-6)
    if sol.nstokes == 3:
        np.testing.assert_allclose(out[1], gold[:, 3], rtol=0, atol=9e-6)
        np.testing.assert_allclose(out[2], gold[:, 4], rtol=0, atol=1e-8)


@pytest.mark.parametrize('kind', ['O', 'W'])
def test_render_brdf_oblique_rays_and_nosurface(kind):
    # rays that are not aligned with the columns, from inside and above the domain, device-resident too
    sol = solved(kind)
    rng = np.random.default_rng(5)
    n = 600
    from at3d_b200.state import Rays
    rays = Rays(rng.uniform(0, 1.0, n), np.zeros(n), rng.uniform(2.0, 40.0, n),
                rng.uniform(0.05, 1.0, n), rng.uniform(0, 2 * np.pi, n))
    dev = DeviceState(sol)
    close(dev.render(rays), O.render(sol, rays))
    close(dev.render(rays, nosurface=True), O.render(sol, rays, nosurface=True))
    import torch
    t = lambda a: torch.from_numpy(a).cuda()
    class R: pass
    r = R(); r.camx, r.camy, r.camz, r.cammu, r.camphi = t(rays.camx), t(rays.camy), t(rays.camz), t(rays.cammu), t(rays.camphi)
    out = dev.render(r).cpu().numpy().T
    close(out, O.render(sol, rays), rtol=2e-4)
    dev.close()


def test_gradient_refuses_general_brdf():
    # the reference stops in SURFACE_BRDF_GRAD for these surfaces (src/surface.f:395-399); the library returns code 3
    from at3d_b200.state import GradInputs
    dev = DeviceState(solved('O'))
    with pytest.raises(Exception) as e:
        dev.attach_gradient(GradInputs(numder=1, maxpg=1))
    assert 'Lambertian' in str(e.value)
    dev.close()


@pytest.mark.parametrize('srctype,variable', [('T', False), ('B', False), ('T', True), ('B', True)])
def test_render_thermal_sources(srctype, variable):
    sc = S.make_scene(nx=7, ny=6, nz=8, nstokes=1, nsplits=3, seed=11, variable_sfc=variable)
    O.finalize_scene(sc)
    st = sc.state
    st.srctype = srctype
    st.units = 'R'
    st.wavelen = 10.5
    st.gndtemp = 291.0
    st.skyrad[...] = 2.7 + 250.0 * (st.skyrad / st.skyrad.max())      # sky "temperatures" for SRCTYPE='T'
    if variable:
        st.sfcgridparms[0] = 7.0 + np.arange(st.nbotpts) * 0.01        # Planck function of the surface temperature
    rays = scenes.ray_set(sc)
    ref = O.render(st, rays)
    dev = DeviceState(st)
    out = dev.render(rays)
    dev.close()
    assert np.abs(ref).max() > 1e-3
    close(out, ref)


def test_core_render_keyword_api_with_ocean_surface():
    # at3d.core.render(**kwargs) as solver.RTE.integrate_to_sensor calls it (at3d/solver.py:681-759), SFCTYPE='VO'
    from at3d_b200 import core
    from test_core_api_gpu import solver_kwargs
    sol = solved('O')
    rays = V.sensor_rays()
    kw = solver_kwargs(sol)
    kw.update(camx=rays.camx, camy=rays.camy, camz=rays.camz, cammu=rays.cammu, camphi=rays.camphi, npix=rays.nrays,
              nosurface=False, correctinterpolate=True, singlescatter=False, sfcgridrad=sol.sfcgridrad)
    assert kw['sfctype'] == 'VO'
    bcrad, stokes, ierr, errmsg = core.render(**kw)
    assert ierr == 0, errmsg
    gold = V.parse_shdom_output(os.path.join(GOLD, 'brdf_O1r.out'))
    np.testing.assert_allclose(stokes[0], gold[:, 2], rtol=0, atol=9e-6)
    core.clear_cache()


@pytest.mark.parametrize('kind,nstokes,bc', [('O', 1, 'periodic'), ('R', 1, 'open'), ('M', 1, 'periodic'), ('L', 3, 'open'),
                                             ('W', 3, 'periodic'), ('D', 3, 'open')])
def test_render_brdf_surface_3d_adaptive_grid(kind, nstokes, bc):
    # 3-D scene with split cells: rays reach the surface obliquely and through open-boundary cells
    sc = S.make_scene(nx=7, ny=6, nz=8, nstokes=nstokes, bc=bc, nsplits=6, seed=21, variable_sfc=True)
    O.finalize_scene(sc)
    st = S.with_brdf_surface(sc.state, kind, seed=4, wavelen=0.55)
    rays = scenes.ray_set(sc)
    ref, tr_ref, _ = O.render(st, rays, trace_cap=64)
    dev = DeviceState(st)
    out, tr = dev.render(rays, trace_cap=64)
    dev.close()
    assert np.array_equal(tr['ncells'], tr_ref['ncells']) and np.array_equal(tr['cells'], tr_ref['cells'])
    close(out, ref)
    lamb = O.render(sc.state, rays)
    assert np.abs(ref - lamb).max() > 1e-3          # the surface does matter for these rays
