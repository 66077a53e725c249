"""GPU parity of LEVISAPPROX_GRADIENT (C-ABI at3d_levisapprox_gradient) against the CPU oracle.

Bars (BASELINE.json north_star): visited-cell sequence / sub-interval counts bit-exact; cost, pixel
Stokes vectors and the gradient within relative 1e-4.  The gradient is compared element-wise with
rtol 1e-4 and an absolute floor of 1e-4 x the largest |gradient| of the same unknown (elements that
are sums of cancelling ray contributions carry the rounding of their largest terms)."""
import numpy as np
import pytest
import scenes

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def run_case(oracle, case, gkw, rays_per_pixel=1, trace=True, bright_only=False, mutate=None, mutate_grad=None, **scene_kw):
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup
    from at3d_b200.state import Rays
    sc = scenes.make(case, oracle, **scene_kw)
    if mutate is not None:
        mutate(sc)
    rays = scenes.ray_set(sc, n_persp=7, res=0.035)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, **gkw)
    if mutate_grad is not None:
        mutate_grad(sc, gi)
    rad = oracle.render(sc.state, rays)
    if bright_only:
        # the log cost function (COSTFUNC='LL') needs I > 0 and a non-zero polarized signal
        keep = rad[0] > 0.02 * rad[0].max()
        if rad.shape[0] > 1:
            keep &= np.hypot(rad[1], rad[2]) > 3e-3 * rad[0]
        idx = np.nonzero(keep)[0]
        rays = Rays(rays.camx[idx], rays.camy[idx], rays.camz[idx], rays.cammu[idx], rays.camphi[idx])
        rad = np.asfortranarray(rad[:, idx])
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=5, rays_per_pixel=rays_per_pixel)
    gfull = gradsetup.with_pixels(gi, pix)
    ref = oracle.levisapprox_gradient(sc.state, rays, gfull, trace_cap=256 if trace else 0)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    out = dev.gradient(rays, pix, trace_cap=256 if trace else 0)
    dev.close()
    return sc, ref, out


def check(ref, out, trace=True):
    gref, cref, soref = ref[:3]
    g, cost, so = out[:3]
    if trace:
        np.testing.assert_array_equal(out[3]['ncells'], ref[3]['ncells'])
        np.testing.assert_array_equal(out[3]['cells'], ref[3]['cells'])
        np.testing.assert_array_equal(out[3]['nsub'], ref[3]['nsub'])
    np.testing.assert_allclose(so, soref, rtol=RTOL, atol=1e-6)
    assert abs(float(cost[0]) - cref) <= RTOL * abs(cref)
    assert np.all(np.isfinite(g))
    for idr in range(gref.shape[1]):
        scale = np.max(np.abs(gref[:, idr]))
        assert scale > 0
        np.testing.assert_allclose(g[:, idr], gref[:, idr], rtol=RTOL, atol=RTOL * scale)


CASES = [
    ('scalar_periodic', dict(numder=1)),
    ('scalar_periodic_split', dict(numder=2)),
    ('scalar_open_split', dict(numder=3, exact_phase_derivative=True)),
    ('scalar_nmu16', dict(numder=2)),
    ('polarized_periodic_split', dict(numder=2)),
    ('polarized_open', dict(numder=3, exact_phase_derivative=True)),
    ('rayleigh_two_species', dict(numder=3, exact_phase_derivative=True)),
    ('polarized_rayleigh_varsfc', dict(numder=2)),
    ('thick_transcut', dict(numder=2)),
]


@pytest.mark.parametrize('case,gkw', CASES, ids=[c[0] for c in CASES])
def test_gradient_matches_oracle(case, gkw, oracle):
    sc, ref, out = run_case(oracle, case, gkw)
    check(ref, out)


@pytest.mark.parametrize('gkw', [dict(numder=2, exact_single_scatter=False),
                                 dict(numder=2, singlescatter=True),
                                 dict(numder=1, costfunc='LL')],
                         ids=['no_exact_ss', 'singlescatter', 'costfunc_LL'])
def test_gradient_flags(gkw, oracle):
    sc, ref, out = run_case(oracle, 'scalar_periodic_split', gkw, bright_only=gkw.get('costfunc') == 'LL')
    check(ref, out)


def test_gradient_polarized_LL(oracle):
    sc, ref, out = run_case(oracle, 'polarized_periodic_split', dict(numder=1, costfunc='LL'), bright_only=True)
    check(ref, out)


def test_gradient_subpixel_rays(oracle):
    sc, ref, out = run_case(oracle, 'scalar_periodic_split', dict(numder=2), rays_per_pixel=3)
    check(ref, out)


def test_gradient_no_deltam(oracle):
    sc, ref, out = run_case(oracle, 'scalar_no_deltam', dict(numder=2))
    check(ref, out)
    sc, ref, out = run_case(oracle, 'polarized_rayleigh_no_deltam', dict(numder=3, exact_phase_derivative=True))
    check(ref, out)


def test_gradient_device_pointers(oracle):
    import torch
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup
    sc = scenes.make('scalar_periodic_split', oracle)
    rays = scenes.ray_set(sc, n_persp=7, res=0.035)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, numder=2)
    rad = oracle.render(sc.state, rays)
    pix = gradsetup.make_pixels(1, rays.nrays, rad, seed=5)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    g, cost, so = dev.gradient(rays, pix)

    class B:
        pass
    r, p = B(), B()
    for k in ('camx', 'camy', 'camz', 'cammu', 'camphi'):
        setattr(r, k, torch.from_numpy(getattr(rays, k)).cuda())
    for k in ('measurements', 'uncertainties', 'rays_per_pixel', 'ray_weights', 'stokes_weights'):
        a = getattr(pix, k)
        # device tensors hold the Fortran-ordered bytes
        setattr(p, k, torch.from_numpy(np.ascontiguousarray(a.T)).cuda())
    p.rays_per_pixel = torch.from_numpy(pix.rays_per_pixel).cuda()
    p.uncertainties = torch.from_numpy(np.ascontiguousarray(pix.uncertainties.transpose(2, 1, 0))).cuda()
    g2, cost2, so2 = dev.gradient(r, p)
    torch.cuda.synchronize()
    scale = np.abs(g).max()
    np.testing.assert_allclose(g2.cpu().numpy().T, g, rtol=1e-9, atol=1e-12 * scale)
    np.testing.assert_allclose(float(cost2.cpu()[0]), float(cost[0]), rtol=1e-12)
    np.testing.assert_array_equal(so2.cpu().numpy().T, so)
    dev.close()


def test_chunked_derivative_pass_matches_single_chunk(oracle, monkeypatch):
    """The derivative pass runs over chunks of rays when the visit records exceed the budget
    (AT3D_B200_REC_GB); a tiny budget (one ray per chunk) must give the same gradient."""
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup
    sc = scenes.make('polarized_periodic_split', oracle)
    rays = scenes.ray_set(sc, n_persp=5, res=0.05)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=3, numder=2)
    rad = oracle.render(sc.state, rays)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=4)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    g1, c1, s1 = dev.gradient(rays, pix)
    monkeypatch.setenv('AT3D_B200_REC_GB', '0.000001')
    g2, c2, s2 = dev.gradient(rays, pix)
    dev.close()
    assert float(c1[0]) == float(c2[0])
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_allclose(g2, g1, rtol=1e-10, atol=1e-12 * np.max(np.abs(g1)))


THERMAL_CASES = [
    ('scalar_periodic_split', 'T', 'R', dict(numder=2)),
    ('scalar_open_split', 'B', 'R', dict(numder=3, exact_phase_derivative=True)),
    ('scalar_no_deltam', 'T', 'R', dict(numder=2)),
    ('scalar_no_deltam', 'B', 'R', dict(numder=2)),
    ('polarized_periodic_split', 'T', 'R', dict(numder=2)),
    ('rayleigh_two_species', 'B', 'R', dict(numder=3, exact_phase_derivative=True)),
    ('scalar_nmu16', 'T', 'T', dict(numder=2)),
]


@pytest.mark.parametrize('case,srctype,units,gkw', THERMAL_CASES, ids=['%s-%s%s' % c[:3] for c in THERMAL_CASES])
def test_thermal_source_gradient_matches_oracle(case, srctype, units, gkw, oracle):
    """SRCTYPE 'T' (thermal) and 'B' (solar + thermal): PLANCK / PLANCK_DERIVATIVE per grid point and the thermal
    component of COMPUTE_SOURCE_GRAD_1CELL (shdomsub4.f:1792-1799, 2009-2016, 3171-3221) with a non-zero DTEMP, the
    Planck sky and the warm Lambertian surface, against the oracle (itself pinned by finite differences,
    tests/test_oracle_golden.py)."""
    def mutate(sc):
        st = sc.state
        st.srctype, st.units, st.wavelen, st.gndtemp = srctype, units, 10.5, 291.0
        if srctype == 'T':
            st.skyrad = np.asfortranarray(2.7 + 200.0 * (st.skyrad / max(float(st.skyrad.max()), 1e-30)))   # sky temperatures [K]
            st.dirflux = np.zeros_like(st.dirflux)
        gp = st.gridpos
        st.temp = (286.0 - 30.0 * gp[2] + 5.0 * np.sin(11.0 * gp[0]) * np.cos(8.0 * gp[1])).astype(np.float32)

    def mutate_grad(sc, gi):
        gi.dtemp = np.asfortranarray(np.random.default_rng(2).uniform(-1.0, 1.0, gi.dext.shape).astype(np.float32))
    sc, ref, out = run_case(oracle, case, gkw, mutate=mutate, mutate_grad=mutate_grad, ssalb=0.85)   # absorbing: emits
    check(ref, out)


def test_thermal_gradient_needs_temp(oracle):
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup
    sc = scenes.make('scalar_periodic_split', oracle)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, numder=1)
    sc.state.srctype, sc.state.units, sc.state.wavelen, sc.state.temp = 'T', 'R', 10.5, None
    dev = DeviceState(sc.state)
    with pytest.raises(Exception) as e:
        dev.attach_gradient(gi)
    assert 'TEMP' in str(e.value)
    dev.close()


@pytest.mark.parametrize('case,gkw', [('scalar_periodic_split', dict(numder=2)),
                                      ('scalar_open_split', dict(numder=3, exact_phase_derivative=True)),
                                      ('polarized_periodic_split', dict(numder=2)),
                                      ('rayleigh_two_species', dict(numder=3, exact_phase_derivative=True))])
def test_streaming_beam_derivative_equals_dense_lists(case, gkw, oracle, monkeypatch):
    """The streaming direct-beam derivative (no DPATH/DPTR in memory: the gradient call walks toward the sun itself) adds
    the terms of COMPUTE_DIRECT_BEAM_DERIV_ADJOINT (shdomsub4.f:4117-4143) in the order of the dense lists: the gradient
    is the same bit for bit, and equal to rounding when the walks run in several passes over point ranges."""
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup, backend as B
    sc = scenes.make(case, oracle)
    rays = scenes.ray_set(sc, n_persp=6, res=0.04)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, **gkw)
    rad = oracle.render(sc.state, rays)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=5)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    g1, c1, s1 = dev.gradient(rays, pix)
    gs = gradsetup.with_streaming_beam(gi, sc.state, sc.pg, B)
    assert gs.dpath is None and gs.dptr is None
    dev.attach_gradient(gs)
    g2, c2, s2 = dev.gradient(rays, pix)
    monkeypatch.setenv('AT3D_B200_BEAM_PAIRS', '5000')
    g3, c3, s3 = dev.gradient(rays, pix)
    dev.close()
    assert np.max(np.abs(g1)) > 0
    np.testing.assert_array_equal(g2, g1)
    np.testing.assert_array_equal(s2, s1)
    assert float(c2[0]) == float(c1[0])
    np.testing.assert_allclose(g3, g1, rtol=1e-10, atol=1e-12 * np.max(np.abs(g1)))


@pytest.mark.parametrize('case,gkw', [('scalar_open_split', dict(numder=2)), ('polarized_periodic_split', dict(numder=2)),
                                      ('rayleigh_two_species', dict(numder=2))])
def test_jacobian_path_matches_oracle_single_sweep(case, gkw, oracle):
    """at3d_levisapprox_gradient_jacobian (MAKEJACOBIAN=.TRUE.) against the oracle's GRAD_INTEGRATE_1RAY path."""
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup
    sc = scenes.make(case, oracle)
    rays = scenes.ray_set(sc, n_persp=4, res=0.07)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, **gkw)
    rad = oracle.render(sc.state, rays)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=5, rays_per_pixel=2)
    g = gradsetup.with_pixels(gi, pix)
    g2, c2, s2 = oracle.levisapprox_gradient(sc.state, rays, g)
    jp = (np.argsort(-np.abs(g2[:, 0]))[:6] + 1).astype(np.int32)
    gref, cref, sref, jref = oracle.levisapprox_jacobian(sc.state, rays, g, jp)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    gout, cost, so, jac = dev.gradient_jacobian(rays, pix, jp)
    dev.close()
    assert abs(float(cost[0]) - cref) <= RTOL * abs(cref)
    np.testing.assert_allclose(so, sref, rtol=RTOL, atol=1e-6)
    for idr in range(gref.shape[1]):
        scale = np.max(np.abs(gref[:, idr]))
        np.testing.assert_allclose(gout[:, idr], gref[:, idr], rtol=RTOL, atol=RTOL * scale)
    for k in range(jref.shape[0]):
        for idr in range(jref.shape[1]):
            scale = np.max(np.abs(jref[k, idr]))
            np.testing.assert_allclose(jac[k, idr], jref[k, idr], rtol=RTOL, atol=RTOL * scale)


def test_streaming_beam_derivative_on_the_jacobian_path(oracle):
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup, backend as B
    sc = scenes.make('scalar_open_split', oracle)
    rays = scenes.ray_set(sc, n_persp=4, res=0.06)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, numder=2)
    rad = oracle.render(sc.state, rays)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=5)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    jp = (np.argsort(-np.abs(dev.gradient(rays, pix)[0][:, 0]))[:6] + 1).astype(np.int32)
    g1, c1, s1, j1 = dev.gradient_jacobian(rays, pix, jp)
    dev.attach_gradient(gradsetup.with_streaming_beam(gi, sc.state, sc.pg, B))
    g2, c2, s2, j2 = dev.gradient_jacobian(rays, pix, jp)
    dev.close()
    np.testing.assert_array_equal(g2, g1)
    assert np.max(np.abs(j1)) > 0
    np.testing.assert_allclose(j2, j1, rtol=1e-5, atol=1e-7 * np.max(np.abs(j1)))


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'polarized_periodic_split'])
def test_gradient_with_orthographic_views(case, oracle):
    """Runs of >= 256 rays with one direction: the forward pass of the gradient reads the per-view source
    (view_source_kernel[_oct]); cost, pixel values and gradient against the oracle at the usual bars."""
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup, synthetic as S
    sc = scenes.make(case, oracle)
    views = [S.orthographic_rays(sc, z, a, 0.016)[0] for z, a in ((0.0, 0.0), (50.0, 200.0))]
    rays = S.concat_rays(views)
    assert min(v.nrays for v in views) >= 256
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, numder=2)
    rad = oracle.render(sc.state, rays)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=5)
    ref = oracle.levisapprox_gradient(sc.state, rays, gradsetup.with_pixels(gi, pix), trace_cap=128)
    dev = DeviceState(sc.state)
    dev.attach_gradient(gi)
    out = dev.gradient(rays, pix, trace_cap=128)
    dev.close()
    check(ref, out)


def test_memory_reuse_between_states_changes_nothing(oracle):
    """at3d_set_memory_reuse: states built and dropped in a loop take their memory from what the previous one left (driver
    pool + parked buffers); radiances and gradient are those of plain allocation, bit for bit."""
    from at3d_b200.device import DeviceState
    from at3d_b200 import gradsetup, backend as B
    sc = scenes.make('scalar_periodic_split', oracle)
    rays = scenes.ray_set(sc, n_persp=6, res=0.04)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, numder=2)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, oracle.render(sc.state, rays), seed=5)

    def evaluate():
        dev = DeviceState(sc.state)
        r = dev.render(rays)
        dev.attach_gradient(gi)
        g, c, s = dev.gradient(rays, pix)
        dev.close()
        return r, g, float(c[0]), s
    r0, g0, c0, s0 = evaluate()
    assert B.memory_reuse(True) is False
    try:
        for _ in range(3):
            r, g, c, s = evaluate()
            np.testing.assert_array_equal(r, r0)
            np.testing.assert_array_equal(g, g0)
            np.testing.assert_array_equal(s, s0)
            assert c == c0
    finally:
        assert B.memory_reuse(False) is True
        B.trim_memory()
    r, g, c, s = evaluate()
    np.testing.assert_array_equal(g, g0)
