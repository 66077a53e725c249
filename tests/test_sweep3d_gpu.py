"""PATH_INTEGRATION on 3-D grids on the GPU (at3d_solver_*: SWEEPING_ORDER + the BACK_INT_GRID3D data-flow sweep,
src/polarized/shdomsub1.f:3261-4036) against the oracle's serial sweep: periodic and open boundaries, split (adaptive)
cells, NSTOKES 1 and 3, two species, Lambertian and general BRDF surfaces; then whole fixed-grid solves."""
import numpy as np
import pytest
import oracle_lib as O
import scenes
from at3d_b200 import solver
from at3d_b200 import synthetic as S

pytestmark = pytest.mark.gpu

CASES = ['scalar_periodic', 'scalar_periodic_split', 'scalar_open_split', 'scalar_nmu16', 'polarized_periodic_split',
         'polarized_open', 'rayleigh_two_species', 'polarized_rayleigh_varsfc', 'thick_transcut']


def wtmu_of(st):
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    return (st.wtdo[:, 0] / delphi).astype(np.float32)


def compare_path_integration(st, rtol=2e-5, transmin=1.0):
    w = wtmu_of(st)
    ref_rad, ref_flux, ref_bc = O.path_integration(st, w, st.shptr, st.source, st.rshptr, transmin=transmin)
    sv = solver.SweepSolver(st, w, transmin)
    for _ in range(2):                       # the solver object is reusable
        rad, flux, bc = sv.path_integration(st.shptr, st.source, st.rshptr)
    sv.close()
    scale = np.abs(ref_rad).max()
    np.testing.assert_allclose(flux, ref_flux, rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(bc, ref_bc, rtol=rtol, atol=1e-7 * scale)
    np.testing.assert_allclose(rad, ref_rad, rtol=1e-4, atol=2e-6 * scale)
    return rad, ref_rad


@pytest.mark.parametrize('case', CASES)
def test_sweep3d_matches_serial_sweep(case):
    sc = scenes.make(case, O)
    compare_path_integration(sc.state)


@pytest.mark.parametrize('transmin', [0.9, 0.3])
def test_sweep3d_transmin_below_one(transmin):
    """TRANSMIN < 1: a ray goes on through valid faces until its transmission falls below TRANSMIN."""
    for case in ('scalar_periodic_split', 'polarized_open'):
        sc = scenes.make(case, O)
        rad, _ = compare_path_integration(sc.state, transmin=transmin)
        base = solver.SweepSolver(sc.state, wtmu_of(sc.state)).path_integration(sc.state.shptr, sc.state.source, sc.state.rshptr)[0]
        assert np.abs(rad - base).max() > 0            # the parameter does change the result


def test_sweep3d_independent_pixel_in_x():
    """IPFLAG=1: the 3-D routine with the cells' IPINX flags (rays never leave through x faces)."""
    sc = S.make_scene(nx=6, ny=7, nz=8, nstokes=1, bc='periodic', nsplits=0, seed=21, ipflag=1)
    O.finalize_scene(sc)
    compare_path_integration(sc.state)


@pytest.mark.parametrize('kind', ['O', 'R', 'W'])
def test_sweep3d_brdf_surfaces(kind):
    sc = scenes.make('polarized_open' if kind == 'W' else 'scalar_periodic_split', O)
    st = S.with_brdf_surface(sc.state, kind, seed=3, wavelen=0.85)
    compare_path_integration(st)


def test_sweep3d_thermal_source():
    sc = scenes.make('scalar_periodic', O)
    st = sc.state.copy()
    st.srctype, st.units, st.wavelen, st.gndtemp = 'T', 'R', 11.0, 295.0
    st.skyrad = np.full_like(st.skyrad, 3.0)        # sky brightness temperature [K] for thermal sources
    compare_path_integration(st.normalize())


@pytest.mark.parametrize('group,levels', [('1', '1'), ('7', '1'), ('1', '0'), ('5', '0')])
def test_sweep3d_result_does_not_depend_on_the_schedule(group, levels, monkeypatch):
    """Ordinates in flight together (ticket order) and the processing order within an ordinate (dependency levels or
    the reference's sweep order) only change the schedule: bit-identical radiances."""
    monkeypatch.delenv('AT3D_SWEEP_GROUP', raising=False)
    monkeypatch.delenv('AT3D_SWEEP_LEVELS', raising=False)
    sc = scenes.make('polarized_periodic_split', O)
    st = sc.state
    w = wtmu_of(st)
    sv = solver.SweepSolver(st, w)
    base = sv.path_integration(st.shptr, st.source, st.rshptr)
    sv.close()
    monkeypatch.setenv('AT3D_SWEEP_GROUP', group)
    monkeypatch.setenv('AT3D_SWEEP_LEVELS', levels)
    sv = solver.SweepSolver(st, w)
    out = sv.path_integration(st.shptr, st.source, st.rshptr)
    sv.close()
    for x, y in zip(out, base):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'polarized_open', 'rayleigh_two_species'])
def test_fixed_grid_solve_3d(case):
    """SOLUTION_ITERATIONS on a fixed 3-D grid: GPU PATH_INTEGRATION + COMPUTE_SOURCE vs the oracle's solve, then the
    GPU RENDER of the GPU solution vs the oracle's RENDER of the oracle's solution."""
    from at3d_b200.device import DeviceState
    sc = scenes.make(case, O)
    st = sc.state
    w = wtmu_of(st)
    sol, iters, solcrit, _ = solver.solve_fixed_grid(st, w, solacc=1e-4, maxiter=50)
    ref, iters_r, solcrit_r = O.solve_fixed_grid(st, w, solacc=1e-4, maxiter=50)
    assert iters == iters_r and solcrit <= 1e-4
    np.testing.assert_array_equal(sol.shptr, ref.shptr)
    np.testing.assert_array_equal(sol.rshptr, ref.rshptr)
    scale = np.abs(ref.radiance).max()
    np.testing.assert_allclose(sol.radiance, ref.radiance, rtol=1e-4, atol=5e-6 * scale)
    np.testing.assert_allclose(sol.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    rays = scenes.ray_set(sc)
    dev = DeviceState(sol)
    out = dev.render(rays)
    dev.close()
    refout = O.render(ref, rays)
    np.testing.assert_allclose(out[0], refout[0], rtol=1e-4, atol=1e-6 * np.abs(refout[0]).max())
    for k in range(1, out.shape[0]):
        np.testing.assert_allclose(out[k], refout[k], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'polarized_open', 'rayleigh_two_species'])
def test_solver_update_medium_equals_a_fresh_solver(case):
    """at3d_solver_update_medium: a live solver object (sweep order, dependency levels, sorted plan kept) given another
    medium on its grid must solve exactly like a solver created for that medium."""
    sc = scenes.make(case, O)
    st = sc.state
    w = wtmu_of(st)
    st2 = st.copy()
    st2.extinct = np.asfortranarray(st.extinct * np.float32(1.7))
    st2.total_ext = st2.extinct.sum(axis=1).astype(np.float32)
    st2.albedo = np.asfortranarray(st.albedo * np.float32(0.93))
    st2.dirflux = (st.dirflux ** np.float32(1.7)).astype(np.float32)
    st2.gndalbedo = 0.5 * st.gndalbedo + 0.2
    st2.normalize()
    sv = solver.SweepSolver(st, w)
    a, ia, _, _ = sv.solve(solacc=1e-4, maxiter=50)
    sv.update_medium(st2)
    b, ib, cb, _ = sv.solve(solacc=1e-4, maxiter=50)
    sv.close()
    fresh = solver.SweepSolver(st2, w)
    c, ic, cc, _ = fresh.solve(solacc=1e-4, maxiter=50)
    fresh.close()
    assert ib == ic and cb == cc
    np.testing.assert_array_equal(b.shptr, c.shptr)
    np.testing.assert_array_equal(b.source, c.source)
    np.testing.assert_array_equal(b.radiance, c.radiance)
    np.testing.assert_array_equal(b.fluxes, c.fluxes)
    np.testing.assert_array_equal(b.bcrad, c.bcrad)
    assert np.abs(a.fluxes - b.fluxes).max() > 1e-3 * np.abs(a.fluxes).max()        # it is another solution
    with pytest.raises(Exception) as e:
        sv = solver.SweepSolver(st, w, 0.5)
        try:
            sv.update_medium(st2)
        finally:
            sv.close()
    assert 'TRANSMIN' in str(e.value)


@pytest.mark.parametrize('case,kw', [('scalar_periodic_split', dict()), ('polarized_open', dict(shacc=0.002)),
                                     ('rayleigh_two_species', dict(accelflag=False))])
def test_solve_continued_from_the_previous_solution(case, kw):
    """at3d_solver_solve_from: the iterations continue from the solution of a nearby medium on the same grid (the
    reference's load_solution + INIT_SOLUTION with INRADFLAG=.FALSE., what an optimisation step does) -- same iterations,
    truncation and fields as the oracle continued from the same solution, and fewer iterations than from the first guess."""
    sc = scenes.make(case, O)
    st = sc.state
    w = wtmu_of(st)
    sv = solver.SweepSolver(st, w)
    prev, it0, _, _ = sv.solve(solacc=1e-4, maxiter=60, **kw)
    st2 = st.copy()
    st2.extinct = np.asfortranarray(st.extinct * np.float32(1.05))
    st2.total_ext = st2.extinct.sum(axis=1).astype(np.float32)
    st2.normalize()
    sv.update_medium(st2)
    warm, it_w, sc_w, _ = sv.solve(solacc=1e-4, maxiter=60, initial=prev, **kw)
    cold, it_c, _, _ = sv.solve(solacc=1e-4, maxiter=60, **kw)
    sv.close()
    ref, it_r, sc_r = O.solve_fixed_grid(st2, w, solacc=1e-4, maxiter=60, initial=prev, **kw)
    assert it_w == it_r and it_w < it_c and sc_w <= 1e-4
    np.testing.assert_array_equal(warm.shptr, ref.shptr)
    np.testing.assert_array_equal(warm.rshptr, ref.rshptr)
    scale = np.abs(ref.radiance).max()
    np.testing.assert_allclose(warm.radiance, ref.radiance, rtol=1e-4, atol=5e-6 * scale)
    np.testing.assert_allclose(warm.source, ref.source, rtol=1e-4, atol=5e-6 * np.abs(ref.source).max())
    np.testing.assert_allclose(warm.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    # both routes end at the same solution within the solution accuracy
    np.testing.assert_allclose(warm.fluxes, cold.fluxes, rtol=5e-3, atol=1e-4 * np.abs(cold.fluxes).max())


@pytest.mark.parametrize('case,kw', [('scalar_open_split', dict()), ('polarized_periodic_split', dict(shacc=0.003)),
                                     ('rayleigh_two_species', dict(accelflag=False)),
                                     ('scalar_nmu16', dict(highorderrad=True, iterfixsh=3))])
def test_device_resident_loop_equals_host_driven_loop(case, kw):
    """at3d_solver_solve (SOURCE/RADIANCE/DELSOURCE resident in HBM, RADIANCE_TRUNCATION and the acceleration on the
    device) against the per-iteration C-ABI calls driven from Python: same iteration count, same truncations, and the
    same solution (the norms are summed in a different order only)."""
    sc = scenes.make(case, O)
    st = sc.state
    w = wtmu_of(st)
    a, ia, ca, ta = solver.solve_fixed_grid(st, w, solacc=1e-4, maxiter=40, device_loop=True, **kw)
    b, ib, cb, _ = solver.solve_fixed_grid(st, w, solacc=1e-4, maxiter=40, device_loop=False, **kw)
    assert ia == ib and ta['loop_ms'] > 0
    np.testing.assert_array_equal(a.shptr, b.shptr)
    np.testing.assert_array_equal(a.rshptr, b.rshptr)
    np.testing.assert_allclose(ca, cb, rtol=1e-3)
    scale = np.abs(b.source).max()
    np.testing.assert_allclose(a.source, b.source, rtol=1e-4, atol=1e-6 * scale)
    np.testing.assert_allclose(a.radiance, b.radiance, rtol=1e-4, atol=1e-6 * np.abs(b.radiance).max())
    np.testing.assert_allclose(a.fluxes, b.fluxes, rtol=1e-5)
    np.testing.assert_allclose(a.bcrad, b.bcrad, rtol=1e-5, atol=1e-8)


def test_device_loop_out_of_sh_memory_is_ierr_2():
    sc = scenes.make('scalar_periodic', O)
    st = sc.state
    with pytest.raises(MemoryError):
        solver.solve_fixed_grid(st, wtmu_of(st), maxiv=5 * st.npts)


TWO_D = [dict(nx=7, ny=1, nz=9, bc='periodic'), dict(nx=6, ny=4, nz=8, bc='periodic', nsplits=5),
         dict(nx=7, ny=1, nz=9, bc='open_x', rayleigh=True), dict(nx=6, ny=3, nz=8, bc='open_x', nstokes=3, nsplits=4)]


@pytest.mark.parametrize('kw', TWO_D)
def test_sweep2d_matches_serial_sweep(kw):
    """IPFLAG=2 (what at3d sets for ny=1, the usual 2-D x-z domains): BACK_INT_GRID2D (shdomsub1.f:4039-4293), rays in
    the X-Z plane and two-point faces, through the same data-flow kernel (the two extra face weights are exact zeros)."""
    sc = S.make_scene(seed=5, ipflag=2, **kw)
    O.finalize_scene(sc)
    compare_path_integration(sc.state)
    compare_path_integration(sc.state, transmin=0.5)


def test_fixed_grid_solve_2d():
    sc = S.make_scene(nx=9, ny=1, nz=10, bc='periodic', seed=8, ipflag=2)
    O.finalize_scene(sc)
    st = sc.state
    w = wtmu_of(st)
    sol, iters, solcrit, _ = solver.solve_fixed_grid(st, w, solacc=1e-4, maxiter=50)
    ref, iters_r, _ = O.solve_fixed_grid(st, w, solacc=1e-4, maxiter=50)
    assert iters == iters_r and solcrit <= 1e-4
    np.testing.assert_array_equal(sol.shptr, ref.shptr)
    np.testing.assert_allclose(sol.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(sol.radiance, ref.radiance, rtol=1e-4, atol=5e-6 * np.abs(ref.radiance).max())


def test_solver_refuses_what_it_does_not_cover():
    from at3d_b200._lib import At3dError
    sc = S.make_scene(nx=5, ny=5, nz=6, seed=1)
    O.finalize_scene(sc)
    st = sc.state.copy()
    st.bcflag = 4                                                    # multi-processor boundary flags
    with pytest.raises(At3dError) as e:
        solver.SweepSolver(st.normalize(), wtmu_of(st))
    assert e.value.code == 3
