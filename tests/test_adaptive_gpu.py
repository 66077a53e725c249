"""The adaptive solve on the GPU (C `at3d_solve_adaptive`: INIT_RADIANCE + SOLUTION_ITERATIONS with SPLIT_GRID) against
SHDOM's own outputs for the reference's 3-D verification case and against the oracle's adaptive solve on small scenes."""
import numpy as np
import pytest
import oracle_lib as O
import shdom_rico as R
from at3d_b200 import solver
from at3d_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def wtmu_of(st):
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    return (st.wtdo[:, 0] / delphi).astype(np.float32)


@pytest.fixture(scope='module')
def rico_gpu():
    st, pg, wtmu, tempp = R.make_state(O)
    sol, iters, solcrit, splitcrit, ms = solver.solve_adaptive(st, pg, wtmu, tempp=tempp, splitacc=0.1, shacc=0.01, solacc=1e-4,
                                                               maxiter=100, timing=True)
    print('rico adaptive solve on the GPU: %d iterations, NPTS=%d NCELLS=%d NSH=%d, ms path/source/split+setup/all = %s'
          % (iters, sol.npts, sol.ncells, int(sol.shptr[sol.npts]), ['%.1f' % m for m in ms]))
    return sol, iters, solcrit, splitcrit


def test_gpu_adaptive_solve_reproduces_the_shdom_run(rico_gpu):
    """RTE.solve with the reference's `Verify_Solver` configuration, entirely through the product: same adaptive grid
    (NPTS, NCELLS), same number of SH terms and iterations as the SHDOM run that wrote the golden files."""
    sol, iters, solcrit, splitcrit = rico_gpu
    assert solcrit <= 1e-4 and splitcrit <= 0.1
    assert (sol.npts, sol.ncells, int(sol.shptr[sol.npts]), iters) == (32809, 33672, 264619, 18)


def test_gpu_adaptive_source_matches_shdom_verification_source_out(rico_gpu):
    sol = rico_gpu[0]
    truth = R.golden_source()
    if sol.npts != truth.shape[1]:
        pytest.skip('adaptive grid differs from the SHDOM run (covered by the test above)')
    testing = sol.source[:, :sol.npts]
    # the reference's assertion (tests/test_shdom.py:267) is allclose(atol=5e-7) [+ numpy's default rtol 1e-5]; the GPU
    # solve differs from the Fortran by summation order in the transforms and sweeps (float32, 18 iterations): 2e-6 + 2e-5 rel
    assert np.allclose(testing, truth, atol=2e-6, rtol=2e-5)
    assert np.sqrt(np.mean((testing - truth) ** 2)) / np.mean(truth) < 3e-5


def test_gpu_adaptive_solution_renders_the_shdom_radiances(rico_gpu):
    from at3d_b200.device import DeviceState
    sol = rico_gpu[0]
    rays = R.sensor_rays()
    gold = R.golden_radiance()
    dev = DeviceState(sol)
    out = dev.render(rays, correctinterpolate=False)
    dev.close()
    assert np.allclose(out[0], gold[:, 2], atol=3e-3)                   # tests/test_shdom.py:269-277
    assert np.allclose(out[1], gold[:, 3], atol=2e-4)
    assert np.allclose(out[2], gold[:, 4], atol=7e-5)
    assert np.abs(out[0] - gold[:, 2]).max() < 6e-5 and np.abs(out[1] - gold[:, 3]).max() < 1.5e-5
    assert np.abs(out[2] - gold[:, 4]).max() < 5e-6


# (scene, SPLITACC): the first case splits until the MAXIG / MAXIC limits stop it (OUTOFMEM path), the others end on SPLITCRIT
CASES = [(dict(nx=8, ny=7, nz=9, nstokes=1, bc='periodic', seed=41, ext_max=40.0), 0.02),
         (dict(nx=7, ny=8, nz=10, nstokes=1, bc='open', seed=42, ext_max=60.0), 0.15),
         (dict(nx=6, ny=6, nz=8, nstokes=3, bc='open', seed=43, ext_max=40.0), 0.1),
         (dict(nx=7, ny=6, nz=9, nstokes=1, bc='periodic', rayleigh=True, seed=44, ext_max=50.0), 0.12)]


@pytest.mark.parametrize('kw,splitacc', CASES)
def test_gpu_adaptive_solve_matches_the_oracle_on_small_scenes(kw, splitacc):
    """Same split grid (cell for cell, point for point), same iteration count, same SHPTR; SOURCE / RADIANCE within 1e-4."""
    sc = S.make_scene(nsplits=0, **kw)
    O.finalize_scene(sc)
    st, pg = sc.state, sc.pg
    w = wtmu_of(st)
    par = dict(splitacc=splitacc, shacc=0.003, solacc=1e-4, maxiter=60, adapt_grid_factor=6.0)
    ref, it_r, sc_r, sp_r = O.solve_adaptive(st, pg, w, **par)
    out, it_g, sc_g, sp_g = solver.solve_adaptive(st, pg, w, **par)
    assert ref.npts > st.npts                                        # the case does split
    assert (out.npts, out.ncells, it_g) == (ref.npts, ref.ncells, it_r)
    for k in ('gridptr', 'neighptr', 'treeptr', 'cellflags', 'gridpos', 'iphase'):
        np.testing.assert_array_equal(getattr(out, k), getattr(ref, k), err_msg=k)
    for k in ('extinct', 'albedo', 'total_ext', 'phaseinterpwt', 'dirflux'):
        np.testing.assert_allclose(getattr(out, k), getattr(ref, k), rtol=1e-6, atol=1e-9, err_msg=k)
    np.testing.assert_array_equal(out.shptr, ref.shptr)
    np.testing.assert_array_equal(out.rshptr, ref.rshptr)
    scale = np.abs(ref.source).max()
    np.testing.assert_allclose(out.source, ref.source, rtol=1e-4, atol=3e-6 * scale)
    np.testing.assert_allclose(out.radiance, ref.radiance, rtol=1e-4, atol=3e-6 * np.abs(ref.radiance).max())
    np.testing.assert_allclose(out.fluxes, ref.fluxes, rtol=1e-4, atol=1e-6)
    assert abs(sc_g - sc_r) <= 2e-3 * sc_r and abs(sp_g - sp_r) <= 1e-4 * sp_r


def test_gpu_adaptive_solve_without_splitting_equals_the_fixed_grid_solve():
    """SPLITACC = 0 and a zero first guess: the adaptive entry point is the fixed-grid loop."""
    sc = S.make_scene(nx=7, ny=7, nz=8, nstokes=1, bc='periodic', nsplits=0, seed=45)
    O.finalize_scene(sc)
    from at3d_b200 import backend as B
    st, pg = sc.state, sc.pg
    w = wtmu_of(st)
    st.dirflux = B.make_direct(st, pg)[0]            # the adaptive entry point runs MAKE_DIRECT itself (INIT_SOLUTION)
    a, it_a, sc_a, _ = solver.solve_adaptive(st, pg, w, splitacc=0.0, solacc=1e-5, maxiter=40, inradflag=False)
    b, it_b, sc_b, _ = solver.solve_fixed_grid(st, w, maxiter=40, solacc=1e-5)
    assert it_a == it_b
    np.testing.assert_array_equal(a.shptr, b.shptr)
    np.testing.assert_allclose(a.source, b.source, rtol=1e-6, atol=1e-9)
