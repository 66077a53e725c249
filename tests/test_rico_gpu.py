"""GPU RENDER / PATH_INTEGRATION / COMPUTE_SOURCE on the reference's 3-D SHDOM verification case (the adaptive, polarized
RICO solve of tests/test_shdom.py:68-277) against SHDOM's own outputs and the oracle.  The solved state comes from the
oracle's adaptive solve (pinned to the same goldens in tests/test_shdom_adaptive.py)."""
import os
import numpy as np
import pytest
import oracle_lib as O
import shdom_rico as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rico():
    st, pg, wtmu, tempp = R.make_state(O)
    sol, iters, solcrit, splitcrit = O.solve_adaptive(st, pg, wtmu, tempp=tempp, splitacc=0.1, shacc=0.01, solacc=1e-4)
    assert (sol.npts, sol.ncells, iters) == (32809, 33672, 18)
    return sol, pg, wtmu


def test_gpu_render_matches_shdom_radiances_and_the_oracle_walk(rico):
    from at3d_b200.device import DeviceState
    sol, pg, wtmu = rico
    rays = R.sensor_rays()
    gold = R.golden_radiance()
    dev = DeviceState(sol)
    out, tr = dev.render(rays, correctinterpolate=False, trace_cap=256)
    dev.close()
    # the reference's own tolerances (tests/test_shdom.py:269-277) and what is actually achieved
    assert np.allclose(out[0], gold[:, 2], atol=3e-3) and np.abs(out[0] - gold[:, 2]).max() < 4e-5
    assert np.allclose(out[1], gold[:, 3], atol=2e-4) and np.abs(out[1] - gold[:, 3]).max() < 1e-5
    assert np.allclose(out[2], gold[:, 4], atol=7e-5) and np.abs(out[2] - gold[:, 4]).max() < 3e-6
    ref, tref, _ = O.render(sol, rays, correctinterpolate=False, trace_cap=256, nthreads=os.cpu_count() or 1)
    np.testing.assert_array_equal(tr['ncells'], tref['ncells'])         # split cells, open boundaries: bit-exact walk
    np.testing.assert_array_equal(tr['cells'], tref['cells'])
    np.testing.assert_array_equal(tr['nsub'], tref['nsub'])
    np.testing.assert_allclose(out[0], ref[0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out[1:], ref[1:], rtol=1e-4, atol=1e-6)


def test_gpu_compute_source_reproduces_shdom_source_from_the_converged_radiance(rico):
    """COMPUTE_SOURCE (NSTOKES=3, fixed truncation of the converged state) applied to the converged RADIANCE gives the
    SOURCE SHDOM wrote, up to the last iteration's change (SOLCRIT < 1e-4) and the acceleration step."""
    from at3d_b200 import backend as B
    sol, pg, wtmu = rico
    npts, nst = sol.npts, sol.nstokes
    maxiv = int(sol.shptr[npts]) + 64
    source = np.zeros((nst, maxiv), np.float32, order='F')
    source[:, :sol.source.shape[1]] = sol.source
    dels = np.zeros((nst, maxiv), np.float32, order='F')
    kw = dict(first=False, accelflag=True, fixsh=True, shacc=0.01, maxiv=maxiv)
    rc_r, shptr_r, src_r, _, _, sums_r = O.compute_source(sol, sol.shptr, source, sol.shptr.copy(), dels, **kw)
    sums64 = O.compute_source_sums64()
    rc_g, shptr_g, src_g, _, _, sums_g = B.compute_source(sol, sol.shptr, source, sol.shptr.copy(), dels, **kw)
    assert rc_r == 0 and rc_g == 0
    np.testing.assert_array_equal(shptr_g, shptr_r)
    n = int(shptr_r[npts])
    np.testing.assert_allclose(src_g[:, :n], src_r[:, :n], rtol=1e-5, atol=1e-6 * np.abs(src_r).max())
    # the norms: 1e-4 against the rounding-free (f64-summed) value; the reference's own REAL running sum over 265 k
    # terms carries ~5e-4 of sequential rounding and is checked at 2e-3
    np.testing.assert_allclose(sums_g, sums64, rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(sums_r, sums64, rtol=2e-3, atol=1e-12)
    gold = R.golden_source()
    assert np.abs(src_g[:, :npts] - gold).max() < 2e-3 * np.abs(gold).max()


def test_gpu_path_integration_on_the_shdom_split_grid(rico):
    """The 3-D data-flow sweep on SHDOM's own adaptive grid (open boundaries, 1284 split cells) vs the oracle's serial sweep."""
    from at3d_b200 import solver
    sol, pg, wtmu = rico
    ref_rad, ref_flux, ref_bc = O.path_integration(sol, wtmu, sol.shptr, sol.source, sol.rshptr)
    sv = solver.SweepSolver(sol, wtmu, 1.0)
    rad, flux, bc = sv.path_integration(sol.shptr, sol.source, sol.rshptr)
    sv.close()
    scale = np.abs(ref_rad).max()
    np.testing.assert_allclose(flux, ref_flux, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(rad, ref_rad, rtol=1e-4, atol=2e-6 * scale)
