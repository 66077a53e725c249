"""The fixed-grid, independent-pixel SHDOM solve on the GPU (at3d_b200/solver.py: PATH_INTEGRATION and COMPUTE_SOURCE
through the C ABI) followed by the GPU RENDER, checked end to end against SHDOM's own verification outputs
(tests/golden/brdf_*1{f,r}.out, reference tests/test_shdom.py:597-805) and against the oracle's solve."""
import os
import numpy as np
import pytest
import oracle_lib as O
import shdom_verification as V
from at3d_b200 import solver
from at3d_b200.device import DeviceState

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def unsolved(kind):
    st, pg, wtmu = V.make_state(O, kind)
    st.nscatangle = 721
    st.phasetab = O.precompute_phase_check(pg.legenp, 721, st.nstokes, st.ml, True)
    return st, wtmu


@pytest.mark.parametrize('kind', ['L', 'O', 'R', 'W', 'D'])
def test_gpu_solve_and_render_reproduce_shdom(kind):
    st, wtmu = unsolved(kind)
    sol, iters, solcrit, _ = solver.solve_ip(st, wtmu, solacc=1e-5)
    assert solcrit <= 1e-5 and iters <= 6
    # the same iteration on the CPU oracle: same truncation, same iteration count, radiance expansion within 1e-4
    ref, iters_r, _ = O.solve_fixed_grid(st, wtmu, solacc=1e-5)
    assert iters == iters_r
    np.testing.assert_array_equal(sol.shptr, ref.shptr)
    np.testing.assert_array_equal(sol.rshptr, ref.rshptr)
    scale = np.abs(ref.radiance).max()
    np.testing.assert_allclose(sol.radiance, ref.radiance, rtol=1e-4, atol=2e-6 * scale)
    np.testing.assert_allclose(sol.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    # SHDOM's printed fluxes and radiances
    gold = V.parse_shdom_output(os.path.join(GOLD, 'brdf_%s1f.out' % kind))
    bot = sol.bcptr[:sol.nbotpts, 1] - 1
    np.testing.assert_allclose(sol.fluxes[0, bot], gold[:, 3], rtol=0, atol=4e-6)
    np.testing.assert_allclose(sol.fluxes[1, bot], gold[:, 2], rtol=0, atol=6e-6)
    dev = DeviceState(sol)
    out = dev.render(V.sensor_rays())
    dev.close()
    gold = V.parse_shdom_output(os.path.join(GOLD, 'brdf_%s1r.out' % kind))
    np.testing.assert_allclose(out[0], gold[:, 2], rtol=0, atol=9e-6)
    if sol.nstokes == 3:
        np.testing.assert_allclose(out[1], gold[:, 3], rtol=0, atol=9e-6)
        np.testing.assert_allclose(out[2], gold[:, 4], rtol=0, atol=1e-8)


@pytest.mark.parametrize('kind', ['L', 'O'])
def test_device_resident_loop_equals_host_driven_loop_ip(kind):
    """The goldens above run through at3d_solver_solve (everything resident in HBM); the same solve driven from Python
    through at3d_path_integration_ip + at3d_compute_source gives the same truncations and the same solution."""
    st, wtmu = unsolved(kind)
    a, ia, ca, ta = solver.solve_fixed_grid(st, wtmu, solacc=1e-5)
    b, ib, cb, _ = solver.solve_fixed_grid(st, wtmu, solacc=1e-5, device_loop=False)
    assert ia == ib and 'loop_ms' in ta
    np.testing.assert_array_equal(a.shptr, b.shptr)
    np.testing.assert_array_equal(a.rshptr, b.rshptr)
    np.testing.assert_allclose(a.source, b.source, rtol=1e-5, atol=1e-7 * np.abs(b.source).max())
    np.testing.assert_allclose(a.radiance, b.radiance, rtol=1e-5, atol=1e-7 * np.abs(b.radiance).max())
    np.testing.assert_allclose(a.fluxes, b.fluxes, rtol=1e-6)
    np.testing.assert_allclose(a.bcrad, b.bcrad, rtol=1e-6, atol=1e-9)


def test_gpu_thermal_slab():
    # Verify_Thermal (reference tests/test_shdom.py:910-982) at the angular resolution the GPU transforms support:
    # GPU solve + GPU RENDER against the oracle at the same resolution, and against the closed form within the
    # quadrature error of NMU=16 (the oracle meets the reference's 3e-4 at its NMU=128, test_shdom_verification.py)
    st, pg, wtmu = V.make_thermal_state(O, 16, 32)
    sol, iters, solcrit, _ = solver.solve_ip(st, wtmu, solacc=1e-5)
    ref, iters_r, _ = O.solve_fixed_grid(st, wtmu, solacc=1e-5)
    assert iters == iters_r == 1
    np.testing.assert_array_equal(sol.shptr, ref.shptr)
    np.testing.assert_allclose(sol.fluxes, ref.fluxes, rtol=1e-5)
    np.testing.assert_allclose(sol.source, ref.source, rtol=1e-5)
    rays = V.nadir_rays()
    dev = DeviceState(sol)
    out = dev.render(rays)
    dev.close()
    np.testing.assert_allclose(out, O.render(ref, rays), rtol=1e-5)
    np.testing.assert_allclose(out[0], V.thermal_slab_radiance(), rtol=0, atol=1.2e-2)


def test_gpu_absorbing_columns_closed_form():
    # Verify_NonuniformGasAbsorption (reference tests/test_shdom.py:855-908) on the GPU, the reference's own atol=2e-7
    st, pg, wtmu = V.make_absorbing_state(O)
    sol, iters, solcrit, _ = solver.solve_ip(st, wtmu, solacc=1e-5)
    assert iters == 1
    dev = DeviceState(sol)
    out = dev.render(V.nadir_rays())
    dev.close()
    tau = np.linspace(0.0, 1.0, 50) * 30.0
    np.testing.assert_allclose(out[0], np.exp(-tau) * 0.04 / np.pi * np.exp(-tau), rtol=0, atol=2e-7)


def test_gpu_combined_source():
    # VerifyCombined (reference tests/test_shdom.py:984-1056), SRCTYPE='B', GPU solve + RENDER vs the oracle
    st, pg, wtmu = V.make_combined_state(O, 16, 32)
    sol, iters, solcrit, _ = solver.solve_ip(st, wtmu, solacc=1e-5)
    ref, iters_r, _ = O.solve_fixed_grid(st, wtmu, solacc=1e-5)
    assert iters == iters_r
    np.testing.assert_allclose(sol.fluxes, ref.fluxes, rtol=1e-5)
    rays = V.nadir_rays()
    dev = DeviceState(sol)
    out = dev.render(rays)
    dev.close()
    np.testing.assert_allclose(out, O.render(ref, rays), rtol=1e-5)
    tr = np.exp(-30.0 * np.linspace(0.001, 0.5, 50))
    np.testing.assert_allclose(out[0], 0.5 * tr * tr / np.pi + V.thermal_slab_radiance(), rtol=0, atol=1.2e-2)
