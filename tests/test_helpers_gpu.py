"""GPU parity of the helper routines of the path against the CPU oracle:
YLMALL, PRECOMPUTE_PHASE_CHECK[_GRAD], PREPARE_DERIV_INTERPS, MAKE_DIRECT, MAKE_DIRECT_DERIVATIVE.
Integer outputs (INTERPPTR, DPTR, LONGEST_PATH_PTS, DI/DJ/DK) are bit-exact; float outputs within 1e-5
relative (device libm differs from glibc in the last bit for cosf/sinf/exp)."""
import numpy as np
import pytest
import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('nstleg,ml,mm', [(1, 7, 7), (1, 15, 15), (6, 7, 7), (6, 15, 15), (6, 9, 5), (1, 40, 31)])
@pytest.mark.parametrize('transpose', [False, True])
def test_ylmall(nstleg, ml, mm, transpose, oracle):
    from at3d_b200 import backend as B
    nlm = sum(2 * min(l, mm) + 1 for l in range(ml + 1))
    for mu, phi in [(0.5, 0.3), (-0.73, 2.9), (1.0, 0.0), (-1.0, 1.0), (0.0, 4.4), (0.9999, 6.2)]:
        ref = oracle.ylmall(transpose, np.float32(mu), np.float32(phi), ml, mm, nstleg, nlm)
        out = B.ylmall(transpose, mu, phi, ml, mm, nstleg, nlm)
        np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize('nstokes', [1, 3])
@pytest.mark.parametrize('grad', [False, True])
def test_precompute_phase_check(nstokes, grad, oracle):
    from at3d_b200 import backend as B
    from at3d_b200 import synthetic as S
    nstleg = 1 if nstokes == 1 else 6
    legenp = S.hg_legendre_table(np.linspace(0.7, 0.88, 7), 200, nstleg)
    if grad:
        legenp = np.asfortranarray(legenp / (2 * np.arange(201) + 1)[None, :, None], np.float32)
    ref = oracle.precompute_phase_check(legenp, 361, nstokes, 15, negcheck=not grad, grad=grad)
    out = B.precompute_phase_check(legenp, 361, nstokes, 15, negcheck=not grad, grad=grad)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-7 * np.abs(ref).max())


def test_precompute_phase_check_negative_is_an_error():
    from at3d_b200 import backend as B
    from at3d_b200._lib import At3dError
    legenp = np.zeros((1, 11, 1), np.float32, order='F')
    legenp[0, 0, 0] = 1.0
    legenp[0, 1, 0] = 3.0 * 1.5       # P = 1 + 4.5 cos(theta) < 0 in the backward hemisphere
    with pytest.raises(At3dError):
        B.precompute_phase_check(legenp, 37, 1, 7, negcheck=True)


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'scalar_open_split', 'polarized_rayleigh_varsfc',
                                  'rayleigh_two_species', 'scalar_no_deltam'])
def test_prepare_deriv_interps(case, oracle):
    from at3d_b200 import backend as B, gradsetup
    sc = scenes.make(case, oracle)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=3, numder=3, exact_phase_derivative=True)
    ref = oracle.prepare_deriv_interps(sc.state, sc.pg, gi)
    out = B.prepare_deriv_interps(sc.state, sc.pg, gi)
    np.testing.assert_array_equal(out[1], ref[1])                       # INTERPPTR
    np.testing.assert_array_equal(out[0], ref[0])                       # OPTINTERPWT (no libm involved)
    for k in (2, 3, 4):
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-6, atol=1e-7 * max(np.abs(ref[k]).max(), 1e-30))


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'scalar_open_split', 'polarized_open',
                                  'rayleigh_two_species'])
@pytest.mark.parametrize('sun', [(-0.5, 0.3), (-0.9, 3.5), (-0.2, 5.1), (-1.0, 0.0)])
def test_make_direct_and_derivative(case, sun, oracle):
    from at3d_b200 import backend as B
    sc = scenes.make(case, oracle)
    st = sc.state
    st.solarmu, st.solaraz = sun
    dref, eref, cref = oracle.make_direct(st, sc.pg)
    dout, eout, cout = B.make_direct(st, sc.pg)
    for k in ('ipdirect', 'di', 'dj', 'dk', 'longest_path_pts'):
        assert cout[k] == cref[k], k
    for k in cref:
        assert cout[k] == cref[k], k                                     # host libm: identical doubles
    np.testing.assert_array_equal(eout, eref)
    np.testing.assert_allclose(dout, dref, rtol=2e-6)
    pref, qref = oracle.make_direct_derivative(st, sc.pg, cref)
    pout, qout = B.make_direct_derivative(st, sc.pg, cout)
    np.testing.assert_array_equal(qout, qref)                            # DPTR bit-exact
    np.testing.assert_array_equal(pout, pref)                            # DPATH: IEEE double arithmetic only


def test_make_direct_derivative_overflow_is_an_error(oracle):
    from at3d_b200 import backend as B
    from at3d_b200._lib import At3dError
    sc = scenes.make('scalar_periodic', oracle)
    _, _, c = B.make_direct(sc.state, sc.pg)
    c['longest_path_pts'] = 8
    with pytest.raises(At3dError):
        B.make_direct_derivative(sc.state, sc.pg, c)


@pytest.mark.parametrize('kw', [dict(bc='periodic', nsplits=6), dict(bc='open', nsplits=5, rayleigh=True),
                                dict(bc='open', nstokes=3, deltam=False), dict(bc='periodic', nstokes=3, rayleigh=True)])
def test_transfer_pa_to_grid(kw, oracle):
    """TRANSFER_PA_TO_GRID on the GPU (property interpolation to base and split grid points, phase-table pointer lists in
    SSORT's order, delta-M scaling) against the ORACLE's TRILIN_INTERP_PROP / PREPARE_PROP (oracle/oracle_prop.c, pinned
    through the SHDOM rico solve): every output bit for bit."""
    from at3d_b200 import backend as B, synthetic as S
    sc = S.make_scene(nx=7, ny=6, nz=9, seed=17, **kw)
    st, pg = sc.state, sc.pg
    ref = oracle.transfer_pa_to_grid(pg, st.gridpos, st.npts, st.ml, bool(st.deltam))
    out = B.transfer_pa_to_grid(pg, st.gridpos, st.npts, st.ml, bool(st.deltam))
    for k in ('iphase', 'extinct', 'albedo', 'total_ext', 'phaseinterpwt', 'legen'):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=k)
    assert out['nleg'] == ref['nleg'] and out['extmin'] == ref['extmin']


def test_transfer_pa_to_grid_on_the_shdom_property_file(oracle):
    """... and on the 31 k base points of the reference's RICO property file (18 tabulated Mie phase functions)."""
    import shdom_rico as R
    from at3d_b200 import backend as B
    st, pg, wtmu, tempp = R.make_state(oracle)
    ref = oracle.transfer_pa_to_grid(pg, st.gridpos, st.npts, st.ml, True)
    out = B.transfer_pa_to_grid(pg, st.gridpos, st.npts, st.ml, True)
    for k in ('iphase', 'extinct', 'albedo', 'total_ext', 'phaseinterpwt', 'legen'):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=k)
