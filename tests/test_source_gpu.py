"""GPU parity of COMPUTE_SOURCE (C-ABI at3d_compute_source) and of the small reductions
(average_subpixel_rays, UPDATE_COSTFUNCTION) against the CPU oracle and the reference's known answers."""
import numpy as np
import pytest
import scenes

pytestmark = pytest.mark.gpu


def _state_for_source(case, oracle, seed=0):
    sc = scenes.make(case, oracle)
    st = sc.state
    rng = np.random.default_rng(seed)
    npts, nst = st.npts, st.nstokes
    # the "old" source: the scene's SOURCE with its adaptive SHPTR; previous DELSOURCE on OSHPTR = SHPTR
    shptr = st.shptr.copy()
    maxiv = int(st.nlm * npts)
    source = np.zeros((nst, maxiv), np.float32, order='F')
    source[:, :st.source.shape[1]] = st.source
    oshptr = shptr.copy()
    delsource = np.zeros((nst, maxiv), np.float32, order='F')
    delsource[:, :shptr[npts]] = 0.01 * rng.standard_normal((nst, shptr[npts])).astype(np.float32)
    return sc, shptr, source, oshptr, delsource, maxiv


CASES = ['scalar_periodic_split', 'scalar_nmu16', 'polarized_periodic_split', 'rayleigh_two_species',
         'polarized_rayleigh_varsfc', 'scalar_no_deltam']


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('mode', ['first', 'accel', 'noaccel', 'fixsh', 'shacc'])
def test_compute_source_matches_oracle(case, mode, oracle):
    from at3d_b200 import backend as B
    sc, shptr, source, oshptr, delsource, maxiv = _state_for_source(case, oracle)
    kw = dict(first=mode == 'first', accelflag=mode != 'noaccel', fixsh=mode == 'fixsh',
              shacc=3e-4 if mode == 'shacc' else 0.0, maxiv=maxiv)
    rc_r, shptr_r, src_r, oshptr_r, del_r, sums_r = oracle.compute_source(sc.state, shptr, source, oshptr, delsource, **kw)
    sums64 = oracle.compute_source_sums64()
    rc_g, shptr_g, src_g, oshptr_g, del_g, sums_g = B.compute_source(sc.state, shptr, source, oshptr, delsource, **kw)
    assert rc_r == 0 and rc_g == 0
    np.testing.assert_array_equal(shptr_g, shptr_r)                 # SHPTR: bit-exact
    np.testing.assert_array_equal(oshptr_g, oshptr_r)
    n = shptr_r[sc.state.npts]
    scale = np.abs(src_r[:, :n]).max()
    np.testing.assert_allclose(src_g[:, :n], src_r[:, :n], rtol=1e-5, atol=1e-6 * scale)
    if kw['accelflag'] and not kw['first']:
        m = oshptr_r[sc.state.npts]
        np.testing.assert_allclose(del_g[:, :m], del_r[:, :m], rtol=1e-4, atol=1e-6 * scale)
    # the four sums are float32 sequential sums in the reference: rtol 1e-4
    np.testing.assert_allclose(sums_g, sums_r, rtol=1e-4, atol=1e-7 * max(abs(sums_r[3]), 1e-30))
    # ... and 1e-5 against the same REAL products summed in f64 (no sequential rounding)
    if not kw['first']:
        np.testing.assert_allclose(sums_g, sums64, rtol=1e-5, atol=1e-7 * max(abs(sums64[3]), 1e-30))


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'polarized_periodic_split', 'rayleigh_two_species'])
@pytest.mark.parametrize('mode', ['accel', 'fixsh', 'first'])
def test_compute_source_device_resident_matches_oracle(case, mode, oracle):
    """at3d_compute_source_device: every array a device pointer (torch tensors), SOURCE / SHPTR double-buffered; twice in a
    row on the same properties (the second call reuses the mixed Legendre rows)."""
    import torch
    from at3d_b200 import backend as B
    sc, shptr, source, oshptr, delsource, maxiv = _state_for_source(case, oracle)
    st = sc.state
    kw = dict(first=mode == 'first', accelflag=True, fixsh=mode == 'fixsh', shacc=0.0, maxiv=maxiv)
    rc_r, shptr_r, src_r, oshptr_r, del_r, sums_r = oracle.compute_source(st, shptr, source, oshptr, delsource, **kw)
    dev = B.DeviceSourceState(st)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a).ravel(order='F'))).cuda()
    for rep in range(2):
        d_shptr, d_src, d_oshptr, d_del = t(shptr), t(source), t(oshptr), t(delsource)
        d_shptr_new = torch.zeros_like(d_shptr)
        d_src_new = torch.zeros_like(d_src)
        rc, total, sums = B.compute_source_device(dev, d_shptr, d_src, d_oshptr, d_del, d_shptr_new, d_src_new, **kw)
        assert rc == 0 and total == shptr_r[st.npts]
        np.testing.assert_array_equal(d_shptr_new.cpu().numpy(), shptr_r)
        n = shptr_r[st.npts]
        src_g = d_src_new.cpu().numpy().reshape(source.shape, order='F')
        scale = np.abs(src_r[:, :n]).max()
        np.testing.assert_allclose(src_g[:, :n], src_r[:, :n], rtol=1e-5, atol=1e-6 * scale)
        if not kw['first']:
            m = oshptr_r[st.npts]
            del_g = d_del.cpu().numpy().reshape(delsource.shape, order='F')
            np.testing.assert_allclose(del_g[:, :m], del_r[:, :m], rtol=1e-4, atol=1e-6 * scale)
            np.testing.assert_allclose(sums, sums_r, rtol=1e-4, atol=1e-7 * max(abs(sums_r[3]), 1e-30))


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'polarized_periodic_split', 'rayleigh_two_species', 'scalar_nmu16'])
@pytest.mark.parametrize('route', ['one_pass', 'two_pass'])
def test_compute_source_adaptive_one_pass_equals_two_pass(case, route, oracle, monkeypatch):
    """The adaptive truncation in one pass (cs_adapt_kernel: chunk offsets by decoupled look-back, DELSOURCE double-buffered)
    against the oracle and against the two-pass route, with an old DELSOURCE whose truncation (OSHPTR) differs from SHPTR."""
    import torch
    from at3d_b200 import backend as B
    sc, shptr, source, oshptr, delsource, maxiv = _state_for_source(case, oracle)
    st = sc.state
    npts, nst = st.npts, st.nstokes
    rng = np.random.default_rng(5)
    # OSHPTR: the truncation of the iteration before, shorter or equal per point
    ns = np.diff(shptr)
    nso = np.maximum(np.minimum(ns, rng.integers(1, st.nlm + 1, npts)), 0).astype(np.int32)
    oshptr = np.concatenate([[0], np.cumsum(nso)]).astype(np.int32)
    delsource = np.zeros((nst, maxiv), np.float32, order='F')
    delsource[:, :oshptr[npts]] = 0.01 * rng.standard_normal((nst, oshptr[npts])).astype(np.float32)
    kw = dict(first=False, accelflag=True, fixsh=False, shacc=2e-4, maxiv=maxiv)
    rc_r, shptr_r, src_r, oshptr_r, del_r, sums_r = oracle.compute_source(st, shptr, source, oshptr, delsource, **kw)
    if route == 'one_pass':
        monkeypatch.setenv('AT3D_B200_CS_ADAPT', 'one')               # opt-in (the two passes are faster)
    dev = B.DeviceSourceState(st)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a).ravel(order='F'))).cuda()
    d_shptr, d_src, d_oshptr, d_del = t(shptr), t(source), t(oshptr), t(delsource)
    d_del_new = torch.zeros_like(d_del)
    d_shptr_new, d_src_new = torch.zeros_like(d_shptr), torch.zeros_like(d_src)
    rc, total, sums = B.compute_source_device(dev, d_shptr, d_src, d_oshptr, d_del, d_shptr_new, d_src_new,
                                              delsource_new=d_del_new, **kw)
    assert rc == 0 and total == shptr_r[npts]
    np.testing.assert_array_equal(d_shptr_new.cpu().numpy(), shptr_r)
    n = shptr_r[npts]
    src_g = d_src_new.cpu().numpy().reshape(source.shape, order='F')
    scale = np.abs(src_r[:, :n]).max()
    np.testing.assert_allclose(src_g[:, :n], src_r[:, :n], rtol=1e-5, atol=1e-6 * scale)
    m = oshptr_r[npts]                                               # the new OSHPTR is the old SHPTR
    assert m == shptr[npts]
    del_g = d_del_new.cpu().numpy().reshape(delsource.shape, order='F')
    np.testing.assert_allclose(del_g[:, :m], del_r[:, :m], rtol=1e-4, atol=1e-6 * scale)
    np.testing.assert_allclose(sums, sums_r, rtol=1e-4, atol=1e-7 * max(abs(sums_r[3]), 1e-30))
    # the old DELSOURCE is untouched
    np.testing.assert_array_equal(d_del.cpu().numpy(), np.asarray(delsource).ravel(order='F'))


def test_compute_source_out_of_sh_memory(oracle):
    from at3d_b200 import backend as B
    sc, shptr, source, oshptr, delsource, maxiv = _state_for_source('scalar_periodic_split', oracle)
    rc_r = oracle.compute_source(sc.state, shptr, source, oshptr, delsource, maxiv=100)[0]
    rc_g = B.compute_source(sc.state, shptr, source, oshptr, delsource, maxiv=100)[0]
    assert rc_r == 2 and rc_g == 2                                   # IERR=2 (at3d/solver.py:619-631 retries)


def test_average_subpixel_rays(oracle):
    from at3d_b200 import backend as B
    rng = np.random.default_rng(3)
    for nstokes, counts in [(1, [1, 1, 1, 1]), (3, [4, 1, 7, 2, 3]), (1, [9]), (3, [2] * 50)]:
        pix = np.repeat(np.arange(len(counts)), counts).astype(np.int32)
        ws = rng.standard_normal((nstokes, pix.size)).astype(np.float32)
        ref = oracle.average_subpixel_rays(ws, pix, len(counts))
        out = B.average_subpixel_rays(ws, pix, len(counts))
        np.testing.assert_array_equal(out, ref)


def test_update_costfunction_known_answers(oracle):
    """Known answers of the reference's own unit tests (tests/test_derivatives.py:75-135)."""
    from at3d_b200 import backend as B
    so = np.ones(4) * 10.0; so[3] = 0.0
    g, c = B.update_costfunction(so, np.ones((4, 10, 1)), np.zeros((10, 1)), 0.0, np.ones((4, 4)) * 5, 'L2',
                                 np.ones(4) * 13.0)
    assert abs(c[0] - 1960.0) < 1e-5 and abs(g[0, 0] + 440.0) < 1e-5
    unc = np.zeros((2, 2)); unc[0, 0] = (1.0 / 0.03) ** 2; unc[1, 1] = (1.0 / 0.005) ** 2
    so = np.ones(3); so[1] = 0.5; so[2] = 0.0
    me = np.ones(3) * 1.25; me[1] = 0.25; me[2] = 0.25
    g, c = B.update_costfunction(so, np.ones((3, 10, 1)), np.zeros((10, 1)), 0.0, unc, 'LL', me)
    assert abs(c[0] - 6519.21) < 1e-2 and abs(g[0, 0] - 45329.43) < 1e-2
    gr, cr = oracle.update_costfunction(so, np.ones((3, 10, 1)), np.zeros((10, 1)), [0.0], unc, 'LL', me)
    np.testing.assert_allclose(g, gr, rtol=1e-12)
