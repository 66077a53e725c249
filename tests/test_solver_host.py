"""Host-side logic of the fixed-grid solve (at3d_b200/solver.py) against the oracle, no GPU needed."""
import numpy as np
import pytest
import oracle_lib as O
import scenes
from at3d_b200 import solver


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'rayleigh_two_species', 'polarized_rayleigh_varsfc',
                                  'scalar_no_deltam'])
@pytest.mark.parametrize('mode', ['adaptive', 'shacc', 'fixsh', 'highorder'])
def test_radiance_truncation_matches_oracle(case, mode):
    st = scenes.make(case, O).state
    kw = dict(fixsh=mode == 'fixsh', shacc=2e-3 if mode == 'shacc' else 0.0, highorderrad=mode == 'highorder',
              maxir=st.nlm * st.npts + st.npts)
    ref = O.radiance_truncation(st, st.shptr, st.radiance, st.rshptr, **kw)
    out = solver.radiance_truncation(st, st.shptr, st.radiance, st.rshptr, **kw)
    np.testing.assert_array_equal(out, ref)
    assert out[st.npts] > 0 and out[st.npts + 1] == out[st.npts]


def test_radiance_truncation_out_of_memory_falls_back():
    st = scenes.make('scalar_periodic_split', O).state
    tight = int(np.sum(np.maximum(4, np.diff(st.shptr[:st.npts + 1])))) + 1     # room for the FIXSH layout only
    ref = O.radiance_truncation(st, st.shptr, st.radiance, st.rshptr, False, 0.0, False, tight)
    out = solver.radiance_truncation(st, st.shptr, st.radiance, st.rshptr, False, 0.0, False, tight)
    np.testing.assert_array_equal(out, ref)
    with pytest.raises(MemoryError):
        solver.radiance_truncation(st, st.shptr, st.radiance, st.rshptr, True, 0.0, False, 10)


@pytest.mark.parametrize('case', ['scalar_periodic', 'scalar_periodic_split', 'scalar_open_split', 'polarized_open',
                                  dict(nx=7, ny=1, nz=9, bc='open_x', ipflag=2, seed=5),
                                  dict(nx=6, ny=4, nz=8, bc='periodic', ipflag=2, nsplits=5, seed=5)])
def test_host_sweeping_order_matches_oracle(case):
    """SWEEPING_ORDER of the library (host C++, at3d_sweeping_order; feeds the rank tables of the 3-D sweep kernel) is
    bit-identical to the oracle's restatement of shdomsub1.f:3261-3352 / :4529-4700, for periodic and open boundaries
    and for split cells; every octant lists every grid point exactly once."""
    import scenes
    from at3d_b200 import solver
    if isinstance(case, dict):
        from at3d_b200 import synthetic as S
        sc = S.make_scene(**case)
        O.finalize_scene(sc)
    else:
        sc = scenes.make(case, O)
    got = solver.sweeping_order(sc.state)
    ref = O.sweeping_order(sc.state)[:, :got.shape[1]]
    np.testing.assert_array_equal(got, ref)
    gp = sc.state.gridptr
    for joct in range(got.shape[1]):
        pts = gp[got[:, joct] & 7, (got[:, joct] >> 3) - 1]
        assert np.array_equal(np.sort(pts), np.arange(1, sc.state.npts + 1))
