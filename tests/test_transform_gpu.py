"""GPU parity of the SH <-> discrete-ordinate transforms (C-ABI at3d_sh_to_do / at3d_do_to_sh) against the oracle's
restatement of SH_TO_DO / DO_TO_SH (the one that reproduces SHDOM's verification outputs through the full solve)."""
import numpy as np
import pytest
import scenes
import oracle_lib as O

pytestmark = pytest.mark.gpu


def wtmu_of(st):
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    return (st.wtdo[:, 0] / delphi).astype(np.float32)


CASES = ['scalar_periodic_split', 'scalar_nmu16', 'polarized_periodic_split', 'polarized_rayleigh_varsfc']


@pytest.mark.parametrize('case', CASES)
def test_sh_to_do_matches_oracle(case):
    from at3d_b200 import backend as B
    st = scenes.make(case, O).state
    w = wtmu_of(st)
    ref = O.sh_to_do(st, w, st.shptr, st.source)
    out = B.sh_to_do(st, w, st.shptr, st.source)
    assert out.shape == ref.shape
    scale = np.abs(ref).max()
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=2e-6 * scale)


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'scalar_nmu16', 'scalar_open_split'])
def test_sh_to_do_tensor_core_variant_matches_oracle(case, monkeypatch):
    """The tcgen05 3xTF32 variant (AT3D_B200_TRANSFORM=tc, NSTOKES=1): same 1e-4 bar against the oracle, and within FP32
    rounding of the FP32-FMA kernels (both use the same basis values)."""
    from at3d_b200 import backend as B
    st = scenes.make(case, O).state
    w = wtmu_of(st)
    ref = O.sh_to_do(st, w, st.shptr, st.source)
    fp32 = B.sh_to_do(st, w, st.shptr, st.source)
    monkeypatch.setenv('AT3D_B200_TRANSFORM', 'tc')
    out = B.sh_to_do(st, w, st.shptr, st.source)
    scale = np.abs(ref).max()
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=2e-6 * scale)
    np.testing.assert_allclose(out, fp32, rtol=2e-5, atol=2e-6 * scale)
    assert np.abs(out).max() > 0


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'scalar_nmu16', 'scalar_open_split'])
def test_do_to_sh_tensor_core_variant_matches_oracle(case, monkeypatch):
    """DO_TO_SH through tcgen05 (3xTF32): the 1e-4 bar against the oracle, FP32 rounding against the FP32-FMA kernel, and
    nothing written beyond a point's RADIANCE length."""
    from at3d_b200 import backend as B
    st = scenes.make(case, O).state
    w = wtmu_of(st)
    rng = np.random.default_rng(5)
    do = O.sh_to_do(st, w, st.rshptr[:st.npts + 1], st.radiance)
    do = do + 0.01 * np.abs(do).max() * rng.standard_normal(do.shape).astype(np.float32)
    ref = O.do_to_sh(st, w, st.rshptr, do)
    fp32 = B.do_to_sh(st, w, st.rshptr, do)
    monkeypatch.setenv('AT3D_B200_TRANSFORM', 'tc')
    out = B.do_to_sh(st, w, st.rshptr, do)
    n = int(st.rshptr[st.npts])
    scale = np.abs(ref[:, :n]).max()
    np.testing.assert_allclose(out[:, :n], ref[:, :n], rtol=1e-4, atol=2e-6 * scale)
    np.testing.assert_allclose(out[:, :n], fp32[:, :n], rtol=2e-5, atol=2e-6 * scale)
    np.testing.assert_array_equal(out[:, n:], fp32[:, n:])


@pytest.mark.parametrize('case', CASES)
def test_do_to_sh_matches_oracle(case):
    from at3d_b200 import backend as B
    st = scenes.make(case, O).state
    w = wtmu_of(st)
    rng = np.random.default_rng(3)
    nang = int(st.nphi0.sum())
    # a smooth, positive discrete-ordinate field plus noise
    do = O.sh_to_do(st, w, st.rshptr[:st.npts + 1], st.radiance)
    do = do + 0.01 * np.abs(do).max() * rng.standard_normal(do.shape).astype(np.float32)
    ref = O.do_to_sh(st, w, st.rshptr, do)
    out = B.do_to_sh(st, w, st.rshptr, do)
    n = int(st.rshptr[st.npts])
    scale = np.abs(ref[:, :n]).max()
    np.testing.assert_allclose(out[:, :n], ref[:, :n], rtol=1e-4, atol=2e-6 * scale)


def test_round_trip():
    # DO_TO_SH(SH_TO_DO(x)) on the GPU equals the oracle's round trip
    from at3d_b200 import backend as B
    import at3d_b200.synthetic as S
    sc = S.make_scene(nx=5, ny=5, nz=6, nmu=8, nphi=16, nstokes=1, truncate=False, seed=2)
    O.finalize_scene(sc)
    st = sc.state
    w = wtmu_of(st)
    src = st.source.copy(order='F')
    do = B.sh_to_do(st, w, st.shptr, src)
    back = B.do_to_sh(st, w, st.shptr, do)
    ref_back = O.do_to_sh(st, w, st.shptr, O.sh_to_do(st, w, st.shptr, src))
    n = int(st.shptr[st.npts])
    np.testing.assert_allclose(back[:, :n], ref_back[:, :n], rtol=1e-4, atol=2e-6 * np.abs(ref_back).max())
    # and it is close to the identity (the reduced Gaussian grid aliases the highest modes only slightly)
    assert np.abs(back[0, :n] - src[0, :n]).max() < 2e-3 * np.abs(src).max()
