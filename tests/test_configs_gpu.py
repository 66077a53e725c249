"""Parity at the SHAPE of every BASELINE.json config that is not covered by tests/test_fullsize_gpu.py (cfg2):

  cfg1   SolveRTE-style scalar solve (NMU=8 / NPHI=16, adaptive grid) + orthographic render of a small cube cloud
  cfg3   NSTOKES=3 at full size (32x37x27 LES-like adaptive grid, NLM=256, 9 x 200x200 perspective rays):
         RENDER (bit-exact walk), Levis gradient, COMPUTE_SOURCE
  cfg4s  NPART=2 (cloud + Rayleigh), 1.05 M grid points, 9 x 256x256 rays: RENDER over the ocean BRDF, gradient with a
         chunked derivative pass (the visit records of all rays do not fit the record budget), additivity over shards

The oracle needs minutes for a whole config on one core, so -- as in test_fullsize_gpu.py -- the GPU is compared with it
on seeded samples of the rays / pixels, and on everything through size-independent properties."""
import os
import sys
import types
import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def sample(rays, idx):
    from at3d_b200.state import Rays
    return Rays(rays.camx[idx], rays.camy[idx], rays.camz[idx], rays.cammu[idx], rays.camphi[idx])


def sample_pixels(pix, idx):
    from at3d_b200.gradsetup import PixelData
    return PixelData(pix.measurements[:, idx], pix.uncertainties[:, :, idx], pix.rays_per_pixel[idx], pix.ray_weights[idx],
                     pix.stokes_weights[:, idx])


def stokes_close(out, ref):
    """north_star: relative 1e-4, absolute 1e-6 on near-zero Q/U (scaled by the largest radiance of the sample)."""
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())


def build(workload):
    import bench
    from at3d_b200 import backend as B
    sc, rays, cfg = bench.build_scene(types.SimpleNamespace(workload=workload, pixels=0))
    B.finalize_scene(sc)
    return sc, rays


# ------------------------------------------------------------------------------------------------------------------
# cfg1
# ------------------------------------------------------------------------------------------------------------------
def test_cfg1_solve_and_orthographic_render(oracle):
    """BASELINE configs[0]: adaptive SHDOM solve (SPLIT_GRID + INIT_RADIANCE, NMU=8 / NPHI=16 -> NLM=64) of a cube cloud
    and 9 orthographic views (the view_source_kernel path: one source evaluation per grid point and view)."""
    from at3d_b200 import solver, synthetic as S
    from at3d_b200.device import DeviceState
    sc = S.make_scene(nx=14, ny=14, nz=13, nmu=8, nphi=16, nstokes=1, bc='periodic', dx=0.04, dy=0.04, dz=0.04,
                      cloud='blob', ext_max=30.0, nsplits=0, seed=101)
    oracle.finalize_scene(sc)
    st, pg = sc.state, sc.pg
    assert st.nlm == 64
    wtmu = (st.wtdo[:, 0] / (np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32))).astype(np.float32)
    par = dict(splitacc=0.1, shacc=0.003, solacc=1e-4, maxiter=60, adapt_grid_factor=6.0)
    ref, it_r, sc_r, sp_r = oracle.solve_adaptive(st, pg, wtmu, **par)
    out, it_g, sc_g, sp_g = solver.solve_adaptive(st, pg, wtmu, **par)
    assert ref.npts > st.npts and (out.npts, out.ncells, it_g) == (ref.npts, ref.ncells, it_r)
    for k in ('gridptr', 'neighptr', 'treeptr', 'cellflags', 'gridpos'):
        np.testing.assert_array_equal(getattr(out, k), getattr(ref, k), err_msg=k)
    np.testing.assert_array_equal(out.shptr, ref.shptr)
    np.testing.assert_allclose(out.source, ref.source, rtol=1e-4, atol=3e-6 * np.abs(ref.source).max())
    views = [S.orthographic_rays(sc, abs(z), 0.0 if z >= 0 else 180.0, 0.02)[0]
             for z in (70.5, 60.0, 45.6, 26.1, 0.0, -26.1, -45.6, -60.0, -70.5)]
    rays = S.concat_rays(views)
    dev = DeviceState(out)
    rad, tr = dev.render(rays, trace_cap=160)
    radref, trref, _ = oracle.render(out, rays, trace_cap=160, nthreads=NT)      # same (GPU-solved) state on both sides
    dev.close()
    np.testing.assert_array_equal(tr['ncells'], trref['ncells'])
    np.testing.assert_array_equal(tr['cells'], trref['cells'])
    stokes_close(rad, radref)
    # and the oracle's own solution rendered by the oracle: the two end-to-end chains agree
    radchain = oracle.render(ref, rays, nthreads=NT)
    np.testing.assert_allclose(rad, radchain, rtol=2e-4, atol=3e-6 * np.abs(radchain).max())


# ------------------------------------------------------------------------------------------------------------------
# cfg3
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def cfg3(oracle):
    from at3d_b200.device import DeviceState
    sc, rays = build('cfg3')
    dev = DeviceState(sc.state)
    yield sc, rays, dev
    dev.close()


def test_cfg3_shape(cfg3):
    sc, rays, dev = cfg3
    st = sc.state
    assert (st.nstokes, st.nlm, rays.nrays) == (3, 256, 9 * 200 * 200) and st.npts > 35000


def test_cfg3_render_sample_matches_oracle_with_bit_exact_walk(cfg3, oracle):
    sc, rays, dev = cfg3
    idx = np.sort(np.random.default_rng(17).choice(rays.nrays, 900, replace=False))
    sub = sample(rays, idx)
    ref, tref, _ = oracle.render(sc.state, sub, trace_cap=192, nthreads=NT)
    out, tr = dev.render(sub, trace_cap=192)
    np.testing.assert_array_equal(tr['ncells'], tref['ncells'])
    np.testing.assert_array_equal(tr['cells'], tref['cells'])
    np.testing.assert_array_equal(tr['nsub'], tref['nsub'])
    stokes_close(out, ref)
    assert np.abs(ref[1:]).max() > 1e-4                                   # Q and U are exercised
    full = dev.render(rays)
    np.testing.assert_array_equal(full[:, idx], out)


@pytest.fixture(scope='module')
def cfg3_gradient(cfg3):
    from at3d_b200 import backend as B, gradsetup
    sc, rays, dev = cfg3
    gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
    dev.attach_gradient(gi)
    rad = dev.render(rays)
    pix = gradsetup.make_pixels(3, rays.nrays, rad, seed=1)
    return gi, pix


def test_cfg3_gradient_pixel_sample_matches_oracle(cfg3, cfg3_gradient, oracle):
    from at3d_b200 import gradsetup
    sc, rays, dev = cfg3
    gi, pix = cfg3_gradient
    idx = np.sort(np.random.default_rng(19).choice(rays.nrays, 700, replace=False))
    sub, sp = sample(rays, idx), sample_pixels(pix, idx)
    g, cost, so = dev.gradient(sub, sp)
    gref, cref, soref = oracle.levisapprox_gradient(sc.state, sub, gradsetup.with_pixels(gi, sp), nthreads=NT)[:3]
    stokes_close(so, soref)
    assert abs(float(cost[0]) - cref) <= 1e-4 * abs(cref)
    np.testing.assert_allclose(g, gref, rtol=1e-4, atol=1e-4 * np.abs(gref).max())
    assert np.count_nonzero(gref) > 500


def test_cfg3_gradient_is_additive_over_pixel_shards_and_reproducible(cfg3, cfg3_gradient):
    from at3d_b200.parallel import shard_for_rank
    sc, rays, dev = cfg3
    gi, pix = cfg3_gradient
    g, cost, so = dev.gradient(rays, pix)
    g2, cost2, so2 = dev.gradient(rays, pix)
    np.testing.assert_array_equal(g, g2)                                   # deterministic accumulation, bit for bit
    assert float(cost[0]) == float(cost2[0])
    gsum, csum = np.zeros_like(g), 0.0
    for rank in range(3):
        r0, r1, p0, p1 = shard_for_rank(pix.rays_per_pixel, rank, 3)
        spx, a, b = pix.slice_pixels(p0, p1)
        gk, ck, sk = dev.gradient(rays.slice(r0, r1), spx)
        gsum += gk; csum += float(ck[0])
        np.testing.assert_array_equal(sk, so[:, p0:p1])
    assert abs(csum - float(cost[0])) <= 1e-12 * abs(float(cost[0]))
    np.testing.assert_allclose(gsum, g, rtol=1e-7, atol=1e-9 * np.abs(g).max())


def test_cfg3_compute_source_matches_oracle(cfg3, oracle):
    """COMPUTE_SOURCE with NSTOKES=3 (Wigner / 6-component Legendre terms) on all points x NLM=256: SHPTR bit-exact,
    SOURCE (I, Q, U) within 1e-5, norms within 1e-4 of the f64-summed value."""
    from at3d_b200 import backend as B
    sc, rays, dev = cfg3
    st = sc.state
    npts, tot, maxiv = st.npts, int(st.shptr[st.npts]), st.nlm * st.npts
    source = np.zeros((3, maxiv), np.float32, order='F')
    source[:, :tot] = st.source[:, :tot]
    delsource = np.zeros((3, maxiv), np.float32, order='F')
    delsource[:, :tot] = 0.01 * st.source[:, :tot]
    args = lambda: (st, st.shptr.copy(), source.copy(order='F'), st.shptr.copy(), delsource.copy(order='F'))
    a = B.compute_source(*args(), maxiv=maxiv, shacc=0.003)
    b = oracle.compute_source(*args(), maxiv=maxiv, shacc=0.003)
    sums64 = oracle.compute_source_sums64()
    assert a[0] == b[0] == 0
    np.testing.assert_array_equal(a[1], b[1])
    n = int(a[1][npts])
    np.testing.assert_allclose(a[2][:, :n], b[2][:, :n], rtol=1e-5, atol=1e-6 * np.abs(b[2][:, :n]).max())
    np.testing.assert_allclose(np.asarray(a[5], np.float64), np.asarray(sums64), rtol=1e-4)


# ------------------------------------------------------------------------------------------------------------------
# cfg4s
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def cfg4s(oracle):
    from at3d_b200.device import DeviceState
    sc, rays = build('cfg4s')
    dev = DeviceState(sc.state)
    yield sc, rays, dev
    dev.close()


def test_cfg4s_shape(cfg4s):
    sc, rays, dev = cfg4s
    st = sc.state
    assert (st.npart, st.nstokes, rays.nrays) == (2, 1, 9 * 256 * 256) and st.npts > 1000000


def test_cfg4s_render_over_ocean_sample_matches_oracle(cfg4s, oracle):
    """BASELINE configs[3] names an ocean BRDF: the Cox-Munk ocean surface (SURFACE_BRDF 'O') under the cfg4s medium."""
    from at3d_b200 import synthetic as S
    from at3d_b200.device import DeviceState
    sc, rays, dev = cfg4s
    st = S.with_brdf_surface(sc.state, 'O', seed=3)
    idx = np.sort(np.random.default_rng(23).choice(rays.nrays, 1200, replace=False))
    sub = sample(rays, idx)
    ref, tref, _ = oracle.render(st, sub, trace_cap=512, nthreads=NT)
    devo = DeviceState(st)
    out, tr = devo.render(sub, trace_cap=512)
    devo.close()
    np.testing.assert_array_equal(tr['ncells'], tref['ncells'])
    np.testing.assert_array_equal(tr['cells'], tref['cells'])
    np.testing.assert_array_equal(tr['nsub'], tref['nsub'])
    stokes_close(out, ref)


def test_cfg4s_render_sample_matches_oracle(cfg4s, oracle):
    sc, rays, dev = cfg4s
    idx = np.sort(np.random.default_rng(29).choice(rays.nrays, 1200, replace=False))
    sub = sample(rays, idx)
    ref, tref, _ = oracle.render(sc.state, sub, trace_cap=512, nthreads=NT)
    out, tr = dev.render(sub, trace_cap=512)
    np.testing.assert_array_equal(tr['cells'], tref['cells'])
    np.testing.assert_array_equal(tr['nsub'], tref['nsub'])
    stokes_close(out, ref)


@pytest.fixture(scope='module')
def cfg4s_gradient(cfg4s):
    from at3d_b200 import backend as B, gradsetup
    sc, rays, dev = cfg4s
    gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
    dev.attach_gradient(gi)
    rad = dev.render(rays)
    pix = gradsetup.make_pixels(1, rays.nrays, rad, seed=1)
    return gi, pix


def test_cfg4s_gradient_pixel_sample_matches_oracle(cfg4s, cfg4s_gradient, oracle):
    from at3d_b200 import gradsetup
    sc, rays, dev = cfg4s
    gi, pix = cfg4s_gradient
    idx = np.sort(np.random.default_rng(31).choice(rays.nrays, 800, replace=False))
    sub, sp = sample(rays, idx), sample_pixels(pix, idx)
    g, cost, so = dev.gradient(sub, sp)
    gref, cref, soref = oracle.levisapprox_gradient(sc.state, sub, gradsetup.with_pixels(gi, sp), nthreads=NT)[:3]
    stokes_close(so, soref)
    assert abs(float(cost[0]) - cref) <= 1e-4 * abs(cref)
    np.testing.assert_allclose(g, gref, rtol=1e-4, atol=1e-4 * np.abs(gref).max())


def test_cfg4s_chunked_derivative_pass_equals_the_single_pass(cfg4s, cfg4s_gradient):
    """All 590 k rays: once with the default record budget and once with a budget that forces the derivative pass
    into many chunks of rays (what cfg4 needs at 2.36 M rays); the pixel values are identical and the gradient
    agrees to the rounding of its FP64 sums (bit for bit with deterministic accumulation)."""
    sc, rays, dev = cfg4s
    gi, pix = cfg4s_gradient
    g, cost, so = dev.gradient(rays, pix)
    old = os.environ.get('AT3D_B200_REC_GB')
    os.environ['AT3D_B200_REC_GB'] = '0.25'
    try:
        gc, costc, soc = dev.gradient(rays, pix)
    finally:
        if old is None:
            del os.environ['AT3D_B200_REC_GB']
        else:
            os.environ['AT3D_B200_REC_GB'] = old
    np.testing.assert_array_equal(soc, so)
    assert float(costc[0]) == float(cost[0])
    np.testing.assert_allclose(gc, g, rtol=1e-9, atol=1e-12 * np.abs(g).max())
    assert np.count_nonzero(g) > 100000 and np.all(np.isfinite(g))


def test_cfg4s_streaming_beam_derivative_equals_the_dense_lists(cfg4s, cfg4s_gradient):
    """All 590 k rays over 1.05 M grid points: the gradient with the direct-beam walks done inside the call (no DPATH / DPTR
    lists: what the cfg4 gradient needs, where they would take 125 GB) equals the one from the dense lists to the rounding of
    its FP64 sums, also with many more passes over point ranges.  Runs last: it re-attaches the derivative tables."""
    from at3d_b200 import backend as B, gradsetup
    sc, rays, dev = cfg4s
    gi, pix = cfg4s_gradient
    g, cost, so = dev.gradient(rays, pix)
    dev.attach_gradient(gradsetup.with_streaming_beam(gi, sc.state, sc.pg, B))
    gs, costs, sos = dev.gradient(rays, pix)
    # ~1e9 terms here: the streaming walks run in passes of 2^29 pairs, each pass summed on its own and added to GRADOUT, so
    # the sums agree to rounding (bit for bit in a single pass: tests/test_gradient_gpu.py)
    np.testing.assert_allclose(gs, g, rtol=1e-9, atol=1e-12 * np.abs(g).max())
    np.testing.assert_array_equal(sos, so)
    old = os.environ.get('AT3D_B200_BEAM_PAIRS')
    os.environ['AT3D_B200_BEAM_PAIRS'] = str(1 << 24)
    try:
        gp, costp, sop = dev.gradient(rays, pix)
    finally:
        if old is None:
            del os.environ['AT3D_B200_BEAM_PAIRS']
        else:
            os.environ['AT3D_B200_BEAM_PAIRS'] = old
    np.testing.assert_allclose(gp, g, rtol=1e-9, atol=1e-12 * np.abs(g).max())
    dev.attach_gradient(gi)
