"""BASELINE.json configs[1] at FULL size (LES-like 32x37x27 adaptive grid, NLM=256, 9 perspective views x 200x200 =
360 k rays, radiance + Levis gradient): the oracle needs minutes for all of it on one core, so parity is checked
(a) against the oracle on seeded random samples of the rays / pixels and (b) through size-independent properties of
the operators: shard invariance (the reference slices the rays across workers, at3d/parallel.py:114-174), exact
linearity of RENDER in the sources, additivity of the gradient and the cost over pixels."""
import os
import sys
import types
import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cfg2(oracle):
    import bench
    from at3d_b200 import backend as B
    from at3d_b200.device import DeviceState
    args = types.SimpleNamespace(workload='cfg2', pixels=0)
    sc, rays, cfg = bench.build_scene(args)
    B.finalize_scene(sc)
    dev = DeviceState(sc.state)
    yield sc, rays, dev
    dev.close()


def sample(rays, idx):
    from at3d_b200.state import Rays
    return Rays(rays.camx[idx], rays.camy[idx], rays.camz[idx], rays.cammu[idx], rays.camphi[idx])


def test_full_size_workload_shape(cfg2):
    sc, rays, dev = cfg2
    assert rays.nrays == 9 * 200 * 200 and sc.state.nlm == 256 and sc.state.npts > 35000


def test_render_sample_matches_oracle_with_bit_exact_walk(cfg2, oracle):
    sc, rays, dev = cfg2
    idx = np.sort(np.random.default_rng(7).choice(rays.nrays, 1500, replace=False))
    sub = sample(rays, idx)
    ref, tref, _ = oracle.render(sc.state, sub, trace_cap=192, nthreads=os.cpu_count() or 1)
    out, tr = dev.render(sub, trace_cap=192)
    np.testing.assert_array_equal(tr['ncells'], tref['ncells'])
    np.testing.assert_array_equal(tr['cells'], tref['cells'])
    np.testing.assert_array_equal(tr['nsub'], tref['nsub'])
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())
    # the same rays inside the full call
    full = dev.render(rays)
    np.testing.assert_array_equal(full[:, idx], out)


def test_render_is_shard_invariant(cfg2):
    sc, rays, dev = cfg2
    from at3d_b200.parallel import shard_for_rank
    full = dev.render(rays)
    rpp = np.ones(rays.nrays, np.int32)
    parts = []
    for rank in range(8):
        r0, r1, p0, p1 = shard_for_rank(rpp, rank, 8)
        parts.append(dev.render(rays.slice(r0, r1)))
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), full)
    assert np.all(np.isfinite(full)) and full.min() >= 0.0


def test_render_is_exactly_linear_in_the_sources(cfg2):
    """Everything that emits (SOURCE, the direct beam behind the single-scattering and the surface terms, the downwelling
    flux behind the Lambertian surface, the sky radiance) times 2 -> every radiance times 2, bit for bit (a power of two
    commutes with every rounding on the path; extinction and the walk are untouched)."""
    sc, rays, dev = cfg2
    from at3d_b200.device import DeviceState
    idx = np.arange(0, rays.nrays, 7)
    sub = sample(rays, idx)
    base = dev.render(sub)
    st2 = sc.state.copy()
    st2.source = np.asfortranarray(st2.source * np.float32(2.0))
    st2.dirflux = (st2.dirflux * np.float32(2.0)).astype(np.float32)
    st2.fluxes = np.asfortranarray(st2.fluxes * np.float32(2.0))
    st2.skyrad = np.asfortranarray(st2.skyrad * np.float32(2.0))
    st2.bcrad = np.asfortranarray(st2.bcrad * np.float32(2.0))
    dev2 = DeviceState(st2.normalize())
    out = dev2.render(sub)
    dev2.close()
    np.testing.assert_array_equal(out, np.float32(2.0) * base)


@pytest.fixture(scope='module')
def cfg2_gradient(cfg2, oracle):
    from at3d_b200 import backend as B, gradsetup
    sc, rays, dev = cfg2
    gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
    dev.attach_gradient(gi)
    rad = dev.render(rays)
    pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=1)
    return gi, pix


def test_gradient_and_cost_are_additive_over_pixel_shards(cfg2, cfg2_gradient):
    sc, rays, dev = cfg2
    gi, pix = cfg2_gradient
    from at3d_b200.parallel import shard_for_rank
    g, cost, so = dev.gradient(rays, pix)
    gsum, csum, parts = np.zeros_like(g), 0.0, []
    for rank in range(4):
        r0, r1, p0, p1 = shard_for_rank(pix.rays_per_pixel, rank, 4)
        sp, a, b = pix.slice_pixels(p0, p1)
        assert (a, b) == (r0, r1)
        gk, ck, sk = dev.gradient(rays.slice(r0, r1), sp)
        gsum += gk; csum += float(ck[0]); parts.append(sk)
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), so)           # pixel values do not depend on the shard
    assert abs(csum - float(cost[0])) <= 1e-12 * abs(float(cost[0]))
    np.testing.assert_allclose(gsum, g, rtol=1e-9, atol=1e-12 * np.abs(g).max())   # FP64 sums in another order
    assert np.count_nonzero(g) > 1000


def test_gradient_pixel_sample_matches_oracle(cfg2, cfg2_gradient, oracle):
    sc, rays, dev = cfg2
    gi, pix = cfg2_gradient
    from at3d_b200 import gradsetup
    from at3d_b200.gradsetup import PixelData
    idx = np.sort(np.random.default_rng(11).choice(rays.nrays, 1200, replace=False))
    sub = sample(rays, idx)
    sp = PixelData(pix.measurements[:, idx], pix.uncertainties[:, :, idx], pix.rays_per_pixel[idx], pix.ray_weights[idx],
                   pix.stokes_weights[:, idx])
    g, cost, so = dev.gradient(sub, sp)
    gref, cref, soref = oracle.levisapprox_gradient(sc.state, sub, gradsetup.with_pixels(gi, sp), nthreads=os.cpu_count() or 1)[:3]
    np.testing.assert_allclose(so, soref, rtol=1e-4, atol=1e-6 * np.abs(soref).max())
    assert abs(float(cost[0]) - cref) <= 1e-4 * abs(cref)
    scale = np.abs(gref).max()
    np.testing.assert_allclose(g, gref, rtol=1e-4, atol=1e-4 * scale)


def test_compute_source_full_size_matches_oracle(cfg2, oracle):
    """COMPUTE_SOURCE on all 38.6 k points x NLM=256: SHPTR (adaptive truncation) bit-exact, SOURCE within 1e-5, the four
    norms within 1e-4 of their f64-summed value (see below)."""
    from at3d_b200 import backend as B
    sc, rays, dev = cfg2
    st = sc.state
    npts = st.npts
    maxiv = st.nlm * npts
    tot = int(st.shptr[npts])
    source = np.zeros((st.nstokes, maxiv), np.float32, order='F')
    source[:, :tot] = st.source[:, :tot]
    delsource = np.zeros((st.nstokes, maxiv), np.float32, order='F')
    delsource[:, :tot] = 0.01 * st.source[:, :tot]
    a = B.compute_source(st, st.shptr.copy(), source.copy(order='F'), st.shptr.copy(), delsource.copy(order='F'), maxiv=maxiv, shacc=0.003)
    b = oracle.compute_source(st, st.shptr.copy(), source.copy(order='F'), st.shptr.copy(), delsource.copy(order='F'), maxiv=maxiv, shacc=0.003)
    sums64 = oracle.compute_source_sums64()
    assert a[0] == b[0] == 0
    np.testing.assert_array_equal(a[1], b[1])
    n = int(a[1][npts])
    np.testing.assert_allclose(a[2][:, :n], b[2][:, :n], rtol=1e-5, atol=1e-6 * np.abs(b[2][:, :n]).max())
    # the reference accumulates the four norms sequentially in REAL over ~10 M terms and loses the
    # small ones (all four oracle values come out low by ~2e-3); the GPU sums per point in f32 and across points in f64
    # -> the bar of 1e-4 is asserted against the oracle's f64-summed norms (the same REAL products, no sequential rounding);
    # the reference's own REAL value is documented by the 5e-3 bound
    np.testing.assert_allclose(np.asarray(a[5], np.float64), np.asarray(sums64), rtol=1e-4)
    np.testing.assert_allclose(np.asarray(b[5], np.float64), np.asarray(sums64), rtol=5e-3)
    assert np.all(np.abs(np.asarray(a[5])) >= np.abs(np.asarray(b[5])) * (1 - 1e-6))
