"""at3d_b200/transforms.py against the formulas of at3d/transforms.py: round trips, chain rule by finite differences where
the reference's gradient is the true derivative, the reference's literal formulas where it is not."""
import numpy as np
import pytest
from at3d_b200 import transforms as TR

RNG = np.random.default_rng(5)


@pytest.mark.parametrize('tr,lo,hi', [(TR.CoordinateTransform(), 0.1, 30.0), (TR.CoordinateTransformLog(), 0.1, 30.0),
                                      (TR.CoordinateTransformScaling(2.0, 0.25), 0.1, 30.0),
                                      (TR.CoordinateTransformHyperBol(0.05), 0.1, 30.0)])
def test_round_trip_and_chain_rule(tr, lo, hi):
    phys = RNG.uniform(lo, hi, 40)
    a = tr.inverse_transform(phys)
    np.testing.assert_allclose(tr(a), phys, rtol=1e-12)
    if isinstance(tr, TR.CoordinateTransformScaling):
        g = RNG.normal(size=40)
        np.testing.assert_array_equal(tr.gradient_transform(a, g), (g - 2.0) * 0.25)     # at3d/transforms.py:203-204
        return
    # d cost / d abstract = d cost / d physical * d physical / d abstract
    g = RNG.normal(size=40)
    h = 1e-6 * np.maximum(1.0, np.abs(a))
    dphys = (tr(a + h) - tr(a - h)) / (2 * h)
    np.testing.assert_allclose(tr.gradient_transform(a, g), g * dphys, rtol=1e-6)


def test_exp_transform_keeps_the_reference_formulas():
    tr = TR.CoordinateTransformExp(10.0)
    phys = RNG.uniform(0.1, 30.0, 20)
    a = tr.inverse_transform(phys)
    np.testing.assert_allclose(a, 1.0 - np.exp(-phys / 10.0), rtol=1e-15)
    np.testing.assert_allclose(tr(a), -phys, rtol=1e-12)                 # at3d/transforms.py:238-239: sign as the reference's
    g = RNG.normal(size=20)
    np.testing.assert_allclose(tr.gradient_transform(a, g), 10.0 * g / (1.0 - a), rtol=1e-15)


def _mask():
    m = np.zeros((4, 3, 6), bool)
    m[1:3, 0:2, 1:5] = True
    m[3, 2, 2] = True
    return m


def test_mask_map():
    m = _mask()
    s2g = TR.StateToGridMask(mask=m)
    assert s2g.state_size == m.sum()
    state = RNG.uniform(1, 2, s2g.state_size)
    g = s2g(state)
    assert g.shape == m.shape and np.all(g[~m] == 0)
    np.testing.assert_array_equal(s2g.inverse_transform(g), state)
    np.testing.assert_array_equal(s2g.gradient_transform(g), state)
    np.testing.assert_array_equal(s2g.inverse_bounds_transform(np.full(m.shape, 3.0)), np.full(m.sum(), 3.0))
    full = TR.StateToGridMask(grid_shape=(2, 3, 4))
    x = RNG.normal(size=24)
    np.testing.assert_array_equal(full(x), x.reshape(2, 3, 4))
    with pytest.raises(ValueError):
        TR.StateToGridMask()
    with pytest.raises(ValueError):
        TR.StateToGridMask(grid_shape=(2, 2, 2), mask=m)


def test_profile_2d_and_uniform_maps():
    m = _mask()
    data = RNG.uniform(1, 2, m.shape)
    prof = TR.StateToGridProfile(mask=m)
    assert prof.state_size == 6
    levels = np.arange(6) + 1.0
    g = prof(levels)
    assert np.all(np.isnan(g[~m]))
    for k in range(6):
        assert np.all(g[..., k][m[..., k]] == levels[k])
    with np.errstate(all='ignore'), pytest.warns(RuntimeWarning):
        back = prof.inverse_transform(data)                       # levels 0 and 5 have no masked point: NaN, as the reference
    for k in range(1, 5):
        assert back[k] == pytest.approx(data[..., k][m[..., k]].mean())
    assert np.isnan(back[0]) and np.isnan(back[5])
    m2 = m.copy(); m2[:, :, 0] = True
    col = TR.StateToGrid2D(mask=m2)
    assert col.state_size == 12
    s = RNG.uniform(1, 2, 12)
    g = col(s)
    np.testing.assert_array_equal(g[..., 0], s.reshape(4, 3))
    assert np.all(np.isnan(g[~m2]))
    np.testing.assert_allclose(col.inverse_transform(g), s, rtol=1e-15)
    np.testing.assert_allclose(col.gradient_transform(data).reshape(4, 3)[1, 1], data[1, 1][m2[1, 1]].mean())
    uni = TR.StateToGridUniform(mask=m)
    assert uni.state_size == 1
    g = uni(np.array([2.5]))
    assert np.all(g[m] == 2.5) and np.all(np.isnan(g[~m]))
    np.testing.assert_allclose(uni.inverse_transform(data), [data[m].mean()])
    np.testing.assert_allclose(uni.inverse_bounds_transform(np.full(m.shape, 7.0)), [7.0])
    with pytest.raises(NotImplementedError):
        uni.inverse_bounds_transform(data)


# ---- the state generator with transforms (at3d/medium.py StateGenerator / StateRepresentation) ----
def _generator(transforms=None, variables=('extinction', 'ssalb')):
    import test_rte_gpu as T
    from at3d_b200.containers import SolversDict, UnknownScatterers
    from at3d_b200.optimize import GridStateGenerator
    params, medium, source, surface = T.make_inputs(6, 5, 7, 'periodic', 1, True)
    unknown = UnknownScatterers()
    unknown.add_unknowns('cloud', list(variables))
    mask = medium['cloud']['extinction'] > 0
    solvers = SolversDict()
    gen = GridStateGenerator(solvers, unknown, {0.672: medium}, {0.672: source}, {0.672: surface}, {0.672: params},
                             {0.672: 1}, mask=mask, transforms=transforms)
    return gen, solvers, medium, mask


def test_generator_state_layout_gradient_and_bounds():
    gen0, _, medium, mask = _generator()
    n = int(mask.sum())
    assert gen0.state_size == 2 * n
    tr = {('cloud', 'extinction'): (TR.CoordinateTransformLog(), None),
          ('cloud', 'ssalb'): (None, TR.StateToGridUniform(mask=mask))}
    gen, _, _, _ = _generator(tr)
    assert gen.state_size == n + 1                                   # variables of unequal size (StateRepresentation)
    x = gen.get_state()
    np.testing.assert_allclose(x[:n], np.log(medium['cloud']['extinction'][mask].astype(np.float64)), rtol=1e-15)
    assert x[n] == pytest.approx(0.999, rel=1e-6)
    g = RNG.normal(size=mask.shape + (2,))
    pg = gen.project_gradient_to_state(x, {'gradient': g})
    np.testing.assert_allclose(pg[:n], g[..., 0][mask] * np.exp(x[:n]), rtol=1e-14)
    assert pg[n] == pytest.approx(g[..., 1][mask].mean())           # the reference's mean projection
    lo, hi = gen.transform_bounds({('cloud', 'extinction'): (1e-3, 200.0), ('cloud', 'ssalb'): (0.5, 1.0)})
    np.testing.assert_allclose(lo[:n], np.log(1e-3)); np.testing.assert_allclose(hi[:n], np.log(200.0))
    assert (lo[n], hi[n]) == (0.5, 1.0)
    # a transform that reverses the order of the bounds: they come back sorted
    gen2, _, _, _ = _generator({('cloud', 'extinction'): (TR.CoordinateTransformScaling(0.0, -1.0), None)}, ('extinction',))
    lo, hi = gen2.transform_bounds({('cloud', 'extinction'): (1.0, 5.0)})
    assert np.all(lo == -5.0) and np.all(hi == -1.0)
    with pytest.raises(KeyError):
        _generator({('rain', 'extinction'): (None, None)})
    with pytest.raises(ValueError):
        _generator({('cloud', 'ssalb'): (None, TR.StateToGridMask(grid_shape=(2, 2, 2)))})


@pytest.mark.gpu
def test_generator_builds_the_solvers_from_a_transformed_state():
    tr = {('cloud', 'extinction'): (TR.CoordinateTransformLog(), None),
          ('cloud', 'ssalb'): (None, None)}
    gen, solvers, medium, mask = _generator(tr)
    n = int(mask.sum())
    base = medium['cloud']['extinction'].copy()
    x = gen.get_state()
    x[:n] += np.log(1.5)
    x[n:] = 0.9
    gen(x)
    rte = solvers[0.672]
    ext, alb = rte.medium['cloud']['extinction'], rte.medium['cloud']['ssalb']
    np.testing.assert_allclose(ext[mask], 1.5 * base[mask], rtol=1e-6)
    np.testing.assert_array_equal(ext[~mask], base[~mask])
    assert np.all(alb[mask] == np.float32(0.9)) and np.all(alb[~mask] == np.float32(0.999))
    np.testing.assert_allclose(gen.get_state(), x, rtol=1e-6)
    # profile map: one extinction per level, NaN outside the mask never reaches the medium
    prof = TR.StateToGridProfile(mask=mask)
    gen2, solvers2, medium2, _ = _generator({('cloud', 'extinction'): (None, prof)}, ('extinction',))
    levels = np.linspace(2.0, 8.0, mask.shape[-1])
    gen2(levels)
    e2 = solvers2[0.672].medium['cloud']['extinction']
    assert np.all(np.isfinite(e2))
    for k in range(mask.shape[-1]):
        assert np.all(e2[..., k][mask[..., k]] == np.float32(levels[k]))
    np.testing.assert_array_equal(e2[~mask], medium2['cloud']['extinction'][~mask])
    solvers2[0.672].solve(maxiter=30)
    assert solvers2[0.672].check_solved(verbose=False)
    for s in (solvers, solvers2):
        for r in s.values():
            r.close()


def test_generator_accepts_a_user_defined_map():
    """A state_to_grid object with only the reference's four methods (no sizes, no mask attribute): one unknown per x-row."""
    class Rows:
        def __init__(self, mask):
            self.mask = mask

        def __call__(self, state):
            return np.where(self.mask, np.asarray(state)[:, None, None], np.nan)

        def inverse_transform(self, gridded):
            return np.nanmean(np.where(self.mask, gridded, np.nan), axis=(1, 2))

        def gradient_transform(self, gridded_gradient):
            return np.nansum(np.where(self.mask, gridded_gradient, np.nan), axis=(1, 2))

        def inverse_bounds_transform(self, gridded_bounds):
            return self.inverse_transform(gridded_bounds)

    _, _, medium, mask = _generator()
    mask = mask.copy(); mask[:, 0, 0] = True                      # every row has a point
    gen, _, _, _ = _generator({('cloud', 'extinction'): (TR.CoordinateTransformScaling(0.0, 0.5), Rows(mask))}, ('extinction',))
    assert gen.state_size == mask.shape[0]
    x = gen.get_state()
    assert x.shape == (mask.shape[0],) and np.all(np.isfinite(x))
    g = RNG.normal(size=mask.shape + (1,))
    np.testing.assert_allclose(gen.project_gradient_to_state(x, {'gradient': g}),
                               np.where(mask, g[..., 0], 0.0).sum(axis=(1, 2)) * 0.5, rtol=1e-13)
