"""at3d_b200/sensor.py and at3d_b200/transforms.py against outputs of the reference's own modules
(tests/golden/make_host_goldens.py ran at3d/sensor.py, transforms.py, uncertainties.py and parallel.py unmodified)."""
import os
import warnings
import numpy as np
import pytest
from at3d_b200 import sensor as SN
from at3d_b200 import transforms as TR

def same(mine, ref, err_msg=''):
    """Equal to the reference's output: exactly for integers / booleans / strings, to 1e-12 relative for floats (the
    goldens match bit for bit on the host that made them; libm / SIMD paths of exp, log, cos differ by an ulp between CPUs)."""
    mine, ref = np.asarray(mine), np.asarray(ref)
    if ref.dtype.kind in 'fc':
        np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=1e-300, equal_nan=True, err_msg=err_msg)
    else:
        np.testing.assert_array_equal(mine, ref, err_msg=err_msg)


GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'host_goldens.npz'))
BOX = {'x': np.linspace(0.0, 0.64, 33), 'y': np.linspace(0.0, 0.72, 37), 'z': np.linspace(0.0, 1.04, 27)}

SENSORS = {
    'ortho': lambda: SN.orthographic_projection(0.672, BOX, 0.08, 0.09, 37.0, 25.0, stokes=['I', 'Q']),
    'ortho_gauss': lambda: SN.orthographic_projection(0.672, BOX, 0.16, 0.12, 200.0, 40.0, altitude=1.5, stokes='I',
                                                      sub_pixel_ray_args={'method': SN.gaussian, 'degree': (2, 3)}),
    'ortho_uniform': lambda: SN.orthographic_projection(1.65, BOX, 0.2, 0.2, 0.0, 0.0,
                                                        sub_pixel_ray_args={'method': SN.uniform, 'nrays': 2}),
    'persp': lambda: SN.perspective_projection(0.672, 17.0, 7, 5, [0.3, -0.2, 5.0], [0.32, 0.36, 0.5], [0, 1, 0],
                                               stokes=['I', 'Q', 'U']),
    'persp_stoch': lambda: SN.perspective_projection(0.672, 30.0, 4, 6, [2.0, 1.5, 3.0], [0.3, 0.3, 0.4], [0, 0, 1], stokes='I',
                                                     sub_pixel_ray_args={'method': SN.stochastic, 'nrays': (3, 2), 'seed': 11}),
    'domaintop': lambda: SN.domaintop_projection(0.672, BOX, 0.1, 0.15, 75.0, 35.0, x_offset=0.01, y_offset=-0.02),
}


@pytest.mark.parametrize('case', sorted(SENSORS))
def test_sensor_matches_the_reference(case):
    got = SENSORS[case]()
    names = [k.split('/', 1)[1] for k in GOLD.files if k.startswith(case + '/')]
    assert set(names) == set(got) and len(names) >= 14
    for name in names:
        ref, mine = GOLD[case + '/' + name], np.asarray(got[name])
        assert mine.shape == ref.shape, name
        assert mine.dtype.kind == ref.dtype.kind, name
        same(mine, ref, err_msg=name)
    # what the ray lists must satisfy whatever the projection
    w = np.bincount(got['pixel_index'], weights=got['ray_weight'])
    np.testing.assert_allclose(w, 1.0, rtol=1e-12)
    assert got['stokes'].dtype == bool and got['stokes'].shape == (4,)


def test_sensor_argument_checks():
    with pytest.raises(ValueError, match='1-D'):
        SN.make_sensor_dataset(np.zeros((2, 2)), np.zeros(4), np.zeros(4), np.ones(4), np.zeros(4), 'I', 0.6)
    with pytest.raises(ValueError, match='same size'):
        SN.make_sensor_dataset(np.zeros(3), np.zeros(4), np.zeros(4), np.ones(4), np.zeros(4), 'I', 0.6)
    with pytest.raises(ValueError, match='altitudes'):
        SN.make_sensor_dataset(np.zeros(2), np.zeros(2), -np.ones(2), np.ones(2), np.zeros(2), 'I', 0.6)
    with pytest.raises(ValueError, match='0.0 are not allowed'):
        SN.make_sensor_dataset(np.zeros(2), np.zeros(2), np.ones(2), np.zeros(2), np.zeros(2), 'I', 0.6)
    with pytest.raises(ValueError, match='Stokes'):
        SN.make_sensor_dataset(np.zeros(2), np.zeros(2), np.ones(2), np.ones(2), np.zeros(2), 'X', 0.6)
    with pytest.raises(KeyError, match='Invalid kwarg'):
        SN.orthographic_projection(0.6, BOX, 0.2, 0.2, 0.0, 0.0, sub_pixel_ray_args={'method': SN.gaussian, 'nrays': 2})
    with pytest.raises(TypeError, match='callable'):
        SN.orthographic_projection(0.6, BOX, 0.2, 0.2, 0.0, 0.0, sub_pixel_ray_args={'method': 'gaussian'})
    s = SN.make_sensor_dataset(np.zeros(2), np.zeros(2), np.ones(2), np.ones(2), np.zeros(2), ['I', 'U'], 0.6,
                               fill_ray_variables=True)
    assert list(s['stokes']) == [True, False, True, False] and s['use_subpixel_rays'] is False
    np.testing.assert_array_equal(s['pixel_index'], [0, 1])


COORD = {'null': TR.CoordinateTransform(), 'log': TR.CoordinateTransformLog(), 'scaling': TR.CoordinateTransformScaling(2.0, 0.25),
         'exp': TR.CoordinateTransformExp(10.0), 'hyperbol': TR.CoordinateTransformHyperBol(0.05)}


@pytest.mark.parametrize('name', sorted(COORD))
def test_coordinate_transform_matches_the_reference(name):
    tr, phys, grad = COORD[name], GOLD['tr/phys'], GOLD['tr/grad']
    a = tr.inverse_transform(phys)
    same(a, GOLD['tr/%s/abstract' % name])
    same(tr(a), GOLD['tr/%s/physical' % name])
    same(tr.gradient_transform(a, grad), GOLD['tr/%s/gradient' % name])


@pytest.mark.parametrize('name,cls', [('mask', TR.StateToGridMask), ('2d', TR.StateToGrid2D),
                                      ('uniform', TR.StateToGridUniform), ('profile', TR.StateToGridProfile)])
def test_state_to_grid_matches_the_reference(name, cls):
    mask, data = GOLD['s2g/mask'], GOLD['s2g/data']
    s2g = cls(mask=mask)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)              # empty levels / columns: NaN, as the reference
        state = s2g.inverse_transform(data)
        same(state, GOLD['s2g/%s/state' % name])
        same(s2g.gradient_transform(data), GOLD['s2g/%s/gradient' % name])
        same(s2g.inverse_bounds_transform(np.full(mask.shape, 3.0)), GOLD['s2g/%s/bounds' % name])
    assert state.shape == (s2g.state_size,)
    if name != 'profile':
        same(s2g(state), GOLD['s2g/%s/gridded' % name])


@pytest.mark.parametrize('case', range(4))
def test_subdivide_raytrace_jobs_matches_the_reference(case):
    from collections import OrderedDict
    from at3d_b200.parallel import subdivide_raytrace_jobs
    sensors = OrderedDict()
    i = 0
    while 'jobs/%d/rpp%d' % (case, i) in GOLD.files:
        sensors[0.4 + 0.1 * i] = {'rays_per_pixel': GOLD['jobs/%d/rpp%d' % (case, i)]}
        i += 1
    n_jobs, job_factor = GOLD['jobs/%d/args' % case]
    keys, rays, pixels = subdivide_raytrace_jobs(sensors, int(n_jobs), int(job_factor))
    np.testing.assert_array_equal(np.array(keys), GOLD['jobs/%d/keys' % case])
    np.testing.assert_array_equal(np.array(rays, np.int64), GOLD['jobs/%d/rays' % case])
    np.testing.assert_array_equal(np.array(pixels, np.int64), GOLD['jobs/%d/pixels' % case])


UNC = {'null': lambda U: U.NullUncertainty('L2', 2.5),
       'radiometric': lambda U: U.RadiometricUncertainty('L2', lambda r: 200.0 * np.sqrt(r / 0.1), 1e-4, 0.03, 0.01, seed=5),
       'radiometric_ll': lambda U: U.RadiometricUncertainty('LL', lambda r: 150.0 + 0.0 * r, 2e-4, 0.02, 0.0, seed=9),
       'tandem': lambda U: U.TandemStereoCamera('L2')}


@pytest.mark.parametrize('name', sorted(UNC))
def test_uncertainty_model_matches_the_reference(name):
    from at3d_b200 import uncertainties as U
    radiance = GOLD['unc/I']
    np.random.seed(123)
    model = UNC[name](U)
    sensor = {'I': radiance.copy(), 'npixels': np.arange(radiance.size), 'stokes': np.array([True, False, False, False])}
    model.calculate_uncertainties(sensor)
    ref = GOLD['unc/%s/uncertainties' % name]
    assert sensor['uncertainties'].shape == ref.shape == (model.num_uncertainty, model.num_uncertainty, radiance.size)
    same(sensor['uncertainties'], ref)
    if name == 'null':
        with pytest.raises(ValueError, match='cannot be used to generate measurement noise'):
            model.add_noise(sensor)
        return
    np.random.seed(77)
    model.add_noise(sensor)
    same(sensor['I'], GOLD['unc/%s/noisy' % name])
    assert np.any(sensor['I'] != radiance)


def test_uncertainty_argument_checks():
    from at3d_b200 import uncertainties as U
    with pytest.raises(NotImplementedError):
        U.NullUncertainty('L1')
    m = U.NullUncertainty('LL')
    assert m.num_uncertainty == 2 and m.cost_function == 'LL' and m.valid_cost_functions == ('L2', 'LL')
    model = U.RadiometricUncertainty('L2', lambda r: 100.0 + 0 * r, 1e-4)
    with pytest.raises(KeyError, match="Stokes component 'Q'"):
        model.add_noise({'I': np.ones(3), 'stokes': np.array([True, True, False, False])})


def test_make_grid():
    from at3d_b200.grid import make_grid
    g = make_grid(0.02, 5, 0.03, 4, [0.0, 0.1, 0.4], nz=7)
    np.testing.assert_array_equal(g['x'], np.linspace(0.0, 0.08, 5)); np.testing.assert_array_equal(g['y'], np.linspace(0.0, 0.09, 4))
    assert g['delx'] == 0.02 and g['dely'] == 0.03 and g['nz'] == 7 and 'nx' not in g
    for bad in ([0.1], [0.0, 0.2, 0.2], [0.3, 0.1], [-0.1, 0.2], [[0.0, 1.0]]):
        with pytest.raises(ValueError, match='strictly increasing'):
            make_grid(0.02, 5, 0.03, 4, bad)
