"""save_forward_model / load_forward_model (at3d_b200/util.py; reference at3d/util.py:512-725): group tree, restored
types, the file-name rule, and -- on the GPU -- that the reloaded solvers reproduce the saved ones' radiances exactly."""
import types
import warnings
import numpy as np
import pytest
from at3d_b200 import util as U
from at3d_b200.containers import SensorsDict, SolversDict
import test_rte_gpu as T


def _sensors():
    sensors = SensorsDict()
    for seed in (0, 1):
        s = T.make_sensor(0.4, 0.35, seed)
        s['wavelength'] = 0.672
        s['use_subpixel_rays'] = True
        s['image_shape'] = np.array([9, 8])
        sensors.add_sensor('MISR', s)
    s = T.make_sensor(0.4, 0.35, 2)
    s['wavelength'] = 0.672
    sensors.add_sensor('MODIS', s)
    return sensors


def _stub_solvers(atmosphere=False):
    params, medium, source, surface = T.make_inputs(5, 4, 6, 'open', 3, True)
    cloud = medium['cloud']
    solvers = SolversDict()
    atm = {'temperature': np.full((5, 4, 6), 280.0, np.float32)} if atmosphere else None
    solvers[0.672] = types.SimpleNamespace(
        medium=medium, numerical_params=params, source=source, surface=surface, _nstokes=3, atmosphere=atm,
        _grid={k: cloud[k] for k in ('x', 'y', 'z', 'delx', 'dely')})
    return solvers


def test_tree_types_and_grid_round_trip(tmp_path):
    name = str(tmp_path / 'model.nc')
    sensors, solvers = _sensors(), _stub_solvers(atmosphere=True)
    assert U.save_forward_model(name, sensors, solvers) == name
    with np.load(name) as f:
        members = set(f.files)
    assert 'sensors/MISR/1/ray_mu' in members and 'sensors/MODIS/0/pixel_index' in members
    assert 'solvers/0.672/medium/rayleigh/legcoef' in members and 'solvers/0.672/numerical_parameters/num_stokes' in members
    assert 'solvers/0.672/atmosphere/temperature' in members and 'solvers/0.672/grid/delx' in members
    back, no_solvers, grid = U.load_forward_model(name, load_solver=False)
    assert isinstance(back, SensorsDict) and isinstance(no_solvers, SolversDict) and len(no_solvers) == 0
    assert list(back) == ['MISR', 'MODIS'] and len(back['MISR']['sensor_list']) == 2
    for key in sensors:
        for a, b in zip(sensors[key]['sensor_list'], back[key]['sensor_list']):
            assert set(a) == set(b)
            for k in a:
                np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))
            assert b['stokes'].dtype == bool and b['pixel_index'].dtype == np.int64 and b['ray_mu'].dtype == np.float64
    assert back['MISR']['sensor_list'][0]['use_subpixel_rays'] is True
    cloud = solvers[0.672].medium['cloud']
    np.testing.assert_array_equal(grid['z'], cloud['z'])
    assert float(grid['delx']) == cloud['delx']


def test_sensor_attributes_survive(tmp_path):
    from at3d_b200 import sensor as SN
    sensors = SensorsDict()
    box = {'x': np.linspace(0, 0.4, 9), 'y': np.linspace(0, 0.3, 7), 'z': np.linspace(0, 0.5, 6)}
    sensors.add_sensor('MISR', SN.orthographic_projection(0.672, box, 0.05, 0.05, 30.0, 40.0,
                                                          sub_pixel_ray_args={'method': SN.gaussian, 'degree': (2, 3)}))
    sensors.add_sensor('MISR', SN.perspective_projection(0.672, 20.0, 5, 4, [0.2, 0.1, 3.0], [0.2, 0.15, 0.2], [0, 1, 0]))
    name = U.save_forward_model(str(tmp_path / 'm.nc'), sensors, _stub_solvers())
    back = U.load_sensors(name)
    for a, b in zip(sensors['MISR']['sensor_list'], back['MISR']['sensor_list']):
        assert set(a) == set(b) and set(a.attrs) == set(b.attrs)
        for k in a:
            np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))
        for k in a.attrs:
            np.testing.assert_array_equal(np.asarray(a.attrs[k]), np.asarray(b.attrs[k]))
    first = back['MISR']['sensor_list'][0]
    assert first.attrs['projection'] == 'Orthographic' and first.attrs['sub_pixel_ray_args_method'] == 'gaussian'
    assert list(first.attrs['sub_pixel_ray_args_degree']) == [2, 3] and first['use_subpixel_rays'] is True


def test_existing_file_gets_a_numbered_name(tmp_path):
    name = str(tmp_path / 'model.nc')
    sensors, solvers = _sensors(), _stub_solvers()
    U.save_forward_model(name, sensors, solvers)
    for expect in ('model_1.nc', 'model_2.nc'):
        with pytest.warns(RuntimeWarning, match='already exists'):
            out = U.save_forward_model(name, sensors, solvers)
        assert out == str(tmp_path / expect)
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        U.save_forward_model(str(tmp_path / 'other.nc'), sensors, solvers)


def test_type_checks(tmp_path):
    with pytest.raises(TypeError):
        U.save_forward_model(str(tmp_path / 'a.nc'), {}, _stub_solvers())
    with pytest.raises(TypeError):
        U.save_forward_model(str(tmp_path / 'a.nc'), _sensors(), {})
    bad = _sensors()
    bad['MODIS']['sensor_list'][0]['note'] = np.array([object()], dtype=object)
    with pytest.raises(TypeError, match='Python objects'):
        U.save_sensors(str(tmp_path / 'b.nc'), bad)


@pytest.mark.gpu
def test_reloaded_solvers_reproduce_the_radiances(tmp_path):
    from at3d_b200.rte import RTE
    name = str(tmp_path / 'model.nc')
    sensors, solvers = _sensors(), SolversDict()
    params, medium, source, surface = T.make_inputs(9, 8, 11, 'open', 3, True)
    solvers.add_solver(0.672, RTE(params, medium, source, surface, num_stokes=3))
    U.save_forward_model(name, sensors, solvers)
    back_sensors, back_solvers, grid = U.load_forward_model(name)
    assert list(back_solvers) == [0.672] and isinstance(back_solvers[0.672], RTE)
    assert back_solvers[0.672]._nstokes == 3 and list(back_solvers[0.672].medium) == ['cloud', 'rayleigh']
    np.testing.assert_array_equal(grid['x'], medium['cloud']['x'])
    solvers.solve(maxiter=40, verbose=False)
    back_solvers.solve(maxiter=40, verbose=False)
    a = solvers[0.672].integrate_to_sensor(sensors['MISR']['sensor_list'][0])
    b = back_solvers[0.672].integrate_to_sensor(back_sensors['MISR']['sensor_list'][0])
    for k in ('I', 'Q', 'U'):
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))
    assert np.asarray(a['I']).max() > 0
