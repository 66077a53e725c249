"""Host logic of the at3d-shaped containers (no GPU): grouping of sensors per solver, measurement bookkeeping."""
import numpy as np
from at3d_b200.containers import SensorsDict, SolversDict, UnknownScatterers
from at3d_b200.rte import RTE


def fake_solver(nstokes=1):
    r = RTE.__new__(RTE)               # bookkeeping only: no device, no grid
    r._nstokes = nstokes
    return r


def sensor(wl, npix, rays_per_pix, seed):
    rng = np.random.default_rng(seed)
    n = npix * rays_per_pix
    return dict(ray_x=rng.random(n), ray_y=rng.random(n), ray_z=np.ones(n), ray_mu=np.full(n, 0.9), ray_phi=np.zeros(n),
                ray_weight=np.full(n, 1.0 / rays_per_pix), pixel_index=np.repeat(np.arange(npix), rays_per_pix),
                stokes=np.array([True, False, False, False]), wavelength=wl)


def test_sort_sensors_groups_by_wavelength_and_builds_pixel_arrays():
    solvers = SolversDict()
    solvers.add_solver(0.66, fake_solver()); solvers.add_solver(0.86, fake_solver())
    fwd, meas = SensorsDict(), SensorsDict()
    for k, (wl, npix, rpp) in enumerate([(0.66, 5, 2), (0.86, 4, 1), (0.66, 3, 4)]):
        s = sensor(wl, npix, rpp, k)
        fwd.add_sensor('a' if k < 2 else 'b', s)
        m = dict(s, I=np.arange(npix, dtype=float) + 10 * k)
        meas.add_sensor('a' if k < 2 else 'b', m)
    assert list(fwd.get_unique_solvers()) == [0.66, 0.86]
    assert fwd.npixels == 12 and fwd.nmeasurements == 12 and fwd.get_minimum_stokes()[0.66] == 1
    rs, mp = fwd.sort_sensors(solvers, meas)
    assert mp[0.66] == [('a', 0), ('b', 0)] and mp[0.86] == [('a', 1)]
    a = rs[0.66]
    assert a['ray_x'].size == 5 * 2 + 3 * 4 and list(a['rays_per_image']) == [10, 12]
    np.testing.assert_array_equal(a['rays_per_pixel'], [2] * 5 + [4] * 3)
    np.testing.assert_array_equal(a['measurement_data'][0], [0, 1, 2, 3, 4, 20, 21, 22])
    assert a['stokes_weights'].shape == (1, 8) and a['uncertainties'].shape == (1, 1, 8) and np.all(a['uncertainties'] == 1)
    # modelled pixel values go back to the right sensors
    out = dict(a, I=np.arange(8, dtype=float))
    fwd.add_measurements_inverse(mp, [out], [0.66])
    np.testing.assert_array_equal(fwd['a']['sensor_list'][0]['I'], [0, 1, 2, 3, 4])
    np.testing.assert_array_equal(fwd['b']['sensor_list'][0]['I'], [5, 6, 7])


def test_unknown_scatterers_optical_unknowns():
    u = UnknownScatterers()
    u.add_unknowns('cloud', ['extinction', 'ssalb'])
    assert u['cloud'].variables == ['extinction', 'ssalb']
    try:
        u.add_unknowns('cloud', ['reff'])
        assert False
    except NotImplementedError:
        pass


def test_solvers_dict_derivative_entry_points():
    from at3d_b200.containers import SolversDict, UnknownScatterers
    calls = []

    class Stub:
        medium = {'cloud': dict(extinction=np.ones((2, 2, 2), np.float32), table_index=np.ones((1, 2, 2, 2), np.int32),
                                phase_weights=np.ones((1, 2, 2, 2), np.float32), legcoef=np.ones((6, 3, 1), np.float32))}

        def calculate_direct_beam_derivative(self):
            calls.append('beam')

        def calculate_microphysical_partial_derivatives(self, info):
            calls.append(sorted(info['cloud']))

    solvers = SolversDict()
    solvers[0.66], solvers[0.86] = Stub(), Stub()
    unknown = UnknownScatterers()
    unknown.add_unknowns('cloud', ['extinction', 'ssalb'])
    solvers.calculate_microphysical_partial_derivatives(unknown)
    solvers.calculate_direct_beam_derivative()
    assert calls == [['extinction', 'ssalb']] * 2 + ['beam'] * 2
    import pytest
    with pytest.raises(TypeError):
        solvers.calculate_microphysical_partial_derivatives({'cloud': ['extinction']})


def test_get_image_reshapes_to_the_image_plane():
    import pytest
    from at3d_b200 import sensor as SN
    from at3d_b200.containers import SensorsDict
    sensors = SensorsDict()
    cam = SN.perspective_projection(0.66, 20.0, 5, 3, [0.2, 0.1, 3.0], [0.2, 0.15, 0.2], [0, 1, 0], stokes=['I', 'Q'])
    cam['I'], cam['Q'] = np.arange(15.0), -np.arange(15.0)
    sensors.add_sensor('cam', cam)
    img = sensors.get_images('cam')[0]
    assert set(img) == {'x', 'y', 'mu', 'phi', 'I', 'Q'} and img['I'].shape == (5, 3)
    np.testing.assert_array_equal(img['I'][:, 0], np.arange(5.0))           # x fastest within the image
    np.testing.assert_array_equal(img['mu'], cam['cam_mu'].reshape((5, 3), order='F'))
    del cam['image_shape']
    with pytest.raises(ValueError, match='image_shape'):
        sensors.get_image('cam', 0)
