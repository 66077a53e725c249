"""source / surface / configuration factories (at3d/source.py, at3d/surface.py, at3d/configuration.py): variables,
defaults and argument checks -- host only."""
import json
import numpy as np
import pytest
from at3d_b200 import configuration, source, surface


def test_sources():
    s = source.solar(0.672, 0.6, 45.0, solarflux=2.0, skyrad=0.1)
    assert s['srctype'] == 'S' and s['units'] == 'R' and s['solarmu'] == -0.6 and s['solarflux'] == 2.0
    assert s['solaraz'] == np.deg2rad(45.0) and np.asarray(s['skyrad']).shape == (1, 1, 1)
    assert source.solar(0.672, -0.6, 0.0)['solarmu'] == -0.6              # the sign is forced
    for bad in (0.0, 1.5):
        with pytest.raises(ValueError, match='solarmu'):
            source.solar(0.672, bad, 0.0)
    t = source.thermal(10.8, skyrad=2.7, units='brightness_temperature')
    assert (t['srctype'], t['units'], t['solarflux'], t['solarmu'], t['skyrad']) == ('T', 'T', 0.0, -0.5, 2.7)
    assert source.thermal(10.8)['units'] == 'R'
    with pytest.raises(ValueError, match='units'):
        source.thermal(10.8, units='kelvin')
    b = source.combined(3.9, 0.3, 10.0)
    assert b['srctype'] == 'B' and b['solarmu'] == -0.3
    with pytest.raises(NotImplementedError):
        source.solar(0.672, 0.5, 0.0, volume_source=object())


def test_configuration(tmp_path):
    cfg = configuration.get_config()
    assert (cfg['num_mu_bins'], cfg['num_phi_bins'], cfg['split_accuracy'], cfg['tautol'], cfg['transcut']) == (16, 32, 0.03, 0.1, 1e-5)
    assert cfg['deltam'] is True and cfg['x_boundary_condition'] == 'open' and cfg['adapt_grid_factor'] == 5 and len(cfg) == 20
    name = str(tmp_path / 'config.json')
    configuration.make_config(name, num_mu_bins=8, x_boundary_condition='periodic')
    assert json.load(open(name))['num_mu_bins'] == {'default_value': 8}
    back = configuration.get_config(name)
    assert back['num_mu_bins'] == 8 and back['x_boundary_condition'] == 'periodic' and back['num_phi_bins'] == 32
    with pytest.raises(TypeError, match='unknown numerical parameters'):
        configuration.make_config_data(num_mu=8)


def test_surfaces():
    fl = surface.lambertian(0.2, ground_temperature=280.0)
    assert (fl['sfctype'], fl['gndalbedo'], fl['gndtemp'], fl['nsfcpar']) == ('FL', 0.2, 280.0, 1) and fl['sfcparms'].size == 0
    for bad in (-0.1, 1.2):
        with pytest.raises(ValueError, match='albedo'):
            surface.lambertian(bad)
    amap = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]])
    with pytest.raises(ValueError, match='delx'):
        surface.lambertian(amap)
    vl = surface.lambertian(amap, delx=0.1, dely=0.2)
    sp = vl['sfcparms'].reshape((2, 3, 4), order='F')
    assert (vl['sfctype'], vl['nxsfc'], vl['nysfc'], vl['nsfcpar']) == ('VL', 2, 3, 2)
    np.testing.assert_array_equal(sp[1, :2, :3], amap.astype(np.float32))
    np.testing.assert_array_equal(sp[:, 2, :], sp[:, 0, :]); np.testing.assert_array_equal(sp[:, :, 3], sp[:, :, 0])
    assert vl['gndalbedo'] == pytest.approx(amap.mean(), rel=1e-6) and vl['gndtemp'] == pytest.approx(298.15, rel=1e-6)
    w = surface.wave_fresnel(1.33, 0.0, 7.0)
    assert (w['sfctype'], w['nsfcpar'], w['delxsfc'], w['gndalbedo']) == ('VW', 4, 0.02, 0.0)
    np.testing.assert_allclose(w['sfcparms'].reshape((4, 2, 2), order='F')[:, 0, 0], [298.15, 1.33, 0.0, 7.0], rtol=1e-7)
    d = surface.diner(0.2, 0.8, 0.3, 0.5, -1.0)
    assert (d['sfctype'], d['nsfcpar']) == ('VD', 6) and d['gndalbedo'] == np.float32(0.2)
    np.testing.assert_allclose(d['sfcparms'][:6], [298.15, 0.2, 0.8, 0.3, 0.5, -1.0], rtol=1e-7)     # A, K, B, ZETA, SIGMA
    r = surface.RPV_unpolarized(0.1, 0.7, -0.24)
    assert (r['sfctype'], r['nsfcpar']) == ('VR', 4) and r['gndalbedo'] == np.float32(0.1)
    with pytest.raises(ValueError, match='same shape'):
        surface.ocean_unpolarized(np.ones((2, 2)), np.ones((2, 3)), delx=0.1, dely=0.1)
    with pytest.raises(ValueError, match='ground temperature'):
        surface.RPV_unpolarized(0.1, 0.7, -0.24, ground_temperature=np.ones((2, 2)))
    with pytest.raises(ValueError, match='delx'):
        surface.ocean_unpolarized(np.ones((2, 2)), np.ones((2, 2)))
    with pytest.raises(ValueError, match='Illegal surface albedo'):
        surface.prep_surface('VL', np.stack([np.full((2, 2), 290.0), np.full((2, 2), 1.5)]))


def test_package_namespace_resolves_the_reference_module_names():
    import at3d_b200 as at3d
    assert at3d.solver.RTE is at3d.rte.RTE and at3d.callback.CallbackFn is at3d.optimize.CallbackFn
    assert callable(at3d.sensor.perspective_projection) and callable(at3d.util.save_forward_model)
    assert at3d.containers.SensorsDict.__name__ == 'SensorsDict' and callable(at3d.grid.make_grid)
    with pytest.raises(AttributeError):
        at3d.visualization
