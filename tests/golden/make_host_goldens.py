"""Golden outputs of the reference's own host-side Python for the containers either side of the path, made by running
the UNMODIFIED reference modules from /root/reference in this container (run once here; /root/reference does not exist
on the GPU box, the .npz travels):

    at3d/sensor.py      orthographic_projection, perspective_projection, domaintop_projection (+ sub-pixel rays)
    at3d/transforms.py  coordinate transforms and state-to-grid maps
    at3d/parallel.py    subdivide_raytrace_jobs (pixel-aligned ray ranges per worker)
    at3d/uncertainties.py  inverse error covariances and noise draws of the radiometric models

xarray is absent from the image, so the sensor module is loaded with a minimal stand-in for the few xarray calls it makes
(Dataset(data_vars, coords), ds[name] = (dims, data) | DataArray, ds.name.data, ds.attrs) -- the arithmetic is the
reference's own.  Usage: python tests/golden/make_host_goldens.py  ->  tests/golden/host_goldens.npz
"""
import importlib.util
import os
import sys
import types
import numpy as np

REF = '/root/reference/at3d'
HERE = os.path.dirname(os.path.abspath(__file__))


class _Var:
    def __init__(self, data):
        self.data = np.asarray(data)

    def __setitem__(self, key, value):
        self.data[key] = value.data if isinstance(value, _Var) else value

    def __getitem__(self, key):
        return _Var(self.data[key])

    def __str__(self):
        return str(self.data)

    def __iadd__(self, other):
        self.data = self.data + other
        return self

    @property
    def size(self):
        return self.data.size


class _Dataset:
    def __init__(self, data_vars=None, coords=None):
        object.__setattr__(self, '_vars', {})
        object.__setattr__(self, 'attrs', {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    @property
    def data_vars(self):
        return self._vars

    def __setitem__(self, key, value):
        if isinstance(value, _Var):
            self._vars[key] = value
        elif isinstance(value, tuple):
            self._vars[key] = _Var(np.array(value[1]))
        else:
            self._vars[key] = _Var(np.array(value))

    def __getitem__(self, key):
        return self._vars[key]

    def __getattr__(self, key):
        try:
            return self._vars[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        object.__setattr__(self, key, value)


def _load(name, stubs):
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location('_ref_' + name, os.path.join(REF, name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    xr = types.ModuleType('xarray')
    xr.Dataset = _Dataset
    xr.DataArray = lambda data, coords=None, dims=None: _Var(data)
    at3d = types.ModuleType('at3d')
    at3d.checks = types.ModuleType('at3d.checks')
    at3d.checks.check_grid = lambda grid: None
    sensor = _load('sensor', {'xarray': xr, 'at3d': at3d, 'at3d.checks': at3d.checks})
    at3d.sensor = sensor                                    # domaintop_projection calls at3d.sensor.orthographic_projection
    transforms = _load('transforms', {})
    out = {}

    def keep(prefix, ds):
        for k, v in ds.data_vars.items():
            out[prefix + '/' + k] = v.data

    box = _Dataset({'x': np.linspace(0.0, 0.64, 33), 'y': np.linspace(0.0, 0.72, 37), 'z': np.linspace(0.0, 1.04, 27)})
    sys.modules['at3d'] = at3d
    try:
        keep('ortho', sensor.orthographic_projection(0.672, box, 0.08, 0.09, 37.0, 25.0, stokes=['I', 'Q']))
        keep('ortho_gauss', sensor.orthographic_projection(
            0.672, box, 0.16, 0.12, 200.0, 40.0, altitude=1.5, stokes='I',
            sub_pixel_ray_args={'method': sensor.gaussian, 'degree': (2, 3)}))
        keep('ortho_uniform', sensor.orthographic_projection(
            1.65, box, 0.2, 0.2, 0.0, 0.0, sub_pixel_ray_args={'method': sensor.uniform, 'nrays': 2}))
        keep('persp', sensor.perspective_projection(0.672, 17.0, 7, 5, [0.3, -0.2, 5.0], [0.32, 0.36, 0.5], [0, 1, 0],
                                                    stokes=['I', 'Q', 'U']))
        keep('persp_stoch', sensor.perspective_projection(
            0.672, 30.0, 4, 6, [2.0, 1.5, 3.0], [0.3, 0.3, 0.4], [0, 0, 1], stokes='I',
            sub_pixel_ray_args={'method': sensor.stochastic, 'nrays': (3, 2), 'seed': 11}))
        keep('domaintop', sensor.domaintop_projection(0.672, box, 0.1, 0.15, 75.0, 35.0, x_offset=0.01, y_offset=-0.02))
    finally:
        sys.modules.pop('at3d', None)

    rng = np.random.default_rng(3)
    phys = rng.uniform(0.1, 30.0, 25)
    grad = rng.normal(size=25)
    out['tr/phys'], out['tr/grad'] = phys, grad
    for name, tr in (('null', transforms.CoordinateTransform()), ('log', transforms.CoordinateTransformLog()),
                     ('scaling', transforms.CoordinateTransformScaling(2.0, 0.25)),
                     ('exp', transforms.CoordinateTransformExp(10.0)), ('hyperbol', transforms.CoordinateTransformHyperBol(0.05))):
        a = tr.inverse_transform(phys)
        out['tr/%s/abstract' % name] = a
        out['tr/%s/physical' % name] = tr(a)
        out['tr/%s/gradient' % name] = tr.gradient_transform(a, grad)
    mask = np.zeros((4, 3, 6), bool)
    mask[1:3, 0:2, 1:5] = True
    mask[3, 2, 2] = True
    mask[:, :, 0] = True
    data = rng.uniform(1.0, 2.0, mask.shape)
    out['s2g/mask'], out['s2g/data'] = mask, data
    for name, cls in (('mask', transforms.StateToGridMask), ('2d', transforms.StateToGrid2D),
                      ('uniform', transforms.StateToGridUniform), ('profile', transforms.StateToGridProfile)):
        s2g = cls(mask=mask)
        state = s2g.inverse_transform(data)
        out['s2g/%s/state' % name] = state
        out['s2g/%s/gradient' % name] = s2g.gradient_transform(data)
        out['s2g/%s/bounds' % name] = s2g.inverse_bounds_transform(np.full(mask.shape, 3.0))
        if name != 'profile':               # the reference's StateToGridProfile.__call__ raises for nz > ny (see transforms.py)
            out['s2g/%s/gridded' % name] = s2g(state)
    # ---- subdivide_raytrace_jobs: merged sensors with 1..4 rays per pixel ----
    parallel = _load('parallel', {})

    class _Merged:
        def __init__(self, rays_per_pixel):
            self.rays_per_pixel = _Var(rays_per_pixel)
            pixel_index = np.repeat(np.arange(rays_per_pixel.size), rays_per_pixel)
            self.sizes = {'nrays': pixel_index.size}
            self.nrays = _Var(np.arange(pixel_index.size))
            self.npixels = _Var(np.arange(rays_per_pixel.size))
            self.pixel_index = types.SimpleNamespace(diff=lambda dim: _Var(np.diff(pixel_index)))

    for case, (sizes, n_jobs, job_factor) in enumerate((((37,), 4, 1), ((50, 13, 29), 8, 1), ((20, 21), 3, 2), ((9,), 4, 1))):
        sensors = {}
        for i, npix in enumerate(sizes):
            rpp = rng.integers(1, 5, npix)
            sensors[0.4 + 0.1 * i] = _Merged(rpp)
            out['jobs/%d/rpp%d' % (case, i)] = rpp
        keys, rays, pixels = parallel.subdivide_raytrace_jobs(sensors, n_jobs, job_factor)
        out['jobs/%d/args' % case] = np.array([n_jobs, job_factor])
        out['jobs/%d/keys' % case] = np.array(keys)
        out['jobs/%d/rays' % case] = np.array(rays, np.int64)
        out['jobs/%d/pixels' % case] = np.array(pixels, np.int64)
    # ---- uncertainty models ----
    at3d.checks.check_sensor = lambda sensor: None
    unc = _load('uncertainties', {'at3d': at3d, 'at3d.checks': at3d.checks})
    radiance = rng.uniform(0.0005, 0.45, 60)
    out['unc/I'] = radiance

    def pixel_sensor():
        return _Dataset({'I': radiance.copy(), 'npixels': np.arange(radiance.size), 'stokes': np.array([True, False, False, False]),
                         'stokes_index': np.array(['I', 'Q', 'U', 'V'])})

    models = {'null': lambda: unc.NullUncertainty('L2', 2.5),
              'radiometric': lambda: unc.RadiometricUncertainty('L2', lambda r: 200.0 * np.sqrt(r / 0.1), 1e-4, 0.03, 0.01, seed=5),
              'radiometric_ll': lambda: unc.RadiometricUncertainty('LL', lambda r: 150.0 + 0.0 * r, 2e-4, 0.02, 0.0, seed=9),
              'tandem': lambda: unc.TandemStereoCamera('L2')}
    for name, make in models.items():
        np.random.seed(123)
        model = make()
        s = pixel_sensor()
        model.calculate_uncertainties(s)
        out['unc/%s/uncertainties' % name] = s['uncertainties'].data
        if name != 'null':
            np.random.seed(77)
            model.add_noise(s)
            out['unc/%s/noisy' % name] = s['I'].data
    with open(os.path.join(HERE, 'host_goldens.npz'), 'wb') as fh:
        np.savez_compressed(fh, **out)
    print(len(out), 'arrays ->', os.path.join(HERE, 'host_goldens.npz'))


if __name__ == '__main__':
    main()
