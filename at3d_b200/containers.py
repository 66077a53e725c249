"""``SensorsDict`` / ``SolversDict`` / ``UnknownScatterers``: the host-side mirror of at3d/containers.py for the B200 hot path.

Same methods, argument meaning and bookkeeping as the reference classes (``SensorsDict`` at3d/containers.py:39-672,
``SolversDict`` :674-835, ``UnknownScatterers`` :837-985), on sensors that are plain mappings ``name -> array`` with the
variable names of at3d/sensor.py:93-107 (``ray_x, ray_y, ray_z, ray_mu, ray_phi, ray_weight, pixel_index, stokes``
[4 booleans], ``wavelength``, optionally ``use_subpixel_rays``, ``image_shape``, ``uncertainties``) -- xarray Datasets work
too: only ``sensor[name]`` (and ``.data``) are used.  Where the reference fans work out over joblib threads or MPI ranks
(``n_jobs``, ``mpi_comm``) the solvers run one after another on the GPU (every RTE owns its device state), and ray lists
are not subdivided: one RENDER / gradient launch takes all rays of a solver.  Multi-GPU: one process per GPU, solvers
dealt round-robin by key (``SolversDict.solve(rank=, world=)``), the gradient all-reduced (at3d_b200/parallel.py).
"""
import copy
import warnings
from collections import OrderedDict
import numpy as np
from . import backend as B
from .rte import RTE, _v, _scalar

STOKES_NAMES = ('I', 'Q', 'U', 'V')


def _npixels(sensor):
    if 'npixels' in sensor and np.ndim(_v(sensor, 'npixels')) == 0:
        return int(_scalar(sensor, 'npixels'))
    pix = _v(sensor, 'pixel_index')
    return int(pix.max()) + 1 if pix.size else 0


def _wavelength(sensor):
    return float(_scalar(sensor, 'wavelength'))


class SensorsDict(OrderedDict):
    """Measurement geometry by instrument; rearranges it per solver for the RTE calls and stores the observables
    (at3d/containers.py:39)."""

    def add_sensor(self, instrument, sensor):
        if instrument not in self:
            self._add_instrument(instrument)
        for k in ('ray_x', 'ray_y', 'ray_z', 'ray_mu', 'ray_phi', 'ray_weight', 'pixel_index', 'stokes', 'wavelength'):
            if k not in sensor:
                raise KeyError("sensor is missing variable '%s' (at3d/sensor.py:93-107)" % k)
        self[instrument]['sensor_list'].append(sensor)

    def _add_instrument(self, key):
        self[key] = {'sensor_list': [], 'uncertainty_model': None}

    def add_uncertainty_model(self, instrument, uncertainty_model):
        if instrument not in self:
            self._add_instrument(instrument)
        self[instrument]['uncertainty_model'] = uncertainty_model

    def calculate_uncertainties(self, instrument):
        for sensor in self[instrument]['sensor_list']:
            self[instrument]['uncertainty_model'].calculate_uncertainties(sensor)

    def add_noise(self, instrument):
        for sensor in self[instrument]['sensor_list']:
            self[instrument]['uncertainty_model'].add_noise(sensor)

    def make_forward_sensors(self, instrument_list=None):
        """Deep copy of the geometry, to hold the forward model's output during an optimization (:104-131)."""
        forward_sensors = SensorsDict()
        for key in (self if instrument_list is None else instrument_list):
            if key not in self:
                raise KeyError("Instrument '{}' is not in SensorsDict".format(key))
            inst = self[key]
            forward_sensors[key] = OrderedDict(sensor_list=[copy.deepcopy(s) for s in inst['sensor_list']],
                                               uncertainty_model=inst['uncertainty_model'])
        return forward_sensors

    def get_images(self, instrument):
        """Every image of `instrument` as 2-D arrays (:389-411)."""
        return [self.get_image(instrument, index) for index in range(len(self[instrument]['sensor_list']))]

    def get_image(self, instrument, sensor_index):
        """Pixel positions, directions and the observed Stokes components of one sensor on its image plane
        (``image_shape``, first image dimension fastest) (:414-459)."""
        sensor = self[instrument]['sensor_list'][sensor_index]
        if 'image_shape' not in sensor:
            raise ValueError("Sensor dataset does not have an 'image_shape' variable. A 2D image cannot be formed.")
        shape = tuple(int(n) for n in _v(sensor, 'image_shape'))
        image = OrderedDict((name, np.asarray(_v(sensor, 'cam_' + name)).reshape(shape, order='F'))
                            for name in ('x', 'y', 'mu', 'phi'))
        for name, observed in zip(STOKES_NAMES, np.asarray(_v(sensor, 'stokes'), bool)):
            if observed:
                image[name] = np.asarray(_v(sensor, name)).reshape(shape, order='F')
        return image

    def get_unique_solvers(self):
        return np.unique([_wavelength(s) for inst in self.values() for s in inst['sensor_list']])

    def get_minimum_stokes(self):
        """Smallest NSTOKES per wavelength that provides the required observables (:363-386)."""
        out = OrderedDict((float(k), 0) for k in self.get_unique_solvers())
        for inst in self.values():
            for s in inst['sensor_list']:
                st = np.asarray(_v(s, 'stokes'), bool)
                n = 4 if np.all(st) else int(np.where(~st)[0][0])
                n = n if n != 2 else 3
                out[_wavelength(s)] = max(out[_wavelength(s)], n)
        return out

    @property
    def nmeasurements(self):
        return int(sum(_npixels(s) * int(np.sum(np.asarray(_v(s, 'stokes'), bool)))
                       for inst in self.values() for s in inst['sensor_list']))

    @property
    def npixels(self):
        return int(sum(_npixels(s) for inst in self.values() for s in inst['sensor_list']))

    # ---- grouping by solver (at3d/containers.py:233-352) ----
    def sort_sensors(self, solvers, measurements=None):
        """Groups the sensors by RTE solver (wavelength).  Returns (rte_sensors, sensor_mappings): per solver key the
        concatenated ray variables, ``stokes`` [nimage, 4], ``rays_per_image``, ``rays_per_pixel`` and -- with
        `measurements` -- ``stokes_weights``, ``measurement_data`` [nstokes, npixels] and ``uncertainties``
        [nstokes, nstokes, npixels]; and the (instrument, index) of every image."""
        if not isinstance(solvers, SolversDict):
            raise TypeError("`solvers` should be of type '{}' not '{}'".format(SolversDict, type(solvers)))
        if measurements is not None and not isinstance(measurements, SensorsDict):
            raise TypeError("`measurements` should be of type '{}' not '{}'".format(SensorsDict, type(measurements)))
        rte_sensors, sensor_mappings = OrderedDict(), OrderedDict()
        var_list = ['ray_x', 'ray_y', 'ray_z', 'ray_mu', 'ray_phi', 'ray_weight', 'pixel_index']
        for key, solver in solvers.items():
            sensor_list, mapping_list = [], []
            for instrument, data in self.items():
                for i, sensor in enumerate(data['sensor_list']):
                    if key == _wavelength(sensor):
                        sensor_list.append(sensor)
                        mapping_list.append((instrument, i))
            output = {}
            if not sensor_list:
                warnings.warn("No sensors found matching solver with key '{}'".format(key))
            else:
                for var in var_list:
                    output[var] = np.concatenate([_v(s, var) for s in sensor_list])
                output['stokes'] = np.stack([np.asarray(_v(s, 'stokes'), bool) for s in sensor_list])
                output['rays_per_image'] = np.array([_v(s, 'ray_x').size for s in sensor_list])
                output['rays_per_pixel'] = np.concatenate([np.unique(_v(s, 'pixel_index'), return_counts=True)[1]
                                                           for s in sensor_list])
                if measurements is not None:
                    meas_list = [s for data in measurements.values() for s in data['sensor_list'] if key == _wavelength(s)]
                    weights, datas, uncs = [], [], []
                    nst = solver._nstokes
                    for s in meas_list:
                        npx = _npixels(s)
                        w, d = np.zeros((nst, npx)), np.zeros((nst, npx))
                        for i, name in enumerate(STOKES_NAMES[:nst]):
                            if name in s:
                                w[i], d[i] = 1.0, _v(s, name)
                        weights.append(w); datas.append(d)
                        if 'uncertainties' in s:
                            uncs.append(np.asarray(_v(s, 'uncertainties'), np.float64)[:nst, :nst])
                        else:                                       # NullUncertainty: unweighted least squares
                            uncs.append(np.repeat(np.eye(nst)[:, :, None], npx, axis=2))
                    output['uncertainties'] = np.concatenate(uncs, axis=-1)
                    output['stokes_weights'] = np.concatenate(weights, axis=-1)
                    output['measurement_data'] = np.concatenate(datas, axis=-1)
            rte_sensors[key] = output
            sensor_mappings[key] = mapping_list
        return rte_sensors, sensor_mappings

    # ---- forward model (at3d/containers.py:133-231) ----
    def get_measurements(self, solvers, n_jobs=1, mpi_comm=None, maxiter=100, verbose=True, init_solution=True,
                         setup_grid=True, destructive=False, overwrite_solver=False):
        """Solves the RTE where needed, renders every sensor's rays and stores the pixel observables in `self`."""
        if not isinstance(solvers, SolversDict):
            raise TypeError("`solvers` should be of type '{}' not '{}'".format(SolversDict, type(solvers)))
        if not isinstance(destructive, bool):
            raise TypeError('`destructive` should be a boolean.')
        if mpi_comm is not None:
            raise NotImplementedError('mpi_comm: one process per GPU with torch.distributed replaces the MPI fan-out')
        rte_sensors, sensor_mappings = self.sort_sensors(solvers)
        solvers.solve(maxiter=maxiter, verbose=verbose, init_solution=init_solution, setup_grid=setup_grid,
                      overwrite_solver=overwrite_solver)
        out = [solvers[key].integrate_to_sensor(rte_sensors[key]) for key in solvers if rte_sensors[key]]
        keys = [key for key in solvers if rte_sensors[key]]
        self.add_measurements_forward(sensor_mappings, out, keys)

    def add_measurements_forward(self, sensor_mappings, measurements, measurement_keys):
        """Splits the rendered rays back into images and stores the pixel-averaged observables (:556-599)."""
        for key in sensor_mappings:
            parts = [m for m, k in zip(measurements, measurement_keys) if k == key]
            if not parts:
                continue
            names = [n for n in STOKES_NAMES if n in parts[0]]
            merged = {n: np.concatenate([np.asarray(_v(p, n)) for p in parts]) for n in names}
            rays_per_image = np.asarray(_v(parts[0], 'rays_per_image'))
            stokes = np.asarray(_v(parts[0], 'stokes'))
            count = 0
            for i, nr in enumerate(rays_per_image):
                rendered = {n: merged[n][count:count + nr] for n in names}
                rendered['stokes'] = stokes[i]
                self._calculate_observables(sensor_mappings[key][i], rendered)
                count += int(nr)

    def _calculate_observables(self, mapping, rendered_rays):
        """Pixel-averaged Stokes components from the ray values (:601-645; util.f90 AVERAGE_SUBPIXEL_RAYS on the GPU)."""
        sensor = self[mapping[0]]['sensor_list'][mapping[1]]
        want = np.asarray(rendered_rays['stokes'], bool)
        names = [n for k, n in enumerate(STOKES_NAMES) if want[k] and n in rendered_rays]
        w = np.asarray(_v(sensor, 'ray_weight'), np.float64)
        use_sub = bool(_scalar(sensor, 'use_subpixel_rays', True))
        if not use_sub:
            for n in names:
                sensor[n] = (w * rendered_rays[n]).astype(np.float32)
            return
        ws = np.asfortranarray(np.stack([w * rendered_rays[n] for n in names]).astype(np.float32))
        obs = B.average_subpixel_rays(ws, np.asarray(_v(sensor, 'pixel_index'), np.int32), _npixels(sensor))
        for i, n in enumerate(names):
            sensor[n] = obs[i]

    def add_measurements_inverse(self, sensor_mappings, measurements, measurement_keys):
        """Stores the pixel observables the gradient evaluation modelled (:456-554): `measurements` hold per-pixel I (Q, U)
        for all images of a solver, concatenated."""
        for key in sensor_mappings:
            parts = [m for m, k in zip(measurements, measurement_keys) if k == key]
            if not parts:
                continue
            names = [n for n in STOKES_NAMES if n in parts[0]]
            merged = {n: np.concatenate([np.asarray(_v(p, n)) for p in parts]) for n in names}
            rpp = np.concatenate([np.asarray(_v(p, 'rays_per_pixel')) for p in parts])
            pixel_inds = np.concatenate([[0], np.cumsum(rpp)]).astype(int)
            rays_per_image = np.asarray(_v(parts[0], 'rays_per_image'))
            stokes = np.asarray(_v(parts[0], 'stokes'))
            ray_ends = np.cumsum(rays_per_image)
            pixel_ends = [int(np.where(pixel_inds == e)[0][0]) for e in ray_ends]
            start = 0
            for i, end in enumerate(pixel_ends):
                fs = self[sensor_mappings[key][i][0]]['sensor_list'][sensor_mappings[key][i][1]]
                for k, n in enumerate(STOKES_NAMES):
                    if n in merged and np.asarray(_v(fs, 'stokes'), bool)[k] and stokes[i][k]:
                        fs[n] = merged[n][start:end]
                start = end


class SolversDict(OrderedDict):
    """The RTE solvers by key (wavelength) (at3d/containers.py:674)."""

    def add_solver(self, key, solver):
        if not isinstance(solver, RTE):
            raise TypeError("solver should be of type '{}'".format(RTE))
        self[key] = solver

    def to_solve(self, overwrite_solver=False):
        """Keys and solvers that still need a solution (:746-766)."""
        keys, solvers = [], []
        for key, solver in self.items():
            if overwrite_solver or not solver.check_solved(verbose=False):
                keys.append(key); solvers.append(solver)
        return keys, solvers

    def solve(self, n_jobs=1, mpi_comm=None, overwrite_solver=False, maxiter=100, verbose=True, init_solution=True,
              setup_grid=True, rank=0, world=1):
        """Solves every unsolved RTE (:690-744).  `n_jobs` is accepted and ignored (the GPU runs one solve at a time at
        full width); with `world` > 1 processes, this rank solves the keys ``rank, rank + world, ...`` (one wavelength
        per GPU, BASELINE.json configs[4])."""
        if mpi_comm is not None:
            raise NotImplementedError('mpi_comm: use rank= / world= (one process per GPU)')
        keys, to_solve = self.to_solve(overwrite_solver)
        for i, solver in enumerate(to_solve):
            if i % world == rank:
                solver.solve(maxiter=maxiter, init_solution=init_solution, verbose=verbose, setup_grid=setup_grid)

    parallel_solve = solve

    def calculate_direct_beam_derivative(self):
        """at3d/containers.py:769-776."""
        for solver in self.values():
            solver.calculate_direct_beam_derivative()

    def calculate_microphysical_partial_derivatives(self, unknown_scatterers):
        """Derivative tables of every solver for the unknowns (at3d/containers.py:778-800)."""
        if not isinstance(unknown_scatterers, UnknownScatterers):
            raise TypeError("`unknown_scatterers` should be of type '{}' not '{}'".format(UnknownScatterers, type(unknown_scatterers)))
        for solver in self.values():
            solver.calculate_microphysical_partial_derivatives(unknown_scatterers.derivative_information(solver))

    @property
    def npixels(self):
        return None


class UnknownScatterers(OrderedDict):
    """Which properties of which scatterers are unknown (at3d/containers.py:837).  Optical unknowns on the property grid:
    ``add_unknowns(name, ['extinction'])`` (and / or ``'ssalb'``); every entry carries ``variables`` and the derivative
    datasets `RTE.calculate_microphysical_partial_derivatives` takes."""

    class _Entry:
        def __init__(self, variables):
            self.variables = list(variables)

    def add_unknowns(self, scatterer_name, variable_names):
        variable_names = [variable_names] if isinstance(variable_names, str) else list(variable_names)
        for v in variable_names:
            if v not in ('extinction', 'ssalb'):
                raise NotImplementedError("unknown '%s': optical unknowns 'extinction' and 'ssalb' are built in; microphysical "
                                          "ones go through RTE.calculate_microphysical_partial_derivatives with the "
                                          "derivative tables of at3d.medium" % v)
        self[scatterer_name] = UnknownScatterers._Entry(variable_names)

    def derivative_information(self, solver):
        """The mapping `RTE.calculate_microphysical_partial_derivatives` takes, for the optical unknowns
        (at3d/medium.py OpticalDerivativeGenerator: d ext / d ext = 1, d ssalb / d ssalb = 1, phase function unchanged)."""
        info = OrderedDict()
        for name, entry in self.items():
            sc = solver.medium[name]
            shape = np.asarray(_v(sc, 'extinction')).shape
            d = OrderedDict()
            for v in entry.variables:
                d[v] = dict(extinction=np.ones(shape, np.float32) if v == 'extinction' else np.zeros(shape, np.float32),
                            ssalb=np.ones(shape, np.float32) if v == 'ssalb' else np.zeros(shape, np.float32),
                            table_index=np.asarray(_v(sc, 'table_index')),
                            phase_weights=np.zeros_like(np.asarray(_v(sc, 'phase_weights'), np.float32)),
                            legcoef=np.asarray(_v(sc, 'legcoef')), derivative_method='table')
            info[name] = d
        return info
