"""A reference-style retrieval script on the B200 path, in the order of the AT3D tutorials: medium -> sensors ->
solvers -> synthetic measurements (+ noise model) -> save / reload the forward model -> state generator with transforms
-> L-BFGS-B on the cloud extinction.  Every call below has the name and arguments of its at3d counterpart; the datasets
are plain mappings.

    python examples/retrieve_extinction.py [--maxiter 8] [--save /tmp/forward_model.nc]
"""
import argparse
import os
import sys
import tempfile
from collections import OrderedDict
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import at3d_b200.configuration as configuration
import at3d_b200.sensor as sensor
import at3d_b200.source as source
import at3d_b200.surface as surface
import at3d_b200.transforms as transforms
import at3d_b200.uncertainties as uncertainties
import at3d_b200.util as util
from at3d_b200.containers import SensorsDict, SolversDict, UnknownScatterers
from at3d_b200.optimize import ObjectiveFunction, Optimizer, GridStateGenerator, CallbackFn
from at3d_b200.rte import RTE


def cloud_scatterer(nx, ny, nz, dx, ztop, peak_extinction):
    """A Gaussian blob of cloud on a regular grid with a small Henyey-Greenstein phase-function table (the variables of
    an optical-property dataset: at3d/medium.py:100-112)."""
    x, y, z = np.arange(nx) * dx, np.arange(ny) * dx, np.linspace(0.0, ztop, nz)
    X, Y, Z = np.meshgrid((np.arange(nx) + 0.5) / nx, (np.arange(ny) + 0.5) / ny, np.arange(nz) / (nz - 1), indexing='ij')
    r2 = ((X - 0.5) / 0.28) ** 2 + ((Y - 0.5) / 0.28) ** 2 + ((Z - 0.5) / 0.3) ** 2
    ext = (peak_extinction * np.exp(-r2)).astype(np.float32)
    ext[r2 > 1.6] = 0.0
    gs = np.array([0.80, 0.85])
    l = np.arange(65)
    legcoef = np.zeros((6, l.size, gs.size), np.float32)
    legcoef[0] = ((2 * l + 1)[:, None] * gs[None, :] ** l[:, None])
    table_index = np.where(Z < 0.5, 1, 2).astype(np.int32)[None]
    return dict(x=x, y=y, z=z, delx=dx, dely=dx, extinction=ext, ssalb=np.full_like(ext, 0.999), table_index=table_index,
                phase_weights=np.ones((1, nx, ny, nz), np.float32), legcoef=legcoef)


def run(maxiter=8, save=None, nx=10, ny=9, nz=11, resolution=14, verbose=True):
    wavelength = 0.672
    truth = cloud_scatterer(nx, ny, nz, 0.05, 0.5, 30.0)
    config = configuration.get_config()
    config['num_mu_bins'], config['num_phi_bins'] = 8, 16
    config['split_accuracy'] = 0.0                       # fixed grid: the solves of the iterations continue each other
    config['solution_accuracy'] = 1e-5

    # ---- sensors: four perspective cameras around the cloud, 2 x 2 Gauss-Legendre rays per pixel ----
    sensors = SensorsDict()
    centre = [0.5 * nx * 0.05, 0.5 * ny * 0.05, 0.25]
    for azimuth in (0.0, 90.0, 180.0, 270.0):
        a = np.deg2rad(azimuth)
        position = [centre[0] + 1.6 * np.cos(a), centre[1] + 1.6 * np.sin(a), 2.4]
        sensors.add_sensor('camera', sensor.perspective_projection(
            wavelength, 16.0, resolution, resolution, position, centre, [0, 0, 1], stokes='I',
            sub_pixel_ray_args={'method': sensor.gaussian, 'degree': 2}))

    # ---- solver of the true medium and its synthetic measurements ----
    def make_solver(scatterer):
        return RTE(numerical_params=config, medium=OrderedDict(cloud=scatterer),
                   source=source.solar(wavelength, 0.5, 35.0), surface=surface.lambertian(0.05), num_stokes=1)
    solvers = SolversDict()
    solvers.add_solver(wavelength, make_solver(truth))
    sensors.get_measurements(solvers, maxiter=100, verbose=False)
    model = uncertainties.RadiometricUncertainty('L2', lambda radiance: 300.0 + 0.0 * radiance, 1e-5,
                                                 camera_to_camera_calibration_uncertainty=0.0, seed=1)
    sensors.add_uncertainty_model('camera', model)

    # ---- the forward model goes to disk and comes back (what a retrieval script starts from) ----
    path = save or os.path.join(tempfile.mkdtemp(), 'forward_model.nc')
    path = util.save_forward_model(path, sensors, solvers)
    for solver in solvers.values():
        solver.close()
    sensors, solvers, rte_grid = util.load_forward_model(path)
    sensors.add_uncertainty_model('camera', model)
    truth_extinction = np.asarray(solvers[wavelength].medium['cloud']['extinction'])

    # ---- the unknown: cloud extinction where the cloud mask is set, optimised in log coordinates ----
    mask = truth_extinction > 0.0
    unknown_scatterers = UnknownScatterers()
    unknown_scatterers.add_unknowns('cloud', ['extinction'])
    first_guess = OrderedDict(solvers[wavelength].medium['cloud'])
    first_guess['extinction'] = np.where(mask, 10.0, 0.0).astype(np.float32)
    inverse_solvers = SolversDict()
    generator = GridStateGenerator(
        inverse_solvers, unknown_scatterers, {wavelength: OrderedDict(cloud=first_guess)},
        {wavelength: solvers[wavelength].source}, {wavelength: solvers[wavelength].surface},
        {wavelength: solvers[wavelength].numerical_params}, {wavelength: 1}, mask=mask,
        transforms={('cloud', 'extinction'): (transforms.CoordinateTransformLog(), transforms.StateToGridMask(mask=mask))})
    initial_state = generator.get_state()
    lower, upper = generator.transform_bounds({('cloud', 'extinction'): (1e-2, 200.0)})

    forward_sensors = sensors.make_forward_sensors()
    objective = ObjectiveFunction.LevisApproxUncorrelatedL2(
        sensors, inverse_solvers, forward_sensors, unknown_scatterers, generator, generator.project_gradient_to_state,
        parallel_solve_kwargs=dict(verbose=False, maxiter=100, init_solution=True),
        gradient_kwargs=dict(cost_function='L2', exact_single_scatter=True), uncertainty_kwargs=dict(add_noise=False),
        min_bounds=lower, max_bounds=upper)

    def report(optimizer):
        retrieved = np.asarray(inverse_solvers[wavelength].medium['cloud']['extinction'])
        error = float(np.sqrt(np.mean((retrieved[mask] - truth_extinction[mask]) ** 2)))
        if verbose:
            print('iteration %3d   cost %.5e   rms extinction error %.3f' % (optimizer.iteration, optimizer.loss_history[-1], error))
        return {'cost': optimizer.loss_history[-1], 'rms_error': error}
    callback = CallbackFn(report)
    optimizer = Optimizer(objective, callback_fn=callback, options=dict(maxiter=maxiter, maxls=10, gtol=1e-16, ftol=1e-16))
    result = optimizer.minimize(initial_state)
    for solver in list(solvers.values()) + list(inverse_solvers.values()):
        solver.close()
    return dict(result=result, history=optimizer.loss_history, output=callback.output, path=path, rte_grid=rte_grid,
                nrays=sum(s['ray_mu'].size for s in sensors['camera']['sensor_list']))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--maxiter', type=int, default=8)
    ap.add_argument('--save', default=None)
    args = ap.parse_args()
    out = run(args.maxiter, args.save)
    print('forward model:', out['path'], '  rays:', out['nrays'])
    print('cost %.4e -> %.4e in %d evaluations' % (out['history'][0], out['history'][-1], len(out['history'])))
