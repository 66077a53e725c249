"""``LevisApproxGradientUncorrelated``: the host-side mirror of at3d/gradient.py for the B200 hot path.

Same constructor arguments, ``__call__`` return values and ``make_gradient_dataset`` layout as the reference
(at3d/gradient.py:21-503): solve every RTE that needs it, build the derivative tables of the unknowns, group the forward
sensors with the measurements per solver, evaluate LEVISAPPROX_GRADIENT per solver on the GPU (at3d_levisapprox_gradient,
all rays of a solver in one call where the reference cuts them into joblib jobs), store the modelled pixel observables in
`forward_sensors`, and return ``(loss, gradient_dataset, jacobian_dataset)`` with loss and gradient summed over solvers
and divided by the number of measurements (:430-438).  Multi-GPU: with `parallel_solve_kwargs` ``rank`` / ``world`` this
process evaluates the solvers ``rank, rank + world, ...`` and the (loss, gradient) sums are all-reduced.
"""
import warnings
import numpy as np
from . import containers
from .rte import _v


class LevisApproxGradient:
    def __init__(self, measurements, solvers, forward_sensors, unknown_scatterers, parallel_solve_kwargs, gradient_kwargs,
                 uncertainty_kwargs):
        if not isinstance(measurements, containers.SensorsDict):
            raise TypeError("`measurements` should be of type '{}' not '{}'".format(containers.SensorsDict, type(measurements)))
        if not isinstance(solvers, containers.SolversDict):
            raise TypeError("`solvers` should be of type '{}' not '{}'".format(containers.SolversDict, type(solvers)))
        if not isinstance(forward_sensors, containers.SensorsDict):
            raise TypeError("`forward_sensors` should be of type '{}' not '{}'".format(containers.SensorsDict, type(forward_sensors)))
        if not isinstance(unknown_scatterers, containers.UnknownScatterers):
            raise TypeError("`unknown_scatterers` should be of type '{}' not '{}'".format(containers.UnknownScatterers, type(unknown_scatterers)))
        if 'add_noise' not in uncertainty_kwargs:
            uncertainty_kwargs['add_noise'] = False
            warnings.warn("'add_noise' flag was unspecified. It has been set to False.")
        elif not isinstance(uncertainty_kwargs['add_noise'], bool):
            raise TypeError("uncertainty_kwargs['add_noise'] should be of boolean type.")
        if 'cost_function' not in gradient_kwargs:
            raise ValueError("'cost_function' must be specified in `gradient_kwargs`. Supported values are 'L2' or 'LL'.")
        if not isinstance(gradient_kwargs['cost_function'], str):
            raise TypeError("gradient_kwargs['cost_function'] should be of string type.")
        if gradient_kwargs.get('indices_for_jacobian') is not None:
            raise NotImplementedError('indices_for_jacobian: use DeviceState.gradient_jacobian (at3d_levisapprox_gradient_jacobian)')
        self.measurements, self.solvers, self.forward_sensors = measurements, solvers, forward_sensors
        self.unknown_scatterers = unknown_scatterers
        self.parallel_solve_kwargs, self.gradient_kwargs, self.uncertainty_kwargs = parallel_solve_kwargs, gradient_kwargs, uncertainty_kwargs
        self._rte_sensors = None
        self._sensor_mapping = None
        for name, instrument in self.measurements.items():
            model = instrument['uncertainty_model']
            if model is None:
                continue                      # NullUncertainty: identity inverse covariance (unweighted least squares)
            if getattr(model, 'cost_function', gradient_kwargs['cost_function']) != gradient_kwargs['cost_function']:
                raise ValueError("Uncertainty model's assumed cost_function '{}' is inconsistent with the one being used '{}'".format(
                    model.cost_function, gradient_kwargs['cost_function']))
            self.measurements.calculate_uncertainties(name)
            if uncertainty_kwargs['add_noise']:
                self.measurements.add_noise(name)

    def _prep_gradient(self):
        kw = dict(self.parallel_solve_kwargs)
        rank, world = int(kw.pop('rank', 0)), int(kw.pop('world', 1))
        kw.pop('n_jobs', None); kw.pop('mpi_comm', None)
        self.solvers.solve(rank=rank, world=world, **kw)
        self.solvers.calculate_microphysical_partial_derivatives(self.unknown_scatterers)      # at3d/gradient.py:140-141
        self.solvers.calculate_direct_beam_derivative()
        rte_sensors, sensor_mapping = self.forward_sensors.sort_sensors(self.solvers, self.measurements)
        self._rte_sensors, self._sensor_mapping = rte_sensors, sensor_mapping
        losses, gradients, outs, keys = [], [], [], []
        for i, (key, solver) in enumerate(self.solvers.items()):
            if i % world != rank or not rte_sensors[key]:
                continue
            loss, gradient, rendered = self.levis_approximation_grad(
                solver, rte_sensors[key], cost_function=self.gradient_kwargs['cost_function'],
                exact_single_scatter=self.gradient_kwargs.get('exact_single_scatter', True))
            losses.append(loss); gradients.append(gradient); outs.append(rendered); keys.append(key)
        self.forward_sensors.add_measurements_inverse(sensor_mapping, outs, keys)
        loss = np.array(losses)
        gradient = np.stack(gradients, axis=-1) if gradients else np.zeros((0, 0, 0))
        if world > 1:
            from .parallel import allreduce_gradient
            g, c = gradient.sum(axis=-1), np.array([loss.sum()])
            g, c = allreduce_gradient(np.ascontiguousarray(g), c)
            gradient, loss = g[..., None], c
        return loss, gradient, None

    def levis_approximation_grad(self, rte_solver, sensor, cost_function='L2', indices_for_jacobian=None,
                                 exact_single_scatter=True):
        """(loss, gradient [nbpts, numder], integrated_rays) for one solver (at3d/gradient.py:168-398)."""
        loss, grad, images = rte_solver.levis_approx_gradient(sensor, exact_single_scatter=exact_single_scatter,
                                                              cost_function=cost_function)
        rendered = dict(sensor)
        for k, n in enumerate(('I', 'Q', 'U')[:rte_solver._nstokes]):
            rendered[n] = images[k]
        return loss, grad.reshape(-1, grad.shape[-1]), rendered

    def __call__(self):
        return self._prep_gradient(), None, None


class LevisApproxGradientUncorrelated(LevisApproxGradient):
    """Different wavelengths are uncorrelated: the default (at3d/gradient.py:413-446)."""

    def __call__(self):
        loss, gradient, other = self._prep_gradient()
        loss = float(np.sum(loss)) / self.forward_sensors.nmeasurements
        gradient = np.sum(gradient, axis=-1) / self.forward_sensors.nmeasurements
        return loss, make_gradient_dataset(gradient, self.unknown_scatterers, self.solvers), None


def make_gradient_dataset(gradient, unknown_scatterers, solvers):
    """The gradient on the property grid with its coordinates (at3d/gradient.py:448-503): a mapping with ``gradient``
    [x, y, z, derivative_index], ``x``, ``y``, ``z`` and ``derivative_index`` = [(scatterer_name, variable_name), ...]."""
    solver = list(solvers.values())[0]
    derivative_index = [(name, v) for name, e in unknown_scatterers.items() for v in e.variables]
    names = np.array(list(solver.medium.keys()))[np.asarray(solver._unknown_scatterer_indices) - 1]
    assert list(names) == [n for n, _ in derivative_index], 'Two different ways of listing unknown scatterer names do not match.'
    grid = next(iter(solver.medium.values()))
    x, y, z = _v(grid, 'x'), _v(grid, 'y'), _v(grid, 'z')
    return dict(gradient=np.asarray(gradient).reshape(x.size, y.size, z.size, -1), x=x, y=y, z=z,
                derivative_index=derivative_index)
