"""``ObjectiveFunction`` / ``Optimizer`` / ``GridStateGenerator``: the host-side mirror of at3d/optimize.py and of the state
<-> solver glue of at3d/medium.py (StateGenerator, :1649-1831) for the B200 hot path.

``ObjectiveFunction.LevisApproxUncorrelatedL2`` and ``Optimizer.minimize`` take the arguments of the reference
(at3d/optimize.py:20-144, 146-230) and drive ``scipy.optimize.minimize(jac=True)`` with the GPU cost + gradient.
``GridStateGenerator`` is ``StateGenerator`` for optical unknowns: the state vector holds the unknown optical property
(extinction and / or ssalb) of the named scatterers, each through its own coordinate transform and state <-> grid map
(at3d_b200/transforms.py; default: no transform, the masked grid points); calling it with a state rebuilds the solvers
(as the reference does every evaluation, at3d/medium.py:1813-1831), ``project_gradient_to_state`` takes the gridded
gradient back to the state (:1860-1890) and ``transform_bounds`` the physical bounds (:1892-1925).
"""
import time
from collections import OrderedDict
import numpy as np
from . import containers
from . import gradient as gradient_mod
from . import transforms as TR
from .rte import RTE, _v


class ObjectiveFunction:
    def __init__(self, measurements, loss_fn, min_bounds=None, max_bounds=None):
        self.measurements = measurements
        self.loss_fn = loss_fn
        self._bounds = list(zip(np.atleast_1d(min_bounds), np.atleast_1d(max_bounds)))
        self._loss = None
        self._total_obj_fn_time = 0.0
        self._ncalls = 0

    def __call__(self, state):
        t = time.perf_counter()
        loss, gradient = self.loss_fn(state, self.measurements)
        self._total_obj_fn_time += time.perf_counter() - t
        self._loss = loss
        self._ncalls += 1
        return loss, gradient

    @classmethod
    def LevisApproxUncorrelatedL2(cls, measurements, solvers, forward_sensors, unknown_scatterers, set_state_fn,
                                  project_gradient_to_state, parallel_solve_kwargs=None, gradient_kwargs=None,
                                  uncertainty_kwargs=None, min_bounds=None, max_bounds=None):
        parallel_solve_kwargs = dict(parallel_solve_kwargs or {'n_jobs': 1, 'mpi_comm': None, 'verbose': True, 'maxiter': 100,
                                                               'init_solution': True})
        gradient_kwargs = dict(gradient_kwargs or {'cost_function': 'L2', 'exact_single_scatter': True})
        uncertainty_kwargs = dict(uncertainty_kwargs or {'add_noise': False})
        gradient_fun = gradient_mod.LevisApproxGradientUncorrelated(
            measurements, solvers, forward_sensors, unknown_scatterers, parallel_solve_kwargs, gradient_kwargs, uncertainty_kwargs)

        def loss_function(state, measurements):
            set_state_fn(state)
            loss, gradient, _ = gradient_fun()
            return loss, project_gradient_to_state(state, gradient)
        return cls(measurements, loss_function, min_bounds=min_bounds, max_bounds=max_bounds)

    @property
    def loss(self):
        return self._loss

    @property
    def bounds(self):
        return self._bounds


class Optimizer:
    """scipy.optimize.minimize around an ObjectiveFunction (at3d/optimize.py:146-260)."""

    def __init__(self, objective_fn, prior_fn=None, callback_fn=None, method='L-BFGS-B', options=None):
        self._method = method
        self._options = dict(options or {'maxiter': 100, 'maxls': 10, 'gtol': 1e-16, 'ftol': 1e-8})      # the reference's, less 'disp'
        self._objective_fn = objective_fn
        self._prior_fn = list(np.atleast_1d(prior_fn)) if prior_fn is not None else []
        self._callback_fn = list(np.atleast_1d(callback_fn)) if callback_fn is not None else []
        self._iteration = 0
        self._state = None
        self._loss_history = []

    def callback(self, state):
        self._iteration += 1
        return [fn(optimizer=self) for fn in self._callback_fn]

    def objective(self, state):
        self._state = state
        loss, gradient = self._objective_fn(state)
        for prior in self._prior_fn:
            p_loss, p_grad = prior(state, self._iteration)          # at3d/optimize.py:196
            loss += p_loss
            gradient = gradient + p_grad
        self._loss_history.append(float(loss))
        return loss, np.asarray(gradient, np.float64)

    def minimize(self, initial_state, iteration_step=0, **kwargs):
        import scipy.optimize
        self._iteration = iteration_step
        args = dict(fun=self.objective, x0=np.asarray(initial_state, np.float64), method=self._method, jac=True,
                    options=self._options, callback=self.callback)
        args.update(kwargs)
        if self._method not in ('CG', 'Newton-CG'):
            b = self._objective_fn.bounds
            if len(b) == 1 and b[0] == (None, None):
                args['bounds'] = [(None, None)] * len(initial_state)
            elif len(b) == len(initial_state):
                args['bounds'] = b
            elif len(b) == 1:
                args['bounds'] = [b[0]] * len(initial_state)
        # every evaluation builds new solvers and render states (like the reference): the freed device memory of one
        # evaluation is kept for the next one while the loop runs
        from . import backend
        with backend.reusing_memory():
            return scipy.optimize.minimize(**args)

    @property
    def objective_fn(self):
        return self._objective_fn

    @property
    def loss_history(self):
        return self._loss_history

    @property
    def iteration(self):
        return self._iteration

    @property
    def method(self):
        return self._method

    @property
    def options(self):
        return self._options

    @property
    def state(self):
        """The state of the latest objective evaluation (what callbacks read through ``optimizer``)."""
        return self._state


class CallbackFn:
    """Calls ``callback_fn(optimizer=...)`` at most every `ckpt_period` seconds and collects the values of the dictionary
    it returns, by name, in ``output`` (at3d/callback.py:32-75)."""

    def __init__(self, callback_fn, ckpt_period=-1):
        self._ckpt_period = ckpt_period
        self._ckpt_time = time.time()
        self._callback_fn = callback_fn
        self.output = {}

    def __call__(self, optimizer=None):
        now = time.time()
        if now - self._ckpt_time > self._ckpt_period:
            self._ckpt_time = now
            out = self._callback_fn(optimizer=optimizer)
            for name, value in (out or {}).items():
                self.output.setdefault(name, []).append(value)


class GridStateGenerator:
    """state vector <-> solvers for optical unknowns on the property grid (at3d.medium.StateGenerator, at3d/medium.py:1649-1925).

    `solvers`: the SolversDict to (re)fill; `unknown_scatterers`; `mediums`: key -> OrderedDict scatterer name -> scatterer
    mapping (the fixed part; the unknown variables are overwritten from the state); `sources`, `surfaces`,
    `numerical_parameters`: key -> mapping; `num_stokes`: key -> int; `mask`: optional boolean [x, y, z] of the grid points
    in the state; `transforms`: optional mapping ``(scatterer name, variable name) -> (coordinate_transform,
    state_to_grid)`` with the objects of at3d_b200/transforms.py (the reference's ``UnknownScatterer.add_variable``,
    :1502-1537; None = no coordinate transform / the mask map).  The state vector is the concatenation of the variables'
    parts, each as long as its state_to_grid map says (``StateRepresentation``, :1596-1647).  Grid points a map leaves
    out (outside its mask) keep the value of the fixed medium."""

    def __init__(self, solvers, unknown_scatterers, mediums, sources, surfaces, numerical_parameters, num_stokes, mask=None,
                 warm_start=True, transforms=None):
        self._solvers, self._unknown = solvers, unknown_scatterers
        self._warm_start = bool(warm_start)
        self._mediums, self._sources, self._surfaces = mediums, sources, surfaces
        self._params, self._num_stokes = numerical_parameters, num_stokes
        grid = next(iter(next(iter(mediums.values())).values()))
        shape = np.asarray(_v(grid, 'extinction')).shape
        self._mask = np.ones(shape, bool) if mask is None else np.asarray(mask, bool)
        self._slots = [(name, v) for name, e in unknown_scatterers.items() for v in e.variables]
        transforms = {} if transforms is None else dict(transforms)
        unknown = set(transforms) - set(self._slots)
        if unknown:
            raise KeyError('`transforms` names variables that are not unknowns: {}'.format(sorted(unknown)))
        self._transforms = []
        for slot in self._slots:
            coordinate, to_grid = transforms.get(slot, (None, None))
            coordinate = TR.CoordinateTransform() if coordinate is None else coordinate
            to_grid = TR.StateToGridMask(mask=self._mask) if to_grid is None else to_grid
            map_shape = getattr(to_grid, '_grid_shape', shape)          # user-defined maps need not carry one
            if tuple(map_shape) != tuple(shape):
                raise ValueError("the state_to_grid map of {} is for grid shape {}, the medium has {}".format(
                    slot, tuple(map_shape), tuple(shape)))
            self._transforms.append((coordinate, to_grid))
        sizes = [getattr(t[1], 'state_size', None) for t in self._transforms]
        for i, (slot, size) in enumerate(zip(self._slots, sizes)):
            if size is None:                 # a map without the property: the length of what it makes of the medium
                sizes[i] = int(np.size(self._transforms[i][1].inverse_transform(
                    np.asarray(_v(next(iter(mediums.values()))[slot[0]], slot[1]), np.float64))))
        ends = np.cumsum(sizes)
        self._bounds = list(zip(np.concatenate(([0], ends[:-1])).tolist(), ends.tolist()))

    @property
    def state_size(self):
        return self._bounds[-1][1] if self._bounds else 0

    def get_state(self, key=None):
        """The state vector of the current mediums (:1834-1858)."""
        key = next(iter(self._mediums)) if key is None else key
        return np.concatenate([np.asarray(ct.inverse_transform(s2g.inverse_transform(
            np.asarray(_v(self._mediums[key][n], v), np.float64))), np.float64).reshape(-1)
            for (n, v), (ct, s2g) in zip(self._slots, self._transforms)])

    def __call__(self, state):
        """set_state_fn: rebuild every solver with the unknown variables taken from `state` (:1787-1832)."""
        state = np.asarray(state, np.float64)
        if state.shape != (self.state_size,):
            raise ValueError('state must have {} entries'.format(self.state_size))
        for key in list(self._mediums):
            medium = OrderedDict()
            for name, sc in self._mediums[key].items():
                sc = dict(sc)
                for (n, v), (ct, s2g), (a, b) in zip(self._slots, self._transforms, self._bounds):
                    if n == name:
                        arr = np.array(_v(sc, v), np.float32)
                        gridded = s2g(ct(state[a:b]))                    # UnknownScatterer.get_grid_data (:1539-1559)
                        where = getattr(s2g, '_where', None)                 # a user-defined map: its finite entries
                        where = np.isfinite(gridded) if where is None else where
                        arr[where] = gridded[where]
                        sc[v] = arr
                medium[name] = sc
            old = self._solvers.get(key)
            previous = None
            if old is not None:
                # the old solution is the first guess of the new solve (at3d/medium.py:1829-1830), where the facade can
                # continue from it: fixed grids
                if self._warm_start and getattr(old, '_solved', None) is not None and old._splitacc == 0.0 \
                        and old._srctype == 'S' and old._sfctype == 'FL':
                    previous = old.save_solution()
                old.close()
            self._mediums[key] = medium
            self._solvers[key] = RTE(self._params[key], medium, self._sources[key], self._surfaces[key],
                                     num_stokes=self._num_stokes[key])
            if previous is not None:
                self._solvers[key].load_solution(previous)

    def project_gradient_to_state(self, state, gradient_dataset):
        """Gridded gradient -> state gradient: state_to_grid, then the chain rule of the coordinate transform on that
        variable's part of the state (:1860-1890)."""
        g = np.asarray(gradient_dataset['gradient'])
        state = np.asarray(state, np.float64)
        return np.concatenate([np.asarray(ct.gradient_transform(state[a:b], s2g.gradient_transform(g[..., i])),
                                          np.float64).reshape(-1)
                               for i, ((ct, s2g), (a, b)) in enumerate(zip(self._transforms, self._bounds))])

    def transform_bounds(self, bounds):
        """``(lower, upper)`` of the state vector from physical bounds (:1892-1925).  `bounds`: mapping ``(scatterer name,
        variable name) -> (lower, upper)``, scalars or gridded arrays; swapped where a transform reverses the order."""
        lower, upper = np.zeros(self.state_size), np.zeros(self.state_size)
        shape = self._mask.shape
        for slot, (ct, s2g), (a, b) in zip(self._slots, self._transforms, self._bounds):
            for value, out in zip(bounds[slot], (lower, upper)):
                out[a:b] = ct.inverse_transform(s2g.inverse_bounds_transform(np.broadcast_to(np.asarray(value, float), shape)))
        return np.minimum(lower, upper), np.maximum(lower, upper)
