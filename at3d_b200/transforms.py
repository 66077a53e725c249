"""Coordinate transforms and state <-> grid maps of the inverse problem: the host-side mirror of at3d/transforms.py.

Same class names, constructor arguments, methods and formulas as the reference (``CoordinateTransform`` :14,
``...Log`` :81, ``...Scaling`` :139, ``...Exp`` :214, ``...HyperBol`` :284; ``StateToGridMask`` :356, ``...Profile`` :461,
``...2D`` :551, ``...Uniform`` :643), used by ``optimize.GridStateGenerator`` the way ``at3d.medium.StateGenerator`` uses
them: state subset -> ``coordinate_transform`` -> ``state_to_grid`` -> gridded unknown, and the chain rule backwards.

Reference behaviours kept on purpose, because results must match: ``CoordinateTransformScaling.gradient_transform``
applies the INVERSE map to the gradient (``(g - offset) * scaling``, :203-204) and the reduced maps (profile, 2-D,
uniform) project gradients with the masked MEAN over the collapsed axes (``nanmean``, :499-502, :529-531), not the sum;
``CoordinateTransformExp.__call__`` returns ``+scaling * log(1 - state)`` (the reference negates twice, :238-239), so it is
not the inverse of ``inverse_transform``.  One reference behaviour NOT kept: ``StateToGridProfile.__call__`` indexes
``gridded_state[np.where(mask[..., i]), i]`` (:481-482), which addresses axes 0 and 1 instead of the level and raises for
nz > ny; here the level's value goes to the masked points of the level, as its docstring says.
"""
import numpy as np


class CoordinateTransform:
    """No transformation (at3d/transforms.py:14-79).  Transforms are square: same length in and out."""

    def __call__(self, abstract_state):
        return abstract_state

    def inverse_transform(self, physical_state):
        return physical_state

    def gradient_transform(self, abstract_state, physical_gradient):
        return physical_gradient


class CoordinateTransformLog(CoordinateTransform):
    """The state is log(physical) (:81-137)."""

    def __call__(self, abstract_state):
        return np.exp(abstract_state)

    def inverse_transform(self, physical_state):
        return np.log(physical_state)

    def gradient_transform(self, abstract_state, physical_gradient):
        return physical_gradient * np.exp(abstract_state)


class CoordinateTransformScaling(CoordinateTransform):
    """state = (physical - offset) * scaling_factor (:139-212)."""

    def __init__(self, offset, scaling_factor):
        self._offset, self._scaling_factor = offset, scaling_factor

    def __call__(self, abstract_state):
        return abstract_state / self._scaling_factor + self._offset

    def inverse_transform(self, physical_state):
        return (physical_state - self._offset) * self._scaling_factor

    def gradient_transform(self, abstract_state, physical_gradient):
        return self.inverse_transform(physical_gradient)           # as the reference does (:203-204)


class CoordinateTransformExp(CoordinateTransform):
    """state = 1 - exp(-physical / scaling), in [0, 1] (:214-282)."""

    def __init__(self, scaling):
        self._scaling = scaling

    def __call__(self, abstract_state):
        # the reference negates twice (:238-239): physical = +scaling * log(1 - state)
        return np.log(1.0 - abstract_state) * self._scaling

    def inverse_transform(self, physical_state):
        return 1.0 - np.exp(-physical_state / self._scaling)

    def gradient_transform(self, abstract_state, physical_gradient):
        return -self._scaling * physical_gradient / (abstract_state - 1.0)


class CoordinateTransformHyperBol(CoordinateTransform):
    """state = s * physical / (1 + s * physical), in [0, 1] (:284-354)."""

    def __init__(self, scaling):
        self._scaling = scaling

    def __call__(self, abstract_state):
        return (abstract_state / (1.0 - abstract_state)) / self._scaling

    def inverse_transform(self, physical_state):
        return self._scaling * physical_state / (1.0 + self._scaling * physical_state)

    def gradient_transform(self, abstract_state, physical_gradient):
        return physical_gradient / (self._scaling * (1.0 - abstract_state) ** 2)


class StateToGridMask:
    """Gridded unknowns <-> the 1-D vector of the grid points where `mask` is true (:356-459); without a mask a reshape."""

    def __init__(self, grid_shape=None, mask=None):
        if mask is None and grid_shape is None:
            raise ValueError("At least one of `grid_shape` or `mask` arguments must be provided.")
        if mask is None:
            mask = np.ones(grid_shape)
        elif grid_shape is None:
            grid_shape = np.shape(mask)
        elif tuple(grid_shape) != tuple(np.shape(mask)):
            raise ValueError("Both `grid_shape` and `mask` arguments were provided."
                             " The shape of `mask` is not consistent with `grid_shape`.")
        self._mask = np.asarray(mask)
        self._grid_shape = tuple(grid_shape)
        self._where = np.asarray(self._mask, bool)

    @property
    def state_size(self):
        """Length of this variable's part of the state vector."""
        return int(self._where.sum())

    def _masked(self, gridded):
        out = np.full(self._grid_shape, np.nan)
        out[self._where] = np.asarray(gridded)[self._where]
        return out

    def __call__(self, state):
        gridded_state = np.zeros(self._grid_shape)
        gridded_state[self._where] = state
        return gridded_state

    def inverse_transform(self, gridded_data):
        return np.asarray(gridded_data)[self._where]

    def gradient_transform(self, gridded_gradient):
        return np.asarray(gridded_gradient)[self._where]

    def inverse_bounds_transform(self, gridded_bounds):
        return np.asarray(gridded_bounds)[self._where]

    def _uniform_bounds(self, gridded_bounds):
        if np.size(np.unique(gridded_bounds)) != 1:
            raise NotImplementedError("Inverse Transform for non-uniform bounds for single variable"
                                      " have not yet been implemented.")
        return self.inverse_transform(gridded_bounds)


class StateToGridProfile(StateToGridMask):
    """One unknown per vertical level, spread over the masked points of the level (:461-549); NaN elsewhere."""

    @property
    def state_size(self):
        return self._grid_shape[-1]

    def __call__(self, state):
        return np.where(self._where, np.asarray(state, float).reshape((1,) * (len(self._grid_shape) - 1) + (-1,)), np.nan)

    def inverse_transform(self, gridded_data):
        return np.nanmean(self._masked(gridded_data), axis=tuple(range(len(self._grid_shape) - 1)))

    def gradient_transform(self, gridded_gradient):
        return self.inverse_transform(gridded_gradient)

    inverse_bounds_transform = StateToGridMask._uniform_bounds


class StateToGrid2D(StateToGridMask):
    """One unknown per column, spread over the masked points of the column (:551-641); NaN elsewhere."""

    @property
    def state_size(self):
        return int(np.prod(self._grid_shape[:2]))

    def __call__(self, state):
        return np.where(self._where, np.asarray(state, float).reshape(self._grid_shape[:2] + (1,)), np.nan)

    def inverse_transform(self, gridded_data):
        return np.nanmean(self._masked(gridded_data), axis=-1).ravel()

    def gradient_transform(self, gridded_gradient):
        return self.inverse_transform(gridded_gradient)

    inverse_bounds_transform = StateToGridMask._uniform_bounds


class StateToGridUniform(StateToGridMask):
    """One unknown for all masked points (:643-733); NaN elsewhere."""

    @property
    def state_size(self):
        return 1

    def __call__(self, state):
        return np.where(self._where, np.asarray(state, float).reshape(-1)[0], np.nan)

    def inverse_transform(self, gridded_data):
        return np.nanmean(self._masked(gridded_data)).ravel()

    def gradient_transform(self, gridded_gradient):
        return self.inverse_transform(gridded_gradient)

    inverse_bounds_transform = StateToGridMask._uniform_bounds
