"""``save_forward_model`` / ``load_forward_model``: sensors and solvers of a forward model in one file.

Mirror of at3d/util.py:512-725 (``save_sensors``, ``load_sensors``, ``save_solvers``, ``load_solvers``,
``save_forward_model``, ``load_forward_model``): same functions, same arguments, the same tree of groups --

    sensors/<instrument>/<j>/<variable>
    solvers/<key>/medium/<scatterer>/<variable>
    solvers/<key>/numerical_parameters/<variable>     (with ``num_stokes`` added, :655-656)
    solvers/<key>/surface | source | grid | atmosphere/<variable>

-- the same rule for a file name that already exists (an integer is appended and a RuntimeWarning raised, :706-720), the
same restoration on load (``stokes`` / ``use_subpixel_rays`` / ``deltam`` / ``high_order_radiance`` back to booleans,
solver keys back to floats, :533-535, :586-587, :611).  The solution itself is not part of the file (:692-693): see
``RTE.save_solution``.

Container: the reference writes netCDF-4 groups through xarray; neither netCDF4 / HDF5 nor xarray exist in this image, so
the tree is stored as one NumPy ``.npz`` archive whose member names are the group paths above (``numpy.load`` lists them;
a file written by the reference is NOT readable here and vice versa; dataset attributes are the members ``@name`` of
their group).  Datasets come back as plain mappings ``name -> array`` with ``attrs`` (at3d_b200/_dataset.py), which is
what ``RTE`` and ``SensorsDict`` of this package take.
"""
import os
import warnings
from collections import OrderedDict
import numpy as np
from ._dataset import Dataset
from .containers import SensorsDict, SolversDict
from .rte import RTE

_SEP = '/'
_ATTR = '@'


def _variables(ds):
    """(name, array) pairs of an xarray.Dataset (data variables and coordinates) or of a plain mapping."""
    names = list(ds.variables) if hasattr(ds, 'variables') else list(ds.keys())
    for name in names:
        x = ds[name]
        x = getattr(x, 'data', x)
        if x is None:
            continue
        a = np.asarray(x)
        if a.dtype == object:
            # the reference pickles such entries (pickle_objects, at3d/util.py:609-610); nothing on the path needs them
            raise TypeError("variable '%s' holds Python objects and cannot be saved" % name)
        yield str(name), a


def _put(tree, group, ds):
    for name, a in _variables(ds):
        if _SEP in name or name.startswith(_ATTR):
            raise ValueError("variable name '%s' contains '%s' or starts with '%s'" % (name, _SEP, _ATTR))
        tree[group + _SEP + name] = a
    # dataset attributes (projection, resolution, sub-pixel ray arguments ...) travel as '@name' members
    for name, value in dict(getattr(ds, 'attrs', None) or {}).items():
        a = np.asarray(value)
        if a.dtype != object and _SEP not in str(name):
            tree[group + _SEP + _ATTR + str(name)] = a


def _read(file_name):
    with np.load(file_name, allow_pickle=False) as f:
        return OrderedDict((k, f[k]) for k in f.files)


def _write(file_name, tree):
    # numpy appends '.npz' to names without it: write through a handle so that the name is the caller's
    with open(file_name, 'wb') as fh:
        np.savez(fh, **tree)


def _append(file_name, tree):
    old = _read(file_name) if os.path.exists(file_name) and os.path.getsize(file_name) > 0 else OrderedDict()
    old.update(tree)
    _write(file_name, old)


def _groups(tree, prefix):
    """Names of the groups directly below `prefix`, in file order."""
    out, n = [], len(prefix) + 1
    for k in tree:
        if k.startswith(prefix + _SEP):
            g = k[n:].split(_SEP, 1)
            if len(g) == 2 and g[0] not in out:
                out.append(g[0])
    return out


def _dataset(tree, group):
    n = len(group) + 1
    ds = Dataset()
    for k, a in tree.items():
        if k.startswith(group + _SEP) and _SEP not in k[n:]:
            value = a[()] if a.ndim == 0 else a
            if k[n:].startswith(_ATTR):
                ds.attrs[k[n + len(_ATTR):]] = value.item() if isinstance(value, np.generic) else value
            else:
                ds[k[n:]] = value
    return ds


def save_sensors(file_name, sensors):
    """at3d/util.py:540-563."""
    if not isinstance(sensors, SensorsDict):
        raise TypeError("`sensors` should be an instance of '{}'".format(SensorsDict))
    tree = OrderedDict()
    for key, sensor in sensors.items():
        for j, image in enumerate(sensor['sensor_list']):
            _put(tree, 'sensors' + _SEP + str(key) + _SEP + str(j), image)
    _append(file_name, tree)


def load_sensors(file_name):
    """at3d/util.py:512-537."""
    tree = _read(file_name)
    sensor_dict = SensorsDict()
    for key in _groups(tree, 'sensors'):
        for i in _groups(tree, 'sensors' + _SEP + key):
            ds = _dataset(tree, 'sensors' + _SEP + key + _SEP + i)
            ds['stokes'] = np.asarray(ds['stokes']).astype(bool)
            if 'use_subpixel_rays' in ds:
                ds['use_subpixel_rays'] = bool(ds['use_subpixel_rays'])
            sensor_dict.add_sensor(key, ds)
    return sensor_dict


def save_solvers(file_name, solvers):
    """at3d/util.py:628-665."""
    if not isinstance(solvers, SolversDict):
        raise TypeError("`solvers` should be an instance of '{}'".format(SolversDict))
    tree = OrderedDict()
    for key, solver in solvers.items():
        base = 'solvers' + _SEP + str(key) + _SEP
        for name, med in solver.medium.items():
            _put(tree, base + 'medium' + _SEP + str(name), med)
        _put(tree, base + 'numerical_parameters', solver.numerical_params)
        tree[base + 'numerical_parameters' + _SEP + 'num_stokes'] = np.asarray(solver._nstokes)
        _put(tree, base + 'surface', solver.surface)
        _put(tree, base + 'source', solver.source)
        _put(tree, base + 'grid', solver._grid)
        if solver.atmosphere is not None:
            _put(tree, base + 'atmosphere', solver.atmosphere)
    _append(file_name, tree)


def load_solvers(file_name):
    """at3d/util.py:566-625: the solvers (not solved) and the grid of the first one."""
    tree = _read(file_name)
    solver_dict = SolversDict()
    keys = _groups(tree, 'solvers')
    for key in keys:
        base = 'solvers' + _SEP + key + _SEP
        numerical_params = _dataset(tree, base + 'numerical_parameters')
        for flag in ('deltam', 'high_order_radiance', 'acceleration_flag'):
            if flag in numerical_params:
                numerical_params[flag] = bool(numerical_params[flag])
        num_stokes = int(numerical_params['num_stokes'])
        mediums = OrderedDict((name, _dataset(tree, base + 'medium' + _SEP + name))
                              for name in _groups(tree, base + 'medium'))
        atmosphere = _dataset(tree, base + 'atmosphere') if 'atmosphere' in _groups(tree, base[:-1]) else None
        solver_dict.add_solver(float(key), RTE(numerical_params=numerical_params, medium=mediums,
                                               source=_dataset(tree, base + 'source'),
                                               surface=_dataset(tree, base + 'surface'),
                                               num_stokes=num_stokes, name=None, atmosphere=atmosphere))
    rte_grid = _dataset(tree, 'solvers' + _SEP + keys[0] + _SEP + 'grid')
    return solver_dict, rte_grid


def load_forward_model(file_name, load_solver=True):
    """at3d/util.py:668-684: ``(sensor_dict, solver_dict, rte_grid)``."""
    sensor_dict = load_sensors(file_name)
    if load_solver:
        solver_dict, rte_grid = load_solvers(file_name)
    else:
        solver_dict = SolversDict()
        tree = _read(file_name)
        rte_grid = _dataset(tree, 'solvers' + _SEP + _groups(tree, 'solvers')[0] + _SEP + 'grid')
    return sensor_dict, solver_dict, rte_grid


def _safe_file_name(file_name):
    """An integer is appended until the name is free (at3d/util.py:706-720; any extension, not only '.nc')."""
    root, ext = os.path.splitext(file_name)
    counter, out = 1, file_name
    while os.path.exists(out):
        out = '{}_{}{}'.format(root, counter, ext)
        counter += 1
    if out != file_name:
        warnings.warn("file_name '{}' already exists, your file is now at '{}'".format(file_name, out),
                      category=RuntimeWarning)
    return out


def save_forward_model(file_name, sensors, solvers):
    """at3d/util.py:687-725.  Returns the name actually written (the reference returns nothing)."""
    if not isinstance(sensors, SensorsDict):
        raise TypeError("`sensors` should be an instance of '{}'".format(SensorsDict))
    if not isinstance(solvers, SolversDict):
        raise TypeError("`solvers` should be an instance of '{}'".format(SolversDict))
    file_name = _safe_file_name(file_name)
    _write(file_name, OrderedDict())
    save_solvers(file_name, solvers)
    save_sensors(file_name, sensors)
    return file_name
