"""at3d_b200 -- B200-native (sm_100a) implementation of the AT3D / SHDOM data-parallel hot path.

The compute path lives in ``lib/libat3d_b200.so`` (hand-written CUDA behind the C-ABI of
``include/at3d_b200.h``); there is no CPU fallback.  Host-side modules mirror the reference's
Python interface for the path (``at3d.core`` keyword API, ``solver.RTE`` state, sensors).
"""
__version__ = '0.1.0'

_SUBMODULES = ('backend', 'callback', 'configuration', 'containers', 'core', 'device', 'gradient', 'gradsetup', 'grid',
               'medium', 'optimize', 'parallel', 'rte', 'sensor', 'solver', 'source', 'state', 'surface', 'synthetic',
               'transforms', 'uncertainties', 'util')


def __getattr__(name):
    """``import at3d_b200 as at3d; at3d.sensor.perspective_projection(...)``: submodules load on first use (the
    reference's ``at3d/__init__.py`` imports them all; here importing the package does not load the CUDA library)."""
    if name in _SUBMODULES:
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError("module '{}' has no attribute '{}'".format(__name__, name))
