"""at3d_b200 -- B200-native (sm_100a) implementation of the AT3D / SHDOM data-parallel hot path.

The compute path lives in ``lib/libat3d_b200.so`` (hand-written CUDA behind the C-ABI of
``include/at3d_b200.h``); there is no CPU fallback.  Host-side modules mirror the reference's
Python interface for the path (``at3d.core`` keyword API, ``solver.RTE`` state, sensors).
"""
__version__ = '0.1.0'
