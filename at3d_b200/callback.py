"""``CallbackFn`` under the reference's module name (at3d/callback.py:32-75); the tensorboard writer of that module is
not mirrored."""
from .optimize import CallbackFn

__all__ = ['CallbackFn']
