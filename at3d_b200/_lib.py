"""ctypes binding of libat3d_b200.so (include/at3d_b200.h).  No fallback: a missing library or a
missing CUDA device raises."""
import ctypes as C
import os
import numpy as np
from .state import StateDesc, GradDesc, i32, f32, f64, P

_HERE = os.path.dirname(os.path.abspath(__file__))
# AT3D_B200_LIB: developer override used to compare kernel build variants on the GPU box
LIB_PATH = os.environ.get('AT3D_B200_LIB') or os.path.join(_HERE, 'lib', 'libat3d_b200.so')
ERRLEN = 600


class At3dError(RuntimeError):
    """Non-zero IERR from the library (at3d.exceptions.SHDOMError equivalent)."""

    def __init__(self, code, msg):
        super().__init__('at3d_b200 error %d: %s' % (code, msg))
        self.code, self.msg = code, msg


class RaysC(C.Structure):
    _fields_ = [('nrays', i32), ('memspace', i32), ('camx', C.c_void_p), ('camy', C.c_void_p),
                ('camz', C.c_void_p), ('cammu', C.c_void_p), ('camphi', C.c_void_p), ('packs', C.c_void_p)]


class CsDeviceDesc(C.Structure):
    """at3d_cs_device_desc (include/at3d_b200.h): COMPUTE_SOURCE on device-resident arrays."""
    _fields_ = [(k, i32) for k in ('npts', 'nstokes', 'nstleg', 'nlm', 'ml', 'mm', 'nleg', 'npart', 'maxnmicro', 'numphase',
                                   'deltam', 'interp_new')] + \
               [('srctype', C.c_char), ('phasemax', f32), ('solarmu', f32)] + \
               [(k, C.c_void_p) for k in ('extinct', 'albedo', 'total_ext', 'legen', 'iphase', 'phaseinterpwt', 'dirflux',
                                          'rshptr', 'radiance', 'ylmsun', 'planck')]


class TraceC(C.Structure):
    _fields_ = [('max_per_ray', i32), ('cells', C.c_void_p), ('ncells', C.c_void_p), ('nsub', C.c_void_p)]


_lib = None

# every symbol include/at3d_b200.h declares (tests/test_capi_symbols.py checks the export list)
SYMBOLS = ['at3d_b200_version', 'at3d_device_count', 'at3d_set_device', 'at3d_state_create',
           'at3d_state_attach_gradient', 'at3d_state_destroy', 'at3d_state_bytes', 'at3d_state_get_bcrad',
           'at3d_ylmall', 'at3d_precompute_phase_check', 'at3d_compute_source', 'at3d_render',
           'at3d_levisapprox_gradient', 'at3d_levisapprox_gradient_jacobian', 'at3d_prepare_deriv_interps', 'at3d_make_direct_derivative',
           'at3d_average_subpixel_rays', 'at3d_update_costfunction', 'at3d_make_direct', 'at3d_state_get_counts',
           'at3d_sh_to_do', 'at3d_do_to_sh', 'at3d_path_integration_ip', 'at3d_solver_create',
           'at3d_solver_path_integration', 'at3d_solver_solve', 'at3d_solver_solve_from', 'at3d_solver_update_medium', 'at3d_solver_destroy', 'at3d_sweeping_order', 'at3d_transfer_pa_to_grid', 'at3d_solve_adaptive', 'at3d_compute_source_device', 'at3d_ray_pack_bytes', 'at3d_make_ray_packs', 'at3d_trim_memory', 'at3d_set_memory_reuse']


class _Missing:
    def __init__(self, name):
        self.name = name

    def __call__(self, *a, **k):
        raise ImportError('libat3d_b200.so does not export %s (stale build?)' % self.name)


class _Binder:
    """Attribute access to the CDLL that tolerates a symbol missing at bind time and fails loudly
    when such a symbol is actually called."""

    def __init__(self, dll):
        object.__setattr__(self, '_dll', dll)

    def __getattr__(self, name):
        try:
            return getattr(self._dll, name)
        except AttributeError:
            m = _Missing(name)
            object.__setattr__(self, name, m)
            return m


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(at3d_b200 has no CPU fallback)' % LIB_PATH)
    L = _Binder(C.CDLL(LIB_PATH))
    L.at3d_b200_version.restype = C.c_char_p
    L.at3d_state_bytes.restype = C.c_int64
    L.at3d_state_bytes.argtypes = [C.c_void_p]
    L.at3d_state_create.argtypes = [P(StateDesc), P(C.c_void_p), C.c_char_p]
    L.at3d_state_attach_gradient.argtypes = [C.c_void_p, P(GradDesc), C.c_char_p]
    L.at3d_state_destroy.argtypes = [C.c_void_p]
    L.at3d_state_get_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    L.at3d_state_get_bcrad.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    L.at3d_render.argtypes = [C.c_void_p, P(RaysC), C.c_void_p, i32, i32, i32, P(TraceC), C.c_void_p,
                              P(f64), C.c_char_p]
    L.at3d_levisapprox_gradient.argtypes = [C.c_void_p, P(RaysC), P(GradDesc), C.c_void_p, C.c_void_p,
                                            C.c_void_p, P(TraceC), C.c_void_p, P(f64), C.c_char_p]
    L.at3d_levisapprox_gradient_jacobian.argtypes = [C.c_void_p, P(RaysC), P(GradDesc), C.c_void_p, C.c_void_p,
                                                     C.c_void_p, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
    L.at3d_compute_source.argtypes = [P(StateDesc), i32, f32, i32, i32, i32, i32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, P(f32), P(f32), P(f32), P(f32), P(f64),
                                      C.c_char_p]
    L.at3d_compute_source_device.argtypes = [P(CsDeviceDesc), i32, f32, C.c_int64, i32, i32] + [C.c_void_p] * 7 + \
        [C.c_int64, i32, C.c_void_p, P(i32), P(f64), C.c_char_p]
    L.at3d_ray_pack_bytes.restype = C.c_int64
    L.at3d_make_ray_packs.argtypes = [C.c_void_p, P(RaysC), C.c_void_p, C.c_char_p]
    L.at3d_ylmall.argtypes = [i32, f32, f32, i32, i32, i32, C.c_void_p, C.c_char_p]
    L.at3d_precompute_phase_check.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, C.c_void_p,
                                              C.c_void_p, i32, i32, i32, C.c_char_p]
    L.at3d_prepare_deriv_interps.argtypes = [P(StateDesc), i32, i32, i32, i32, f32, f32, f32, f32,
                                             C.c_void_p, P(GradDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_char_p]
    L.at3d_make_direct_derivative.argtypes = [i32, i32, i32, i32, i32, f32, f32, f32, f32, C.c_void_p,
                                              C.c_void_p, i32, i32, i32, i32] + [f64] * 13 + \
                                             [C.c_void_p, C.c_void_p, i32, C.c_char_p]
    L.at3d_make_direct.argtypes = [i32, i32, i32, i32, i32, i32, i32, f32, f32, f32, C.c_void_p,
                                   i32, i32, i32, f32, f32, f32, f32, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_void_p, i32, i32, i32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_char_p]
    L.at3d_path_integration_ip.argtypes = [P(StateDesc)] + [C.c_void_p] * 7 + [P(C.c_double), C.c_char_p]
    L.at3d_solver_create.argtypes = [P(StateDesc), C.c_void_p, f32, P(C.c_void_p), C.c_char_p]
    L.at3d_solver_path_integration.argtypes = [C.c_void_p] * 7 + [P(C.c_double), C.c_char_p]
    L.at3d_solver_solve.argtypes = [C.c_void_p, P(StateDesc), i32, f32, f32, i32, i32, i32, i32] + [C.c_void_p] * 6 + \
        [P(i32), P(f32), C.c_void_p, C.c_char_p]
    L.at3d_solver_solve_from.argtypes = [C.c_void_p, P(StateDesc), i32, f32, f32, i32, i32, i32, i32, i32] + [C.c_void_p] * 6 + \
        [P(i32), P(f32), C.c_void_p, C.c_char_p]
    L.at3d_solver_update_medium.argtypes = [C.c_void_p, P(StateDesc), C.c_char_p]
    L.at3d_solver_destroy.argtypes = [C.c_void_p]
    L.at3d_sweeping_order.argtypes = [P(StateDesc), C.c_void_p, C.c_char_p]
    L.at3d_transfer_pa_to_grid.argtypes = [i32, C.c_void_p, i32, i32, i32, f32, f32, f32, f32, C.c_void_p, i32, i32] + \
        [C.c_void_p] * 4 + [i32, C.c_void_p, i32, i32, f32] + [C.c_void_p] * 5 + [C.c_char_p]
    L.at3d_solve_adaptive.argtypes = [P(StateDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
    L.at3d_sh_to_do.argtypes = [i32] * 8 + [C.c_void_p] * 7 + [P(C.c_double), C.c_char_p]
    L.at3d_do_to_sh.argtypes = [i32] * 8 + [C.c_void_p] * 7 + [P(C.c_double), C.c_char_p]
    L.at3d_average_subpixel_rays.argtypes = [i32, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
    L.at3d_update_costfunction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           i32, i32, i32, i32, C.c_void_p, i32, C.c_char_p]
    _lib = L
    return L


def check(code, buf):
    if code != 0:
        raise At3dError(code, buf.value.decode(errors='replace').strip())


def errbuf():
    return C.create_string_buffer(ERRLEN)


def vp(a):
    """void* of a numpy array, a torch tensor (data_ptr) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()
