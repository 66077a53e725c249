"""Host-side containers for the SHDOM state arrays and their C-ABI descriptors.

``ShdomState`` holds the arrays of a solved ``at3d.solver.RTE`` (the private ``_gridpos``,
``_gridptr`` ... attributes, at3d/solver.py:106-247, Appendix A of SURVEY.md) in the reference
layout: Fortran order, 1-based index contents, REAL=float32, INTEGER=int32, INTEGER*2=int16.
``GradInputs`` holds the extra arguments of ``core.levisapprox_gradient`` (at3d/gradient.py:262-398).

``STATE_FIELDS`` / ``GRAD_FIELDS`` mirror ``at3d_state_desc`` / ``at3d_grad_desc`` of
include/at3d_b200.h field by field.
"""
import ctypes as C
import numpy as np

i32, f32, f64 = C.c_int32, C.c_float, C.c_double
P = C.POINTER

# (name, ctype, numpy dtype or None for scalars)
STATE_FIELDS = [
    ('nstokes', i32, None), ('nstleg', i32, None), ('nx', i32, None), ('ny', i32, None),
    ('nz', i32, None), ('npts', i32, None), ('ncells', i32, None),
    ('ml', i32, None), ('mm', i32, None), ('nlm', i32, None), ('nleg', i32, None),
    ('numphase', i32, None), ('npart', i32, None), ('maxnmicro', i32, None),
    ('bcflag', i32, None), ('ipflag', i32, None),
    ('nmu', i32, None), ('nphi0max', i32, None), ('nang', i32, None),
    ('maxnbc', i32, None), ('ntoppts', i32, None), ('nbotpts', i32, None), ('nsfcpar', i32, None),
    ('nscatangle', i32, None), ('nstphase', i32, None),
    ('deltam', i32, None), ('srctype', i32, None), ('units', i32, None),
    ('sfctype0', i32, None), ('sfctype1', i32, None), ('interp_new', i32, None),
    ('solarmu', f32, None), ('solaraz', f32, None), ('solarflux', f32, None), ('wavelen', f32, None),
    ('gndtemp', f32, None), ('gndalbedo', f32, None), ('phasemax', f32, None),
    ('waveno0', f32, None), ('waveno1', f32, None),
    ('tautol', f64, None), ('transcut', f64, None),
    ('gridptr', P(i32), np.int32), ('neighptr', P(i32), np.int32), ('treeptr', P(i32), np.int32),
    ('cellflags', P(C.c_int16), np.int16),
    ('xgrid', P(f32), np.float32), ('ygrid', P(f32), np.float32), ('zgrid', P(f32), np.float32),
    ('gridpos', P(f32), np.float32),
    ('extinct', P(f32), np.float32), ('albedo', P(f32), np.float32), ('total_ext', P(f32), np.float32),
    ('legen', P(f32), np.float32), ('iphase', P(i32), np.int32), ('phaseinterpwt', P(f32), np.float32),
    ('dirflux', P(f32), np.float32), ('fluxes', P(f32), np.float32),
    ('shptr', P(i32), np.int32), ('source', P(f32), np.float32),
    ('rshptr', P(i32), np.int32), ('radiance', P(f32), np.float32),
    ('ylmsun', P(f32), np.float32), ('phasetab', P(f32), np.float32),
    ('planck', P(f32), np.float32), ('temp', P(f32), np.float32),
    ('nphi0', P(i32), np.int32), ('mu', P(f32), np.float32), ('phi', P(f32), np.float32),
    ('wtdo', P(f32), np.float32), ('skyrad', P(f32), np.float32),
    ('bcptr', P(i32), np.int32), ('bcrad', P(f32), np.float32),
    ('sfcgridparms', P(f32), np.float32), ('sfcgridrad', P(f32), np.float32),
]

GRAD_FIELDS = [
    ('npix', i32, None), ('maxpg', i32, None), ('numder', i32, None), ('dnumphase', i32, None),
    ('deriv_maxnmicro', i32, None), ('longest_path_pts', i32, None),
    ('nuncertainty', i32, None), ('maxsubgridints', i32, None),
    ('exact_single_scatter', i32, None), ('singlescatter', i32, None), ('costfunc_ll', i32, None),
    ('extmin', f64, None), ('scatmin', f64, None),
    ('partder', P(i32), np.int32), ('doexact', P(i32), np.int32),
    ('measurements', P(f32), np.float32), ('uncertainties', P(f64), np.float64),
    ('rays_per_pixel', P(i32), np.int32), ('ray_weights', P(f64), np.float64),
    ('stokes_weights', P(f64), np.float64),
    ('dext', P(f32), np.float32), ('dalb', P(f32), np.float32), ('dextm', P(f32), np.float32),
    ('dalbm', P(f32), np.float32), ('dfj', P(f32), np.float32),
    ('optinterpwt', P(f32), np.float32), ('interpptr', P(i32), np.int32),
    ('dleg', P(f32), np.float32), ('dphasetab', P(f32), np.float32),
    ('diphasep', P(i32), np.int32), ('dphasewtp', P(f32), np.float32),
    ('iphasep', P(i32), np.int32), ('phasewtp', P(f32), np.float32),
    ('extinctp', P(f32), np.float32), ('albedop', P(f32), np.float32),
    ('dtemp', P(f32), np.float32),
    ('dpath', P(f32), np.float32), ('dptr', P(i32), np.int32),
    ('beam_npx', i32, None), ('beam_npy', i32, None), ('beam_npz', i32, None),
    ('beam_xstart', f32, None), ('beam_ystart', f32, None),
    ('beam_zlevels', P(f32), np.float32), ('beam_d', P(f64), np.float64), ('beam_i', P(i32), np.int32),
]


def make_struct(name, fields):
    return type(name, (C.Structure,), {'_fields_': [(n, t) for n, t, _ in fields]})


StateDesc = make_struct('StateDesc', STATE_FIELDS)
GradDesc = make_struct('GradDesc', GRAD_FIELDS)


def farray(a, dtype):
    """Array in the reference layout: given dtype, Fortran order, own memory."""
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def _ptr(a, ctype):
    if a is None:
        return C.cast(None, ctype)
    return a.ctypes.data_as(ctype)


class _Fields:
    """Attribute bag validated against a field table; builds the matching ctypes struct."""
    _FIELDS = ()
    _CHARS = ()

    def __init__(self, **kw):
        for n, _, dt in self._FIELDS:
            setattr(self, n, None if dt is not None else 0)
        for k, v in kw.items():
            setattr(self, k, v)

    def normalize(self):
        for n, _, dt in self._FIELDS:
            v = getattr(self, n)
            if dt is not None and v is not None:
                a = farray(v, dt)
                setattr(self, n, a)
        return self

    def fill(self, struct):
        """Fill a ctypes struct with identical field names; returns it (arrays must stay alive)."""
        self.normalize()
        for n, ct, dt in self._FIELDS:
            v = getattr(self, n)
            if dt is None:
                if n in self._CHARS and isinstance(v, (str, bytes)):
                    v = ord(v)
                setattr(struct, n, v)
            else:
                setattr(struct, n, _ptr(v, ct))
        return struct

    def copy(self):
        out = type(self)()
        for n, _, dt in self._FIELDS:
            v = getattr(self, n)
            setattr(out, n, v.copy(order='F') if isinstance(v, np.ndarray) else v)
        for k, v in self.__dict__.items():
            if not hasattr(out, k) or getattr(out, k) is None:
                setattr(out, k, v.copy() if isinstance(v, np.ndarray) else v)
        return out


class ShdomState(_Fields):
    _FIELDS = STATE_FIELDS
    _CHARS = ('srctype', 'units', 'sfctype0', 'sfctype1')

    def desc(self):
        return self.fill(StateDesc())


class GradInputs(_Fields):
    _FIELDS = GRAD_FIELDS

    def desc(self):
        return self.fill(GradDesc())


class Rays:
    """Sensor rays (CAMX..CAMPHI of RENDER).  camx/y/z are down-cast to float32 exactly as f2py
    does for the reference."""

    def __init__(self, camx, camy, camz, cammu, camphi):
        self.camx = np.ascontiguousarray(camx, dtype=np.float32)
        self.camy = np.ascontiguousarray(camy, dtype=np.float32)
        self.camz = np.ascontiguousarray(camz, dtype=np.float32)
        self.cammu = np.ascontiguousarray(cammu, dtype=np.float64)
        self.camphi = np.ascontiguousarray(camphi, dtype=np.float64)
        self.nrays = int(self.camx.shape[0])

    def slice(self, lo, hi):
        return Rays(self.camx[lo:hi], self.camy[lo:hi], self.camz[lo:hi], self.cammu[lo:hi], self.camphi[lo:hi])
