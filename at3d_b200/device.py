"""The solved SHDOM state resident in HBM, and the ray kernels that read it.

``DeviceState`` owns an ``at3d_state`` handle (include/at3d_b200.h).  It replaces the per-call array
marshalling of the reference's f2py boundary (at3d/solver.py:681-759, at3d/gradient.py:262-398):
the state is uploaded and re-packed once per solve, sensor rays stream through per call.
Ray inputs / outputs may be numpy arrays (host, copies inside the call) or torch CUDA tensors
(device pointers, no copies).
"""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import RaysC, TraceC, vp


def _is_torch(x):
    return x is not None and not isinstance(x, np.ndarray) and hasattr(x, 'data_ptr')


class DeviceState:
    def __init__(self, state):
        """state: ``ShdomState`` (host arrays in the reference layout)."""
        L = _lib.lib()
        self._L = L
        self.state = state
        self._desc = state.desc()
        self.nstokes = int(state.nstokes)
        self._h = C.c_void_p()
        buf = _lib.errbuf()
        _lib.check(L.at3d_state_create(C.byref(self._desc), C.byref(self._h), buf), buf)
        self._grad = None

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._L.at3d_state_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def hbm_bytes(self):
        return int(self._L.at3d_state_bytes(self._h))

    def counts(self):
        """Work counters of the last render / gradient call (include/at3d_b200.h)."""
        out = np.zeros(8, np.int64)
        buf = _lib.errbuf()
        _lib.check(self._L.at3d_state_get_counts(self._h, vp(out), buf), buf)
        return dict(cells=int(out[0]), points=int(out[1]), sum_ns=int(out[2]), sum_nr=int(out[3]),
                    subintervals=int(out[4]), rays=int(out[5]), surface_hits=int(out[6]))

    def bcrad(self):
        s = self.state
        lamb = s.sfctype1 in ('L', ord('L'))
        out = np.zeros((s.nstokes, s.ntoppts + s.nbotpts * (1 if lamb else 1 + s.nang // 2)), np.float32, order='F')
        buf = _lib.errbuf()
        _lib.check(self._L.at3d_state_get_bcrad(self._h, vp(out), buf), buf)
        return out

    # ------------------------------------------------------------------
    def _rays_struct(self, camx, camy, camz, cammu, camphi, packs=None):
        dev = _is_torch(camx)
        r = RaysC()
        r.nrays = int(camx.shape[0])
        r.memspace = 1 if dev else 0
        r.camx, r.camy, r.camz, r.cammu, r.camphi = vp(camx), vp(camy), vp(camz), vp(cammu), vp(camphi)
        r.packs = vp(packs) if (dev and packs is not None) else None
        return r, dev

    def make_ray_packs(self, rays, device='cuda'):
        """Per-ray setup records of HOST rays (at3d_make_ray_packs: host libm, bit-exact walk) as a CUDA uint8 tensor, to
        be attached as ``.packs`` to device-resident rays of the same geometry."""
        import torch
        r, dev = self._rays_struct(rays.camx, rays.camy, rays.camz, rays.cammu, rays.camphi)
        if dev:
            raise ValueError('make_ray_packs takes host (numpy) rays')
        nb = int(self._L.at3d_ray_pack_bytes())
        out = np.zeros(r.nrays * nb, np.uint8)
        buf = _lib.errbuf()
        _lib.check(self._L.at3d_make_ray_packs(self._h, C.byref(r), vp(out), buf), buf)
        return torch.from_numpy(out).to(device)

    def render(self, rays, correctinterpolate=True, singlescatter=False, nosurface=False,
               trace_cap=0, out=None, stream=None, timing=False):
        """RENDER (src/polarized/shdomsub4.f:93).  ``rays``: ``state.Rays`` or any object with
        camx, camy, camz (float32) and cammu, camphi (float64) numpy arrays / torch CUDA tensors.
        Returns stokes [nstokes, nrays] (numpy F-order, or the torch tensor ``out``), plus the
        trace dict when ``trace_cap`` > 0 and the kernel milliseconds when ``timing``."""
        r, dev = self._rays_struct(rays.camx, rays.camy, rays.camz, rays.cammu, rays.camphi, getattr(rays, 'packs', None))
        n = r.nrays
        if out is None:
            if dev:
                import torch
                out = torch.empty((n, self.nstokes), dtype=torch.float32, device=rays.camx.device)
            else:
                out = np.zeros((self.nstokes, n), np.float32, order='F')
        tr = None
        trace = None
        if trace_cap > 0:
            if dev:
                raise ValueError('tracing is only supported with host arrays')
            trace = dict(cells=np.zeros((trace_cap, n), np.int32, order='F'),
                         ncells=np.zeros(n, np.int32), nsub=np.zeros(n, np.int32))
            tr = TraceC(trace_cap, vp(trace['cells']), vp(trace['ncells']), vp(trace['nsub']))
        ms = C.c_double(0.0)
        buf = _lib.errbuf()
        code = self._L.at3d_render(self._h, C.byref(r), vp(out), int(correctinterpolate), int(singlescatter),
                                   int(nosurface), C.byref(tr) if tr is not None else None,
                                   C.c_void_p(stream) if stream else None,
                                   C.byref(ms) if timing else None, buf)
        _lib.check(code, buf)
        res = [out]
        if trace is not None:
            res.append(trace)
        if timing:
            res.append(ms.value)
        return res[0] if len(res) == 1 else tuple(res)

    # ------------------------------------------------------------------
    def attach_gradient(self, grad):
        """Upload the derivative tables of a ``GradInputs`` (once per gradient evaluation)."""
        self._grad = grad
        self._gdesc = grad.desc()
        buf = _lib.errbuf()
        _lib.check(self._L.at3d_state_attach_gradient(self._h, C.byref(self._gdesc), buf), buf)

    def gradient_jacobian(self, rays, pix, jacobianptr):
        """LEVISAPPROX_GRADIENT with MAKEJACOBIAN=.TRUE. (shdomsub4.f:536-631), host arrays.
        ``jacobianptr``: 1-based property-grid indices.  Returns (gradout, cost, stokesout,
        jacobian [nstokes,numder,njac,npix] f32)."""
        if self._grad is None:
            raise RuntimeError('attach_gradient() first')
        g = self._grad
        r, dev = self._rays_struct(rays.camx, rays.camy, rays.camz, rays.cammu, rays.camphi, getattr(rays, 'packs', None))
        if dev:
            raise NotImplementedError('gradient_jacobian takes host (numpy) arrays')
        gd = g.desc()
        npix = int(pix.rays_per_pixel.shape[0])
        gd.npix = npix
        gd.nuncertainty = int(pix.uncertainties.shape[0])
        gd.measurements = C.cast(vp(pix.measurements), type(gd.measurements))
        gd.uncertainties = C.cast(vp(pix.uncertainties), type(gd.uncertainties))
        gd.rays_per_pixel = C.cast(vp(pix.rays_per_pixel), type(gd.rays_per_pixel))
        gd.ray_weights = C.cast(vp(pix.ray_weights), type(gd.ray_weights))
        gd.stokes_weights = C.cast(vp(pix.stokes_weights), type(gd.stokes_weights))
        jp = np.ascontiguousarray(jacobianptr, np.int32).ravel()
        gradout = np.zeros((g.maxpg, g.numder), np.float64, order='F')
        stokesout = np.zeros((self.nstokes, npix), np.float32, order='F')
        cost = np.zeros(1, np.float64)
        jac = np.zeros((self.nstokes, g.numder, jp.size, npix), np.float32, order='F')
        buf = _lib.errbuf()
        code = self._L.at3d_levisapprox_gradient_jacobian(self._h, C.byref(r), C.byref(gd), vp(gradout), vp(cost),
                                                          vp(stokesout), int(jp.size), vp(jp), vp(jac), None, buf)
        _lib.check(code, buf)
        return gradout, cost, stokesout, jac

    def gradient(self, rays, pix, gradout=None, stokesout=None, cost=None, trace_cap=0, stream=None,
                 timing=False):
        """LEVISAPPROX_GRADIENT, MAKEJACOBIAN=.FALSE. (shdomsub4.f:288).  ``pix``: object with
        measurements [nstokes,npix] f32, uncertainties [nunc,nunc,npix] f64, rays_per_pixel [npix] i32,
        ray_weights [nrays] f64, stokes_weights [nstokes,npix] f64 (numpy or torch CUDA, same memory
        space as the rays).  Returns (gradout [maxpg,numder] f64, cost f64, stokesout [nstokes,npix])."""
        if self._grad is None:
            raise RuntimeError('attach_gradient() first')
        g = self._grad
        r, dev = self._rays_struct(rays.camx, rays.camy, rays.camz, rays.cammu, rays.camphi, getattr(rays, 'packs', None))
        gd = g.desc()
        npix = int(pix.rays_per_pixel.shape[0])
        gd.npix = npix
        # torch tensors hold the Fortran-ordered bytes, i.e. carry the reversed shape
        gd.nuncertainty = int(pix.uncertainties.shape[-1] if dev else pix.uncertainties.shape[0])
        gd.measurements = C.cast(vp(pix.measurements), type(gd.measurements))
        gd.uncertainties = C.cast(vp(pix.uncertainties), type(gd.uncertainties))
        gd.rays_per_pixel = C.cast(vp(pix.rays_per_pixel), type(gd.rays_per_pixel))
        gd.ray_weights = C.cast(vp(pix.ray_weights), type(gd.ray_weights))
        gd.stokes_weights = C.cast(vp(pix.stokes_weights), type(gd.stokes_weights))
        if dev:
            import torch
            d = rays.camx.device
            if gradout is None:
                gradout = torch.empty((g.numder, g.maxpg), dtype=torch.float64, device=d)
            if stokesout is None:
                stokesout = torch.empty((npix, self.nstokes), dtype=torch.float32, device=d)
            if cost is None:
                cost = torch.zeros(1, dtype=torch.float64, device=d)
        else:
            if gradout is None:
                gradout = np.zeros((g.maxpg, g.numder), np.float64, order='F')
            if stokesout is None:
                stokesout = np.zeros((self.nstokes, npix), np.float32, order='F')
            if cost is None:
                cost = np.zeros(1, np.float64)
        tr = None
        trace = None
        if trace_cap > 0:
            n = r.nrays
            trace = dict(cells=np.zeros((trace_cap, n), np.int32, order='F'),
                         ncells=np.zeros(n, np.int32), nsub=np.zeros(n, np.int32))
            tr = TraceC(trace_cap, vp(trace['cells']), vp(trace['ncells']), vp(trace['nsub']))
        ms = (C.c_double * 8)()
        buf = _lib.errbuf()
        code = self._L.at3d_levisapprox_gradient(self._h, C.byref(r), C.byref(gd), vp(gradout), vp(cost),
                                                 vp(stokesout), C.byref(tr) if tr is not None else None,
                                                 C.c_void_p(stream) if stream else None,
                                                 C.cast(ms, C.POINTER(C.c_double)) if timing else None, buf)
        _lib.check(code, buf)
        res = [gradout, cost, stokesout]
        if trace is not None:
            res.append(trace)
        if timing:
            res.append(list(ms))
        return tuple(res)
