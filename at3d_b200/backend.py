"""Array-level Python entry points of libat3d_b200.so for the helper routines of the path
(special-function tables and the per-evaluation gradient input preparation).

Same call shapes as the CPU oracle binding used by the parity tests (tests/oracle_lib.py), so that
``gradsetup.make_gradient_inputs`` can be driven by either; every function here runs CUDA kernels
behind the C-ABI of include/at3d_b200.h and raises ``At3dError`` on a non-zero IERR.
"""
import ctypes as C
import numpy as np
from . import _lib
from ._lib import vp
from .state import i32, f64


def _call(fn, *args):
    buf = _lib.errbuf()
    _lib.check(fn(*args, buf), buf)


def memory_reuse(on=True):
    """Switch the library's memory reuse (`at3d_set_memory_reuse`): freed device memory is kept for the next state / solver
    object instead of going back to the system.  For loops that build states per step (inversions).  Returns the previous
    setting; `trim_memory()` releases what is kept."""
    return bool(_lib.lib().at3d_set_memory_reuse(int(bool(on))))


def trim_memory():
    _lib.lib().at3d_trim_memory()


class reusing_memory:
    """Context manager: memory reuse on inside the block, previous setting (and a trim) afterwards."""

    def __enter__(self):
        self._was = memory_reuse(True)
        return self

    def __exit__(self, *exc):
        memory_reuse(self._was)
        if not self._was:
            trim_memory()
        return False


def ylmall(transpose, mu, phi, ml, mm, nstleg, nlm):
    """YLMALL (shdomsub2.f:4244): YR[nstleg,nlm]."""
    yr = np.zeros((nstleg, nlm), np.float32, order='F')
    _call(_lib.lib().at3d_ylmall, int(transpose), float(mu), float(phi), ml, mm, nstleg, vp(yr))
    return yr


def precompute_phase_check(legen, nscatangle, nstokes, ml, deltam=True, negcheck=True, grad=False):
    """PRECOMPUTE_PHASE_CHECK[_GRAD] (shdomsub4.f:2388,2493): PHASETAB[nstphase,numphase,nscatangle]."""
    legen = np.asfortranarray(legen, np.float32)
    nstleg, nlegp1, numphase = legen.shape
    nstphase = 1 if nstokes == 1 else 2
    tab = np.zeros((nstphase, numphase, nscatangle), np.float32, order='F')
    _call(_lib.lib().at3d_precompute_phase_check, nscatangle, numphase, nstphase, nstokes, ml, 0, nstleg,
          nlegp1 - 1, vp(legen), vp(tab), int(deltam), int(negcheck), int(grad))
    return tab


def finalize_scene(scene):
    """Fill YLMSUN and PHASETAB of a synthetic scene (INIT_SOLUTION / _precompute_phase equivalents)."""
    st = scene.state
    st.ylmsun = ylmall(True, np.float32(st.solarmu), np.float32(st.solaraz), st.ml, st.mm, st.nstleg, st.nlm)
    st.phasetab = precompute_phase_check(scene.pg.legenp, st.nscatangle, st.nstokes, st.ml, bool(st.deltam))
    return scene


def prepare_deriv_interps(state, pg, grad):
    """PREPARE_DERIV_INTERPS (shdomsub4.f:2917): (optinterpwt, interpptr, dalbm, dextm, dfj)."""
    st = state.copy().normalize()
    d = st.desc()
    gd = grad.desc()
    npts = st.npts
    optw = np.zeros((8, npts), np.float32, order='F')
    iptr = np.zeros((8, npts), np.int32, order='F')
    dalbm = np.zeros((8, npts, grad.numder), np.float32, order='F')
    dextm = np.zeros((pg.maxpg, grad.numder), np.float32, order='F')
    dfj = np.zeros((8, npts, grad.numder), np.float32, order='F')
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    _call(_lib.lib().at3d_prepare_deriv_interps, C.byref(d), pg.npx, pg.npy, pg.npz, pg.maxpg, pg.delx, pg.dely,
          pg.xstart, pg.ystart, vp(zl), C.byref(gd), vp(optw), vp(iptr), vp(dalbm), vp(dextm), vp(dfj))
    return optw, iptr, dalbm, dextm, dfj


_BEAM_NAMES = ['cx', 'cy', 'cz', 'cxinv', 'cyinv', 'czinv', 'epss', 'epsz', 'xdomain', 'ydomain',
               'uniformzlev', 'delxd', 'delyd']


def make_direct(state, pg, nzckd=0, zckd=None, gasabs=None):
    """MAKE_DIRECT (shdomsub2.f:393): (dirflux[npts], extdirp[maxpg], dict of beam constants)."""
    st = state
    extdirp = np.zeros(pg.maxpg, np.float32)
    dirflux = np.zeros(st.npts, np.float32)
    od = (f64 * 13)()
    oi = (i32 * 5)()
    gp = np.asfortranarray(st.gridpos, np.float32)
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    _call(_lib.lib().at3d_make_direct, st.npts, st.bcflag, st.ipflag, int(st.deltam), st.ml, st.nstleg, pg.nlegp,
          st.solarflux, st.solarmu, st.solaraz, vp(gp), pg.npx, pg.npy, pg.npz, pg.delx, pg.dely, pg.xstart,
          pg.ystart, vp(zl), vp(pg.extinctp), vp(pg.albedop), vp(pg.legenp), pg.numphase, vp(pg.iphasep),
          vp(pg.phasewtp), pg.maxnmicro, pg.npart, nzckd, vp(zckd), vp(gasabs), vp(extdirp), vp(dirflux), od, oi)
    c = dict(zip(_BEAM_NAMES, list(od)))
    c.update(ipdirect=oi[0], di=oi[1], dj=oi[2], dk=oi[3], longest_path_pts=max(oi[4], 1))
    return dirflux, extdirp, c


def make_direct_derivative(state, pg, c):
    """MAKE_DIRECT_DERIVATIVE (shdomsub5.f:1553): (dpath, dptr)[longest_path_pts,npts]."""
    npts = state.npts
    lpp = c['longest_path_pts']
    dpath = np.zeros((lpp, npts), np.float32, order='F')
    dptr = np.zeros((lpp, npts), np.int32, order='F')
    gp = np.asfortranarray(state.gridpos, np.float32)
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    _call(_lib.lib().at3d_make_direct_derivative, npts, state.bcflag, pg.npx, pg.npy, pg.npz, pg.delx, pg.dely,
          pg.xstart, pg.ystart, vp(gp), vp(zl), c['ipdirect'], c['di'], c['dj'], c['dk'], c['cx'], c['cy'],
          c['cz'], c['cxinv'], c['cyinv'], c['czinv'], c['epss'], c['epsz'], c['xdomain'], c['ydomain'],
          c['uniformzlev'], c['delxd'], c['delyd'], vp(dpath), vp(dptr), lpp)
    return dpath, dptr


def transfer_pa_to_grid(pg, gridpos, npts, ml, deltam, phasemax=0.999):
    """TRANSFER_PA_TO_GRID on the GPU (C `at3d_transfer_pa_to_grid`): the dict of `medium.transfer_pa_to_grid`
    (extinct, albedo, total_ext, legen, iphase, phaseinterpwt, nleg, extmin, scatmin).  The Legendre table (a few kB)
    is scaled on the host; everything per grid point runs in one kernel."""
    npart, mnm = pg.npart, pg.maxnmicro
    nq = 8 * mnm
    zl32 = np.asarray(pg.zlevels, np.float32)
    extmin = float(np.float32(1.0e-5) / ((zl32[-1] - zl32[0]) / np.float32(pg.npz)))     # REAL arithmetic (shdom90.f90:73)
    nleg = ml + 1 if deltam else ml
    nleg = min(max(nleg, 1), pg.nlegp) if not deltam else nleg
    if pg.nlegp < nleg:
        raise ValueError('property Legendre table shorter than ML+1')
    l = np.arange(nleg + 1, dtype=np.float32)
    legen = np.asfortranarray(pg.legenp[:, :nleg + 1, :] / (2 * l + 1)[None, :, None], dtype=np.float32)
    ftab = np.ascontiguousarray(legen[0, ml + 1, :]) if deltam else None
    extinct = np.zeros((npts, npart), np.float32, order='F')
    albedo = np.zeros((npts, npart), np.float32, order='F')
    total_ext = np.zeros(npts, np.float32)
    iphase = np.ones((nq, npts, npart), np.int32, order='F')
    pwt = np.zeros((nq, npts, npart), np.float32, order='F')
    gp = np.asfortranarray(np.asarray(gridpos, np.float32)[:, :npts])
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    _call(_lib.lib().at3d_transfer_pa_to_grid, npts, vp(gp), pg.npx, pg.npy, pg.npz, pg.delx, pg.dely, pg.xstart, pg.ystart,
          vp(zl), npart, mnm, vp(pg.extinctp), vp(pg.albedop), vp(pg.iphasep), vp(pg.phasewtp), pg.numphase, vp(ftab), ml,
          int(deltam), phasemax, vp(extinct), vp(albedo), vp(total_ext), vp(iphase), vp(pwt))
    if deltam:
        legen[0, :ml + 1, :] -= ftab[None, :]
        if pg.nstleg > 1:
            legen[1:4, :ml + 1, :] -= ftab[None, None, :]
    return dict(extinct=extinct, albedo=albedo, total_ext=total_ext, legen=legen, iphase=iphase, phaseinterpwt=pwt,
                nleg=nleg, extmin=extmin, scatmin=float(np.float32(0.1)) * extmin)


def compute_source(state, shptr, source, oshptr, delsource, fixsh=False, shacc=0.0, maxiv=None, first=False,
                   accelflag=True, newmethod=True, timing=False):
    """COMPUTE_SOURCE (shdomsub1.f:967).  Returns (ierr, shptr, source, oshptr, delsource,
    [deljdot, deljold, deljnew, jnorm]) (+ kernel ms when ``timing``); ierr=2 is "out of SH memory"."""
    st = state.copy().normalize()
    d = st.desc()
    shptr = np.array(shptr, np.int32); oshptr = np.array(oshptr, np.int32)
    source = np.array(source, np.float32, order='F'); delsource = np.array(delsource, np.float32, order='F')
    if maxiv is None:
        maxiv = source.shape[1]
    o = [C.c_float(0), C.c_float(0), C.c_float(0), C.c_float(0)]
    ms = C.c_double(0.0)
    buf = _lib.errbuf()
    code = _lib.lib().at3d_compute_source(C.byref(d), int(fixsh), shacc, int(maxiv), int(first), int(accelflag),
                                          int(newmethod), vp(shptr), vp(source), vp(oshptr), vp(delsource),
                                          C.byref(o[0]), C.byref(o[1]), C.byref(o[2]), C.byref(o[3]),
                                          C.byref(ms), buf)
    if code not in (0, 2):
        _lib.check(code, buf)
    res = (code, shptr, source, oshptr, delsource, [x.value for x in o])
    return res + (ms.value,) if timing else res


class DeviceSourceState:
    """The arrays COMPUTE_SOURCE reads, resident on the GPU as torch tensors with the reference's (Fortran) layout: what a
    GPU-resident solver holds between iterations (RTE._extinct, _albedo, _legen, _iphase, ... at3d/solver.py:1609-1760).
    ``radiance`` / ``rshptr`` are the current SH radiance; SOURCE / SHPTR / DELSOURCE / OSHPTR are passed per call."""

    FIELDS = ('extinct', 'albedo', 'total_ext', 'legen', 'iphase', 'phaseinterpwt', 'dirflux', 'rshptr', 'radiance',
              'ylmsun', 'planck')

    def __init__(self, state=None, device='cuda', **arrays):
        import torch
        self._t = {}
        if state is not None:
            st = state
            self.meta = dict(npts=st.npts, nstokes=st.nstokes, nstleg=st.nstleg, nlm=st.nlm, ml=st.ml, mm=st.mm, nleg=st.nleg,
                             npart=st.npart, maxnmicro=st.maxnmicro, numphase=st.numphase, deltam=int(st.deltam),
                             interp_new=int(st.interp_new), srctype=st.srctype, phasemax=st.phasemax, solarmu=st.solarmu)
            for k in self.FIELDS:
                a = getattr(st, k, None)
                if a is None:
                    continue
                # bytes in Fortran order: ravel(order='K') of the F-ordered host array
                self._t[k] = torch.from_numpy(np.ascontiguousarray(np.asarray(a).ravel(order='F'))).to(device)
        else:
            self.meta = arrays.pop('meta')
            self._t = dict(arrays)

    def tensor(self, k):
        return self._t.get(k)

    def desc(self):
        d = _lib.CsDeviceDesc()
        for k, v in self.meta.items():
            setattr(d, k, v.encode() if k == 'srctype' else v)
        for k in self.FIELDS:
            t = self._t.get(k)
            setattr(d, k, t.data_ptr() if t is not None else None)
        return d


def compute_source_device(dstate, shptr_old, source_old, oshptr_old, delsource, shptr_new, source_new, fixsh=False,
                          shacc=0.0, maxiv=None, first=False, accelflag=True, properties_changed=False, timing=False,
                          delsource_new=None):
    """COMPUTE_SOURCE (shdomsub1.f:967) on device-resident torch tensors (C: at3d_compute_source_device).  SOURCE and SHPTR
    are double-buffered by the caller (``*_old`` in, ``*_new`` out); DELSOURCE is updated in place at the old SHPTR offsets,
    or written to ``delsource_new`` when given -- a separate buffer lets the adaptive truncation run in one pass
    (cs_adapt_kernel) instead of two.  Returns (ierr, total_new, [deljdot, deljold, deljnew, jnorm]) (+ kernel ms when
    ``timing``)."""
    nst = dstate.meta['nstokes']
    cap = source_new.numel() // nst
    if maxiv is None:
        maxiv = cap
    norms = np.zeros(4, np.float32)
    tot = C.c_int32(0)
    ms = C.c_double(0.0)
    buf = _lib.errbuf()
    d = dstate.desc()
    code = _lib.lib().at3d_compute_source_device(
        C.byref(d), int(fixsh), shacc, int(maxiv), int(first), int(accelflag), vp(shptr_old), vp(source_old),
        vp(oshptr_old), vp(delsource), vp(delsource if delsource_new is None else delsource_new), vp(shptr_new),
        vp(source_new), int(cap), int(properties_changed),
        vp(norms), C.byref(tot), C.byref(ms), buf)
    if code not in (0, 2):
        _lib.check(code, buf)
    res = (code, tot.value, [float(x) for x in norms])
    return res + (ms.value,) if timing else res


def _angles(state, wtmu):
    return (np.ascontiguousarray(state.nphi0, np.int32), np.ascontiguousarray(state.mu, np.float32),
            np.asfortranarray(state.phi, np.float32), np.ascontiguousarray(wtmu, np.float32))


def sh_to_do(state, wtmu, shptr, indata, timing=False):
    """SH_TO_DO (shdomsub1.f:2789) for all ordinates: DOFIELD[npts, nstokes, nang] (Fortran order)."""
    nphi0, mu, phi, wt = _angles(state, wtmu)
    shptr = np.ascontiguousarray(shptr, np.int32)
    indata = np.asfortranarray(indata, np.float32)
    out = np.zeros((state.npts, state.nstokes, int(nphi0.sum())), np.float32, order='F')
    ms = C.c_double(0.0)
    buf = _lib.errbuf()
    _lib.check(_lib.lib().at3d_sh_to_do(state.npts, state.nstokes, state.nstleg, state.ml, state.mm, state.nlm,
                                        state.nmu, state.nphi0max, vp(nphi0), vp(mu), vp(phi), vp(wt), vp(shptr),
                                        vp(indata), vp(out), C.byref(ms), buf), buf)
    return (out, ms.value) if timing else out


def do_to_sh(state, wtmu, rshptr, dofield, timing=False):
    """DO_TO_SH (shdomsub1.f:3041) summed over all zenith angles: OUTDATA[nstokes, rshptr[npts]]."""
    nphi0, mu, phi, wt = _angles(state, wtmu)
    rshptr = np.ascontiguousarray(rshptr, np.int32)
    dofield = np.asfortranarray(dofield, np.float32)
    out = np.zeros((state.nstokes, max(int(rshptr[state.npts]), 1)), np.float32, order='F')
    ms = C.c_double(0.0)
    buf = _lib.errbuf()
    _lib.check(_lib.lib().at3d_do_to_sh(state.npts, state.nstokes, state.nstleg, state.ml, state.mm, state.nlm,
                                        state.nmu, state.nphi0max, vp(nphi0), vp(mu), vp(phi), vp(wt), vp(rshptr),
                                        vp(dofield), vp(out), C.byref(ms), buf), buf)
    return (out, ms.value) if timing else out


def average_subpixel_rays(weighted_stokes, pixel_index, npixels):
    """average_subpixel_rays (src/util.f90:484)."""
    ws = np.asfortranarray(weighted_stokes, np.float32)
    nstokes, nrays = ws.shape
    pi = np.ascontiguousarray(pixel_index, np.int32)
    out = np.zeros((nstokes, npixels), np.float32, order='F')
    _call(_lib.lib().at3d_average_subpixel_rays, npixels, nrays, nstokes, vp(ws), vp(pi), vp(out))
    return out


def update_costfunction(stokesout, raygrad_pixel, gradout, cost, uncertainties, costfunc, measurement):
    """UPDATE_COSTFUNCTION (shdomsub4.f:13): returns (gradout, cost)."""
    nstokes = len(stokesout)
    rg = np.asfortranarray(raygrad_pixel, np.float64)
    _, maxpg, numder = rg.shape
    gradout = np.array(gradout, np.float64, order='F')
    cost = np.atleast_1d(np.array(cost, np.float64))
    unc = np.asfortranarray(uncertainties, np.float64)
    so = np.ascontiguousarray(stokesout, np.float64); me = np.ascontiguousarray(measurement, np.float64)
    _call(_lib.lib().at3d_update_costfunction, vp(so), vp(rg), vp(gradout), vp(cost), vp(unc),
          1 if costfunc == 'LL' else 0, nstokes, maxpg, numder, vp(me), unc.shape[0])
    return gradout, cost
