"""Drop-in for the hot-path entry points of the reference's f2py module ``at3d.core``.

Same keyword names (the Fortran dummy-argument names), same return-tuple orders and the same
``(ierr, errmsg)`` convention as the call sites in the reference:

=============================  ==========================================  ===========================================
function                       reference call site                         returns
=============================  ==========================================  ===========================================
``render``                     at3d/solver.py:681-759                      ``(bcrad, stokes, ierr, errmsg)``
``levisapprox_gradient``       at3d/gradient.py:262-398                    ``(gradient, loss, images, jacobian, ierr, errmsg)``
``precompute_phase_check``     at3d/solver.py:2805-2823                    ``(phasetab, ierr, errmsg)``
``precompute_phase_check_grad``at3d/solver.py:1428-1442                    ``(dphasetab, ierr, errmsg)``
``prepare_deriv_interps``      at3d/solver.py:1474-1515                    ``(optinterpwt, interpptr, ierr, errmsg, dalbm, dextm, dfj)``
``make_direct_derivative``     at3d/solver.py:1280-1311                    ``(dpath, dptr, ierr, errmsg)``
``average_subpixel_rays``      at3d/containers.py:642-647                  ``observables``
``update_costfunction``        reference tests/test_derivatives.py:84-93   ``(gradout, cost, ierr, errmsg)``
``ylmall``                     at3d/solver.py (INIT_SOLUTION)              ``yr``
=============================  ==========================================  ===========================================

Everything runs on the GPU through libat3d_b200.so (include/at3d_b200.h); there is no CPU fallback.
Unlike f2py, the solved state is not re-marshalled on every call: ``render`` / ``levisapprox_gradient``
keep the state resident in HBM (``DeviceState``) and re-use it while the caller passes the same arrays
(the reference calls these once per sensor chunk / per L-BFGS evaluation with unchanged solver arrays).
Configurations the GPU path does not implement (gradients with a thermal source or a non-Lambertian surface, band-integrated
Planck units) return ``ierr=3``.
"""
import numpy as np
from . import backend as B
from ._lib import At3dError
from .device import DeviceState
from .gradsetup import PixelData
from .state import ShdomState, GradInputs, Rays

_CACHE = []          # [(key, DeviceState)] most recent first
_CACHE_SIZE = 4


def _ch(v):
    if isinstance(v, (bytes, np.bytes_)):
        v = v.decode()
    return str(np.asarray(v).item() if isinstance(v, np.ndarray) else v)


def _errmsg(msg):
    return msg.encode()[:600].ljust(600)


def _state_from_kwargs(kw, with_radiance):
    sfctype = _ch(kw['sfctype'])
    interp = _ch(kw['interpmethod'])
    npts = int(kw['npts'])
    nstokes = int(kw['nstokes'])
    shptr = np.asarray(kw['shptr'], np.int32)[:npts + 1]
    st = ShdomState(
        nstokes=nstokes, nstleg=int(kw['nstleg']), nx=int(kw['nx']), ny=int(kw['ny']), nz=int(kw['nz']),
        npts=npts, ncells=int(kw['ncells']), ml=int(kw['ml']), mm=int(kw['mm']), nlm=int(kw['nlm']),
        nleg=np.asarray(kw['legen']).shape[1] - 1, numphase=int(kw['numphase']), npart=int(kw['npart']),
        maxnmicro=int(kw['maxnmicro']), bcflag=int(kw['bcflag']), ipflag=int(kw['ipflag']), nmu=int(kw['nmu']),
        nphi0max=int(kw['nphi0max']), nang=int(kw.get('nang', 0)), maxnbc=int(kw['maxnbc']),
        ntoppts=int(kw['ntoppts']), nbotpts=int(kw['nbotpts']), nsfcpar=int(kw['nsfcpar']),
        nscatangle=int(kw['nscatangle']), nstphase=int(kw['nstphase']), deltam=int(bool(kw['deltam'])),
        srctype=_ch(kw['srctype']), units=_ch(kw['units']), sfctype0=sfctype[0], sfctype1=sfctype[1],
        interp_new=int(interp[1] == 'N'), solarmu=float(kw['solarmu']), solaraz=float(kw['solaraz']),
        solarflux=float(kw.get('solarflux', 1.0)), wavelen=float(kw['wavelen']), gndtemp=float(kw['gndtemp']),
        gndalbedo=float(kw['gndalbedo']), phasemax=float(kw['phasemax']), waveno0=0.0, waveno1=0.0,
        tautol=float(kw['tautol']), transcut=float(kw['transcut']),
        gridptr=np.asarray(kw['gridptr'])[:, :int(kw['ncells'])], neighptr=np.asarray(kw['neighptr'])[:, :int(kw['ncells'])],
        treeptr=np.asarray(kw['treeptr'])[:, :int(kw['ncells'])], cellflags=np.asarray(kw['cellflags'])[:int(kw['ncells'])],
        xgrid=kw['xgrid'], ygrid=kw['ygrid'], zgrid=kw['zgrid'], gridpos=np.asarray(kw['gridpos'])[:, :npts],
        extinct=np.asarray(kw['extinct'])[:npts], albedo=np.asarray(kw['albedo'])[:npts],
        total_ext=np.asarray(kw['total_ext'])[:npts], legen=kw['legen'],
        iphase=np.asarray(kw['iphase'])[:, :npts], phaseinterpwt=np.asarray(kw['phaseinterpwt'])[:, :npts],
        dirflux=np.asarray(kw['dirflux'])[:npts], fluxes=np.asarray(kw['fluxes'])[:, :npts], shptr=shptr,
        source=np.asarray(kw['source'])[:, :max(int(shptr[npts]), 1)], ylmsun=kw['ylmsun'], phasetab=kw['phasetab'],
        nphi0=kw['nphi0'], mu=kw['mu'], phi=np.asarray(kw['phi']).reshape(int(kw['nmu']), -1),
        wtdo=np.asarray(kw['wtdo']).reshape(int(kw['nmu']), -1), skyrad=_skyrad(kw), bcptr=kw['bcptr'],
        bcrad=np.asarray(kw['bcrad'])[:, :_nbcrad(kw)].copy(order='F'),
        sfcgridparms=kw['sfcgridparms'], sfcgridrad=_sfcgridrad(kw))
    if kw.get('temp') is not None and _ch(kw['srctype']) != 'S':
        st.temp = np.asarray(kw['temp'], np.float32)[:npts]       # TEMP of LEVISAPPROX_GRADIENT (thermal component)
    if with_radiance:
        rshptr = np.asarray(kw['rshptr'], np.int32)[:npts + 2]
        st.rshptr = rshptr
        st.radiance = np.asarray(kw['radiance'])[:, :max(int(rshptr[npts]), 1)]
    return st.normalize()


def _nbcrad(kw):
    """Columns of BCRAD in use: top + bottom points, plus the stored downwelling radiance of the NANG/2 downward
    ordinates at every bottom point for general BRDF surfaces (shdomsub1.f:2139-2145)."""
    lamb = _ch(kw['sfctype'])[1] == 'L'
    return int(kw['ntoppts']) + int(kw['nbotpts']) * (1 if lamb else 1 + int(kw['nang']) // 2)


def _sfcgridrad(kw):
    s = kw.get('sfcgridrad')
    if s is None or 'nang' not in kw:
        return None
    s = np.asarray(s, np.float32)
    nrow, nbot = int(kw['nang']) // 2 + 1, int(kw['nbotpts'])
    if s.size < nrow * nbot or not s.any():
        return None
    return s.reshape((nrow, -1), order='F')[:, :nbot]


def _skyrad(kw):
    """SKYRAD as the Fortran sees it: a scalar isotropic radiance or an array [nstokes,nmu/2,nphi0max]."""
    nst, nmu, nph = int(kw['nstokes']), int(kw['nmu']), int(kw['nphi0max'])
    s = np.asarray(kw['skyrad'], np.float32)
    if s.size == 1:
        out = np.zeros((nst, nmu // 2, nph), np.float32, order='F')
        out[0] = float(s.ravel()[0])
        return out
    return s.reshape((nst, nmu // 2, nph), order='F')


def _cache_key(kw, with_radiance):
    names = ['source', 'shptr', 'gridptr', 'gridpos', 'extinct', 'albedo', 'total_ext', 'dirflux', 'legen', 'phasetab',
             'iphase', 'phaseinterpwt', 'fluxes', 'ylmsun']
    if with_radiance:
        names += ['radiance', 'rshptr']
    if _ch(kw['sfctype'])[1] != 'L':
        names += ['bcrad', 'sfcgridparms']
    key = []
    for n in names:
        a = np.asarray(kw[n])
        key.append((a.__array_interface__['data'][0], a.shape, a.dtype.str,
                    float(a.ravel()[0]) if a.size else 0.0, float(a.ravel()[-1]) if a.size else 0.0))
    key.append(tuple(_ch(kw[n]) for n in ('srctype', 'sfctype', 'interpmethod')))
    key.append(tuple(float(kw[n]) for n in ('solarmu', 'solaraz', 'gndalbedo', 'tautol', 'transcut', 'phasemax')))
    key.append(with_radiance)
    return tuple(key)


def _device_state(kw, with_radiance):
    key = _cache_key(kw, with_radiance)
    for i, (k, dev) in enumerate(_CACHE):
        if k == key or (with_radiance is False and k[:-1] == key[:-1]):
            _CACHE.insert(0, _CACHE.pop(i))
            return dev
    dev = DeviceState(_state_from_kwargs(kw, with_radiance))
    _CACHE.insert(0, (key, dev))
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.pop()[1].close()
    return dev


def clear_cache():
    while _CACHE:
        _CACHE.pop()[1].close()


def render(**kw):
    """RENDER (src/polarized/shdomsub4.f:93).  Returns ``(bcrad, stokes, ierr, errmsg)``."""
    bcrad = np.array(kw['bcrad'], np.float32, order='F')
    nstokes, npix = int(kw['nstokes']), int(kw['npix'])
    stokes = np.zeros((nstokes, npix), np.float32, order='F')
    try:
        dev = _device_state(kw, False)
        rays = Rays(np.asarray(kw['camx'])[:npix], np.asarray(kw['camy'])[:npix], np.asarray(kw['camz'])[:npix],
                    np.asarray(kw['cammu'])[:npix], np.asarray(kw['camphi'])[:npix])
        stokes = dev.render(rays, correctinterpolate=bool(kw.get('correctinterpolate', True)),
                            singlescatter=bool(kw.get('singlescatter', False)),
                            nosurface=bool(kw.get('nosurface', False)))
        nb = int(kw['ntoppts']) + int(kw['nbotpts'])
        bcrad[:, :nb] = dev.bcrad()[:, :nb]
    except At3dError as e:
        return bcrad, stokes, e.code, _errmsg(e.msg)
    return bcrad, stokes, 0, _errmsg('')


def _grad_from_kwargs(kw, st):
    return GradInputs(
        npix=0, maxpg=int(kw['maxpg']), numder=int(kw['numder']), dnumphase=int(kw['dnumphase']),
        deriv_maxnmicro=int(kw['deriv_maxnmicro']), longest_path_pts=int(kw['longest_path_pts']),
        nuncertainty=int(kw['nuncertainty']), maxsubgridints=int(kw['maxsubgridints']),
        exact_single_scatter=int(bool(kw['exact_single_scatter'])), singlescatter=int(bool(kw['singlescatter'])),
        costfunc_ll=1 if _ch(kw['costfunc']) == 'LL' else 0, extmin=float(kw['extmin']), scatmin=float(kw['scatmin']),
        partder=kw['partder'], doexact=kw['doexact'], dext=kw['dext'], dalb=kw['dalb'], dextm=kw['dextm'],
        dalbm=kw['dalbm'], dfj=kw['dfj'], optinterpwt=kw['optinterpwt'], interpptr=kw['interpptr'], dleg=kw['dleg'],
        dphasetab=kw['dphasetab'], diphasep=kw['diphasep'], dphasewtp=kw['dphasewtp'], iphasep=kw['iphasep'],
        phasewtp=kw['phasewtp'], extinctp=kw['extinctp'], albedop=kw['albedop'], dpath=kw['dpath'], dptr=kw['dptr'],
        dtemp=kw.get('dtemp')).normalize()


def levisapprox_gradient(**kw):
    """LEVISAPPROX_GRADIENT (shdomsub4.f:288), default adjoint path.
    Returns ``(gradient[maxpg,numder,1], loss[1], images[nstokes,npixels], jacobian, ierr, errmsg)``."""
    nstokes, maxpg, numder = int(kw['nstokes']), int(kw['maxpg']), int(kw['numder'])
    rpp = np.ascontiguousarray(kw['rays_per_pixel'], np.int32)
    npixels = rpp.size
    gradient = np.zeros((maxpg, numder, 1), np.float64, order='F')
    loss = np.zeros(1, np.float64)
    images = np.zeros((nstokes, npixels), np.float32, order='F')
    jacobian = kw.get('jacobian')
    makejac = bool(kw.get('makejacobian', False))
    try:
        dev = _device_state(kw, True)
        dev.attach_gradient(_grad_from_kwargs(kw, dev.state))
        npix = int(kw['npix'])
        rays = Rays(np.asarray(kw['camx'])[:npix], np.asarray(kw['camy'])[:npix], np.asarray(kw['camz'])[:npix],
                    np.asarray(kw['cammu'])[:npix], np.asarray(kw['camphi'])[:npix])
        pix = PixelData(np.asarray(kw['measurements'])[:nstokes], kw['uncertainties'], rpp, kw['ray_weights'],
                        np.asarray(kw['stokes_weights'])[:nstokes])
        if makejac:
            # single-sweep semantics (shdomsub4.f:536-631): gradient, cost, images and the per-pixel Jacobian
            g, c, so, jacobian = dev.gradient_jacobian(rays, pix, np.asarray(kw['jacobianptr']).ravel()
                                                       [:int(kw.get('num_jacobian_pts', np.size(kw['jacobianptr'])))])
        else:
            g, c, so = dev.gradient(rays, pix)
        gradient[:, :, 0] = g
        loss[0] = c[0]
        images = so
    except At3dError as e:
        return gradient, loss, images, jacobian, e.code, _errmsg(e.msg)
    return gradient, loss, images, jacobian, 0, _errmsg('')


def _wrap(fn, nout, *a, **k):
    try:
        out = fn(*a, **k)
        return (out if isinstance(out, tuple) else (out,)) + (0, _errmsg(''))
    except At3dError as e:
        return (None,) * nout + (e.code, _errmsg(e.msg))


def precompute_phase_check(negcheck, nstphase, nstleg, nscatangle, nstokes, numphase, ml, nlm, nleg, legen, deltam):
    return _wrap(B.precompute_phase_check, 1, legen, nscatangle, nstokes, ml, deltam=bool(deltam), negcheck=bool(negcheck))


def precompute_phase_check_grad(negcheck, nstphase, nstleg, nscatangle, nstokes, dnumphase, ml, nlm, nleg, dleg, deltam):
    return _wrap(B.precompute_phase_check, 1, dleg, nscatangle, nstokes, ml, deltam=bool(deltam), negcheck=bool(negcheck),
                 grad=True)


class _PG:
    pass


def _pg_from(kw):
    pg = _PG()
    pg.npx, pg.npy, pg.npz = int(kw['npx']), int(kw['npy']), int(kw['npz'])
    pg.maxpg = pg.npx * pg.npy * pg.npz
    pg.delx, pg.dely = np.float32(kw['delx']), np.float32(kw['dely'])
    pg.xstart, pg.ystart = np.float32(kw['xstart']), np.float32(kw['ystart'])
    pg.zlevels = np.ascontiguousarray(kw['zlevels'], np.float32)[:pg.npz]
    return pg


def prepare_deriv_interps(**kw):
    """PREPARE_DERIV_INTERPS (shdomsub4.f:2917).  Returns
    ``(optinterpwt, interpptr, ierr, errmsg, dalbm, dextm, dfj)`` (the order at at3d/solver.py:1474-1476)."""
    npts = int(kw['npts'])
    interp = _ch(kw['interpmethod'])
    st = ShdomState(npts=npts, nstleg=int(kw['nstleg']), nleg=int(kw['nleg']), ml=int(kw['ml']), npart=int(kw['npart']),
                    maxnmicro=int(kw['maxnmicro']), numphase=int(kw['numphase']), deltam=int(bool(kw['deltam'])),
                    interp_new=int(interp[1] == 'N'), phasemax=float(kw['phasemax']), gridpos=np.asarray(kw['gridpos'])[:, :npts],
                    legen=kw['legen'], albedo=np.asarray(kw['albedo'])[:npts], iphase=np.asarray(kw['iphase'])[:, :npts],
                    phaseinterpwt=np.asarray(kw['phaseinterpwt'])[:, :npts])
    gi = GradInputs(maxpg=int(kw['maxpg']), numder=int(kw['numder']), dnumphase=int(kw['dnumphase']),
                    deriv_maxnmicro=int(kw['deriv_maxnmicro']), partder=kw['partder'], doexact=kw['doexact'],
                    dext=kw['dext'], dalb=kw['dalb'], dleg=kw['dleg'], diphasep=kw['diphasep'], dphasewtp=kw['dphasewtp'],
                    iphasep=kw['iphasep'], phasewtp=kw['phasewtp'], extinctp=kw['extinctp'], albedop=kw['albedop'])
    try:
        optw, iptr, dalbm, dextm, dfj = B.prepare_deriv_interps(st, _pg_from(kw), gi.normalize())
    except At3dError as e:
        return None, None, e.code, _errmsg(e.msg), None, None, None
    return optw, iptr, 0, _errmsg(''), dalbm, dextm, dfj


def make_direct_derivative(**kw):
    """MAKE_DIRECT_DERIVATIVE (src/shdomsub5.f:1553).  Returns ``(dpath, dptr, ierr, errmsg)``."""
    npts = int(kw['npts'])
    st = ShdomState(npts=npts, bcflag=int(kw['bcflag']), gridpos=np.asarray(kw['gridpos'])[:, :npts])
    c = {k: float(kw[k]) for k in ('cx', 'cy', 'cz', 'cxinv', 'cyinv', 'czinv', 'epss', 'epsz', 'xdomain', 'ydomain',
                                   'uniformzlev', 'delxd', 'delyd')}
    c.update({k: int(kw[k]) for k in ('ipdirect', 'di', 'dj', 'dk', 'longest_path_pts')})
    return _wrap(B.make_direct_derivative, 2, st.normalize(), _pg_from(kw), c)


def average_subpixel_rays(pixel_index, nstokes, weighted_stokes, nrays, npixels):
    """average_subpixel_rays (src/util.f90:484), keyword order of at3d/containers.py:642-647."""
    return B.average_subpixel_rays(np.asarray(weighted_stokes)[:nstokes, :nrays], np.asarray(pixel_index)[:nrays], npixels)


def update_costfunction(cost, gradout, stokesout, measurement, raygrad_pixel, uncertainties, costfunc):
    """UPDATE_COSTFUNCTION (shdomsub4.f:13).  Returns ``(gradout, cost, ierr, errmsg)``."""
    shape = np.shape(gradout)
    rg = np.asarray(raygrad_pixel, np.float64)
    try:
        g, c = B.update_costfunction(stokesout, rg.reshape(rg.shape[0], rg.shape[1], -1, order='F'),
                                     np.asarray(gradout, np.float64).reshape(rg.shape[1], -1, order='F'), cost,
                                     uncertainties, _ch(costfunc), measurement)
    except At3dError as e:
        return gradout, cost, e.code, _errmsg(e.msg)
    return g.reshape(shape, order='F'), float(c[0]), 0, _errmsg('')


def ylmall(transpose, mu, phi, ml, mm, nstleg, nlm=None):
    """YLMALL (shdomsub2.f:4244): YR[nstleg,nlm]."""
    if nlm is None:
        nlm = sum(2 * min(l, mm) + 1 for l in range(ml + 1))
    return B.ylmall(bool(transpose), mu, phi, ml, mm, nstleg, nlm)
