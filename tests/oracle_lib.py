"""ctypes binding of the CPU parity oracle (oracle/libshdom_oracle.so).  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess
import numpy as np
from at3d_b200.state import STATE_FIELDS, GRAD_FIELDS, make_struct, i32, f32, f64, P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, 'oracle')
LIB = os.path.join(ORACLE_DIR, 'libshdom_oracle.so')

OracleState = make_struct('OracleState', STATE_FIELDS)      # identical field order to oracle_state
OracleGrad = make_struct('OracleGrad', GRAD_FIELDS)         # identical field order to oracle_grad_in


class OracleRays(C.Structure):
    _fields_ = [('nrays', i32), ('camx', C.c_void_p), ('camy', C.c_void_p), ('camz', C.c_void_p),
                ('cammu', C.c_void_p), ('camphi', C.c_void_p)]


class OracleTrace(C.Structure):
    _fields_ = [('max_per_ray', i32), ('cells', C.c_void_p), ('ncells', C.c_void_p), ('nsub', C.c_void_p)]


_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_ylmall.argtypes = [i32, f32, f32, i32, i32, i32, C.c_void_p]
        _lib.oracle_ylmall.restype = None
        _lib.oracle_make_direct_derivative.argtypes = [i32, i32, i32, i32, i32, f32, f32, f32, f32,
                                                       C.c_void_p, C.c_void_p, i32, i32, i32, i32] + \
            [f64] * 13 + [C.c_void_p, C.c_void_p, i32, C.c_char_p]
        _lib.oracle_make_direct.argtypes = [i32, i32, i32, i32, i32, i32, i32, f32, f32, f32, C.c_void_p,
                                            i32, i32, i32, f32, f32, f32, f32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, i32,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_char_p]
        _lib.oracle_prepare_deriv_interps.argtypes = [P(OracleState), i32, i32, i32, i32, f32, f32, f32, f32,
                                                      C.c_void_p, P(OracleGrad), C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        _lib.oracle_compute_source.argtypes = [P(OracleState), i32, f32, i32, i32, i32, i32, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, P(f32), P(f32), P(f32),
                                               P(f32), C.c_char_p]
        _lib.oracle_render.argtypes = [P(OracleState), P(OracleRays), C.c_void_p, i32, i32, i32,
                                       P(OracleTrace), i32, C.c_char_p]
        _lib.oracle_levisapprox_gradient.argtypes = [P(OracleState), P(OracleRays), P(OracleGrad), C.c_void_p,
                                                     C.c_void_p, C.c_void_p, P(OracleTrace), i32, C.c_char_p]
        _lib.oracle_levisapprox_jacobian.argtypes = [P(OracleState), P(OracleRays), P(OracleGrad), i32, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        _lib.oracle_precompute_phase_check.argtypes = [i32, i32, i32, i32, i32, i32, i32, C.c_void_p,
                                                       C.c_void_p, i32, i32, C.c_char_p]
        _lib.oracle_precompute_phase_check_grad.argtypes = _lib.oracle_precompute_phase_check.argtypes
        _lib.oracle_update_costfunction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    i32, i32, i32, i32, C.c_void_p, i32]
        _lib.oracle_average_subpixel_rays.argtypes = [i32, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_average_subpixel_rays.restype = None
    return _lib


class OracleError(RuntimeError):
    pass


def _check(code, buf):
    if code:
        raise OracleError('oracle ierr=%d: %s' % (code, buf.value.decode(errors='replace')))


def _vp(a):
    return None if a is None else a.ctypes.data


def ylmall(transpose, mu, phi, ml, mm, nstleg, nlm):
    yr = np.zeros((nstleg, nlm), np.float32, order='F')
    lib().oracle_ylmall(int(transpose), mu, phi, ml, mm, nstleg, _vp(yr))
    return yr


def precompute_phase_check(legen, nscatangle, nstokes, ml, deltam=True, negcheck=True, grad=False):
    legen = np.asfortranarray(legen, np.float32)
    nstleg, nlegp1, numphase = legen.shape
    nstphase = 1 if nstokes == 1 else 2
    tab = np.zeros((nstphase, numphase, nscatangle), np.float32, order='F')
    buf = C.create_string_buffer(600)
    fn = lib().oracle_precompute_phase_check_grad if grad else lib().oracle_precompute_phase_check
    _check(fn(nscatangle, numphase, nstphase, nstokes, ml, nstleg, nlegp1 - 1, _vp(legen), _vp(tab),
              int(deltam), int(negcheck), buf), buf)
    return tab


def finalize_scene(scene):
    """Fill YLMSUN and PHASETAB of a synthetic scene with the oracle's special functions."""
    st = scene.state
    st.ylmsun = ylmall(True, np.float32(st.solarmu), np.float32(st.solaraz), st.ml, st.mm, st.nstleg, st.nlm)
    st.phasetab = precompute_phase_check(scene.pg.legenp, st.nscatangle, st.nstokes, st.ml, bool(st.deltam))
    return scene


def _rays(rays):
    r = OracleRays(rays.nrays, _vp(rays.camx), _vp(rays.camy), _vp(rays.camz), _vp(rays.cammu), _vp(rays.camphi))
    return r


def render(state, rays, correctinterpolate=True, singlescatter=False, nosurface=False, trace_cap=0,
           nthreads=1):
    st = state.copy().normalize()        # BCRAD is mutated
    d = st.fill(OracleState())
    out = np.zeros((st.nstokes, rays.nrays), np.float32, order='F')
    tr = None
    trace = None
    if trace_cap:
        trace = dict(cells=np.zeros((trace_cap, rays.nrays), np.int32, order='F'),
                     ncells=np.zeros(rays.nrays, np.int32), nsub=np.zeros(rays.nrays, np.int32))
        tr = OracleTrace(trace_cap, _vp(trace['cells']), _vp(trace['ncells']), _vp(trace['nsub']))
    buf = C.create_string_buffer(600)
    r = _rays(rays)
    _check(lib().oracle_render(C.byref(d), C.byref(r), _vp(out), int(correctinterpolate), int(singlescatter),
                               int(nosurface), C.byref(tr) if tr is not None else None, nthreads, buf), buf)
    if trace is not None:
        return out, trace, st.bcrad
    return out


def compute_source(state, shptr, source, oshptr, delsource, fixsh=False, shacc=0.0, maxiv=None, first=False,
                   accelflag=True, newmethod=True, timing=False):
    import time
    st = state.copy().normalize()
    d = st.fill(OracleState())
    shptr = np.array(shptr, np.int32); oshptr = np.array(oshptr, np.int32)
    source = np.array(source, np.float32, order='F'); delsource = np.array(delsource, np.float32, order='F')
    if maxiv is None:
        maxiv = source.shape[1]
    o = [f32(0), f32(0), f32(0), f32(0)]
    buf = C.create_string_buffer(600)
    t = time.perf_counter()
    code = lib().oracle_compute_source(C.byref(d), int(fixsh), shacc, maxiv, int(first), int(accelflag),
                                       int(newmethod), _vp(shptr), _vp(source), _vp(oshptr), _vp(delsource),
                                       C.byref(o[0]), C.byref(o[1]), C.byref(o[2]), C.byref(o[3]), buf)
    ms = 1e3 * (time.perf_counter() - t)
    res = (code, shptr, source, oshptr, delsource, [x.value for x in o])
    return res + (ms,) if timing else res


def compute_source_sums64():
    """The four norms of the last compute_source call with f64 running sums (the REAL products of the reference,
    summed without the sequential f32 rounding of shdomsub1.f:1229-1246)."""
    out = (f64 * 4)()
    lib().oracle_compute_source_sums64.restype = None
    lib().oracle_compute_source_sums64(out)
    return list(out)


def prepare_deriv_interps(state, pg, grad):
    st = state.copy().normalize()
    d = st.fill(OracleState())
    gd = grad.fill(OracleGrad())
    npts = st.npts
    optw = np.zeros((8, npts), np.float32, order='F')
    iptr = np.zeros((8, npts), np.int32, order='F')
    dalbm = np.zeros((8, npts, grad.numder), np.float32, order='F')
    dextm = np.zeros((pg.maxpg, grad.numder), np.float32, order='F')
    dfj = np.zeros((8, npts, grad.numder), np.float32, order='F')
    buf = C.create_string_buffer(600)
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    _check(lib().oracle_prepare_deriv_interps(C.byref(d), pg.npx, pg.npy, pg.npz, pg.maxpg, pg.delx, pg.dely,
                                              pg.xstart, pg.ystart, _vp(zl), C.byref(gd), _vp(optw), _vp(iptr),
                                              _vp(dalbm), _vp(dextm), _vp(dfj), buf), buf)
    return optw, iptr, dalbm, dextm, dfj


def make_direct(state, pg, nzckd=0, zckd=None, gasabs=None):
    """MAKE_DIRECT: returns dirflux, extdirp, dict of beam constants."""
    st = state
    npts = st.npts
    extdirp = np.zeros(pg.maxpg, np.float32)
    dirflux = np.zeros(npts, np.float32)
    od = (f64 * 13)()
    oi = (i32 * 5)()
    buf = C.create_string_buffer(600)
    gp = np.asfortranarray(st.gridpos, np.float32)
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    _check(lib().oracle_make_direct(npts, st.bcflag, st.ipflag, int(st.deltam), st.ml, st.nstleg, pg.nlegp,
                                    st.solarflux, st.solarmu, st.solaraz, _vp(gp), pg.npx, pg.npy, pg.npz,
                                    pg.delx, pg.dely, pg.xstart, pg.ystart, _vp(zl), _vp(pg.extinctp),
                                    _vp(pg.albedop), _vp(pg.legenp), _vp(pg.iphasep), _vp(pg.phasewtp),
                                    pg.maxnmicro, pg.npart, nzckd, _vp(zckd), _vp(gasabs), _vp(extdirp),
                                    _vp(dirflux), od, oi, buf), buf)
    names = ['cx', 'cy', 'cz', 'cxinv', 'cyinv', 'czinv', 'epss', 'epsz', 'xdomain', 'ydomain',
             'uniformzlev', 'delxd', 'delyd']
    c = dict(zip(names, list(od)))
    c.update(ipdirect=oi[0], di=oi[1], dj=oi[2], dk=oi[3], longest_path_pts=max(oi[4], 1))
    return dirflux, extdirp, c


def make_direct_derivative(state, pg, c):
    npts = state.npts
    lpp = c['longest_path_pts']
    dpath = np.zeros((lpp, npts), np.float32, order='F')
    dptr = np.zeros((lpp, npts), np.int32, order='F')
    gp = np.asfortranarray(state.gridpos, np.float32)
    zl = np.ascontiguousarray(pg.zlevels, np.float32)
    buf = C.create_string_buffer(600)
    _check(lib().oracle_make_direct_derivative(
        npts, state.bcflag, pg.npx, pg.npy, pg.npz, pg.delx, pg.dely, pg.xstart, pg.ystart, _vp(gp), _vp(zl),
        c['ipdirect'], c['di'], c['dj'], c['dk'], c['cx'], c['cy'], c['cz'], c['cxinv'], c['cyinv'],
        c['czinv'], c['epss'], c['epsz'], c['xdomain'], c['ydomain'], c['uniformzlev'], c['delxd'],
        c['delyd'], _vp(dpath), _vp(dptr), lpp, buf), buf)
    return dpath, dptr


def levisapprox_gradient(state, rays, grad, trace_cap=0, nthreads=1):
    st = state.copy().normalize()
    d = st.fill(OracleState())
    gd = grad.fill(OracleGrad())
    gradout = np.zeros((grad.maxpg, grad.numder), np.float64, order='F')
    cost = np.zeros(1, np.float64)
    stokesout = np.zeros((st.nstokes, grad.npix), np.float32, order='F')
    tr = None
    trace = None
    if trace_cap:
        trace = dict(cells=np.zeros((trace_cap, rays.nrays), np.int32, order='F'),
                     ncells=np.zeros(rays.nrays, np.int32), nsub=np.zeros(rays.nrays, np.int32))
        tr = OracleTrace(trace_cap, _vp(trace['cells']), _vp(trace['ncells']), _vp(trace['nsub']))
    buf = C.create_string_buffer(600)
    r = _rays(rays)
    _check(lib().oracle_levisapprox_gradient(C.byref(d), C.byref(r), C.byref(gd), _vp(gradout), _vp(cost),
                                             _vp(stokesout), C.byref(tr) if tr is not None else None,
                                             nthreads, buf), buf)
    if trace is not None:
        return gradout, cost[0], stokesout, trace
    return gradout, cost[0], stokesout


def levisapprox_jacobian(state, rays, grad, jacobianptr):
    """Single-sweep (MAKEJACOBIAN=.TRUE.) path: returns gradout, cost, stokesout, jacobian[nstokes,numder,njac,npix]."""
    st = state.copy().normalize()
    d = st.fill(OracleState())
    gd = grad.fill(OracleGrad())
    jp = np.ascontiguousarray(jacobianptr, np.int32)
    gradout = np.zeros((grad.maxpg, grad.numder), np.float64, order='F')
    cost = np.zeros(1, np.float64)
    stokesout = np.zeros((st.nstokes, grad.npix), np.float32, order='F')
    jac = np.zeros((st.nstokes, grad.numder, jp.size, grad.npix), np.float32, order='F')
    buf = C.create_string_buffer(600)
    r = _rays(rays)
    _check(lib().oracle_levisapprox_jacobian(C.byref(d), C.byref(r), C.byref(gd), int(jp.size), _vp(jp),
                                             _vp(gradout), _vp(cost), _vp(stokesout), _vp(jac), buf), buf)
    return gradout, cost[0], stokesout, jac


def update_costfunction(stokesout, raygrad_pixel, gradout, cost, uncertainties, costfunc, measurement):
    nstokes = len(stokesout)
    raygrad_pixel = np.asfortranarray(raygrad_pixel, np.float64)
    _, maxpg, numder = raygrad_pixel.shape
    gradout = np.array(gradout, np.float64, order='F')
    cost = np.array(cost, np.float64)
    unc = np.asfortranarray(uncertainties, np.float64)
    so = np.ascontiguousarray(stokesout, np.float64); me = np.ascontiguousarray(measurement, np.float64)
    lib().oracle_update_costfunction(_vp(so), _vp(raygrad_pixel), _vp(gradout), _vp(cost), _vp(unc),
                                     1 if costfunc == 'LL' else 0, nstokes, maxpg, numder, _vp(me), unc.shape[0])
    return gradout, cost


def average_subpixel_rays(weighted_stokes, pixel_index, npixels):
    ws = np.asfortranarray(weighted_stokes, np.float32)
    nstokes, nrays = ws.shape
    pi = np.ascontiguousarray(pixel_index, np.int32)
    out = np.zeros((nstokes, npixels), np.float32, order='F')
    lib().oracle_average_subpixel_rays(npixels, nrays, nstokes, _vp(ws), _vp(pi), _vp(out))
    return out


def solve_fixed_grid(state, wtmu, maxiter=100, solacc=1e-5, shacc=0.0, accelflag=True, highorderrad=False,
                     iterfixsh=30, maxiv=None, initial=None):
    """Fixed-grid SHDOM solution iterations (oracle/oracle_solver.c).  Fills and returns a copy of `state`
    with shptr/source/rshptr/radiance/fluxes/bcrad of the converged solution, plus (iters, solcrit).  `initial`: a solved
    state on the same grid whose SHPTR / SOURCE / RSHPTR / RADIANCE the iterations continue from (INIT_SOLUTION with
    INRADFLAG=.FALSE. after RTE.load_solution)."""
    st = state.copy().normalize()
    npts, ns = st.npts, st.nstokes
    if maxiv is None:
        maxiv = npts * st.nlm
    lamb = st.sfctype1 == 'L' or st.sfctype1 == ord('L')
    nbc = st.ntoppts + st.nbotpts * (1 if lamb else 1 + st.nang // 2)
    st.shptr = np.zeros(npts + 1, np.int32)
    st.source = np.zeros((ns, maxiv), np.float32, order='F')
    st.rshptr = np.zeros(npts + 2, np.int32)
    st.radiance = np.zeros((ns, maxiv + npts), np.float32, order='F')
    if initial is not None:
        ts, tr = int(initial.shptr[npts]), int(initial.rshptr[npts])
        st.shptr[:] = initial.shptr[:npts + 1]
        st.rshptr[:npts + 1] = initial.rshptr[:npts + 1]
        st.source[:, :ts] = initial.source[:, :ts]
        st.radiance[:, :tr] = initial.radiance[:, :tr]
    st.fluxes = np.zeros((2, npts), np.float32, order='F')
    st.bcrad = np.zeros((ns, nbc), np.float32, order='F')
    d = st.fill(OracleState())
    wtmu = np.ascontiguousarray(wtmu, np.float32)
    iters, solcrit = i32(0), f32(0)
    buf = C.create_string_buffer(600)
    fn = lib().oracle_solve_fixed_grid_from
    fn.argtypes = [P(OracleState), C.c_void_p, i32, f32, f32, i32, i32, i32, i32, i32, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(i32), P(f32), C.c_char_p]
    _check(fn(C.byref(d), _vp(wtmu), maxiter, solacc, shacc, int(accelflag), int(highorderrad), iterfixsh, maxiv,
              int(initial is not None), _vp(st.shptr), _vp(st.source), _vp(st.rshptr), _vp(st.radiance), _vp(st.fluxes), _vp(st.bcrad),
              C.byref(iters), C.byref(solcrit), buf), buf)
    tot = int(st.shptr[npts])
    st.source = np.asfortranarray(st.source[:, :max(tot, 1)])
    st.radiance = np.asfortranarray(st.radiance[:, :max(int(st.rshptr[npts]), 1)])
    return st, iters.value, solcrit.value


def sweeping_order(state):
    st = state.copy().normalize()
    out = np.zeros((st.npts, 8), np.int32, order='F')
    d = st.fill(OracleState())
    fn = lib().oracle_sweeping_order
    fn.argtypes = [P(OracleState), C.c_void_p]
    if fn(C.byref(d), _vp(out)) != 0:
        raise RuntimeError('SWEEPING_ORDER failed')
    return out


def path_integration(state, wtmu, shptr, source, rshptr, transmin=1.0):
    """One PATH_INTEGRATION (oracle/oracle_solver.c): returns (radiance, fluxes, bcrad)."""
    lib().oracle_set_transmin.argtypes = [f32]
    lib().oracle_set_transmin.restype = None
    lib().oracle_set_transmin(transmin)
    st = state.copy().normalize()
    npts, ns = st.npts, st.nstokes
    lamb = st.sfctype1 == 'L' or st.sfctype1 == ord('L')
    nbc = st.ntoppts + st.nbotpts * (1 if lamb else 1 + st.nang // 2)
    rad = np.zeros((ns, max(int(rshptr[npts]), 1)), np.float32, order='F')
    fluxes = np.zeros((2, npts), np.float32, order='F')
    bcrad = np.zeros((ns, nbc), np.float32, order='F')
    st.bcrad = bcrad
    d = st.fill(OracleState())
    wtmu = np.ascontiguousarray(wtmu, np.float32)
    shptr = np.ascontiguousarray(shptr, np.int32); rshptr = np.ascontiguousarray(rshptr, np.int32)
    source = np.asfortranarray(source, np.float32)
    buf = C.create_string_buffer(600)
    fn = lib().oracle_path_integration_once
    fn.argtypes = [P(OracleState)] + [C.c_void_p] * 7 + [C.c_char_p]
    try:
        _check(fn(C.byref(d), _vp(wtmu), _vp(shptr), _vp(source), _vp(rshptr), _vp(rad), _vp(fluxes), _vp(bcrad), buf), buf)
    finally:
        lib().oracle_set_transmin(1.0)
    return rad, fluxes, bcrad


def surface_brdf(sfctype, refparms, wavelen, mu2, phi2, mu1, phi1, nstokes):
    """SURFACE_BRDF: REFLECT(1:nstokes,1:nstokes)."""
    refl = np.zeros((4, 4), np.float32, order='F')
    parms = np.ascontiguousarray(refparms, np.float32)
    fn = lib().oracle_surface_brdf
    fn.argtypes = [i32, C.c_void_p, f32, f32, f32, f32, f32, i32, C.c_void_p]
    code = fn(ord(sfctype), _vp(parms), wavelen, mu2, phi2, mu1, phi1, nstokes, _vp(refl))
    if code:
        raise OracleError('SURFACE_BRDF: unsupported call')
    return refl[:nstokes, :nstokes].copy()


def sh_to_do(state, wtmu, shptr, indata):
    st = state.copy().normalize()
    d = st.fill(OracleState())
    wtmu = np.ascontiguousarray(wtmu, np.float32)
    shptr = np.ascontiguousarray(shptr, np.int32)
    indata = np.asfortranarray(indata, np.float32)
    out = np.zeros((st.npts, st.nstokes, int(np.sum(st.nphi0))), np.float32, order='F')
    fn = lib().oracle_sh_to_do_all
    fn.argtypes = [P(OracleState), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fn(C.byref(d), _vp(wtmu), _vp(shptr), _vp(indata), _vp(out))
    return out


def do_to_sh(state, wtmu, rshptr, dofield):
    st = state.copy().normalize()
    d = st.fill(OracleState())
    wtmu = np.ascontiguousarray(wtmu, np.float32)
    rshptr = np.ascontiguousarray(rshptr, np.int32)
    dofield = np.asfortranarray(dofield, np.float32)
    out = np.zeros((st.nstokes, max(int(rshptr[st.npts]), 1)), np.float32, order='F')
    fn = lib().oracle_do_to_sh_all
    fn.argtypes = [P(OracleState), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fn(C.byref(d), _vp(wtmu), _vp(rshptr), _vp(dofield), _vp(out))
    return out


def radiance_truncation(state, shptr, radiance, rshptr, fixsh, shacc, highorderrad, maxir):
    st = state.copy().normalize()
    d = st.fill(OracleState())
    shptr = np.ascontiguousarray(shptr, np.int32)
    radiance = np.asfortranarray(radiance, np.float32)
    out = np.array(rshptr, np.int32)
    fn = lib().oracle_radiance_truncation
    fn.argtypes = [P(OracleState), i32, C.c_void_p, C.c_void_p, i32, i32, f32, C.c_void_p]
    rc = fn(C.byref(d), int(highorderrad), _vp(shptr), _vp(radiance), int(maxir), int(fixsh), shacc, _vp(out))
    if rc:
        raise OracleError('RADIANCE_TRUNCATION: out of memory')
    return out


class OracleProp(C.Structure):
    _fields_ = [(n, i32) for n in ('npx', 'npy', 'npz', 'numphase', 'nlegp', 'maxnmicro', 'npart', 'nzckd', 'nstleg')] + \
               [(n, f32) for n in ('delx', 'dely', 'xstart', 'ystart')] + \
               [(n, C.c_void_p) for n in ('zlevels', 'tempp', 'extinctp', 'albedop', 'legenp', 'iphasep', 'phasewtp',
                                          'zckd', 'gasabs')]


def _prop(pg, tempp=None, zckd=None, gasabs=None):
    """oracle_prop view of a PropertyGrid; returns (struct, keep-alive list)."""
    keep = [np.ascontiguousarray(pg.zlevels, np.float32), None if tempp is None else np.ascontiguousarray(tempp, np.float32),
            None if zckd is None else np.ascontiguousarray(zckd, np.float32),
            None if gasabs is None else np.ascontiguousarray(gasabs, np.float32)]
    p = OracleProp(pg.npx, pg.npy, pg.npz, pg.numphase, pg.nlegp, pg.maxnmicro, pg.npart,
                   0 if zckd is None else len(zckd), pg.nstleg, pg.delx, pg.dely, pg.xstart, pg.ystart,
                   _vp(keep[0]), _vp(keep[1]), _vp(pg.extinctp), _vp(pg.albedop), _vp(pg.legenp), _vp(pg.iphasep),
                   _vp(pg.phasewtp), _vp(keep[2]), _vp(keep[3]))
    return p, keep


def transfer_pa_to_grid(pg, gridpos, npts, ml, deltam, interp_new=True, phasemax=0.999, srctype='S', units='R',
                        wavelen=0.0, tempp=None, zckd=None, gasabs=None):
    """TRANSFER_PA_TO_GRID (oracle/oracle_prop.c) -> dict like at3d_b200.medium.transfer_pa_to_grid, plus temp/planck."""
    p, keep = _prop(pg, tempp, zckd, gasabs)
    nleg = ml + 1 if deltam else ml
    nq = 8 * pg.maxnmicro
    gp = np.asfortranarray(gridpos[:, :npts], np.float32)
    out = dict(temp=np.zeros(npts, np.float32), planck=np.zeros((npts, pg.npart), np.float32, order='F'),
               extinct=np.zeros((npts, pg.npart), np.float32, order='F'),
               albedo=np.zeros((npts, pg.npart), np.float32, order='F'),
               legen=np.zeros((pg.nstleg, nleg + 1, pg.numphase), np.float32, order='F'),
               iphase=np.zeros((nq, npts, pg.npart), np.int32, order='F'),
               phaseinterpwt=np.zeros((nq, npts, pg.npart), np.float32, order='F'),
               total_ext=np.zeros(npts, np.float32))
    extmin, scatmin, albmax = f64(0), f64(0), f32(0)
    waveno = np.zeros(2, np.float32)
    buf = C.create_string_buffer(600)
    fn = lib().oracle_transfer_pa_to_grid
    fn.argtypes = [P(OracleProp), i32, C.c_void_p, i32, i32, i32, i32, f32, i32, i32, C.c_void_p, f32] + \
        [C.c_void_p] * 8 + [P(f64), P(f64), P(f32), C.c_char_p]
    _check(fn(C.byref(p), npts, _vp(gp), ml, nleg, int(deltam), int(interp_new), phasemax, ord(srctype), ord(units),
              _vp(waveno), wavelen, _vp(out['temp']), _vp(out['planck']), _vp(out['extinct']), _vp(out['albedo']),
              _vp(out['legen']), _vp(out['iphase']), _vp(out['phaseinterpwt']), _vp(out['total_ext']),
              C.byref(extmin), C.byref(scatmin), C.byref(albmax), buf), buf)
    out.update(nleg=nleg, extmin=extmin.value, scatmin=scatmin.value, albmax=albmax.value)
    return out


def solve_adaptive(state, pg, wtmu, tempp=None, splitacc=0.03, shacc=0.0, solacc=1e-4, maxiter=100, accelflag=True,
                   highorderrad=False, iterfixsh=30, adapt_grid_factor=5.0, num_sh_term_factor=1.0,
                   cell_to_point_ratio=1.5, inradflag=True, temp=None):
    """INIT_SOLUTION + SOLUTION_ITERATIONS with adaptive cell splitting (oracle/oracle_solver.c).  `state` holds the
    base grid and the optical properties on it; array capacities follow RTE._setup_memory (at3d/solver.py:2286-2322).
    Returns (solved state with the split grid, iters, solcrit, splitcrit)."""
    st = state.copy().normalize()
    ns, nbpts, nbcells, npart, nq = st.nstokes, st.npts, st.ncells, st.npart, 8 * st.maxnmicro
    maxig = int(adapt_grid_factor * nbpts)
    maxic = max(int(cell_to_point_ratio * maxig), nbcells)
    maxiv = max(int(num_sh_term_factor * st.nlm * maxig), nbpts * 4)
    maxido = maxig * st.nphi0max
    maxnbc = int(maxig * 3 / st.nz)
    lamb = st.sfctype1 in ('L', ord('L'))
    maxbcrad = 2 * maxnbc if lamb else int((2 + st.nmu * st.nphi0max / 2) * maxnbc)

    def grow(a, shape, dtype):
        out = np.zeros(shape, dtype, order='F')
        if a is not None:
            a = np.asarray(a)
            out[tuple(slice(0, s) for s in a.shape)] = a
        return out
    st.gridpos = grow(st.gridpos, (3, maxig), np.float32)
    st.gridptr = grow(st.gridptr, (8, maxic), np.int32)
    st.neighptr = grow(st.neighptr, (6, maxic), np.int32)
    st.treeptr = grow(st.treeptr, (2, maxic), np.int32)
    st.cellflags = grow(st.cellflags, (maxic,), np.int16)
    st.extinct = grow(st.extinct, (maxig, npart), np.float32)
    st.albedo = grow(st.albedo, (maxig, npart), np.float32)
    st.planck = grow(st.planck, (maxig, npart), np.float32)
    st.total_ext = grow(st.total_ext, (maxig,), np.float32)
    st.iphase = grow(st.iphase, (nq, maxig, npart), np.int32)
    st.iphase[st.iphase == 0] = 1
    st.phaseinterpwt = grow(st.phaseinterpwt, (nq, maxig, npart), np.float32)
    st.dirflux = np.zeros(maxig, np.float32)
    st.fluxes = np.zeros((2, maxig), np.float32, order='F')
    st.shptr = np.zeros(maxig + 1, np.int32)
    st.rshptr = np.zeros(maxig + 2, np.int32)
    st.source = np.zeros((ns, maxiv), np.float32, order='F')
    st.radiance = np.zeros((ns, maxiv + maxig), np.float32, order='F')
    st.bcptr = np.zeros((maxnbc, 2), np.int32, order='F')
    st.maxnbc = maxnbc
    st.bcrad = np.zeros((ns, maxbcrad), np.float32, order='F')
    tgrid = grow(temp, (maxig,), np.float32)
    extdirp = np.zeros(pg.maxpg, np.float32)
    p, keep = _prop(pg, tempp)
    d = st.fill(OracleState())
    wtmu = np.ascontiguousarray(wtmu, np.float32)
    iters, solcrit, splitcrit = i32(0), f32(0), f32(0)
    buf = C.create_string_buffer(600)
    fn = lib().oracle_solve_adaptive
    fn.argtypes = [P(OracleState), P(OracleProp), C.c_void_p, C.c_void_p] + [i32] * 8 + [f32, f32, f32, i32, i32, i32, i32,
                                                                                   C.c_void_p, P(i32), P(f32), P(f32),
                                                                                   C.c_char_p]
    _check(fn(C.byref(d), C.byref(p), _vp(wtmu), _vp(tgrid), maxig, maxic, maxiv, maxido, maxbcrad, nbpts, nbcells,
              maxiter, solacc, splitacc, shacc, int(accelflag), int(highorderrad), iterfixsh, int(inradflag),
              _vp(extdirp), C.byref(iters), C.byref(solcrit), C.byref(splitcrit), buf), buf)
    npts, ncells = d.npts, d.ncells
    st.npts, st.ncells, st.ntoppts, st.nbotpts = npts, ncells, d.ntoppts, d.nbotpts
    st.gridpos = np.asfortranarray(st.gridpos[:, :npts])
    for n in ('gridptr', 'neighptr', 'treeptr'):
        setattr(st, n, np.asfortranarray(getattr(st, n)[:, :ncells]))
    st.cellflags = st.cellflags[:ncells].copy()
    for n in ('extinct', 'albedo', 'planck'):
        setattr(st, n, np.asfortranarray(getattr(st, n)[:npts, :]))
    for n in ('iphase', 'phaseinterpwt'):
        setattr(st, n, np.asfortranarray(getattr(st, n)[:, :npts, :]))
    st.total_ext = st.total_ext[:npts].copy()
    st.dirflux = st.dirflux[:npts].copy()
    st.fluxes = np.asfortranarray(st.fluxes[:, :npts])
    st.shptr = st.shptr[:npts + 1].copy()
    st.rshptr = st.rshptr[:npts + 2].copy()
    st.source = np.asfortranarray(st.source[:, :max(int(st.shptr[npts]), 1)])
    st.radiance = np.asfortranarray(st.radiance[:, :max(int(st.rshptr[npts]), 1)])
    st.temp = tgrid[:npts].copy()
    st.extdirp = extdirp
    return st, iters.value, solcrit.value, splitcrit.value
