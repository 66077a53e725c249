"""Fixed-grid SHDOM solution iterations on the GPU.

Mirrors ``at3d.solver.RTE.solve`` -> ``core.solution_iterations`` (``SOLUTION_ITERATIONS``,
src/polarized/shdomsub1.f:445-822) for ``split_accuracy=0``.  ``solve_fixed_grid`` runs the whole loop resident in HBM
through ``at3d_solver_create`` / ``at3d_solver_solve`` (``RADIANCE_TRUNCATION``, ``PATH_INTEGRATION`` -- independent
columns for ``ip_flag=3``, ``BACK_INT_GRID2D`` for ``ip_flag=2``, the ``BACK_INT_GRID3D`` data-flow sweep otherwise --
``COMPUTE_SOURCE``, ``CALC_ACCEL_SOLCRIT`` / ``ACCELERATE_SOLUTION``); with ``device_loop=False`` the same iteration is
driven from Python through the per-routine C-ABI calls (``at3d_path_integration_ip`` / ``at3d_solver_path_integration``,
``at3d_compute_source``) with ``RADIANCE_TRUNCATION`` on the host.  The first guess is a zero radiance field
(``INIT_RADIANCE``'s Eddington field is not restated; the fixed point does not depend on it).  Adaptive cell splitting
(``SPLIT_GRID``) is SURVEY.md 8f "next".
"""
import ctypes as C
import numpy as np
from . import _lib
from . import backend as B
from . import grid as G
from ._lib import vp
from .state import i32


def radiance_truncation(st, shptr, radiance, rshptr, fixsh, shacc, highorderrad, maxir):
    """RADIANCE_TRUNCATION (shdomsub1.f:1615-1805): the new RSHPTR[npts+2] (host; integer results)."""
    npts, ml, mm, nq = st.npts, st.ml, st.mm, 8 * st.maxnmicro
    f32 = np.float32
    full = (ml * (ml + 1)) + ml + 1 if ml <= mm else (2 * mm + 1) * ml - (mm * (1 + (mm - 1))) + mm + 1
    ns = np.diff(shptr[:npts + 1]).astype(np.int64)
    if not fixsh:
        lofj = G.lofj(ml, mm)
        nro = np.diff(rshptr[:npts + 1]).astype(np.int64)
        # NOTEND turns false at the first point without radiance terms and stays false
        zero = np.nonzero(nro == 0)[0]
        notend = np.ones(npts, bool)
        if zero.size:
            notend[zero[0]:] = False
        rad0 = radiance[0, np.minimum(rshptr[:npts], max(radiance.shape[1] - 1, 0))].astype(f32)
        ext = st.total_ext[:npts].astype(f32)
        nlt = st.nstleg * (st.nleg + 1)
        legen_flat = np.asfortranarray(st.legen, f32).ravel(order='F')
        iph = st.iphase.reshape(nq, npts, st.npart, order='F')
        pwt = st.phaseinterpwt.reshape(nq, npts, st.npart, order='F')
        single = pwt[0] >= f32(st.phasemax)                               # [npts, npart]
        fs = np.ones((npts, st.npart), f32)
        if st.interp_new and st.deltam:
            for ipa in range(st.npart):
                fsum = np.zeros(npts, f32)
                for q in range(nq):
                    # the reference indexes LEGEN(Q,ML+1,IPHASE(Q,.)) with the mixture index Q (shdomsub1.f:1672)
                    off = np.minimum(nlt * (iph[q, :, ipa].astype(np.int64) - 1) + st.nstleg * (ml + 1) + q,
                                     legen_flat.size - 1)
                    use = pwt[q, :, ipa] > f32(1e-5)
                    fsum = np.where(use, fsum + legen_flat[off] * pwt[q, :, ipa], fsum).astype(f32)
                f1 = legen_flat[nlt * (iph[0, :, ipa].astype(np.int64) - 1) + st.nstleg * (ml + 1)]
                f = np.where(single[:, ipa], f1, fsum).astype(f32)
                fs[:, ipa] = (f32(1.0) / (f32(1.0) - f)).astype(f32)
        lr = np.ones(npts, np.int64)
        for l in range(1, ml + 1):
            rad = np.zeros(npts, f32)
            for ipa in range(st.npart):
                e = st.extinct.reshape(npts, st.npart, order='F')[:, ipa].astype(f32)
                with np.errstate(divide='ignore', invalid='ignore'):
                    w = np.where(ext == 0, f32(1.0), e / ext).astype(f32)
                alb = st.albedo.reshape(npts, st.npart, order='F')[:, ipa].astype(f32)
                leg1 = legen_flat[nlt * (iph[0, :, ipa].astype(np.int64) - 1) + st.nstleg * l]
                if st.interp_new:
                    mix = np.zeros(npts, f32)
                    for q in range(nq):
                        lq = legen_flat[nlt * (iph[q, :, ipa].astype(np.int64) - 1) + st.nstleg * l]
                        mix = np.where(pwt[q, :, ipa] > f32(1e-5), mix + lq * pwt[q, :, ipa], mix).astype(f32)
                    legent = np.where(single[:, ipa], leg1, mix).astype(f32)
                    if st.deltam:
                        legent = (legent * fs[:, ipa]).astype(f32)
                else:
                    legent = leg1
                term = (w * alb * legent * rad0).astype(f32)
                rad = np.where(w == 0, rad, rad + term).astype(f32)
            lr = np.where(rad > f32(shacc), l, lr)
        lr = np.where(notend, lr, ml)
        ls = lofj[np.maximum(ns, 1) - 1]
        lr = np.minimum(lr, ls + ml // 8 + 2)
        if highorderrad:
            lr = np.full(npts, ml)
        nr = np.where(lr <= mm, lr * (lr + 1) + lr + 1, (2 * mm + 1) * lr - (mm * (1 + (mm - 1))) + mm + 1)
        out = np.zeros(npts + 2, np.int32)
        csum = np.cumsum(nr)
        if csum[-1] <= maxir:
            out[1:npts + 1] = csum
            out[npts + 1] = csum[-1]
            return out
    nr = np.maximum(4, ns)
    if highorderrad:
        nr = np.full(npts, full)
    csum = np.cumsum(nr)
    if csum[-1] > maxir:
        raise MemoryError('RADIANCE_TRUNCATION: Really out of memory for more radiance terms. Increase MAXIV.')
    out = np.zeros(npts + 2, np.int32)
    out[1:npts + 1] = csum
    out[npts + 1] = csum[-1]
    return out


def path_integration_ip(st, wtmu, shptr, source, rshptr, timing=False):
    """PATH_INTEGRATION on the GPU: returns (radiance[nstokes, rshptr[npts]], fluxes[2,npts], bcrad)."""
    lamb = st.sfctype1 in ('L', ord('L'))
    nbc = st.ntoppts + st.nbotpts * (1 if lamb else 1 + st.nang // 2)
    rad = np.zeros((st.nstokes, max(int(rshptr[st.npts]), 1)), np.float32, order='F')
    fluxes = np.zeros((2, st.npts), np.float32, order='F')
    bcrad = np.zeros((st.nstokes, nbc), np.float32, order='F')
    d = st.desc()
    wt = np.ascontiguousarray(wtmu, np.float32)
    shptr = np.ascontiguousarray(shptr, np.int32); rshptr = np.ascontiguousarray(rshptr, np.int32)
    source = np.asfortranarray(source, np.float32)
    ms = C.c_double(0.0)
    buf = _lib.errbuf()
    _lib.check(_lib.lib().at3d_path_integration_ip(C.byref(d), vp(wt), vp(shptr), vp(source), vp(rshptr), vp(rad),
                                                   vp(fluxes), vp(bcrad), C.byref(ms), buf), buf)
    return (rad, fluxes, bcrad, ms.value) if timing else (rad, fluxes, bcrad)


def sweeping_order(st):
    """SWEEPING_ORDER (shdomsub1.f:3261-3352) of a 3-D (8 octants) or 2-D (IPFLAG=2, 4 octants) grid: SWEEPORD[npts,
    noct] (Fortran order), host only."""
    out = np.zeros((st.npts, 4 if (st.ipflag & 2) else 8), np.int32, order='F')
    d = st.desc()
    buf = _lib.errbuf()
    _lib.check(_lib.lib().at3d_sweeping_order(C.byref(d), vp(out), buf), buf)
    return out


class SweepSolver:
    """PATH_INTEGRATION / SOLUTION_ITERATIONS on a fixed grid: the device-resident solver object of at3d_solver_create
    (3-D grids, IPFLAG 0 or 1, and 2-D grids, IPFLAG=2: topology, SWEEPING_ORDER, ordinate geometry, transform tables,
    discrete-ordinate fields; IPFLAG=3: independent columns)."""

    def __init__(self, st, wtmu, transmin=1.0):
        self.st = st
        self._keep = st.desc()
        self.h = C.c_void_p()
        wt = np.ascontiguousarray(wtmu, np.float32)
        buf = _lib.errbuf()
        _lib.check(_lib.lib().at3d_solver_create(C.byref(self._keep), vp(wt), float(transmin), C.byref(self.h), buf), buf)

    def update_medium(self, st):
        """A new medium on the same grid (`at3d_solver_update_medium`): the sweep order, dependency levels and the sorted
        sweep plan of the object are kept; only extinction, direct beam and surface parameters are replaced.  `st` is a
        ShdomState of the solver's grid; the next `solve` reads its optical properties."""
        if st.npts != self.st.npts or st.ncells != self.st.ncells:
            raise ValueError('update_medium: the state is on another grid')
        keep = st.desc()
        buf = _lib.errbuf()
        _lib.check(_lib.lib().at3d_solver_update_medium(self.h, C.byref(keep), buf), buf)
        self.st, self._keep = st, keep

    def path_integration(self, shptr, source, rshptr, timing=False):
        st = self.st
        lamb = st.sfctype1 in ('L', ord('L'))
        nbc = st.ntoppts + st.nbotpts * (1 if lamb else 1 + st.nang // 2)
        rad = np.zeros((st.nstokes, max(int(rshptr[st.npts]), 1)), np.float32, order='F')
        fluxes = np.zeros((2, st.npts), np.float32, order='F')
        bcrad = np.zeros((st.nstokes, nbc), np.float32, order='F')
        shptr = np.ascontiguousarray(shptr, np.int32); rshptr = np.ascontiguousarray(rshptr, np.int32)
        source = np.asfortranarray(source, np.float32)
        ms = C.c_double(0.0)
        buf = _lib.errbuf()
        _lib.check(_lib.lib().at3d_solver_path_integration(self.h, vp(shptr), vp(source), vp(rshptr), vp(rad), vp(fluxes),
                                                           vp(bcrad), C.byref(ms), buf), buf)
        return (rad, fluxes, bcrad, ms.value) if timing else (rad, fluxes, bcrad)

    def solve(self, maxiter=100, solacc=1e-4, shacc=0.0, accelflag=True, highorderrad=False, iterfixsh=30, maxiv=None,
              initial=None):
        """The whole fixed-grid solve on the device (at3d_solver_solve).  Returns (solved copy of the state, iterations,
        solcrit, timings).  `initial`: a solved state on the same grid (e.g. the previous step of an optimisation): the
        iterations continue from its SHPTR / SOURCE / RSHPTR / RADIANCE instead of the first guess
        (at3d_solver_solve_from; the reference's load_solution + INRADFLAG=.FALSE.)."""
        st = self.st.copy().normalize()
        npts, ns = st.npts, st.nstokes
        if maxiv is None:
            maxiv = npts * st.nlm
        lamb = st.sfctype1 in ('L', ord('L'))
        nbc = st.ntoppts + st.nbotpts * (1 if lamb else 1 + st.nang // 2)
        shptr = np.zeros(npts + 1, np.int32)
        source = np.empty((ns, maxiv), np.float32, order='F')              # only [:, :shptr[npts]] is written and kept
        rshptr = np.zeros(npts + 2, np.int32)
        radiance = np.empty((ns, maxiv + npts), np.float32, order='F')
        fluxes = np.zeros((2, npts), np.float32, order='F')
        bcrad = np.zeros((ns, nbc), np.float32, order='F')
        iters, solcrit = C.c_int32(0), C.c_float(0.0)
        ms = np.zeros(3, np.float64)
        buf = _lib.errbuf()
        if initial is not None:
            if initial.npts != npts or initial.nstokes != ns:
                raise ValueError('solve(initial=...): the initial solution is on another grid')
            ts, tr = int(initial.shptr[npts]), int(initial.rshptr[npts])
            if ts > maxiv or tr > maxiv + npts:
                raise MemoryError('solve(initial=...): the initial solution does not fit maxiv')
            shptr[:] = np.asarray(initial.shptr)[:npts + 1]
            rshptr[:npts + 1] = np.asarray(initial.rshptr)[:npts + 1]
            source[:, :ts] = np.asarray(initial.source)[:, :ts]
            radiance[:, :tr] = np.asarray(initial.radiance)[:, :tr]
        rc = _lib.lib().at3d_solver_solve_from(self.h, C.byref(self._keep), int(maxiter), float(solacc), float(shacc),
                                               int(accelflag), int(highorderrad), int(iterfixsh), int(maxiv),
                                               int(initial is not None), vp(shptr), vp(source), vp(rshptr), vp(radiance),
                                               vp(fluxes), vp(bcrad), C.byref(iters), C.byref(solcrit), vp(ms), buf)
        if rc == 2:
            raise MemoryError(buf.value.decode(errors='replace'))
        _lib.check(rc, buf)
        st.shptr, st.rshptr = shptr, rshptr
        st.source = np.asfortranarray(source[:, :max(int(shptr[npts]), 1)])
        st.radiance = np.asfortranarray(radiance[:, :max(int(rshptr[npts]), 1)])
        st.fluxes, st.bcrad = fluxes, bcrad
        return st, int(iters.value), float(solcrit.value), dict(path_integration_ms=float(ms[0]), compute_source_ms=float(ms[1]),
                                                                   loop_ms=float(ms[2]))

    def close(self):
        if self.h:
            _lib.lib().at3d_solver_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def solve_ip(state, wtmu, **kw):
    """The fixed-grid solve for independent-pixel grids (IPFLAG=3); see solve_fixed_grid (device_loop=False drives the
    iterations from Python through at3d_path_integration_ip and at3d_compute_source)."""
    return solve_fixed_grid(state, wtmu, **kw)


def solve_fixed_grid(state, wtmu, maxiter=100, solacc=1e-4, shacc=0.0, accelflag=True, highorderrad=False, iterfixsh=30,
                     maxiv=None, verbose=False, transmin=1.0, device_loop=True):
    """SOLUTION_ITERATIONS on a fixed grid (shdomsub1.f:445-822): RADIANCE_TRUNCATION, PATH_INTEGRATION (GPU: independent
    columns for IPFLAG=3, the BACK_INT_GRID3D sweep otherwise), COMPUTE_SOURCE (GPU), sequence acceleration.
    Returns (solved copy of `state` with shptr/source/rshptr/radiance/fluxes/bcrad, iters, solcrit, timings)."""
    st = state.copy().normalize()
    sweep = None if ((st.ipflag & 3) == 3 and not device_loop) else SweepSolver(st, wtmu, transmin)
    if device_loop and not verbose:
        # the whole iteration loop stays on the device (independent-pixel and 3-D grids)
        try:
            return sweep.solve(maxiter=maxiter, solacc=solacc, shacc=shacc, accelflag=accelflag, highorderrad=highorderrad,
                               iterfixsh=iterfixsh, maxiv=maxiv)
        finally:
            sweep.close()
    npts, ns = st.npts, st.nstokes
    f32 = np.float32
    if maxiv is None:
        maxiv = npts * st.nlm
    maxir = maxiv + npts
    rshptr = np.zeros(npts + 2, np.int32)
    rshptr[:npts + 1] = 4 * np.arange(npts + 1)
    rshptr[npts + 1] = rshptr[npts]
    st.rshptr = rshptr
    st.radiance = np.zeros((ns, int(rshptr[npts])), np.float32, order='F')
    shptr = np.zeros(npts + 1, np.int32)
    source = np.zeros((ns, maxiv), np.float32, order='F')
    delsource = np.zeros((ns, maxiv), np.float32, order='F')
    rc, shptr, source, oshptr, delsource, sums = B.compute_source(st, shptr, source, shptr.copy(), delsource, fixsh=False,
                                                                 shacc=shacc, maxiv=maxiv, first=True, accelflag=accelflag)
    if rc:
        raise MemoryError('COMPUTE_SOURCE: out of spherical-harmonic memory')
    if accelflag:
        oshptr = shptr.copy()
        delsource[:, :int(oshptr[npts])] = 0.0
    albmax = float(np.max(st.albedo))
    solcrit, a, it, fixsh = f32(1.0), f32(0.0), 0, False
    t_path = t_src = 0.0
    while it < maxiter and solcrit > solacc:
        it += 1
        st.rshptr = radiance_truncation(st, shptr, st.radiance, st.rshptr, fixsh, shacc, highorderrad, maxir)
        st.shptr, st.source = shptr, source
        if sweep is None:
            rad, fluxes, bcrad, ms = path_integration_ip(st, wtmu, shptr, source, st.rshptr, timing=True)
        else:
            rad, fluxes, bcrad, ms = sweep.path_integration(shptr, source, st.rshptr, timing=True)
        t_path += ms
        st.radiance, st.fluxes, st.bcrad = rad, fluxes, bcrad
        if solcrit < 0.001 or it > iterfixsh:
            fixsh = True
        res = B.compute_source(st, shptr, source, oshptr, delsource, fixsh=fixsh, shacc=shacc, maxiv=maxiv,
                               first=False, accelflag=accelflag, timing=True)
        rc, shptr, source, oshptr, delsource, sums = res[:6]
        t_src += res[6]
        if rc:
            raise MemoryError('COMPUTE_SOURCE: out of spherical-harmonic memory')
        deljdot, deljold, deljnew, jnorm = (f32(x) for x in sums)
        # CALC_ACCEL_SOLCRIT
        if accelflag and a == 0 and deljnew < deljold:
            r = f32(np.sqrt(deljnew / deljold))
            theta = f32(np.arccos(deljdot / f32(np.sqrt(deljold * deljnew))))
            a = f32((1 - r * f32(np.cos(theta)) + r ** f32(1 + 0.5 * 3.14159 / theta)) /
                    (1 + r * r - 2 * r * f32(np.cos(theta))) - 1.0)
            a = f32(min(10.0, max(0.0, float(a))))
        else:
            a = f32(0.0)
        if jnorm > 0:
            solcrit = f32(np.sqrt(deljnew / jnorm))
        elif deljnew == 0:
            solcrit = f32(0.0)
        # ACCELERATE_SOLUTION
        if a > 0:
            nsn, nsd = np.diff(shptr[:npts + 1]), np.diff(oshptr[:npts + 1])
            nsc = np.minimum(nsn, nsd)
            idx_s = np.concatenate([np.arange(s0, s0 + n) for s0, n in zip(shptr[:npts], nsc)]) if nsc.sum() else np.zeros(0, int)
            idx_d = np.concatenate([np.arange(s0, s0 + n) for s0, n in zip(oshptr[:npts], nsc)]) if nsc.sum() else np.zeros(0, int)
            source[:, idx_s] = (source[:, idx_s] + a * delsource[:, idx_d]).astype(f32)
        if albmax < solacc:
            solcrit = f32(solacc)
        if verbose:
            print('  %4d %8.3f %8d' % (it, np.log10(max(float(solcrit), 1e-20)), npts))
    if sweep is not None:
        sweep.close()
    tot = int(shptr[npts])
    st.shptr = shptr
    st.source = np.asfortranarray(source[:, :max(tot, 1)])
    return st, it, float(solcrit), dict(path_integration_ms=t_path, compute_source_ms=t_src)


# ---------------------------------------------------------------------------------------------------------------------
# the adaptive solve (C `at3d_solve_adaptive`): RTE.solve of at3d/solver.py:279 with split_accuracy > 0
# ---------------------------------------------------------------------------------------------------------------------
class PropDesc(C.Structure):
    """at3d_prop_desc of include/at3d_b200.h."""
    _fields_ = [(n, i32) for n in ('npx', 'npy', 'npz', 'numphase', 'nlegp', 'maxnmicro', 'npart', 'nzckd', 'nstleg')] + \
               [(n, C.c_float) for n in ('delx', 'dely', 'xstart', 'ystart')] + \
               [(n, C.c_void_p) for n in ('zlevels', 'tempp', 'extinctp', 'albedop', 'legenp', 'iphasep', 'phasewtp',
                                          'zckd', 'gasabs')]


class AdaptIO(C.Structure):
    """at3d_adapt_io of include/at3d_b200.h."""
    _fields_ = [(n, i32) for n in ('maxig', 'maxic', 'maxiv', 'maxido', 'maxnbc', 'maxbcrad', 'nbpts', 'nbcells',
                                   'maxiter', 'accelflag', 'highorderrad', 'iterfixsh', 'inradflag')] + \
               [(n, C.c_float) for n in ('solacc', 'splitacc', 'shacc', 'transmin')] + \
               [('nxsfc', i32), ('nysfc', i32), ('delxsfc', C.c_float), ('delysfc', C.c_float), ('sfcparms', C.c_void_p)] + \
               [(n, C.c_void_p) for n in ('gridpos', 'gridptr', 'neighptr', 'treeptr', 'cellflags', 'temp', 'planck',
                                          'extinct', 'albedo', 'total_ext', 'dirflux', 'fluxes', 'iphase', 'phaseinterpwt',
                                          'shptr', 'rshptr', 'source', 'radiance', 'bcptr', 'bcrad', 'sfcgridparms',
                                          'extdirp')] + \
               [(n, i32) for n in ('npts', 'ncells', 'ntoppts', 'nbotpts', 'iters', 'nsplit_calls')] + \
               [('solcrit', C.c_float), ('splitcrit', C.c_float)]


def memory_sizes(nbpts, nbcells, nlm, nz, nmu, nphi0max, lambertian, adapt_grid_factor=5.0, num_sh_term_factor=1.0,
                 cell_to_point_ratio=1.5):
    """MAXIG, MAXIC, MAXIV, MAXIDO, MAXNBC, MAXBCRAD of RTE._setup_memory (at3d/solver.py:2286-2322, :1942-1944)."""
    maxig = int(adapt_grid_factor * nbpts)
    maxic = max(int(cell_to_point_ratio * maxig), nbcells)
    maxiv = max(int(num_sh_term_factor * nlm * maxig), nbpts * 4)
    maxnbc = int(maxig * 3 / nz)
    maxbcrad = 2 * maxnbc if lambertian else int((2 + nmu * nphi0max / 2) * maxnbc)
    return maxig, maxic, maxiv, maxig * nphi0max, maxnbc, maxbcrad


def solve_adaptive(state, pg, wtmu, tempp=None, temp=None, splitacc=0.03, shacc=0.0, solacc=1e-4, maxiter=100, accelflag=True,
                   highorderrad=False, iterfixsh=30, adapt_grid_factor=5.0, num_sh_term_factor=1.0, cell_to_point_ratio=1.5,
                   inradflag=True, transmin=1.0, sfcparms=None, delxsfc=0.0, delysfc=0.0, zckd=None, gasabs=None, timing=False):
    """INIT_SOLUTION + SOLUTION_ITERATIONS with adaptive cell splitting on the GPU (C `at3d_solve_adaptive`).

    ``state``: ShdomState of the BASE grid with the optical properties on it (TRANSFER_PA_TO_GRID) and YLMSUN; ``pg``: the
    PropertyGrid; array capacities follow RTE._setup_memory.  Returns (solved state on the split grid, iters, solcrit,
    splitcrit) and the ms[4] timings when ``timing``.  ``sfcparms``: SFCPARMS[nsfcpar, nxsfc+1, nysfc+1] of a variable surface."""
    st = state.copy().normalize()
    ns, nbpts, nbcells, npart, nq = st.nstokes, st.npts, st.ncells, st.npart, 8 * st.maxnmicro
    lamb = st.sfctype1 in ('L', ord('L'))
    maxig, maxic, maxiv, maxido, maxnbc, maxbcrad = memory_sizes(nbpts, nbcells, st.nlm, st.nz, st.nmu, st.nphi0max, lamb,
                                                                 adapt_grid_factor, num_sh_term_factor, cell_to_point_ratio)
    if splitacc <= 0.0:
        maxig, maxic = nbpts, nbcells
        maxiv = max(int(num_sh_term_factor * st.nlm * maxig), nbpts * 4)
        maxido, maxnbc = maxig * st.nphi0max, max(int(maxig * 3 / st.nz), st.ntoppts, st.nbotpts)
        maxbcrad = 2 * maxnbc if lamb else int((2 + st.nmu * st.nphi0max / 2) * maxnbc)

    def grow(a, shape, dtype):
        out = np.zeros(shape, dtype, order='F')
        if a is not None:
            a = np.asarray(a)
            out[tuple(slice(0, s) for s in a.shape)] = a
        return out
    arr = dict(
        gridpos=grow(st.gridpos, (3, maxig), np.float32), gridptr=grow(st.gridptr, (8, maxic), np.int32),
        neighptr=grow(st.neighptr, (6, maxic), np.int32), treeptr=grow(st.treeptr, (2, maxic), np.int32),
        cellflags=grow(st.cellflags, (maxic,), np.int16), temp=grow(temp if temp is not None else st.temp, (maxig,), np.float32),
        planck=grow(st.planck, (maxig, npart), np.float32), extinct=grow(st.extinct, (maxig, npart), np.float32),
        albedo=grow(st.albedo, (maxig, npart), np.float32), total_ext=grow(st.total_ext, (maxig,), np.float32),
        dirflux=np.zeros(maxig, np.float32), fluxes=np.zeros((2, maxig), np.float32, order='F'),
        iphase=grow(st.iphase, (nq, maxig, npart), np.int32), phaseinterpwt=grow(st.phaseinterpwt, (nq, maxig, npart), np.float32),
        shptr=np.zeros(maxig + 1, np.int32), rshptr=np.zeros(maxig + 2, np.int32),
        source=np.zeros((ns, maxiv), np.float32, order='F'), radiance=np.zeros((ns, maxiv + maxig), np.float32, order='F'),
        bcptr=np.zeros((maxnbc, 2), np.int32, order='F'), bcrad=np.zeros((ns, maxbcrad), np.float32, order='F'),
        sfcgridparms=np.zeros((max(st.nsfcpar, 1), maxnbc), np.float32, order='F'), extdirp=np.zeros(pg.maxpg, np.float32))
    arr['iphase'][arr['iphase'] == 0] = 1
    if st.sfcgridparms is not None:
        sg = np.asarray(st.sfcgridparms)
        arr['sfcgridparms'][:sg.shape[0], :sg.shape[1]] = sg
    keep = [np.ascontiguousarray(pg.zlevels, np.float32),
            None if tempp is None else np.ascontiguousarray(tempp, np.float32),
            None if zckd is None else np.ascontiguousarray(zckd, np.float32),
            None if gasabs is None else np.ascontiguousarray(gasabs, np.float32),
            None if sfcparms is None else np.asfortranarray(sfcparms, np.float32)]
    p = PropDesc(pg.npx, pg.npy, pg.npz, pg.numphase, pg.nlegp, pg.maxnmicro, pg.npart, 0 if zckd is None else len(zckd),
                 pg.nstleg, pg.delx, pg.dely, pg.xstart, pg.ystart, _lib.vp(keep[0]), _lib.vp(keep[1]), _lib.vp(pg.extinctp),
                 _lib.vp(pg.albedop), _lib.vp(pg.legenp), _lib.vp(pg.iphasep), _lib.vp(pg.phasewtp), _lib.vp(keep[2]),
                 _lib.vp(keep[3]))
    io = AdaptIO()
    io.maxig, io.maxic, io.maxiv, io.maxido, io.maxnbc, io.maxbcrad = maxig, maxic, maxiv, maxido, maxnbc, maxbcrad
    io.nbpts, io.nbcells, io.maxiter, io.accelflag, io.highorderrad = nbpts, nbcells, maxiter, int(accelflag), int(highorderrad)
    io.iterfixsh, io.inradflag, io.solacc, io.splitacc, io.shacc, io.transmin = iterfixsh, int(inradflag), solacc, splitacc, shacc, transmin
    if sfcparms is not None:
        io.nxsfc, io.nysfc, io.delxsfc, io.delysfc = keep[4].shape[1] - 1, keep[4].shape[2] - 1, delxsfc, delysfc
        io.sfcparms = _lib.vp(keep[4])
    for k, v in arr.items():
        setattr(io, k, _lib.vp(v))
    d = st.desc()
    wtmu = np.ascontiguousarray(wtmu, np.float32)
    ms = (C.c_double * 4)()
    buf = _lib.errbuf()
    _lib.check(_lib.lib().at3d_solve_adaptive(C.byref(d), C.byref(p), _lib.vp(wtmu), C.byref(io), ms, buf), buf)
    npts, ncells = io.npts, io.ncells
    st.npts, st.ncells, st.ntoppts, st.nbotpts, st.maxnbc = npts, ncells, io.ntoppts, io.nbotpts, maxnbc
    st.gridpos = np.asfortranarray(arr['gridpos'][:, :npts])
    for n in ('gridptr', 'neighptr', 'treeptr'):
        setattr(st, n, np.asfortranarray(arr[n][:, :ncells]))
    st.cellflags = arr['cellflags'][:ncells].copy()
    for n in ('extinct', 'albedo', 'planck'):
        setattr(st, n, np.asfortranarray(arr[n][:npts, :]))
    for n in ('iphase', 'phaseinterpwt'):
        setattr(st, n, np.asfortranarray(arr[n][:, :npts, :]))
    st.total_ext = arr['total_ext'][:npts].copy()
    st.dirflux = arr['dirflux'][:npts].copy()
    st.temp = arr['temp'][:npts].copy()
    st.fluxes = np.asfortranarray(arr['fluxes'][:, :npts])
    st.shptr = arr['shptr'][:npts + 1].copy()
    st.rshptr = arr['rshptr'][:npts + 2].copy()
    st.source = np.asfortranarray(arr['source'][:, :max(int(st.shptr[npts]), 1)])
    st.radiance = np.asfortranarray(arr['radiance'][:, :max(int(st.rshptr[npts]), 1)])
    st.bcptr = arr['bcptr']
    nbc = io.ntoppts + io.nbotpts * (1 if lamb else 1 + st.nang // 2)
    st.bcrad = np.asfortranarray(arr['bcrad'][:, :max(nbc, 1)])
    st.sfcgridparms = np.asfortranarray(arr['sfcgridparms'][:, :max(io.nbotpts, 1)])
    st.extdirp = arr['extdirp']
    res = (st, io.iters, io.solcrit, io.splitcrit)
    return res + (list(ms),) if timing else res


def __getattr__(name):
    # the reference keeps its RTE class in at3d/solver.py: ``at3d.solver.RTE`` resolves here too (at3d_b200/rte.py)
    if name == 'RTE':
        from .rte import RTE
        return RTE
    raise AttributeError("module '{}' has no attribute '{}'".format(__name__, name))
