"""Host-side assembly of the LEVISAPPROX_GRADIENT inputs for a synthetic scene.

Mirrors what ``at3d.solver.RTE.calculate_microphysical_partial_derivatives`` (at3d/solver.py:1327-1515)
and ``RTE.calculate_direct_beam_derivative`` (:1266-1325) put together before
``at3d.gradient.LevisApproxGradient.levis_approximation_grad`` calls ``core.levisapprox_gradient``
(at3d/gradient.py:262-398): partial derivatives of extinction / single-scattering albedo / phase
function on the property grid, the derivative phase tables, the property->RTE interpolation tables
(``prepare_deriv_interps``) and the direct-beam path lists (``make_direct_derivative``), plus the
per-pixel measurement arrays that ``SensorsDict.sort_sensors`` adds (at3d/containers.py:296-349).

The numerical helpers come from a ``backend`` object exposing ``precompute_phase_check``,
``prepare_deriv_interps``, ``make_direct`` and ``make_direct_derivative``: ``at3d_b200.core`` (the CUDA
library) for product use, the CPU oracle binding in the parity tests so that both sides of a comparison
read identical inputs.
"""
import numpy as np
from .state import GradInputs


class PixelData:
    """Per-pixel / per-ray measurement arrays of the merged sensor (containers.py:296-349)."""

    def __init__(self, measurements, uncertainties, rays_per_pixel, ray_weights, stokes_weights):
        self.measurements = np.asfortranarray(measurements, np.float32)
        self.uncertainties = np.asfortranarray(uncertainties, np.float64)
        self.rays_per_pixel = np.ascontiguousarray(rays_per_pixel, np.int32)
        self.ray_weights = np.ascontiguousarray(ray_weights, np.float64)
        self.stokes_weights = np.asfortranarray(stokes_weights, np.float64)
        self.npix = int(self.rays_per_pixel.shape[0])

    def slice_pixels(self, p0, p1):
        """Pixels [p0,p1) and the (contiguous) rays that belong to them: (PixelData, ray_lo, ray_hi)."""
        starts = np.concatenate([[0], np.cumsum(self.rays_per_pixel)])
        r0, r1 = int(starts[p0]), int(starts[p1])
        return PixelData(self.measurements[:, p0:p1], self.uncertainties[:, :, p0:p1],
                         self.rays_per_pixel[p0:p1], self.ray_weights[r0:r1],
                         self.stokes_weights[:, p0:p1]), r0, r1


def make_pixels(nstokes, nrays, radiances, seed=0, rays_per_pixel=1, noise=0.05, nunc=None):
    """Group rays into pixels and fabricate measurements = perturbed radiances.

    rays_per_pixel: int (uniform; the remainder goes to the last pixel) or an int array."""
    rng = np.random.default_rng(seed)
    if np.isscalar(rays_per_pixel):
        k = int(rays_per_pixel)
        npix = max(nrays // k, 1)
        rpp = np.full(npix, k, np.int32)
        rpp[-1] += nrays - k * npix
    else:
        rpp = np.asarray(rays_per_pixel, np.int32)
        npix = rpp.size
        assert rpp.sum() == nrays
    starts = np.concatenate([[0], np.cumsum(rpp)])
    ray_weights = np.empty(nrays, np.float64)
    for p in range(npix):
        n = rpp[p]
        w = 0.5 + rng.random(n)
        ray_weights[starts[p]:starts[p + 1]] = w / w.sum()
    stokes_weights = np.ones((nstokes, npix), np.float64, order='F')
    pixrad = np.zeros((nstokes, npix))
    rad = np.asarray(radiances, np.float64)
    for k in range(nstokes):
        pixrad[k] = np.add.reduceat(rad[k] * ray_weights, starts[:-1])
    meas = pixrad * (1.0 + noise * rng.standard_normal((nstokes, npix)))
    meas[0] = np.maximum(meas[0], 1e-4 * max(np.abs(pixrad[0]).max(), 1e-12))
    nunc = nstokes if nunc is None else nunc
    unc = np.zeros((nunc, nunc, npix), np.float64, order='F')
    for k in range(nunc):
        unc[k, k, :] = 1.0 / (0.03 * max(np.abs(pixrad[0]).max(), 1e-12)) ** 2 * (1.0 + 0.3 * rng.random(npix))
    if nunc > 1:
        unc[0, 1, :] = unc[1, 0, :] = 0.1 * unc[0, 0, :]
    return PixelData(meas.astype(np.float32), unc, rpp, ray_weights, stokes_weights)


def make_gradient_inputs(scene, backend, seed=0, numder=2, exact_single_scatter=True, singlescatter=False,
                         costfunc='L2', exact_phase_derivative=False, maxsubgridints=0, stream_beam=False):
    """Derivative tables for ``numder`` unknowns of the scene.

    Unknown 1 is the extinction of species 1 (dext=1 where there is cloud); unknown 2 varies extinction,
    albedo and the phase-table weights together (an effective-radius-like variable, derivative_method
    'table'); unknown 3, when ``exact_phase_derivative``, uses derivative_method 'exact' with its own
    derivative Legendre table; with a Rayleigh species present the last unknown belongs to species 2.
    """
    st, pg = scene.state, scene.pg
    rng = np.random.default_rng(seed + 1000)
    maxpg, npart, mnm = pg.maxpg, pg.npart, pg.maxnmicro
    nstleg, nleg, nlegp = st.nstleg, st.nleg, pg.nlegp
    partder = np.ones(numder, np.int32)
    doexact = np.zeros(numder, np.int32)
    if npart > 1 and numder > 1:
        partder[-1] = 2
    if exact_phase_derivative:
        doexact[min(2, numder - 1)] = 1
    dmnm = mnm
    dext = np.zeros((maxpg, numder), np.float32, order='F')
    dalb = np.zeros((maxpg, numder), np.float32, order='F')
    diphasep = np.ones((dmnm, maxpg, numder), np.int32, order='F')
    dphasewtp = np.zeros((dmnm, maxpg, numder), np.float32, order='F')
    cloudy = pg.extinctp[:, 0] > 0
    ndleg = 0
    for i in range(numder):
        ipa = partder[i] - 1
        if i == 0:
            dext[:, i] = np.where(cloudy, 1.0, 0.0)
        else:
            dext[:, i] = (pg.extinctp[:, ipa] * (0.2 + 0.3 * rng.random(maxpg))).astype(np.float32)
            dalb[:, i] = (-0.02 * rng.random(maxpg) * (pg.albedop[:, ipa] > 0)).astype(np.float32)
        if doexact[i]:
            ndleg += 2
            diphasep[:, :, i] = rng.integers(ndleg - 1, ndleg + 1, (dmnm, maxpg))
        else:
            diphasep[:, :, i] = pg.iphasep[:dmnm, :, ipa]
            if i > 0:
                dphasewtp[:, :, i] = (0.3 * rng.standard_normal((dmnm, maxpg))).astype(np.float32)
    dnumphase = max(ndleg, 1)
    # derivative Legendre tables (coefficients include 2l+1 like the property tables)
    l = np.arange(nlegp + 1, dtype=np.float64)
    dlegp = np.zeros((nstleg, nlegp + 1, dnumphase), np.float32, order='F')
    for k in range(ndleg):
        gk = 0.8 + 0.03 * k
        dlegp[0, :, k] = (2 * l + 1) * l * gk ** np.maximum(l - 1, 0) * 0.05
        if nstleg > 1:
            dlegp[1, :, k] = 0.9 * dlegp[0, :, k] * (l >= 2)
            dlegp[2, :, k] = 0.8 * dlegp[0, :, k] * (l >= 2)
            dlegp[3, :, k] = 0.8 * dlegp[0, :, k]
            dlegp[4, :, k] = -0.2 * dlegp[0, :, k] * (l >= 2)
            dlegp[5, :, k] = 0.05 * dlegp[0, :, k] * (l >= 2)
    # solver.py:1417-1419: dleg[0,0,:]=0; dleg/(2l+1); phase LUT from the full table, then truncate
    dleg_full = np.asfortranarray(dlegp / (2 * l + 1)[None, :, None], dtype=np.float32)
    dleg_full[0, 0, :] = 0.0
    if ndleg:
        dphasetab = backend.precompute_phase_check(dleg_full, st.nscatangle, st.nstokes, st.ml,
                                                   deltam=bool(st.deltam), negcheck=False, grad=True)
    else:
        dphasetab = np.zeros((st.nstphase, dnumphase, st.nscatangle), np.float32, order='F')
    dleg = np.asfortranarray(dleg_full[:, :nleg + 1, :])
    gi = GradInputs(
        npix=0, maxpg=maxpg, numder=numder, dnumphase=dnumphase, deriv_maxnmicro=dmnm,
        longest_path_pts=1, nuncertainty=st.nstokes, maxsubgridints=maxsubgridints,
        exact_single_scatter=int(exact_single_scatter), singlescatter=int(singlescatter),
        costfunc_ll=1 if costfunc == 'LL' else 0,
        extmin=scene.meta['extmin'], scatmin=scene.meta['scatmin'],
        partder=partder, doexact=doexact, dext=dext, dalb=dalb,
        dleg=dleg, dphasetab=dphasetab, diphasep=diphasep, dphasewtp=dphasewtp,
        iphasep=pg.iphasep, phasewtp=pg.phasewtp, extinctp=pg.extinctp, albedop=pg.albedop,
        dtemp=np.zeros((maxpg, numder), np.float32, order='F'))
    gi.normalize()
    optw, iptr, dalbm, dextm, dfj = backend.prepare_deriv_interps(st, pg, gi)
    gi.optinterpwt, gi.interpptr, gi.dalbm, gi.dextm, gi.dfj = optw, iptr, dalbm, dextm, dfj
    _beam_lists(gi, st, pg, backend, exact_single_scatter, stream_beam)
    return gi.normalize()


_BEAM_NAMES = ['cx', 'cy', 'cz', 'cxinv', 'cyinv', 'czinv', 'epss', 'epsz', 'xdomain', 'ydomain', 'uniformzlev', 'delxd', 'delyd']
STREAM_BEAM_BYTES = 4 << 30     # dense DPATH/DPTR lists larger than this are not built: the beam walks stream instead


def _beam_lists(gi, st, pg, backend, exact_single_scatter, stream_beam):
    """Direct-beam derivative inputs: the dense lists of MAKE_DIRECT -> MAKE_DIRECT_DERIVATIVE (shdomsub5.f:1553), or --
    `stream_beam` True, or None with the CUDA backend -- only the beam constants, so that the gradient call walks the paths
    itself (at3d_grad_desc.beam_*; the oracle takes the lists).  STREAM_BEAM_BYTES: the size above which bench.py does not
    build the lists for its CPU arm."""
    gi.longest_path_pts = 1
    if not exact_single_scatter:
        gi.dpath = np.zeros((1, st.npts), np.float32, order='F')
        gi.dptr = np.zeros((1, st.npts), np.int32, order='F')
        return
    _, _, c = backend.make_direct(st, pg)
    if stream_beam is None:
        # the CUDA backend walks the paths inside the gradient call: no LONGEST_PATH_PTS x NPTS lists to build, read back and
        # upload again per cost-function evaluation (185 MB and ~40 ms at BASELINE configs[1] for +0.2 ms per gradient call)
        stream_beam = getattr(backend, '__name__', '').startswith('at3d_b200')
    if stream_beam:
        _set_streaming(gi, pg, c)
        return
    dpath, dptr = backend.make_direct_derivative(st, pg, c)
    gi.longest_path_pts = int(dpath.shape[0])
    gi.dpath, gi.dptr = dpath, dptr


def _set_streaming(gi, pg, c):
    gi.longest_path_pts = int(c['longest_path_pts'])
    gi.dpath = gi.dptr = None
    gi.beam_npx, gi.beam_npy, gi.beam_npz = int(pg.npx), int(pg.npy), int(pg.npz)
    gi.beam_xstart, gi.beam_ystart = float(pg.xstart), float(pg.ystart)
    gi.beam_zlevels = np.ascontiguousarray(pg.zlevels, np.float32)
    gi.beam_d = np.array([c[k] for k in _BEAM_NAMES], np.float64)
    gi.beam_i = np.array([c['ipdirect'], c['di'], c['dj'], c['dk'], c['longest_path_pts']], np.int32)


def with_streaming_beam(gi, state, pg, backend):
    """Copy of ``gi`` without the dense DPATH/DPTR lists: the gradient call walks the direct-beam paths itself
    (at3d_grad_desc.beam_*; same terms in the same order)."""
    out = gi.copy()
    _, _, c = backend.make_direct(state, pg)
    _set_streaming(out, pg, c)
    return out


def optical_gradient_inputs(state, pg, backend, partder, doexact, dext, dalb, diphasep, dphasewtp, dleg, dphasetab,
                            extmin, scatmin, exact_single_scatter=True, costfunc='L2', maxsubgridints=0, stream_beam=None):
    """The derivative tables LEVISAPPROX_GRADIENT takes (what at3d.solver.RTE.calculate_microphysical_partial_derivatives
    leaves on the solver, at3d/solver.py:1327-1517) from the partial derivatives of the optical properties on the property
    grid: `partder` [numder] species (1-based) of every unknown, `doexact` [numder] 1 where the phase derivative comes from
    the DLEG table, `dext` / `dalb` [maxpg, numder], `diphasep` / `dphasewtp` [deriv_maxnmicro, maxpg, numder],
    `dleg` [nstleg, nleg+1, dnumphase] and its `dphasetab` [nstphase, dnumphase, nscatangle].  Runs PREPARE_DERIV_INTERPS and
    (exact single scatter) MAKE_DIRECT_DERIVATIVE on the GPU."""
    st = state
    numder = int(np.asarray(partder).size)
    maxpg = pg.maxpg
    gi = GradInputs(
        npix=0, maxpg=maxpg, numder=numder, dnumphase=int(dleg.shape[2]), deriv_maxnmicro=int(diphasep.shape[0]),
        longest_path_pts=1, nuncertainty=st.nstokes, maxsubgridints=maxsubgridints,
        exact_single_scatter=int(exact_single_scatter), singlescatter=0,
        costfunc_ll=1 if costfunc == 'LL' else 0, extmin=extmin, scatmin=scatmin,
        partder=np.asarray(partder, np.int32), doexact=np.asarray(doexact, np.int32),
        dext=np.asfortranarray(dext, np.float32), dalb=np.asfortranarray(dalb, np.float32),
        dleg=np.asfortranarray(dleg, np.float32), dphasetab=np.asfortranarray(dphasetab, np.float32),
        diphasep=np.asfortranarray(diphasep, np.int32), dphasewtp=np.asfortranarray(dphasewtp, np.float32),
        iphasep=pg.iphasep, phasewtp=pg.phasewtp, extinctp=pg.extinctp, albedop=pg.albedop,
        dtemp=np.zeros((maxpg, numder), np.float32, order='F'))
    gi.normalize()
    optw, iptr, dalbm, dextm, dfj = backend.prepare_deriv_interps(st, pg, gi)
    gi.optinterpwt, gi.interpptr, gi.dalbm, gi.dextm, gi.dfj = optw, iptr, dalbm, dextm, dfj
    _beam_lists(gi, st, pg, backend, exact_single_scatter, stream_beam)
    return gi.normalize()


def extinction_gradient_inputs(state, pg, backend, species, extmin, scatmin, exact_single_scatter=True,
                               costfunc='L2', maxsubgridints=0, variables=None, stream_beam=None):
    """Derivative tables for the optical unknowns "extinction (or single-scattering albedo) of species k" (k in
    `species`, 0-based; `variables` per unknown, default all 'extinction', the unknown of BASELINE.json configs[1]):
    d(extinction)/d(unknown) = 1 on the property grid for 'extinction', d(ssalb)/d(unknown) = 1 for 'ssalb', phase function
    unchanged (what at3d.solver.RTE.calculate_microphysical_partial_derivatives produces for the optical variables
    'extinction' and 'ssalb', at3d/solver.py:1327-1517; at3d/medium.py OpticalDerivativeGenerator)."""
    st = state
    numder = len(species)
    variables = list(variables) if variables is not None else ['extinction'] * numder
    maxpg, mnm = pg.maxpg, pg.maxnmicro
    dext = np.zeros((maxpg, numder), np.float32, order='F')
    dalb = np.zeros((maxpg, numder), np.float32, order='F')
    diphasep = np.ones((mnm, maxpg, numder), np.int32, order='F')
    dphasewtp = np.zeros((mnm, maxpg, numder), np.float32, order='F')
    for i, (k, var) in enumerate(zip(species, variables)):
        if var == 'extinction':
            dext[:, i] = 1.0
        elif var == 'ssalb':
            dalb[:, i] = 1.0
        else:
            raise NotImplementedError("optical unknown '%s' (supported: 'extinction', 'ssalb'; phase-function unknowns go "
                                      "through RTE.calculate_microphysical_partial_derivatives)" % var)
        diphasep[:, :, i] = pg.iphasep[:, :, k]
    return optical_gradient_inputs(
        st, pg, backend, [k + 1 for k in species], np.zeros(numder, np.int32), dext, dalb, diphasep, dphasewtp,
        np.zeros((st.nstleg, st.nleg + 1, 1), np.float32, order='F'),
        np.zeros((st.nstphase, 1, st.nscatangle), np.float32, order='F'), extmin, scatmin,
        exact_single_scatter=exact_single_scatter, costfunc=costfunc, maxsubgridints=maxsubgridints, stream_beam=stream_beam)


def with_pixels(gi, pix):
    """Copy of ``gi`` carrying the per-pixel arrays (for backends that take one flat structure)."""
    out = gi.copy()
    out.npix = pix.npix
    out.nuncertainty = int(pix.uncertainties.shape[0])
    out.measurements = pix.measurements
    out.uncertainties = pix.uncertainties
    out.rays_per_pixel = pix.rays_per_pixel
    out.ray_weights = pix.ray_weights
    out.stokes_weights = pix.stokes_weights
    return out.normalize()
