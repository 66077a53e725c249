"""Property grid -> RTE grid preparation (host side, numpy, vectorised).

Follows the semantics of
* TRILIN_INTERP_PROP  (src/polarized/shdom90.f90:17-346): trilinear interpolation of extinction /
  scattering from the property grid, phase-table pointers ``IPHASE`` and weights ``PHASEINTERPWT``
  (8*MAXNMICRO entries per point per species, duplicates merged, sorted by descending weight),
  tabulated ``LEGEN = LEGENP/(2l+1)``;
* PREPARE_PROP (src/polarized/shdomsub2.f:479-608): delta-M scaling in the 'N' interpolation mode
  (diagonal elements stored as chi-f, divide by 1-f at use) and TOTAL_EXT;
* MAKE_ANGLE_SET (shdomsub2.f:1061-1144): reduced Gaussian ordinate set (ITYPE=2).

This is input preparation: float results are not claimed bit-exact
with the Fortran; everything downstream (oracle and CUDA kernels) consumes the same arrays.
"""
import numpy as np
from .grid import btest

_GOODNFFT = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16, 16, 18, 18, 20, 20, 24, 24, 24, 24,
             25, 27, 27, 30, 30, 30, 32, 32, 36, 36, 36, 36, 40, 40, 40, 40, 45, 45, 45, 45, 45, 48,
             48, 48, 50, 50, 54, 54, 54, 54, 60, 60, 60, 60, 60, 60, 64, 64, 64, 64]


def make_angle_set(nmu, nphi, itype=2):
    """mu[nmu], phi[nmu,nphi], wtdo[nmu,nphi], nphi0[nmu], nang (first NMU/2 mu's are downwelling)."""
    x, w = np.polynomial.legendre.leggauss(nmu)
    mu = x.astype(np.float32)
    wtmu = w.astype(np.float32)
    nphi0 = np.zeros(nmu, dtype=np.int32)
    phi = np.zeros((nmu, nphi), dtype=np.float32, order='F')
    wtdo = np.zeros((nmu, nphi), dtype=np.float32, order='F')
    nang = 0
    for j in range(nmu):
        if itype == 1:
            n = nphi
        else:
            n = int(0.9 + nphi * np.sqrt(1.0 - float(mu[j]) ** 2))
            n = max(1, min(nphi, _GOODNFFT[n - 1] if n <= len(_GOODNFFT) else n))
        nphi0[j] = n
        delphi = np.float32(2.0 * np.pi / n)
        phi[j, :n] = np.arange(n, dtype=np.float32) * delphi
        wtdo[j, :n] = delphi * wtmu[j]
        nang += n
    return mu, phi, wtdo, nphi0, nang


class PropertyGrid:
    """The regular property grid (``ShdomPropertyArrays`` of at3d/solver.py:25-104)."""

    def __init__(self, npx, npy, npz, delx, dely, zlevels, extinctp, albedop, iphasep, phasewtp,
                 legenp, nlegp, nstleg, xstart=0.0, ystart=0.0):
        self.npx, self.npy, self.npz = npx, npy, npz
        self.delx, self.dely = np.float32(delx), np.float32(dely)
        self.xstart, self.ystart = np.float32(xstart), np.float32(ystart)
        self.zlevels = np.asarray(zlevels, dtype=np.float32)
        self.maxpg = npx * npy * npz
        self.extinctp = np.asfortranarray(extinctp, dtype=np.float32)      # [maxpg,npart]
        self.albedop = np.asfortranarray(albedop, dtype=np.float32)        # [maxpg,npart]
        self.iphasep = np.asfortranarray(iphasep, dtype=np.int32)          # [maxnmicro,maxpg,npart]
        self.phasewtp = np.asfortranarray(phasewtp, dtype=np.float32)      # [maxnmicro,maxpg,npart]
        self.legenp = np.asfortranarray(legenp, dtype=np.float32)          # [nstleg,0:nlegp,numphase]
        self.nlegp, self.nstleg = nlegp, nstleg
        self.npart = self.extinctp.shape[1]
        self.maxnmicro = self.iphasep.shape[0]
        self.numphase = self.legenp.shape[2]


def interp_weights(pg, x, y, z):
    """COMPUTE_INTERP_WEIGHTS / TRILIN location logic, vectorised: (interpptr[8,n], wt[8,n] f64)."""
    x = np.asarray(x, np.float32); y = np.asarray(y, np.float32); z = np.asarray(z, np.float32)
    zl = pg.zlevels
    iz = np.clip(np.searchsorted(zl, z, side='right'), 1, pg.npz - 1)
    w = (z - zl[iz - 1]).astype(np.float64) / (zl[iz] - zl[iz - 1])
    w = np.clip(w, 0.0, 1.0)
    ix = ((x - pg.xstart) / pg.delx).astype(np.int64) + 1
    ix[np.abs(x - pg.xstart - np.float32(pg.npx) * pg.delx) < np.float32(0.01) * pg.delx] = pg.npx
    if np.any((ix < 1) | (ix > pg.npx)):
        raise ValueError('TRILIN: Beyond X domain')
    ixp = ix % pg.npx + 1
    u = (x - pg.xstart - pg.delx * (ix - 1).astype(np.float32)).astype(np.float64) / pg.delx
    u = np.clip(u, 0.0, 1.0); u[u < 1e-5] = 0.0; u[u > 1 - 1e-5] = 1.0
    iy = ((y - pg.ystart) / pg.dely).astype(np.int64) + 1
    iy[np.abs(y - pg.ystart - np.float32(pg.npy) * pg.dely) < np.float32(0.01) * pg.dely] = pg.npy
    if np.any((iy < 1) | (iy > pg.npy)):
        raise ValueError('TRILIN: Beyond Y domain')
    iyp = iy % pg.npy + 1
    v = (y - pg.ystart - pg.dely * (iy - 1).astype(np.float32)).astype(np.float64) / pg.dely
    v = np.clip(v, 0.0, 1.0); v[v < 1e-5] = 0.0; v[v > 1 - 1e-5] = 1.0
    npz, npy = pg.npz, pg.npy
    i1 = iz + npz * (iy - 1) + npz * npy * (ix - 1)
    i2 = iz + npz * (iy - 1) + npz * npy * (ixp - 1)
    i3 = iz + npz * (iyp - 1) + npz * npy * (ix - 1)
    i4 = iz + npz * (iyp - 1) + npz * npy * (ixp - 1)
    ptr = np.stack([i1, i2, i3, i4, i1 + 1, i2 + 1, i3 + 1, i4 + 1]).astype(np.int32)
    wt = np.stack([(1 - u) * (1 - v) * (1 - w), u * (1 - v) * (1 - w), (1 - u) * v * (1 - w), u * v * (1 - w),
                   (1 - u) * (1 - v) * w, u * (1 - v) * w, (1 - u) * v * w, u * v * w])
    return np.asfortranarray(ptr), wt


def transfer_pa_to_grid(pg, gridpos, npts, ml, deltam, phasemax=0.999):
    """Property grid -> RTE grid arrays (INTERPMETHOD 'ON').

    Returns dict(extinct, albedo, total_ext, legen, iphase, phaseinterpwt, nleg, extmin, scatmin).
    """
    x, y, z = gridpos[0, :npts], gridpos[1, :npts], gridpos[2, :npts]
    ptr, wt = interp_weights(pg, x, y, z)
    npart, mnm = pg.npart, pg.maxnmicro
    nq = 8 * mnm
    extmin = 1.0e-5 / ((float(pg.zlevels[-1]) - float(pg.zlevels[0])) / pg.npz)
    scatmin = 0.1 * extmin
    nleg = ml + 1 if deltam else ml
    nleg = min(max(nleg, 1), pg.nlegp) if not deltam else nleg
    if pg.nlegp < nleg:
        raise ValueError('property Legendre table shorter than ML+1')
    l = np.arange(nleg + 1, dtype=np.float32)
    legen = np.asfortranarray(pg.legenp[:, :nleg + 1, :] / (2 * l + 1)[None, :, None], dtype=np.float32)
    extinct = np.zeros((npts, npart), np.float32, order='F')
    albedo = np.zeros((npts, npart), np.float32, order='F')
    iphase = np.ones((nq, npts, npart), np.int32, order='F')
    pwt = np.zeros((nq, npts, npart), np.float32, order='F')
    for ipa in range(npart):
        e = pg.extinctp[:, ipa].astype(np.float64)[ptr - 1]           # [8,npts]
        a = pg.albedop[:, ipa].astype(np.float64)[ptr - 1]
        ext = (wt * e).sum(0)
        scat8 = wt * e * a
        scatter = scat8.sum(0)
        alb = np.where(ext > extmin, scatter / np.maximum(ext, 1e-300), scatter / extmin)
        extinct[:, ipa] = ext
        albedo[:, ipa] = alb
        denom = np.where(scatter >= scatmin, scatter, scatmin)
        ip = np.zeros((nq, npts), np.int64)
        pw = np.zeros((nq, npts), np.float64)
        for c in range(8):
            ip[c * mnm:(c + 1) * mnm] = pg.iphasep[:, :, ipa][:, ptr[c] - 1]
            pw[c * mnm:(c + 1) * mnm] = pg.phasewtp[:, :, ipa][:, ptr[c] - 1] * (scat8[c] / denom)[None, :]
        # merge duplicates (first occurrence keeps the sum), then sort by descending weight
        for q in range(nq):
            for q2 in range(q + 1, nq):
                same = ip[q] == ip[q2]
                pw[q] = np.where(same, pw[q] + pw[q2], pw[q])
                pw[q2] = np.where(same, 0.0, pw[q2])
        order = np.argsort(-pw, axis=0, kind='stable')
        iphase[:, :, ipa] = np.take_along_axis(ip, order, 0)
        pwt[:, :, ipa] = np.take_along_axis(pw, order, 0)
    total_ext = extinct.sum(1).astype(np.float32)
    if deltam:
        f_tab = legen[0, ml + 1, :].copy()
        legen[0, :ml + 1, :] -= f_tab[None, :]
        if pg.nstleg > 1:
            legen[1:4, :ml + 1, :] -= f_tab[None, None, :]
        total_ext = np.maximum(total_ext - extinct.sum(1), 0).astype(np.float32)
        for ipa in range(npart):
            f = np.where(pwt[0, :, ipa] >= phasemax, f_tab[iphase[0, :, ipa] - 1],
                         (f_tab[iphase[:, :, ipa] - 1] * pwt[:, :, ipa]).sum(0)).astype(np.float32)
            a0 = albedo[:, ipa].copy()
            extinct[:, ipa] = (np.float32(1.0) - a0 * f) * extinct[:, ipa]
            albedo[:, ipa] = (np.float32(1.0) - f) * a0 / (np.float32(1.0) - a0 * f)
            total_ext = total_ext + extinct[:, ipa]
    return dict(extinct=extinct, albedo=albedo, total_ext=np.asarray(total_ext, np.float32), legen=legen,
                iphase=iphase, phaseinterpwt=pwt, nleg=nleg, extmin=extmin, scatmin=scatmin)


def direct_beam_ip(pg, gridpos, npts, solarflux, solarmu, extdirp):
    """Independent-column direct beam: F0*exp(-tau_above/|mu0|) (synthetic scenes only; the
    reference's 3-D DIRECT_BEAM_PROP walk is restated in oracle/oracle_direct.c)."""
    ptr, wt = interp_weights(pg, gridpos[0, :npts], gridpos[1, :npts], gridpos[2, :npts])
    e = extdirp.reshape(pg.npx, pg.npy, pg.npz).astype(np.float64)
    dz = np.diff(pg.zlevels.astype(np.float64))
    layer = 0.5 * (e[:, :, 1:] + e[:, :, :-1]) * dz[None, None, :]
    tau_above = np.zeros_like(e)
    tau_above[:, :, :-1] = np.cumsum(layer[:, :, ::-1], axis=2)[:, :, ::-1]
    t = tau_above.reshape(-1)[ptr - 1]
    tau = (wt * t).sum(0)
    return (solarflux * np.exp(-tau / abs(solarmu))).astype(np.float32)


def extdirp_from_properties(pg, ml, deltam):
    """Delta-M scaled property-grid extinction EXTDIRP (DIRECT_BEAM_PROP INIT=1, shdom90.f90:453-483)."""
    out = np.zeros(pg.maxpg, np.float64)
    l = ml + 1
    for ipa in range(pg.npart):
        ext = pg.extinctp[:, ipa].astype(np.float64)
        alb = pg.albedop[:, ipa].astype(np.float64)
        if deltam:
            f = (pg.phasewtp[:, :, ipa].astype(np.float64) *
                 pg.legenp[0, l, :][pg.iphasep[:, :, ipa] - 1] / (2 * l + 1)).sum(0)
            ext = (1.0 - alb * f) * ext
        out += ext
    return out.astype(np.float32)
