"""Synthetic SHDOM states and sensor rays for parity tests and benchmarks.

No solver is involved: grids follow the reference structure exactly (``grid.py``), optical
properties go through ``medium.transfer_pa_to_grid`` (delta-M scaled like PREPARE_PROP), and
``SOURCE`` / ``RADIANCE`` are smooth random spherical-harmonic fields with power-law decay in l and
random adaptive truncation (``SHPTR``), cf. SURVEY.md section 8(c,d).  Throughput of the ray kernels
depends on the values only through TRANSCUT termination and the TAUTOL sub-stepping.
"""
import numpy as np
from . import grid as G
from . import medium as M
from .state import ShdomState, Rays


class Scene:
    def __init__(self, state, pg, meta):
        self.state, self.pg, self.meta = state, pg, meta


def hg_legendre_table(gs, nlegp, nstleg):
    """Property Legendre table LEGENP[nstleg,0:nlegp,numphase] (coefficients include 2l+1) of
    Henyey-Greenstein phase functions; the polarized elements are a smooth synthetic phase matrix."""
    numphase = len(gs)
    l = np.arange(nlegp + 1, dtype=np.float64)
    out = np.zeros((nstleg, nlegp + 1, numphase), dtype=np.float32, order='F')
    for k, g in enumerate(gs):
        a1 = (2 * l + 1) * g ** l
        out[0, :, k] = a1
        if nstleg > 1:
            lge2 = (l >= 2)
            out[1, :, k] = 0.95 * a1 * lge2
            out[2, :, k] = 0.90 * a1 * lge2
            out[3, :, k] = 0.90 * a1
            out[4, :, k] = -0.35 * (2 * l + 1) * (0.6 * g) ** l * lge2
            out[5, :, k] = 0.10 * (2 * l + 1) * (0.5 * g) ** l * lge2
    return out


def cloud_field(nx, ny, nz, kind, rng, ext_max):
    x = (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    z = (np.arange(nz)) / max(nz - 1, 1)
    X, Y, Z = np.meshgrid(x, y, z, indexing='ij')
    if kind == 'blob':
        r2 = ((X - 0.5) / 0.28) ** 2 + ((Y - 0.5) / 0.28) ** 2 + ((Z - 0.5) / 0.3) ** 2
        ext = ext_max * np.exp(-r2)
        ext[r2 > 1.6] = 0.0
    elif kind == 'les':
        # cumulus-like: power-law random field thresholded, confined to a cloud layer
        kx = np.fft.fftfreq(nx)[:, None, None]; ky = np.fft.fftfreq(ny)[None, :, None]
        kz = np.fft.fftfreq(nz)[None, None, :]
        k = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2); k[0, 0, 0] = 1.0
        spec = k ** (-11.0 / 6.0); spec[0, 0, 0] = 0.0
        f = np.fft.ifftn(spec * np.exp(2j * np.pi * rng.random((nx, ny, nz)))).real
        f = (f - f.mean()) / f.std()
        layer = np.exp(-((Z - 0.45) / 0.22) ** 4)
        ext = ext_max * np.clip(f - 0.6, 0, None) / 2.0 * layer
    else:
        ext = np.full((nx, ny, nz), ext_max)
    return ext


def make_scene(nx=8, ny=8, nz=9, nmu=8, nphi=16, nstokes=1, bc='periodic', dx=0.05, dy=0.05, dz=0.04,
               cloud='blob', ext_max=20.0, ssalb=1.0, rayleigh=False, numphase=5, mix_fraction=0.3,
               deltam=True, solarmu=-0.5, solaraz=0.3, solarflux=1.0, gndalbedo=0.05,
               nsplits=0, truncate=True, seed=0, tautol=0.1, transcut=1e-5, with_radiance=True,
               variable_sfc=False, ipflag=0):
    """Build a complete synthetic ``ShdomState`` (+ its ``PropertyGrid``)."""
    rng = np.random.default_rng(seed)
    nstleg = 1 if nstokes == 1 else 6
    ml, mm, nlm = G.sh_sizes(nmu, nphi)
    bcflag = 0
    if bc == 'open':
        bcflag = 3
    elif bc == 'open_x':          # open in X only (what at3d sets for open boundaries with independent pixels in Y)
        bcflag = 1
    npx, npy, npz = nx, ny, nz
    zlevels = (np.arange(nz) * dz).astype(np.float32)
    # ---- property grid ----
    npart = 2 if rayleigh else 1
    maxpg = npx * npy * npz
    nlegp = max(2 * (ml + 1), 180)
    gs = np.linspace(0.80, 0.87, numphase)
    legenp_cloud = hg_legendre_table(gs, nlegp, nstleg)
    tables = [legenp_cloud]
    if rayleigh:
        ray = np.zeros((nstleg, nlegp + 1, 1), np.float32, order='F')
        ray[0, 0, 0] = 1.0; ray[0, 2, 0] = 0.5
        if nstleg > 1:
            ray[1, 2, 0] = 3.0; ray[3, 1, 0] = 1.5; ray[4, 2, 0] = np.sqrt(1.5)
        tables.append(ray)
    legenp = np.asfortranarray(np.concatenate(tables, axis=2))
    extp = np.zeros((maxpg, npart), np.float32, order='F')
    albp = np.zeros((maxpg, npart), np.float32, order='F')
    iphp = np.ones((1, maxpg, npart), np.int32, order='F')
    pwp = np.ones((1, maxpg, npart), np.float32, order='F')
    e = cloud_field(npx, npy, npz, cloud, rng, ext_max)
    extp[:, 0] = e.reshape(-1)
    albp[:, 0] = ssalb
    # droplet-size like variation of the table index with height and randomly
    iz = np.tile(np.arange(npz), npx * npy)
    iphp[0, :, 0] = 1 + (iz * numphase // max(npz, 1) + rng.integers(0, 2, maxpg)) % numphase
    if rayleigh:
        zz = np.tile(zlevels, npx * npy)
        extp[:, 1] = 0.02 * np.exp(-zz / 8.0)
        albp[:, 1] = 1.0
        iphp[0, :, 1] = numphase + 1
    pg = M.PropertyGrid(npx, npy, npz, dx, dy, zlevels, extp, albp, iphp, pwp, legenp, nlegp, nstleg)
    # ---- RTE grid ----
    nx1, ny1, nbpts, nbcells = G.grid_sizes(nx, ny, nz, bcflag, ipflag)
    xg, yg, zg = G.new_grids(bcflag, 'P', npx, npy, npz, nx, ny, nz, 0.0, 0.0, dx, dy, zlevels)
    npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags = G.init_cell_structure(
        bcflag, ipflag, nx, ny, nz, nx1, ny1, xg[:nx1], yg[:ny1], zg,
        maxic=nbcells + 2 * nsplits + 2, maxig=nbpts + 4 * nsplits + 4)
    tree = G.CellTree(npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags)
    for _ in range(nsplits):
        # split a random end cell that is not an open-boundary (zero width) cell
        for _try in range(50):
            ic = int(rng.integers(1, tree.ncells + 1))
            if tree.treeptr[1, ic - 1] == 0 and (int(tree.cellflags[ic - 1]) & 3) == 0:
                tree.divide_cell(ic, int(rng.integers(1, 4)))
                break
    npts, ncells = tree.npts, tree.ncells
    gridpos = np.asfortranarray(tree.gridpos[:, :npts])
    gridptr = np.asfortranarray(tree.gridptr[:, :ncells])
    neighptr = np.asfortranarray(tree.neighptr[:, :ncells])
    treeptr = np.asfortranarray(tree.treeptr[:, :ncells])
    cellflags = tree.cellflags[:ncells].copy()
    # ---- optical properties on the RTE grid ----
    t = M.transfer_pa_to_grid(pg, gridpos, npts, ml, deltam)
    if mix_fraction > 0:
        # force some points to genuinely mix two phase tables (PHASEINTERPWT(1) < PHASEMAX branch)
        sel = np.nonzero(rng.random(npts) < mix_fraction)[0]
        w = rng.uniform(0.55, 0.9, sel.size).astype(np.float32)
        other = 1 + (t['iphase'][0, sel, 0] + rng.integers(0, numphase - 1, sel.size)) % numphase
        other = np.where(other == t['iphase'][0, sel, 0], 1 + other % numphase, other)
        t['phaseinterpwt'][0, sel, 0] = w
        t['phaseinterpwt'][1, sel, 0] = np.float32(1.0) - w
        t['iphase'][1, sel, 0] = other
    nleg = t['nleg']
    numphase_tot = legenp.shape[2]
    extdirp = M.extdirp_from_properties(pg, ml, deltam)
    dirflux = M.direct_beam_ip(pg, gridpos, npts, solarflux, solarmu, extdirp)
    # ---- angle set, boundary lists ----
    mu, phi, wtdo, nphi0, nang = M.make_angle_set(nmu, nphi)
    ntop, nbot, bcptr = G.boundary_pnts(npts, gridpos, zg[0], zg[-1])
    fluxes = np.zeros((2, npts), np.float32, order='F')
    fluxes[0] = (0.3 + 0.2 * rng.random(npts)).astype(np.float32) * abs(solarmu) * solarflux
    fluxes[1] = (0.2 + 0.2 * rng.random(npts)).astype(np.float32) * abs(solarmu) * solarflux
    bcrad = np.zeros((nstokes, ntop + nbot), np.float32, order='F')
    skyrad = np.zeros((nstokes, nmu // 2, nphi), np.float32, order='F')
    skyrad[0] = 0.01 + 0.005 * rng.random((nmu // 2, nphi))
    # ---- spherical-harmonic fields ----
    lj = G.lofj(ml, mm)

    def sh_field(scale_q):
        # per-point truncation to whole l-shells (shdomsub1.f:1583-1588)
        if truncate:
            ltr = np.where(rng.random(npts) < 0.35, rng.integers(0, ml + 1, npts), ml)
        else:
            ltr = np.full(npts, ml)
        ltr[t['total_ext'] <= 0] = np.where(rng.random(np.count_nonzero(t['total_ext'] <= 0)) < 0.5, -1, 0)
        ns = np.where(ltr < 0, 0,
                      np.where(ltr <= mm, ltr * (ltr + 1) + ltr + 1, (2 * mm + 1) * ltr - mm * mm + mm + 1))
        ptr = np.zeros(npts + 2, np.int32)
        ptr[1:npts + 1] = np.cumsum(ns)
        ptr[npts + 1] = ptr[npts]
        tot = int(ptr[npts])
        a0 = (0.05 + 0.25 * rng.random(npts)).astype(np.float32)
        arr = np.zeros((nstokes, max(tot, 1)), np.float32, order='F')
        pidx = np.repeat(np.arange(npts), ns)
        jidx = np.arange(tot) - np.repeat(ptr[:npts], ns)
        decay = (0.55 ** lj[jidx]).astype(np.float32)
        arr[0, :tot] = a0[pidx] * decay * (rng.standard_normal(tot).astype(np.float32) * 0.25)
        first = jidx == 0
        arr[0, :tot][first] = a0[pidx][first] * 3.5449077
        if nstokes > 1:
            qmask = (jidx >= 4)
            arr[1, :tot] = scale_q * a0[pidx] * decay * rng.standard_normal(tot).astype(np.float32) * qmask
            arr[2, :tot] = scale_q * a0[pidx] * decay * rng.standard_normal(tot).astype(np.float32) * qmask
        return ptr, arr

    shptr2, source = sh_field(0.05)
    rshptr, radiance = sh_field(0.03) if with_radiance else (None, None)
    st = ShdomState(
        nstokes=nstokes, nstleg=nstleg, nx=nx, ny=ny, nz=nz, npts=npts, ncells=ncells,
        ml=ml, mm=mm, nlm=nlm, nleg=nleg, numphase=numphase_tot, npart=npart, maxnmicro=1,
        bcflag=bcflag, ipflag=ipflag, nmu=nmu, nphi0max=nphi, nang=nang,
        maxnbc=bcptr.shape[0], ntoppts=ntop, nbotpts=nbot, nsfcpar=2,
        nscatangle=max(36, min(721, 2 * nlegp)), nstphase=1 if nstokes == 1 else 2,
        deltam=int(deltam), srctype='S', units='R',
        sfctype0='V' if variable_sfc else 'F', sfctype1='L', interp_new=1,
        solarmu=solarmu, solaraz=solaraz, solarflux=solarflux, wavelen=0.66, gndtemp=0.0,
        gndalbedo=gndalbedo, phasemax=0.999, waveno0=0.0, waveno1=0.0, tautol=tautol, transcut=transcut,
        gridptr=gridptr, neighptr=neighptr, treeptr=treeptr, cellflags=cellflags,
        xgrid=xg if not (bcflag & 5) else xg[:nx], ygrid=yg if not (bcflag & 10) else yg[:ny], zgrid=zg,
        gridpos=gridpos, extinct=t['extinct'], albedo=t['albedo'], total_ext=t['total_ext'],
        legen=t['legen'], iphase=t['iphase'], phaseinterpwt=t['phaseinterpwt'],
        dirflux=dirflux, fluxes=fluxes, shptr=shptr2[:npts + 1], source=source,
        rshptr=rshptr, radiance=radiance, ylmsun=None, phasetab=None,
        planck=np.zeros((npts, npart), np.float32, order='F'), temp=None,
        nphi0=nphi0, mu=mu, phi=phi, wtdo=wtdo, skyrad=skyrad, bcptr=bcptr, bcrad=bcrad,
        sfcgridparms=np.asfortranarray(np.stack([np.zeros(nbot, np.float32),
                                                 (gndalbedo * (0.5 + rng.random(nbot))).astype(np.float32)])),
        sfcgridrad=None)
    meta = dict(dx=dx, dy=dy, dz=dz, npx=npx, npy=npy, npz=npz, nlegp=nlegp, extdirp=extdirp,
                extmin=t['extmin'], scatmin=t['scatmin'], nbpts=nbpts, nbcells=nbcells,
                xmax=float(dx * (nx - 1 if bcflag & 1 else nx)), ymax=float(dy * (ny - 1 if bcflag & 2 else ny)),
                zmax=float(zg[-1]))
    return Scene(st.normalize(), pg, meta)


# ------------------------------------------------------------------------------------------
# sensor rays (geometry of at3d/sensor.py:188-300 orthographic, :367-470 perspective)
# ------------------------------------------------------------------------------------------
def orthographic_rays(scene, zenith_deg, azimuth_deg, resolution, altitude=None):
    m = scene.meta
    mu = np.cos(np.deg2rad(zenith_deg))
    phi = np.deg2rad(azimuth_deg)
    alt = m['zmax'] if altitude is None else altitude
    alpha = np.sqrt(1 - mu ** 2) * np.cos(phi) / mu
    beta = np.sqrt(1 - mu ** 2) * np.sin(phi) / mu
    xs, ys = [], []
    for xx in (0.0, m['xmax']):
        for yy in (0.0, m['ymax']):
            for zz in (0.0, m['zmax']):
                xs.append(xx - alpha * zz + alpha * alt)
                ys.append(yy - beta * zz + beta * alt)
    x = np.arange(min(xs), max(xs) + resolution, resolution)
    y = np.arange(min(ys), max(ys) + resolution, resolution)
    X, Y = np.meshgrid(x, y)
    n = X.size
    return Rays(X.ravel(), Y.ravel(), np.full(n, alt), np.full(n, mu), np.full(n, phi)), (x.size, y.size)


def perspective_rays(position, lookat, fov_deg, nx, ny, up=(0.0, 1.0, 0.0)):
    position = np.asarray(position, np.float64); lookat = np.asarray(lookat, np.float64)
    up = np.asarray(up, np.float64)
    zaxis = lookat - position; zaxis /= np.linalg.norm(zaxis)
    xaxis = np.cross(up, zaxis); xaxis /= np.linalg.norm(xaxis)
    yaxis = np.cross(zaxis, xaxis)
    rot = np.stack((xaxis, yaxis, zaxis), axis=1)
    mx = max(nx, ny)
    R = np.array([nx, ny]) / mx
    dxp, dyp = 2 * R[0] / nx, 2 * R[1] / ny
    xs, ys, zs = np.meshgrid(np.linspace(-R[0] + dxp / 2, R[0] - dxp / 2, nx),
                             np.linspace(-R[1] + dyp / 2, R[1] - dyp / 2, ny), 1.0)
    focal = 1.0 / np.tan(np.deg2rad(fov_deg) / 2.0)
    hom = np.stack([xs.ravel() / focal, ys.ravel() / focal, zs.ravel()])
    v = rot @ hom
    v /= np.linalg.norm(v, axis=0)
    mu = -v[2]
    phi = np.arctan2(v[1], v[0]) + np.pi
    n = nx * ny
    return Rays(np.full(n, position[0]), np.full(n, position[1]), np.full(n, position[2]), mu, phi), (nx, ny)


def concat_rays(rays_list):
    return Rays(*[np.concatenate([getattr(r, k) for r in rays_list])
                  for k in ('camx', 'camy', 'camz', 'cammu', 'camphi')])


# ------------------------------------------------------------------------------------------
# general BRDF surfaces (SURFACE_BRDF types, src/polarized/shdomsub2.f:1222-1301)
# ------------------------------------------------------------------------------------------
BRDF_PARAMETERS = {
    # type: rows of SFCGRIDPARMS after the Planck term: (low, high) ranges sampled per bottom point
    'L': [(0.02, 0.4)],
    'W': [(1.33, 1.33), (0.0, 0.0), (2.0, 12.0)],
    'D': [(0.1, 0.3), (0.7, 0.9), (0.1, 0.4), (0.0, 1.0), (-1.0, -1.0)],
    'O': [(2.0, 12.0), (0.0, 0.3)],
    'R': [(0.05, 0.3), (0.5, 1.0), (-0.3, -0.1)],
    'M': [(0.05, 0.3), (0.0, 0.05), (0.0, 0.1)],
}


def with_brdf_surface(state, kind, seed=0, wavelen=None):
    """Copy of a synthetic ``ShdomState`` over a variable surface of SURFACE_BRDF type `kind`: per-point BRDF
    parameters and a stored downwelling radiance BCRAD(:, ibc, 2:) for the NANG/2 downward ordinates."""
    rng = np.random.default_rng(seed)
    st = state.copy()
    nst, nbot, ntop, nh = st.nstokes, st.nbotpts, st.ntoppts, st.nang // 2
    rows = BRDF_PARAMETERS[kind]
    parms = np.zeros((1 + len(rows), nbot), np.float32, order='F')
    for i, (lo, hi) in enumerate(rows):
        parms[1 + i] = rng.uniform(lo, hi, nbot)
    bcrad = np.zeros((nst, ntop + nbot * (1 + nh)), np.float32, order='F')
    down = 0.02 + 0.06 * rng.random((nbot * nh))
    bcrad[0, ntop + nbot:] = down
    if nst > 1:
        bcrad[1, ntop + nbot:] = 0.1 * down * rng.standard_normal(nbot * nh)
        bcrad[2, ntop + nbot:] = 0.1 * down * rng.standard_normal(nbot * nh)
    st.sfctype0, st.sfctype1 = 'V', kind
    st.nsfcpar = parms.shape[0]
    st.sfcgridparms = parms
    st.bcrad = bcrad
    if wavelen is not None:
        st.wavelen = wavelen
    return st.normalize()
