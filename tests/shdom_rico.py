"""The reference's 3-D SHDOM verification case (tests/test_shdom.py:68-277, `solve_prop` + `Verify_Solver`):
polarized (NSTOKES=3) solve of the 32 x 36 x 27 RICO LES cloud given as an SHDOM property file with 18 tabulated
Mie phase functions, NMU=8 / NPHI=16, open horizontal boundaries, adaptive cell splitting (SPLITACC=0.1) and adaptive
SH truncation (SHACC=0.01), Lambertian surface (albedo 0.05), sun at mu0=-0.5.  The golden outputs are SHDOM's own:
``shdom_verification_source_out.out`` (first NPTS entries of SOURCE) and ``rico32x36x26w672ar.out`` (I, Q, U of
5 directions x 47 x 53 pixels at z = 1 km).

TEST INFRASTRUCTURE: host-side restatements of the reference's input preparation, each citing the routine."""
import gzip
import os
import numpy as np
from at3d_b200 import grid as G, medium as M
from at3d_b200.state import ShdomState, Rays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def read_properties(path, nstleg=6, min_nleg=8):
    """READ_PROPERTY_SIZE + READ_PROPERTIES for the 'P' (polarized tabulated phase function) format
    (src/polarized/shdomsub3.f:286-363, 367-552).  Returns a dict with LEGENP[nstleg, 0:maxleg, numphase]."""
    with gzip.open(path, 'rt') as fh:
        tok = fh.read().split('\n', 1)
    assert tok[0].startswith('P')
    vals = tok[1].split()
    pos = 0

    def take(n, conv=float):
        nonlocal pos
        out = [conv(v) for v in vals[pos:pos + n]]
        pos += n
        return out
    npx, npy, npz = take(3, int)
    delx, dely = take(2)
    zlevels = np.array(take(npz), np.float32)
    numphase = take(1, int)[0]
    tables = []
    maxleg = min_nleg
    for _ in range(numphase):
        rows = []
        for j in range(6):
            js, numl = take(2, int)
            assert js == j + 1
            rows.append(take(numl + 1))
            maxleg = max(maxleg, numl)
        tables.append(rows)
    legenp = np.zeros((nstleg, maxleg + 1, numphase), np.float32, order='F')
    for i, rows in enumerate(tables):
        for j in range(nstleg):
            legenp[j, :len(rows[j]), i] = rows[j]
    n = npx * npy * npz
    rest = np.array(vals[pos:], np.float64).reshape(-1, 7)
    ix, iy, iz = (rest[:, k].astype(np.int64) for k in range(3))
    k = (iz - 1) + npz * (iy - 1) + npz * npy * (ix - 1)
    tempp = np.zeros(n, np.float32); extinctp = np.zeros(n, np.float32); albedop = np.zeros(n, np.float32)
    iphasep = np.ones(n, np.int32)
    tempp[k] = rest[:, 3]; extinctp[k] = rest[:, 4]; albedop[k] = rest[:, 5]; iphasep[k] = rest[:, 6].astype(np.int32)
    return dict(npx=npx, npy=npy, npz=npz, delx=delx, dely=dely, zlevels=zlevels, numphase=numphase, maxleg=maxleg,
                legenp=legenp, tempp=tempp, extinctp=extinctp, albedop=albedop, iphasep=iphasep)


def make_state(oracle, nmu=8, nphi=16, nstokes=3):
    """Base-grid ShdomState + PropertyGrid + wtmu + TEMPP exactly as `solve_prop` sets the solver up."""
    pr = read_properties(os.path.join(GOLDEN, 'rico32x36x26w672.prp.gz'))
    nstleg = 1 if nstokes == 1 else 6
    ml, mm, nlm = G.sh_sizes(nmu, nphi)
    nx, ny, nz = pr['npx'], pr['npy'], pr['npz']
    bcflag, ipflag = 3, 0                                   # default_config.json: open boundaries
    maxpg = nx * ny * nz
    pg = M.PropertyGrid(nx, ny, nz, pr['delx'], pr['dely'], pr['zlevels'], pr['extinctp'].reshape(maxpg, 1),
                        pr['albedop'].reshape(maxpg, 1), pr['iphasep'].reshape(1, maxpg, 1),
                        np.ones((1, maxpg, 1), np.float32), pr['legenp'][:nstleg], pr['maxleg'], nstleg)
    nx1, ny1, nbpts, nbcells = G.grid_sizes(nx, ny, nz, bcflag, ipflag)
    xg, yg, zg = G.new_grids(bcflag, 'P', nx, ny, nz, nx, ny, nz, 0.0, 0.0, pg.delx, pg.dely, pr['zlevels'])
    npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags = G.init_cell_structure(
        bcflag, ipflag, nx, ny, nz, nx1, ny1, xg[:nx1], yg[:ny1], zg)
    assert npts == nbpts and ncells == nbcells
    t = oracle.transfer_pa_to_grid(pg, gridpos, npts, ml, True, interp_new=False, tempp=pr['tempp'], wavelen=0.672)
    mu, phi, wtdo, nphi0, nang = M.make_angle_set(nmu, nphi)
    wtmu = (wtdo[:, 0] / (np.float32(2.0 * np.pi) / nphi0.astype(np.float32))).astype(np.float32)
    ntop, nbot, bcptr = G.boundary_pnts(npts, gridpos, zg[0], zg[-1])
    solarmu, solaraz = -0.5, 0.0
    st = ShdomState(
        nstokes=nstokes, nstleg=nstleg, nx=nx, ny=ny, nz=nz, npts=npts, ncells=ncells,
        ml=ml, mm=mm, nlm=nlm, nleg=t['nleg'], numphase=pr['numphase'], npart=1, maxnmicro=1,
        bcflag=bcflag, ipflag=ipflag, nmu=nmu, nphi0max=nphi, nang=nang,
        maxnbc=bcptr.shape[0], ntoppts=ntop, nbotpts=nbot, nsfcpar=2,
        nscatangle=max(36, min(721, 2 * pr['maxleg'])), nstphase=1 if nstokes == 1 else 2,
        deltam=1, srctype='S', units='R', sfctype0='F', sfctype1='L', interp_new=0,
        solarmu=solarmu, solaraz=solaraz, solarflux=1.0, wavelen=0.672, gndtemp=288.0,
        gndalbedo=0.05, phasemax=0.999, waveno0=0.0, waveno1=0.0, tautol=0.1, transcut=1e-5,
        gridptr=np.asfortranarray(gridptr[:, :ncells]), neighptr=np.asfortranarray(neighptr[:, :ncells]),
        treeptr=np.asfortranarray(treeptr[:, :ncells]), cellflags=cellflags[:ncells].copy(),
        xgrid=xg, ygrid=yg, zgrid=zg, gridpos=np.asfortranarray(gridpos[:, :npts]),
        extinct=t['extinct'], albedo=t['albedo'], total_ext=t['total_ext'], legen=t['legen'],
        iphase=t['iphase'], phaseinterpwt=t['phaseinterpwt'],
        dirflux=None, fluxes=np.zeros((2, npts), np.float32, order='F'),
        shptr=np.zeros(npts + 1, np.int32), source=np.zeros((nstokes, 1), np.float32, order='F'),
        rshptr=np.zeros(npts + 2, np.int32), radiance=np.zeros((nstokes, 1), np.float32, order='F'),
        ylmsun=None, phasetab=None, planck=np.zeros((npts, 1), np.float32, order='F'), temp=t['temp'],
        nphi0=nphi0, mu=mu, phi=phi, wtdo=wtdo,
        skyrad=np.zeros((nstokes, nmu // 2, nphi), np.float32, order='F'),
        bcptr=bcptr, bcrad=np.zeros((nstokes, ntop + nbot), np.float32, order='F'),
        sfcgridparms=np.zeros((2, nbot), np.float32, order='F'), sfcgridrad=None)
    st.normalize()
    st.ylmsun = oracle.ylmall(True, np.float32(solarmu), np.float32(solaraz), ml, mm, nstleg, nlm)
    st.phasetab = oracle.precompute_phase_check(pg.legenp, st.nscatangle, nstokes, ml, True)
    return st, pg, wtmu, pr['tempp']


def sensor_rays():
    """tests/test_shdom.py:194-210: 47 x 53 pixel positions at z = 1 km, five (mu, phi) directions."""
    x, y = np.meshgrid(np.arange(0.0, 0.62, 0.0133), np.arange(0.0, 0.70, 0.0133))
    x = x.ravel(); y = y.ravel()
    mu = np.array([1.0, 0.5, 0.2, 0.5, 0.2])
    phi = np.array([0.0, 0.0, 0.0, 90.0, 90.0])
    return Rays(np.tile(x, 5), np.tile(y, 5), np.ones(5 * x.size), np.repeat(mu, x.size),
                np.deg2rad(np.repeat(phi, x.size)))


def golden_source():
    """SOURCE(1:3, 1:NPTS) written by SHDOM (tests/test_shdom.py:255-256)."""
    with gzip.open(os.path.join(GOLDEN, 'shdom_verification_source_out.out.gz'), 'rt') as fh:
        rows = [ln.split() for ln in fh if '*' not in ln and ln.strip()]
    return np.array(rows, np.float64).T


def golden_radiance():
    """x, y, I, Q, U rows of rico32x36x26w672ar.out (tests/test_shdom.py:264)."""
    with gzip.open(os.path.join(GOLDEN, 'rico32x36x26w672ar.out.gz'), 'rt') as fh:
        rows = [ln.split() for ln in fh if '!' not in ln and ln.strip()]
    return np.array(rows, np.float64)
