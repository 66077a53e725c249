"""Inputs of the reference's SHDOM surface verification cases (tests/test_shdom.py:29-66, 597-805):
a 50 x 1 x 11 independent-pixel Rayleigh atmosphere (0.85 um, sun at mu0=-0.707) over a surface whose BRDF
parameters vary along x.  The golden outputs (tests/golden/brdf_*1{f,r}.out) are SHDOM's own.

TEST INFRASTRUCTURE: host-side restatements of the reference's input preparation, each citing the routine."""
import numpy as np
from at3d_b200 import grid as G, medium as M
from at3d_b200.state import ShdomState, Rays

ZLEV = np.array([0, 3.0, 6.0, 9.0, 12.0, 15.0, 18.0, 21.0, 24.0, 27.0, 30.0])
TEMP = np.array([288.0, 269.0, 249.0, 230.0, 217.0, 217.0, 217.0, 218.0, 221.0, 224.0, 227.0])
NX = 50


def rayleigh_extinct(zlevels, temp, raysfcpres, raylcoef):
    """RAYLEIGH_EXTINCT (src/shdomsub5.f:1457-1495) in REAL arithmetic."""
    f = np.float32
    z = np.asarray(zlevels, f); t = np.asarray(temp, f)
    n = z.size
    ext = np.zeros(n, f)
    pres = f(raysfcpres)
    lapse = f(6.5) * f(0.001)
    pres = pres * (t[0] / (t[0] + lapse * z[0] * f(1000.))) ** f(f(9.8) / (f(287.) * lapse))
    for i in range(n - 1):
        ext[i] = f(raylcoef) * pres / t[i]
        dz = f(1000.) * (z[i + 1] - z[i])
        lapse = (t[i] - t[i + 1]) / dz
        if abs(lapse) > f(0.00001):
            pres = pres * f((t[i + 1] / t[i]) ** f(f(9.8) / (f(287.) * lapse)))
        else:
            pres = pres * f(np.exp(f(-9.8) * dz / (f(287.) * t[i])))
    ext[n - 1] = f(raylcoef) * pres / t[n - 1]
    return ext


def rayleigh_coefficient(wavelength, surface_pressure):
    """at3d/rayleigh.py:163-166 (Bodhaine et al. 1999)."""
    w = float(wavelength)
    return 0.03370 * (surface_pressure / 1013.25) * 0.0021520 * (
        1.0455996 - 341.29061 / w ** 2 - 0.90230850 * w ** 2) / (1 + 0.0027059889 / w ** 2 - 85.968563 * w ** 2)


def rayleigh_phase_function(wavelen, nlegp, nstleg):
    """RAYLEIGH_PHASE_FUNCTION (src/polarized/shdomsub4.f:2350-2384) -> LEGENP[nstleg,0:nlegp,1]."""
    f = np.float32
    fking = f(1.0469541 + 3.2503153e-04 / f(wavelen) ** 2 + 3.8622851e-05 / f(wavelen) ** 4)
    depol = f(6.0 * (float(fking) - 1.0) / (3.0 + 7.0 * float(fking)))
    delta = (f(1.0) - depol) / (f(1.0) + f(0.5) * depol)
    deltap = (f(1.0) - f(2.0) * depol) / (f(1.0) - depol)
    t = np.zeros((6, nlegp + 1, 1), f, order='F')
    t[0, 0] = 1.0; t[0, 2] = f(0.5) * delta
    t[1, 2] = f(3.0) * delta
    t[3, 1] = f(1.5) * deltap * delta
    t[4, 2] = f(np.sqrt(f(1.5))) * delta
    return np.asfortranarray(t[:nstleg])


def prep_surface(parms_in):
    """PREP_SURFACE (src/surface.f:1-288): parms_in[npar, nxsfc] (uniform in y, NYSFC=1) ->
    SFCPARMS[npar, nxsfc+1, nysfc+1] with the periodic edge copies."""
    npar, nxs = parms_in.shape
    out = np.zeros((npar, nxs + 1, 2), np.float32, order='F')
    out[:, :nxs, 0] = parms_in
    out[:, :nxs, 1] = out[:, :nxs, 0]
    out[:, nxs, :] = out[:, 0, :]
    return out


def surface_parm_interp(bcptr_bot, nbot, gridpos, sfcparms, delxsfc, delysfc, srctype='S'):
    """SURFACE_PARM_INTERP (src/polarized/shdomsub1.f:2221-2275) for a solar source (Planck term = 0)."""
    f = np.float32
    npar, nx1, ny1 = sfcparms.shape
    nxs, nys = nx1 - 1, ny1 - 1
    out = np.zeros((npar, nbot), f, order='F')
    for ibc in range(nbot):
        i = bcptr_bot[ibc] - 1
        rx = f(gridpos[0, i]) / f(delxsfc)
        ry = f(gridpos[1, i]) / f(delysfc)
        ix = max(1, min(nxs, int(rx) + 1))
        iy = max(1, min(nys, int(ry) + 1))
        u = f(max(0.0, min(1.0, rx - f(ix - 1))))
        v = f(max(0.0, min(1.0, ry - f(iy - 1))))
        out[:, ibc] = ((1 - u) * (1 - v) * sfcparms[:, ix - 1, iy - 1] + (1 - u) * v * sfcparms[:, ix - 1, iy]
                       + u * (1 - v) * sfcparms[:, ix, iy - 1] + u * v * sfcparms[:, ix, iy])
        assert srctype == 'S'
        out[0, ibc] = 0.0
    return out


def ramp(lo, hi):
    return np.linspace(lo, hi - (hi - lo) / NX, NX)


SURFACES = {
    # name: (SFCTYPE(2:2), NSTOKES, parameter rows after the temperature)     tests/test_shdom.py:597-790
    'L': ('L', 1, [ramp(0.0, 0.3)]),
    'O': ('O', 1, [ramp(4.0, 12.0), np.zeros(NX)]),
    'R': ('R', 1, [np.full(NX, 0.1), ramp(0.5, 1.0), np.full(NX, -0.24)]),
    'W': ('W', 3, [np.full(NX, 1.33), np.zeros(NX), ramp(4.0, 12.0)]),
    'D': ('D', 3, [np.full(NX, 0.2), np.full(NX, 0.8), np.full(NX, 0.3), ramp(0.0, 1.0), np.full(NX, -1.0)]),
}


def make_state(oracle, kind, nmu=16, nphi=32, wavelen=0.85, solarmu=-0.707, solaraz=0.0):
    """Unsolved ShdomState + PropertyGrid + wtmu for surface case `kind`."""
    sfc1, nstokes, rows = SURFACES[kind]
    nstleg = 1 if nstokes == 1 else 6
    ml, mm, nlm = G.sh_sizes(nmu, nphi)
    dx = 0.02
    nx, ny, nz = NX, 1, ZLEV.size
    bcflag, ipflag = 0, 3
    # --- medium (get_basic_state_for_surface, tests/test_shdom.py:29-51): Rayleigh, extinction rounded to 4 digits
    ext = rayleigh_extinct(ZLEV, TEMP, 1013.25, rayleigh_coefficient(wavelen, 1013.25))
    ext = np.round(ext.astype(np.float32), 4)
    nlegp = ml + 1
    legenp = rayleigh_phase_function(wavelen, nlegp, nstleg)
    maxpg = nx * ny * nz
    extp = np.tile(ext, nx * ny).reshape(maxpg, 1)
    albp = np.ones((maxpg, 1), np.float32)
    pg = M.PropertyGrid(nx, ny, nz, dx, dx, ZLEV, extp, albp, np.ones((1, maxpg, 1), np.int32),
                        np.ones((1, maxpg, 1), np.float32), legenp, nlegp, nstleg)
    nx1, ny1, nbpts, nbcells = G.grid_sizes(nx, ny, nz, bcflag, ipflag)
    xg, yg, zg = G.new_grids(bcflag, 'P', nx, ny, nz, nx, ny, nz, 0.0, 0.0, dx, dx, ZLEV)
    npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags = G.init_cell_structure(
        bcflag, ipflag, nx, ny, nz, nx1, ny1, xg[:nx1], yg[:ny1], zg)
    t = M.transfer_pa_to_grid(pg, gridpos, npts, ml, True)
    mu, phi, wtdo, nphi0, nang = M.make_angle_set(nmu, nphi)
    wtmu = (wtdo[:, 0] / (np.float32(2.0 * np.pi) / nphi0.astype(np.float32))).astype(np.float32)
    ntop, nbot, bcptr = G.boundary_pnts(npts, gridpos, zg[0], zg[-1])
    parms_in = np.stack([np.full(NX, 288.0)] + rows).astype(np.float32)
    sfcparms = prep_surface(parms_in)
    sfcgridparms = surface_parm_interp(bcptr[:, 1], nbot, gridpos, sfcparms, 0.02, 0.04)
    st = ShdomState(
        nstokes=nstokes, nstleg=nstleg, nx=nx, ny=ny, nz=nz, npts=npts, ncells=ncells,
        ml=ml, mm=mm, nlm=nlm, nleg=t['nleg'], numphase=1, npart=1, maxnmicro=1,
        bcflag=bcflag, ipflag=ipflag, nmu=nmu, nphi0max=nphi, nang=nang,
        maxnbc=bcptr.shape[0], ntoppts=ntop, nbotpts=nbot, nsfcpar=parms_in.shape[0],
        nscatangle=max(36, min(721, 2 * nlegp)), nstphase=1 if nstokes == 1 else 2,
        deltam=1, srctype='S', units='R', sfctype0='V', sfctype1=sfc1, interp_new=1,
        solarmu=solarmu, solaraz=solaraz, solarflux=1.0, wavelen=wavelen, gndtemp=288.0,
        gndalbedo=float(np.mean(rows[0])), phasemax=0.999, waveno0=0.0, waveno1=0.0, tautol=0.1, transcut=1e-5,
        gridptr=np.asfortranarray(gridptr[:, :ncells]), neighptr=np.asfortranarray(neighptr[:, :ncells]),
        treeptr=np.asfortranarray(treeptr[:, :ncells]), cellflags=cellflags[:ncells].copy(),
        xgrid=xg, ygrid=yg, zgrid=zg, gridpos=np.asfortranarray(gridpos[:, :npts]),
        extinct=t['extinct'], albedo=t['albedo'], total_ext=t['total_ext'], legen=t['legen'],
        iphase=t['iphase'], phaseinterpwt=t['phaseinterpwt'],
        dirflux=None, fluxes=np.zeros((2, npts), np.float32, order='F'),
        shptr=np.zeros(npts + 1, np.int32), source=np.zeros((nstokes, 1), np.float32, order='F'),
        rshptr=np.zeros(npts + 2, np.int32), radiance=np.zeros((nstokes, 1), np.float32, order='F'),
        ylmsun=None, phasetab=None, planck=np.zeros((npts, 1), np.float32, order='F'), temp=None,
        nphi0=nphi0, mu=mu, phi=phi, wtdo=wtdo,
        skyrad=np.zeros((nstokes, nmu // 2, nphi), np.float32, order='F'),
        bcptr=bcptr, bcrad=np.zeros((nstokes, ntop + nbot * (1 + nang // 2)), np.float32, order='F'),
        sfcgridparms=sfcgridparms, sfcgridrad=np.zeros((nang // 2 + 1, bcptr.shape[0]), np.float32, order='F'))
    st.normalize()
    st.ylmsun = oracle.ylmall(True, np.float32(solarmu), np.float32(solaraz), ml, mm, nstleg, nlm)
    st.phasetab = oracle.precompute_phase_check(legenp, st.nscatangle, nstokes, ml, True)
    st.dirflux, _, _ = oracle.make_direct(st, pg)
    return st, pg, wtmu


def sensor_rays():
    """The 19 x 50 rays of get_basic_state_for_surface (tests/test_shdom.py:53-66)."""
    x = np.linspace(0, 1.0 - 1.0 / 50, 50)
    mu = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9] + [1.0] + [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9][::-1])
    phi = np.array([180.0] * 9 + [0.0] * 10)
    return Rays(np.tile(x, 19), np.zeros(950), np.full(950, 30.0), np.repeat(mu, 50), np.deg2rad(np.repeat(phi, 50)))


def parse_shdom_output(filename, comment='!'):
    """tests/test_shdom.py:17-27"""
    rows = []
    with open(filename) as fh:
        for line in fh:
            if comment not in line and line.strip():
                rows.append(np.array(line.split(), dtype=np.float64))
    return np.array(rows)


def planck_radiance(temperature, wavelength):
    """at3d.util.planck_function (at3d/util.py:336-347): W / m^2 / micron / sr."""
    c, h, k = 2.99792458e8, 6.62606876e-34, 1.3806503e-23
    w = wavelength * 1e-6
    return 2 * h * c ** 2 / w ** 5 / (np.exp(h * c / (w * k * temperature)) - 1.0) * 1e-6


def make_thermal_state(oracle, nmu=16, nphi=32, wavelen=11.0, tair=288.0, tsfc=300.0, albedo=0.5):
    """The isothermal absorbing slab of the reference's Verify_Thermal (tests/test_shdom.py:910-982): 50 columns
    with extinction 0.001 ... 0.5 /km (no scattering) at 288 K over a Lambertian surface (albedo 0.5) at 300 K,
    thermal source at 11 um.  The reference runs it with ip_flag=1 (y is a single periodic column); ip_flag=3 is
    the same medium for the 1-D sweep."""
    nstokes, nstleg = 1, 1
    ml, mm, nlm = G.sh_sizes(nmu, nphi)
    dx = 0.02
    nx, ny, nz = NX, 1, ZLEV.size
    bcflag, ipflag = 0, 3
    nlegp = ml
    legenp = rayleigh_phase_function(wavelen, max(nlegp, 2), nstleg)
    maxpg = nx * ny * nz
    extp = np.repeat(np.linspace(0.001, 0.5, nx), nz).reshape(maxpg, 1).astype(np.float32)
    albp = np.zeros((maxpg, 1), np.float32)
    pg = M.PropertyGrid(nx, ny, nz, dx, dx, ZLEV, extp, albp, np.ones((1, maxpg, 1), np.int32),
                        np.ones((1, maxpg, 1), np.float32), legenp, max(nlegp, 2), nstleg)
    nx1, ny1, nbpts, nbcells = G.grid_sizes(nx, ny, nz, bcflag, ipflag)
    xg, yg, zg = G.new_grids(bcflag, 'P', nx, ny, nz, nx, ny, nz, 0.0, 0.0, dx, dx, ZLEV)
    npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags = G.init_cell_structure(
        bcflag, ipflag, nx, ny, nz, nx1, ny1, xg[:nx1], yg[:ny1], zg)
    t = M.transfer_pa_to_grid(pg, gridpos, npts, ml, False)
    mu, phi, wtdo, nphi0, nang = M.make_angle_set(nmu, nphi)
    wtmu = (wtdo[:, 0] / (np.float32(2.0 * np.pi) / nphi0.astype(np.float32))).astype(np.float32)
    ntop, nbot, bcptr = G.boundary_pnts(npts, gridpos, zg[0], zg[-1])
    # PREPARE_PROP (shdomsub2.f:575-578): PLANCK = (1-ALBEDO)*B(TEMP), B from PLANCK_FUNCTION in REAL
    f = np.float32
    bb = f(f(1.1911e8) / f(wavelen) ** 5 / (np.exp(f(1.4388e4) / (f(wavelen) * f(tair))) - f(1)))
    planck = ((f(1.0) - t['albedo']) * bb).astype(np.float32)
    st = ShdomState(
        nstokes=nstokes, nstleg=nstleg, nx=nx, ny=ny, nz=nz, npts=npts, ncells=ncells,
        ml=ml, mm=mm, nlm=nlm, nleg=t['nleg'], numphase=1, npart=1, maxnmicro=1,
        bcflag=bcflag, ipflag=ipflag, nmu=nmu, nphi0max=nphi, nang=nang,
        maxnbc=bcptr.shape[0], ntoppts=ntop, nbotpts=nbot, nsfcpar=2,
        nscatangle=36, nstphase=1, deltam=0, srctype='T', units='R', sfctype0='F', sfctype1='L', interp_new=1,
        solarmu=-1.0, solaraz=0.0, solarflux=0.0, wavelen=wavelen, gndtemp=tsfc, gndalbedo=albedo,
        phasemax=0.999, waveno0=0.0, waveno1=0.0, tautol=0.1, transcut=1e-5,
        gridptr=np.asfortranarray(gridptr[:, :ncells]), neighptr=np.asfortranarray(neighptr[:, :ncells]),
        treeptr=np.asfortranarray(treeptr[:, :ncells]), cellflags=cellflags[:ncells].copy(),
        xgrid=xg, ygrid=yg, zgrid=zg, gridpos=np.asfortranarray(gridpos[:, :npts]),
        extinct=t['extinct'], albedo=t['albedo'], total_ext=t['total_ext'], legen=t['legen'],
        iphase=t['iphase'], phaseinterpwt=t['phaseinterpwt'],
        dirflux=np.zeros(npts, np.float32), fluxes=np.zeros((2, npts), np.float32, order='F'),
        shptr=np.zeros(npts + 1, np.int32), source=np.zeros((nstokes, 1), np.float32, order='F'),
        rshptr=np.zeros(npts + 2, np.int32), radiance=np.zeros((nstokes, 1), np.float32, order='F'),
        ylmsun=np.zeros((nstleg, nlm), np.float32, order='F'), phasetab=np.zeros((1, 1, 36), np.float32, order='F'),
        planck=planck, temp=np.full(npts, tair, np.float32),
        nphi0=nphi0, mu=mu, phi=phi, wtdo=wtdo,
        skyrad=np.zeros((nstokes, nmu // 2, nphi), np.float32, order='F'),
        bcptr=bcptr, bcrad=np.zeros((nstokes, ntop + nbot), np.float32, order='F'),
        sfcgridparms=np.zeros((2, nbot), np.float32, order='F'), sfcgridrad=None)
    st.normalize()
    return st, pg, wtmu


def thermal_slab_radiance(wavelen=11.0, tair=288.0, tsfc=300.0, albedo=0.5):
    """Closed form of tests/test_shdom.py:965-975 for the nadir radiance at the top of the slab."""
    from scipy.special import exp1
    tau = 30.0 * np.linspace(0.001, 0.5, NX)
    tr = np.exp(-tau)
    ba, bs = planck_radiance(tair, wavelen), planck_radiance(tsfc, wavelen)
    return (1.0 - albedo) * bs * tr + ba * (1.0 - tr) + tr * (albedo * ba * (-tr * (1.0 - tau) - tau ** 2 * exp1(tau) + 1))


def nadir_rays():
    x = np.linspace(0, 1.0 - 1.0 / 50, 50)
    return Rays(x, np.zeros(50), np.full(50, 30.0), np.ones(50), np.zeros(50))


def make_absorbing_state(oracle, nmu=16, nphi=32):
    """Verify_NonuniformGasAbsorption of the reference (tests/test_shdom.py:855-908): purely absorbing columns with
    absorption 0 ... 1 /km, overhead sun, Lambertian albedo 0.04, TRANSCUT=0; I = e^-tau * 0.04/pi * e^-tau.
    The absorber is entered as a non-scattering species (the reference enters it as gas absorption)."""
    st, pg, wtmu = make_thermal_state(oracle, nmu, nphi, wavelen=0.85)
    npts, nz = st.npts, ZLEV.size
    ext = np.repeat(np.linspace(0.0, 1.0, NX), nz).astype(np.float32)
    pg.extinctp[:, 0] = ext
    st.extinct = np.asfortranarray(ext.reshape(npts, 1))
    st.total_ext = ext.copy()
    st.planck = np.zeros((npts, 1), np.float32, order='F')
    st.srctype, st.solarmu, st.solaraz, st.solarflux = 'S', -1.0, 0.0, 1.0
    st.gndalbedo, st.gndtemp, st.transcut = 0.04, 288.0, 0.0      # delta-M is moot without scattering: left off
    st.ylmsun = oracle.ylmall(True, np.float32(-1.0), np.float32(0.0), st.ml, st.mm, 1, st.nlm)
    st.nscatangle = 36
    st.phasetab = oracle.precompute_phase_check(pg.legenp, 36, 1, st.ml, False)
    st.normalize()
    st.dirflux, _, _ = oracle.make_direct(st, pg)
    return st, pg, wtmu


def make_combined_state(oracle, nmu=16, nphi=32):
    """VerifyCombined of the reference (tests/test_shdom.py:984-1056): the thermal slab lit by an overhead sun of unit
    flux (SRCTYPE='B'); the closed form gains 0.5 * T^2 / pi."""
    st, pg, wtmu = make_thermal_state(oracle, nmu, nphi)
    st.srctype, st.solarmu, st.solaraz, st.solarflux = 'B', -1.0, 0.0, 1.0
    st.ylmsun = oracle.ylmall(True, np.float32(-1.0), np.float32(0.0), st.ml, st.mm, 1, st.nlm)
    st.normalize()
    st.dirflux, _, _ = oracle.make_direct(st, pg)
    return st, pg, wtmu
