"""The ``at3d.core``-compatible keyword API (at3d_b200/core.py): the calls at3d/solver.py:681-759 and
at3d/gradient.py:262-398 make, with the reference's keyword names and return-tuple orders, checked
against the CPU oracle."""
import numpy as np
import pytest
import scenes

pytestmark = pytest.mark.gpu


def solver_kwargs(st):
    """The keyword arguments solver.RTE passes to core.render (at3d/solver.py:681-759)."""
    return dict(
        sfcgridrad=np.zeros(1, np.float32), nang=st.nang, transcut=st.transcut, tautol=st.tautol,
        maxnmicro=st.maxnmicro, interpmethod='ON' if st.interp_new else 'OO', phaseinterpwt=st.phaseinterpwt,
        phasemax=st.phasemax, nstphase=st.nstphase, ylmsun=st.ylmsun, phasetab=st.phasetab, nscatangle=st.nscatangle,
        ncs=1, nstokes=st.nstokes, nstleg=st.nstleg, nx=st.nx, ny=st.ny, nz=st.nz, bcflag=st.bcflag, ipflag=st.ipflag,
        npts=st.npts, ncells=st.ncells, ml=st.ml, mm=st.mm, nlm=st.nlm, numphase=st.numphase, nmu=st.nmu,
        nphi0max=st.nphi0max, nphi0=st.nphi0, maxnbc=st.maxnbc, ntoppts=st.ntoppts, nbotpts=st.nbotpts,
        nsfcpar=st.nsfcpar, gridptr=st.gridptr, neighptr=st.neighptr, treeptr=st.treeptr, shptr=st.shptr,
        bcptr=st.bcptr, cellflags=st.cellflags, iphase=st.iphase, deltam=bool(st.deltam), solarmu=st.solarmu,
        solaraz=st.solaraz, gndtemp=st.gndtemp, gndalbedo=st.gndalbedo, skyrad=st.skyrad, waveno=np.zeros(2, np.float32),
        wavelen=st.wavelen, mu=st.mu, phi=st.phi, wtdo=st.wtdo, xgrid=st.xgrid, ygrid=st.ygrid, zgrid=st.zgrid,
        gridpos=st.gridpos, sfcgridparms=st.sfcgridparms, bcrad=st.bcrad.copy(order='F'), extinct=st.extinct,
        albedo=st.albedo, legen=st.legen, dirflux=st.dirflux, fluxes=st.fluxes, source=st.source,
        srctype=chr(st.srctype) if isinstance(st.srctype, int) else st.srctype,
        sfctype=(chr(st.sfctype0) if isinstance(st.sfctype0, int) else st.sfctype0) +
                (chr(st.sfctype1) if isinstance(st.sfctype1, int) else st.sfctype1),
        units=chr(st.units) if isinstance(st.units, int) else st.units, total_ext=st.total_ext, npart=st.npart)


@pytest.mark.parametrize('case', ['scalar_periodic_split', 'polarized_rayleigh_varsfc'])
def test_core_render_keyword_api(case, oracle):
    from at3d_b200 import core
    sc = scenes.make(case, oracle)
    st = sc.state
    rays = scenes.ray_set(sc)
    ref, _, bcrad_ref = oracle.render(st, rays, trace_cap=4)
    kw = solver_kwargs(st)
    kw.update(camx=rays.camx, camy=rays.camy, camz=rays.camz, cammu=rays.cammu, camphi=rays.camphi, npix=rays.nrays,
              nosurface=False, correctinterpolate=True, singlescatter=False)
    for _ in range(2):                     # the second call re-uses the state resident in HBM
        bcrad, stokes, ierr, errmsg = core.render(**kw)
        assert ierr == 0 and len(errmsg) == 600
        assert stokes.shape == (st.nstokes, rays.nrays) and stokes.flags.f_contiguous
        np.testing.assert_allclose(stokes[0], ref[0], rtol=1e-4, atol=1e-6 * ref[0].max())
        nt = st.ntoppts
        np.testing.assert_allclose(bcrad[:, nt:nt + st.nbotpts], bcrad_ref[:, nt:], rtol=1e-6)
    # error convention: a ray starting below the domain -> ierr=1 and a message, no exception
    kw['camz'] = np.full(rays.nrays, -1.0, np.float32)
    _, _, ierr, errmsg = core.render(**kw)
    assert ierr == 1 and b'below domain' in errmsg
    core.clear_cache()


def test_core_gradient_keyword_api(oracle):
    from at3d_b200 import core, gradsetup
    sc = scenes.make('scalar_periodic_split', oracle)
    st = sc.state
    rays = scenes.ray_set(sc, n_persp=7, res=0.035)
    gi = gradsetup.make_gradient_inputs(sc, oracle, seed=11, numder=2)
    rad = oracle.render(st, rays)
    pix = gradsetup.make_pixels(1, rays.nrays, rad, seed=5, rays_per_pixel=2)
    gref, cref, soref = oracle.levisapprox_gradient(st, rays, gradsetup.with_pixels(gi, pix))
    kw = solver_kwargs(st)
    kw.update(camx=rays.camx, camy=rays.camy, camz=rays.camz, cammu=rays.cammu, camphi=rays.camphi, npix=rays.nrays,
              costfunc='L2', ncost=1, ngrad=1, nuncertainty=1, uncertainties=pix.uncertainties,
              rays_per_pixel=pix.rays_per_pixel, ray_weights=pix.ray_weights, stokes_weights=pix.stokes_weights,
              exact_single_scatter=True, measurements=pix.measurements, jacobian=np.zeros((1, 1, 1, 1), np.float32),
              jacobianptr=np.zeros((2, 1), np.int32), num_jacobian_pts=1, makejacobian=False, singlescatter=False,
              maxsubgridints=gi.maxsubgridints, longest_path_pts=gi.longest_path_pts, dextm=gi.dextm, dalbm=gi.dalbm,
              dfj=gi.dfj, dpath=gi.dpath, dptr=gi.dptr, partder=gi.partder, numder=gi.numder, dext=gi.dext, dalb=gi.dalb,
              dtemp=gi.dtemp, diphasep=gi.diphasep, dphasewtp=gi.dphasewtp, iphasep=gi.iphasep, phasewtp=gi.phasewtp,
              deriv_maxnmicro=gi.deriv_maxnmicro, albedop=gi.albedop, extinctp=gi.extinctp, extmin=gi.extmin,
              scatmin=gi.scatmin, optinterpwt=gi.optinterpwt, interpptr=gi.interpptr, doexact=gi.doexact, dleg=gi.dleg,
              dphasetab=gi.dphasetab, dnumphase=gi.dnumphase, maxpg=gi.maxpg, solarflux=st.solarflux,
              rshptr=st.rshptr, radiance=st.radiance)
    gradient, loss, images, jac, ierr, errmsg = core.levisapprox_gradient(**kw)
    assert ierr == 0
    assert gradient.shape == (gi.maxpg, gi.numder, 1) and images.shape == (1, pix.npix)
    assert abs(loss[0] - cref) <= 1e-4 * abs(cref)
    np.testing.assert_allclose(images, soref, rtol=1e-4, atol=1e-6)
    for idr in range(gi.numder):
        np.testing.assert_allclose(gradient[:, idr, 0], gref[:, idr], rtol=1e-4, atol=1e-4 * np.abs(gref[:, idr]).max())
    # MAKEJACOBIAN=.TRUE.: gradient / cost as above plus the per-pixel Jacobian at the selected property points
    jp = (np.argsort(-np.abs(gref[:, 0]))[:3] + 1).astype(np.int32)
    g1, c1, s1, jref = oracle.levisapprox_jacobian(st, rays, gradsetup.with_pixels(gi, pix), jp)
    kw.update(makejacobian=True, jacobianptr=jp, num_jacobian_pts=jp.size,
              jacobian=np.zeros((1, gi.numder, jp.size, pix.npix), np.float32, order='F'))
    gradient, loss, images, jac, ierr, errmsg = core.levisapprox_gradient(**kw)
    assert ierr == 0 and jac.shape == (1, gi.numder, jp.size, pix.npix)
    assert abs(loss[0] - c1) <= 1e-4 * abs(c1)
    np.testing.assert_allclose(jac, jref, rtol=1e-4, atol=1e-4 * np.abs(jref).max())
    core.clear_cache()


def test_core_small_routines_keyword_api(oracle):
    from at3d_b200 import core
    gradout, cost, ierr, errmsg = core.update_costfunction(
        cost=0.0, gradout=np.zeros((10, 1, 1)), stokesout=np.array([10.0, 10.0, 10.0, 0.0]),
        measurement=np.ones(4) * 13.0, raygrad_pixel=np.ones((4, 10, 1)), uncertainties=np.ones((4, 4)) * 5, costfunc='L2')
    assert ierr == 0 and abs(cost - 1960.0) < 1e-5 and abs(gradout[0, 0, 0] + 440.0) < 1e-5
    ws = np.arange(24, dtype=np.float32).reshape(2, 12, order='F')
    pix = np.repeat(np.arange(4), 3).astype(np.int32)
    out = core.average_subpixel_rays(pixel_index=pix, nstokes=2, weighted_stokes=ws, nrays=12, npixels=4)
    np.testing.assert_array_equal(out, oracle.average_subpixel_rays(ws, pix, 4))
    yr = core.ylmall(False, 0.4, 1.1, 7, 7, 6)
    np.testing.assert_allclose(yr, oracle.ylmall(False, np.float32(0.4), np.float32(1.1), 7, 7, 6, 64), rtol=2e-5, atol=2e-6)
