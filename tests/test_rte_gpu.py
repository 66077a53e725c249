"""The `at3d.solver.RTE`-shaped facade (at3d_b200/rte.py): constructor from medium / source / surface / numerical
parameter containers with the reference's variable names, `solve`, `integrate_to_sensor`, sub-pixel averaging -- end to
end on the GPU, checked against the oracle run on the same prepared state (MAKE_DIRECT, fixed-grid solve, RENDER)."""
import numpy as np
import pytest
import oracle_lib as O

pytestmark = pytest.mark.gpu


def hg_legcoef(gs, nleg, polarized):
    l = np.arange(nleg + 1)
    out = np.zeros((6, nleg + 1, len(gs)), np.float32)
    for k, g in enumerate(gs):
        out[0, :, k] = (2 * l + 1) * g ** l
        if polarized:
            out[1, :, k] = out[0, :, k] * (l >= 2)
            out[2, :, k] = 0.9 * out[0, :, k] * (l >= 2)
            out[3, :, k] = 0.9 * out[0, :, k]
            out[4, :, k] = -0.1 * out[0, :, k] * (l >= 2)
            out[5, :, k] = 0.03 * out[0, :, k] * (l >= 2)
    return out


def make_inputs(nx, ny, nz, bc, nstokes, two_species, seed=0):
    rng = np.random.default_rng(seed)
    dx = dy = 0.05
    x, y, z = np.arange(nx) * dx, np.arange(ny) * dy, np.linspace(0.0, 0.5, nz)
    X, Y, Z = np.meshgrid((np.arange(nx) + 0.5) / nx, (np.arange(ny) + 0.5) / ny, np.arange(nz) / (nz - 1), indexing='ij')
    r2 = ((X - 0.5) / 0.3) ** 2 + ((Y - 0.5) / 0.3) ** 2 + ((Z - 0.5) / 0.3) ** 2
    ext = (25.0 * np.exp(-r2)).astype(np.float32)
    ext[r2 > 1.5] = 0.0
    gs = [0.80, 0.84, 0.87]
    cloud = dict(x=x, y=y, z=z, delx=dx, dely=dy, extinction=ext, ssalb=np.full_like(ext, 0.999),
                 table_index=rng.integers(1, len(gs) + 1, (1, nx, ny, nz)).astype(np.int32),
                 phase_weights=np.ones((1, nx, ny, nz), np.float32), legcoef=hg_legcoef(gs, 180, nstokes > 1))
    medium = {'cloud': cloud}
    if two_species:
        ray = np.zeros((6, 3, 1), np.float32)
        ray[0, 0, 0] = 1.0; ray[0, 2, 0] = 0.5
        ray[1, 2, 0] = 3.0; ray[3, 1, 0] = 1.5; ray[4, 2, 0] = np.sqrt(1.5)
        rext = (0.03 * np.exp(-Z * 0.5 / 8.0)).astype(np.float32)
        medium['rayleigh'] = dict(x=x, y=y, z=z, delx=dx, dely=dy, extinction=rext, ssalb=np.ones_like(rext),
                                  table_index=np.ones((1, nx, ny, nz), np.int32),
                                  phase_weights=np.ones((1, nx, ny, nz), np.float32), legcoef=ray)
    params = dict(num_mu_bins=8, num_phi_bins=16, split_accuracy=0.0, deltam=True, spherical_harmonics_accuracy=0.0,
                  solution_accuracy=1e-4, acceleration_flag=True, high_order_radiance=False, ip_flag=0, iterfixsh=30,
                  tautol=0.2, transcut=5e-5, transmin=1.0, angle_set=2, x_boundary_condition=bc, y_boundary_condition=bc)
    source = dict(wavelength=0.672, srctype='S', solarflux=1.0, solarmu=-0.5, solaraz=0.2, skyrad=0.0, units='R')
    surface = dict(sfctype='FL', gndalbedo=0.05, gndtemp=298.15)
    return params, medium, source, surface


def make_sensor(xmax, ymax, seed=0):
    """Two orthographic views with 2x2 sub-pixel rays per pixel (variable names of at3d/sensor.py:93-107)."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.linspace(0.02, xmax - 0.02, 9), np.linspace(0.02, ymax - 0.02, 8), indexing='ij')
    npix1 = xs.size
    rx, ry, rmu, rphi, pix, w = [], [], [], [], [], []
    for iv, (mu, phi) in enumerate(((1.0, 0.0), (0.6, 1.1))):
        for sx, sy in ((-0.004, -0.004), (0.004, -0.004), (-0.004, 0.004), (0.004, 0.004)):
            rx.append(xs.ravel() + sx); ry.append(ys.ravel() + sy)
            rmu.append(np.full(npix1, mu)); rphi.append(np.full(npix1, phi))
            pix.append(np.arange(npix1) + iv * npix1); w.append(np.full(npix1, 0.25))
    order = np.argsort(np.concatenate(pix), kind='stable')
    cat = lambda a: np.concatenate(a)[order]
    return dict(ray_x=cat(rx), ray_y=cat(ry), ray_z=np.full(order.size, 0.5), ray_mu=cat(rmu), ray_phi=cat(rphi),
                ray_weight=cat(w), pixel_index=cat(pix).astype(np.int64), stokes=np.array([True, True, True, False]))


@pytest.mark.parametrize('bc,nstokes,two', [('periodic', 1, False), ('open', 3, True)])
def test_rte_solve_and_integrate_to_sensor(bc, nstokes, two):
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    params, medium, source, surface = make_inputs(9, 8, 11, bc, nstokes, two)
    rte = RTE(params, medium, source, surface, num_stokes=nstokes)
    rte.solve(maxiter=60)
    assert rte.check_solved() and 2 < rte.num_iterations < 60
    st0 = rte._unsolved
    # the prepared state: direct beam against the oracle's MAKE_DIRECT
    dirflux_ref = O.make_direct(st0, rte._pg)[0]
    np.testing.assert_allclose(st0.dirflux, dirflux_ref, rtol=1e-5, atol=1e-7)
    assert st0.bcflag == (3 if bc == 'open' else 0) and st0.npart == (2 if two else 1)
    # the solve (Eddington first guess, no splitting) and the rendering against the oracle on the same prepared state
    ref, iters, solcrit, _ = O.solve_adaptive(st0, rte._pg, rte._wtmu, splitacc=0.0, solacc=1e-4, maxiter=60)
    assert iters == rte.num_iterations
    np.testing.assert_array_equal(rte._solved.shptr, ref.shptr)
    np.testing.assert_allclose(rte._solved.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    xmax = 0.05 * (8 if bc == 'open' else 9); ymax = 0.05 * (7 if bc == 'open' else 8)
    sensor = make_sensor(xmax, ymax)
    if nstokes == 1:
        sensor['stokes'] = np.array([True, False, False, False])
    out = rte.integrate_to_sensor(sensor)
    rays = Rays(sensor['ray_x'], sensor['ray_y'], sensor['ray_z'], sensor['ray_mu'], sensor['ray_phi'])
    refrad = O.render(ref, rays)
    np.testing.assert_allclose(out['I'], refrad[0], rtol=1e-4, atol=1e-6 * refrad[0].max())
    if nstokes == 3:
        np.testing.assert_allclose(out['Q'], refrad[1], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(out['U'], refrad[2], rtol=1e-4, atol=1e-6)
    # pixel observables: 4 sub-pixel rays of weight 1/4 each
    obs = rte.average_subpixel_rays(out)
    npix = obs.shape[1]
    want = (refrad[:, :4 * npix].reshape(nstokes, npix, 4) * 0.25).sum(axis=2)
    np.testing.assert_allclose(obs, want, rtol=2e-4, atol=1e-6)
    assert rte.fluxes.shape[0] == 2 and np.all(rte.fluxes >= 0)
    rte.close()


def test_rte_levis_gradient_wrt_extinction():
    """RTE.levis_approx_gradient (unknown = cloud extinction) against the oracle's LEVISAPPROX_GRADIENT with the
    derivative tables of the same unknown (the parity bar), and a sanity check against a finite difference of the cost
    function through the whole chain (new medium -> solve -> render)."""
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    from at3d_b200 import gradsetup
    params, medium, source, surface = make_inputs(7, 6, 9, 'periodic', 1, False)
    params['solution_accuracy'] = 1e-5
    truth = RTE(params, medium, source, surface)
    truth.solve(maxiter=100)
    sensor = make_sensor(0.05 * 7, 0.05 * 6)
    sensor['stokes'] = np.array([True, False, False, False])
    obs = truth.average_subpixel_rays(truth.integrate_to_sensor(dict(sensor)))
    truth.close()
    npix = obs.shape[1]
    merged = dict(sensor, rays_per_pixel=np.full(npix, 4, np.int32), stokes_weights=np.ones((1, npix)),
                  measurement_data=obs, uncertainties=np.full((1, 1, npix), 1.0 / (0.02 * obs.max()) ** 2))

    def cost_of(scale_field):
        med = {'cloud': dict(medium['cloud'], extinction=(medium['cloud']['extinction'] * scale_field).astype(np.float32))}
        r = RTE(params, med, source, surface)
        r.solve(maxiter=100)
        return r

    guess = np.full_like(medium['cloud']['extinction'], 0.8)
    rte = cost_of(guess)
    loss, grad, images = rte.levis_approx_gradient(merged, ['cloud'])
    # oracle on the same state and tables
    gi = gradsetup.extinction_gradient_inputs(rte._solved, rte._pg, O, [0], rte._t['extmin'], rte._t['scatmin'])
    rays = Rays(sensor['ray_x'], sensor['ray_y'], sensor['ray_z'], sensor['ray_mu'], sensor['ray_phi'])
    pix = gradsetup.PixelData(merged['measurement_data'], merged['uncertainties'], merged['rays_per_pixel'],
                              merged['ray_weight'], merged['stokes_weights'])
    gref, cref, soref = O.levisapprox_gradient(rte._solved, rays, gradsetup.with_pixels(gi, pix))[:3]
    assert abs(loss - cref) <= 1e-4 * abs(cref) and loss > 0
    np.testing.assert_allclose(images, soref, rtol=1e-4)
    np.testing.assert_allclose(grad.reshape(-1), gref[:, 0], rtol=1e-4, atol=1e-4 * np.abs(gref).max())
    # finite difference along the direction "scale the whole cloud": d cost / d s = sum_i grad_i * ext_i
    ext = medium['cloud']['extinction'].astype(np.float64)
    directional = float(np.sum(grad[..., 0] * ext))
    eps = 0.01
    up, dn = cost_of(guess + eps), cost_of(guess - eps)
    lu = up.levis_approx_gradient(merged, ['cloud'])[0]
    ld = dn.levis_approx_gradient(merged, ['cloud'])[0]
    fd = (lu - ld) / (2 * eps)
    for r in (rte, up, dn):
        r.close()
    # the approximation holds the diffuse source fixed, so for "more cloud everywhere" it under-estimates the response:
    # same sign, same order of magnitude
    assert np.sign(fd) == np.sign(directional) and 0.2 * abs(fd) <= abs(directional) <= 1.5 * abs(fd)


def test_rte_two_dimensional_domain():
    """ny = 1: the reference switches to independent pixels in Y (IPFLAG bit 1, at3d/solver.py:2111) -> BACK_INT_GRID2D."""
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    params, medium, source, surface = make_inputs(10, 1, 9, 'periodic', 1, False)
    rte = RTE(params, medium, source, surface)
    rte.solve(maxiter=60)
    st0 = rte._unsolved
    assert st0.ipflag == 2 and rte.check_solved()
    ref, iters, solcrit, _ = O.solve_adaptive(st0, rte._pg, rte._wtmu, splitacc=0.0, solacc=1e-4, maxiter=60)
    assert iters == rte.num_iterations
    np.testing.assert_allclose(rte._solved.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    n = 40
    rays = Rays(np.linspace(0.01, 0.49, n), np.zeros(n), np.full(n, 0.5), np.full(n, 0.7), np.zeros(n))
    sensor = dict(ray_x=rays.camx, ray_y=rays.camy, ray_z=rays.camz, ray_mu=rays.cammu, ray_phi=rays.camphi,
                  stokes=np.array([True, False, False, False]))
    out = rte.integrate_to_sensor(sensor)
    refrad = O.render(ref, rays)
    np.testing.assert_allclose(out['I'], refrad[0], rtol=1e-4, atol=1e-6 * refrad[0].max())
    rte.close()


def test_rte_save_and_load_solution():
    """save_solution / load_solution with the reference's variable names (at3d/solver.py:1519-1686): a second RTE that
    loads the saved fields renders bit-identical radiances without solving; a saved ADAPTIVE grid (cells split, as an
    adaptive solve of the reference leaves them) is adopted with the properties re-interpolated to its points."""
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    from at3d_b200 import grid as G
    params, medium, source, surface = make_inputs(8, 7, 9, 'periodic', 3, False)
    a = RTE(params, medium, source, surface, num_stokes=3)
    a.solve(maxiter=60)
    ds = a.save_solution()
    assert set(('gridptr', 'neighptr', 'treeptr', 'cellflags', 'shptr', 'rshptr', 'source', 'radiance', 'fluxes')) <= set(ds)
    assert ds['gridptr'].min() >= 1                                    # 1-based contents, as the reference stores them
    sensor = make_sensor(0.05 * 8, 0.05 * 7)
    ia = a.integrate_to_sensor(dict(sensor))
    b = RTE(params, medium, source, surface, num_stokes=3)
    b.load_solution(ds)
    ib = b.integrate_to_sensor(dict(sensor))
    for k in ('I', 'Q', 'U'):
        np.testing.assert_array_equal(ia[k], ib[k])
    # an adaptive grid: split some cells of the saved grid the way DIVIDE_CELL does, interpolate the fields to the new points
    st = a._solved
    tree = G.CellTree(st.npts, st.ncells, np.pad(st.gridpos, ((0, 0), (0, 200))), np.pad(st.gridptr, ((0, 0), (0, 100))),
                      np.pad(st.neighptr, ((0, 0), (0, 100))), np.pad(st.treeptr, ((0, 0), (0, 100))),
                      np.pad(st.cellflags, (0, 100)))
    rng = np.random.default_rng(3)
    for _ in range(12):
        ic = int(rng.integers(1, tree.ncells + 1))
        if tree.treeptr[1, ic - 1] == 0:
            tree.divide_cell(ic, int(rng.integers(1, 4)))
    nnew = tree.npts - st.npts
    assert nnew > 0
    ds2 = dict(ds, npts=tree.npts, ncells=tree.ncells, gridpos=tree.gridpos[:, :tree.npts], gridptr=tree.gridptr[:, :tree.ncells],
               neighptr=tree.neighptr[:, :tree.ncells], treeptr=tree.treeptr[:, :tree.ncells], cellflags=tree.cellflags[:tree.ncells])
    # new points: 4 SH terms each (zero), fluxes zero -- enough for a consistency check of the adopted topology
    shptr = np.concatenate([st.shptr[:st.npts + 1], st.shptr[st.npts] + 4 * np.arange(1, nnew + 1)]).astype(np.int32)
    rshptr = np.concatenate([st.rshptr[:st.npts + 1], st.rshptr[st.npts] + 4 * np.arange(1, nnew + 1)]).astype(np.int32)
    rshptr = np.concatenate([rshptr, rshptr[-1:]])
    ds2.update(shptr=shptr, rshptr=rshptr, source=np.pad(ds['source'], ((0, 0), (0, 4 * nnew))),
               radiance=np.pad(ds['radiance'], ((0, 0), (0, 4 * nnew))), fluxes=np.pad(ds['fluxes'], ((0, 0), (0, nnew))))
    c = RTE(params, medium, source, surface, num_stokes=3)
    c.load_solution(ds2)
    assert c._solved.npts == tree.npts and c._solved.ncells == tree.ncells
    rays = Rays(sensor['ray_x'], sensor['ray_y'], sensor['ray_z'], sensor['ray_mu'], sensor['ray_phi'])
    ic_ = c.integrate_to_sensor(dict(sensor))
    ref = O.render(c._solved, rays)
    np.testing.assert_allclose(ic_['I'], ref[0], rtol=1e-4, atol=1e-6 * ref[0].max())
    np.testing.assert_allclose(ic_['Q'], ref[1], rtol=1e-4, atol=1e-6)
    for r in (a, b, c):
        r.close()


def test_rte_default_config_adaptive_solve():
    """The reference's default numerical parameters (default_config.json: split_accuracy 0.03, open boundaries, adapt grid
    factor 5 ...) through RTE.solve: the grid is split on the way exactly as the oracle's SPLIT_GRID does, and
    integrate_to_sensor works on the split grid; a second solve starts again from the base grid."""
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    params, medium, source, surface = make_inputs(8, 7, 10, 'open', 1, True)
    params.update(split_accuracy=0.03, adapt_grid_factor=5, num_sh_term_factor=1, cell_to_point_ratio=1.5)
    rte = RTE(params, medium, source, surface)
    rte.solve(maxiter=60)
    st0 = rte._unsolved
    ref, iters, solcrit, splitcrit = O.solve_adaptive(st0, rte._pg, rte._wtmu, splitacc=0.03, solacc=1e-4, maxiter=60)
    assert ref.npts > st0.npts and rte.check_solved()
    assert (rte._solved.npts, rte._solved.ncells, rte.num_iterations) == (ref.npts, ref.ncells, iters)
    np.testing.assert_array_equal(rte._solved.gridptr, ref.gridptr)
    np.testing.assert_array_equal(rte._solved.shptr, ref.shptr)
    sensor = make_sensor(0.05 * 7, 0.05 * 6)
    sensor['stokes'] = np.array([True, False, False, False])
    out = rte.integrate_to_sensor(sensor)
    rays = Rays(sensor['ray_x'], sensor['ray_y'], sensor['ray_z'], sensor['ray_mu'], sensor['ray_phi'])
    refrad = O.render(ref, rays)
    np.testing.assert_allclose(out['I'], refrad[0], rtol=1e-4, atol=1e-6 * refrad[0].max())
    ds = rte.save_solution()
    assert ds['npts'] == ref.npts and ds['nbcells'] == st0.ncells
    npts1 = rte._solved.npts
    rte.solve(maxiter=60)
    assert rte._solved.npts == npts1 and rte._unsolved.npts == st0.npts
    done = rte.num_iterations
    rte.solve(maxiter=5, init_solution=False)            # `maxiter` is a cap on the total: already used up, nothing runs
    assert rte.num_iterations == done and rte.check_solved(verbose=False)
    with pytest.raises(NotImplementedError):
        rte.solve(maxiter=200, init_solution=False)      # continuing with cell splitting
    rte.close()


def test_rte_load_solution_validates_before_mutating():
    from at3d_b200.rte import RTE
    params, medium, source, surface = make_inputs(6, 6, 7, 'periodic', 1, False)
    a = RTE(params, medium, source, surface)
    a.solve(maxiter=40)
    ds = a.save_solution()
    before = (a._npts, a._ncells, a._solved.npts)
    bad = dict(ds, shptr=ds['shptr'][:-3])
    with pytest.raises(ValueError):
        a.load_solution(bad)
    bad = dict(ds, nstokes=3)
    with pytest.raises(ValueError):
        a.load_solution(bad)
    assert (a._npts, a._ncells, a._solved.npts) == before
    a.close()


def test_rte_thermal_source_solve_render_and_gradient():
    """A thermal source through the facade (at3d.source.thermal + `atmosphere` temperature, at3d/solver.py:1762-1845):
    PLANCK on the grid, the thermal solve, RENDER and the Levis gradient with its thermal component -- each against the
    oracle on the same prepared state."""
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    from at3d_b200 import gradsetup
    params, medium, source, surface = make_inputs(7, 6, 9, 'periodic', 1, False)
    medium['cloud']['ssalb'] = np.full_like(medium['cloud']['extinction'], 0.7)
    source = dict(wavelength=10.5, srctype='T', solarflux=0.0, solarmu=-0.5, solaraz=0.0, skyrad=2.7, units='R')
    surface = dict(sfctype='FL', gndalbedo=0.03, gndtemp=296.0)
    c = medium['cloud']
    X, Y, Z = np.meshgrid(c['x'], c['y'], c['z'], indexing='ij')
    atmosphere = dict(x=c['x'], y=c['y'], z=c['z'], temperature=(292.0 - 30.0 * Z + 3.0 * np.sin(20.0 * X) * np.cos(15.0 * Y)).astype(np.float32))
    with pytest.raises(KeyError):
        RTE(params, medium, source, surface)                          # thermal source without a temperature field
    rte = RTE(params, medium, source, surface, atmosphere=atmosphere)
    rte.solve(maxiter=60)
    st0 = rte._unsolved
    assert st0.srctype in ('T', ord('T')) and rte.check_solved()
    ref, iters, solcrit, _ = O.solve_adaptive(st0, rte._pg, rte._wtmu, tempp=rte._tempp, splitacc=0.0, solacc=1e-4, maxiter=60)
    assert iters == rte.num_iterations
    np.testing.assert_allclose(rte._solved.temp, ref.temp, rtol=1e-6)
    np.testing.assert_allclose(rte._solved.planck, ref.planck, rtol=1e-5)
    np.testing.assert_allclose(rte._solved.fluxes, ref.fluxes, rtol=1e-4, atol=1e-6 * ref.fluxes.max())
    sensor = make_sensor(0.05 * 7, 0.05 * 6)
    sensor['stokes'] = np.array([True, False, False, False])
    out = rte.integrate_to_sensor(dict(sensor))
    rays = Rays(sensor['ray_x'], sensor['ray_y'], sensor['ray_z'], sensor['ray_mu'], sensor['ray_phi'])
    refrad = O.render(ref, rays)
    assert refrad[0].min() > 1.0                                       # W m-2 sr-1 um-1 near 10 um
    np.testing.assert_allclose(out['I'], refrad[0], rtol=1e-4)
    obs = rte.average_subpixel_rays(out)
    npix = obs.shape[1]
    merged = dict(sensor, rays_per_pixel=np.full(npix, 4, np.int32), stokes_weights=np.ones((1, npix)),
                  measurement_data=(obs * 1.03).astype(np.float32), uncertainties=np.full((1, 1, npix), 1.0 / (0.02 * obs.max()) ** 2))
    loss, grad, images = rte.levis_approx_gradient(merged, ['cloud'])
    gi = gradsetup.extinction_gradient_inputs(rte._solved, rte._pg, O, [0], rte._t['extmin'], rte._t['scatmin'])
    pix = gradsetup.PixelData(merged['measurement_data'], merged['uncertainties'], merged['rays_per_pixel'],
                              merged['ray_weight'], merged['stokes_weights'])
    gref, cref, soref = O.levisapprox_gradient(rte._solved, rays, gradsetup.with_pixels(gi, pix))[:3]
    assert abs(loss - cref) <= 1e-4 * abs(cref) and loss > 0
    np.testing.assert_allclose(grad.reshape(-1), gref[:, 0], rtol=1e-4, atol=1e-4 * np.abs(gref).max())
    rte.close()


def test_rte_variable_lambertian_surface_equals_the_fixed_one():
    """'VL' with uniform parameters is the 'FL' surface: SURFACE_PARM_INTERP + VARIABLE_LAMBERTIAN_BOUNDARY through the
    facade must reproduce the fixed-Lambertian solve and radiances."""
    from at3d_b200.rte import RTE
    params, medium, source, surface = make_inputs(8, 7, 9, 'periodic', 1, False)
    surface['gndalbedo'] = 0.3
    fl = RTE(params, medium, source, surface)
    fl.solve(maxiter=60)
    nxs, nys = 3, 2
    sp = np.zeros((2, nxs + 1, nys + 1), np.float32, order='F')
    sp[0], sp[1] = 298.15, 0.3
    vsurf = dict(sfctype='VL', gndalbedo=0.3, gndtemp=298.15, nsfcpar=2, nxsfc=nxs, nysfc=nys, delxsfc=0.05 * 8 / nxs,
                 delysfc=0.05 * 7 / nys, sfcparms=sp.ravel(order='F'))
    vl = RTE(params, medium, source, vsurf)
    vl.solve(maxiter=60)
    assert vl.num_iterations == fl.num_iterations
    np.testing.assert_allclose(vl._solved.fluxes, fl._solved.fluxes, rtol=1e-5, atol=1e-7)
    sensor = make_sensor(0.05 * 8, 0.05 * 7)
    sensor['stokes'] = np.array([True, False, False, False])
    a, b = fl.integrate_to_sensor(dict(sensor))['I'], vl.integrate_to_sensor(dict(sensor))['I']
    np.testing.assert_allclose(b, a, rtol=1e-5, atol=1e-7)
    fl.close(); vl.close()


def test_rte_ocean_surface_matches_the_oracle():
    """'VO' (ocean_unpolarized) through the facade: the solve with the stored downwelling radiances per ordinate and the
    BRDF integration in RENDER against the oracle on the same prepared state (SFCGRIDPARMS interpolated on the host)."""
    from at3d_b200.rte import RTE
    from at3d_b200.state import Rays
    import shdom_verification as V
    params, medium, source, surface = make_inputs(7, 6, 9, 'periodic', 1, False)
    nxs, nys = 2, 2
    sp = np.zeros((3, nxs + 1, nys + 1), np.float32, order='F')
    sp[0] = 290.0
    sp[1] = 4.0 + 3.0 * np.arange(nxs + 1)[:, None] + 1.0 * np.arange(nys + 1)[None, :]      # wind speed [m/s]
    sp[2] = 0.1 + 0.05 * np.arange(nys + 1)[None, :]                                         # pigment
    dxs, dys = 0.05 * 7 / nxs, 0.05 * 6 / nys
    vsurf = dict(sfctype='VO', gndalbedo=0.0, gndtemp=290.0, nsfcpar=3, nxsfc=nxs, nysfc=nys, delxsfc=dxs, delysfc=dys,
                 sfcparms=sp.ravel(order='F'))
    rte = RTE(params, medium, source, vsurf)
    rte.solve(maxiter=60)
    st0 = rte._unsolved.copy()
    st0.sfcgridparms = np.asfortranarray(V.surface_parm_interp(st0.bcptr[:, 1], st0.nbotpts, st0.gridpos, sp, dxs, dys))
    np.testing.assert_allclose(rte._solved.sfcgridparms[:, :st0.nbotpts], st0.sfcgridparms, rtol=1e-6, atol=1e-7)
    ref, iters, solcrit, _ = O.solve_adaptive(st0, rte._pg, rte._wtmu, splitacc=0.0, solacc=1e-4, maxiter=60)
    assert iters == rte.num_iterations
    np.testing.assert_allclose(rte._solved.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    sensor = make_sensor(0.05 * 7, 0.05 * 6)
    sensor['stokes'] = np.array([True, False, False, False])
    out = rte.integrate_to_sensor(dict(sensor))
    rays = Rays(sensor['ray_x'], sensor['ray_y'], sensor['ray_z'], sensor['ray_mu'], sensor['ray_phi'])
    refrad = O.render(ref, rays)
    np.testing.assert_allclose(out['I'], refrad[0], rtol=1e-4, atol=1e-6 * refrad[0].max())
    rte.close()


def test_rte_solve_continues_from_a_loaded_solution():
    """The warm start of an optimisation step (at3d/medium.py:1829-1830: the new solver loads the old solution, then
    solves): `load_solution` + `solve` continues SOLUTION_ITERATIONS from the loaded SOURCE / RADIANCE on this grid --
    fewer iterations than from INIT_RADIANCE, the same solution within the solution accuracy, and the same iterations as the
    oracle continued from the same fields."""
    from at3d_b200.rte import RTE
    params, medium, source, surface = make_inputs(8, 7, 9, 'periodic', 1, False)
    params['solution_accuracy'] = 1e-5
    a = RTE(params, medium, source, surface)
    a.solve(maxiter=100)
    ds = a.save_solution()
    med2 = {'cloud': dict(medium['cloud'], extinction=(medium['cloud']['extinction'] * 1.04).astype(np.float32))}
    cold = RTE(params, med2, source, surface)
    cold.solve(maxiter=100)
    warm = RTE(params, med2, source, surface)
    warm.load_solution(ds)
    assert not warm.check_solved()                       # a starting point, not this medium's solution
    st0 = warm._restore
    warm.solve(maxiter=100)
    assert warm.check_solved() and warm.num_iterations < cold.num_iterations
    ref, iters, solcrit = O.solve_fixed_grid(st0, warm._wtmu, solacc=1e-5, maxiter=100, initial=st0)
    assert iters == warm.num_iterations
    np.testing.assert_array_equal(warm._solved.shptr, ref.shptr)
    np.testing.assert_allclose(warm._solved.fluxes, ref.fluxes, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(warm._solved.fluxes, cold._solved.fluxes, rtol=2e-3, atol=1e-4 * cold._solved.fluxes.max())
    # init_solution=False on a solved object: nothing left to do, one iteration confirms it
    before = warm.num_iterations
    warm.solve(maxiter=100, init_solution=False)
    assert before < warm.num_iterations <= before + 2 and warm.check_solved()       # the count carries on
    # `maxiter` caps the total: nothing runs when it is already used up
    warm.solve(maxiter=before, init_solution=False)
    assert warm.num_iterations <= before + 2 and warm.check_solved()
    for r in (a, cold, warm):
        r.close()


def test_rte_from_the_factory_datasets():
    """`configuration.get_config`, `source.solar`, `surface.lambertian` / `surface.ocean_unpolarized` and the `sensor`
    projections give the datasets the reference's scripts pass to RTE: same result as the hand-written mappings."""
    from at3d_b200 import configuration, source as SRC, surface as SFC, sensor as SNS
    from at3d_b200.rte import RTE
    params, medium, source, surface = make_inputs(8, 7, 9, 'open', 1, False)
    cfg = configuration.get_config()
    assert cfg['split_accuracy'] == 0.03 and cfg['num_mu_bins'] == 16 and cfg['x_boundary_condition'] == 'open'
    for k, v in params.items():
        cfg[k] = v
    a = RTE(params, medium, source, surface)
    b = RTE(cfg, medium, SRC.solar(0.672, 0.5, np.rad2deg(0.2)), SFC.lambertian(0.05))
    a.solve(maxiter=60); b.solve(maxiter=60)
    assert a.num_iterations == b.num_iterations
    grid = medium['cloud']
    cams = [SNS.orthographic_projection(0.672, grid, 0.04, 0.04, 20.0, 30.0, altitude=0.5,
                                        sub_pixel_ray_args={'method': SNS.gaussian, 'degree': 2}),
            SNS.perspective_projection(0.672, 12.0, 12, 10, [0.2, 0.15, 2.5], [0.2, 0.17, 0.25], [0, 1, 0])]
    for cam in cams:
        ia = a.integrate_to_sensor(cam.copy())['I']
        ib = b.integrate_to_sensor(cam.copy())['I']
        np.testing.assert_allclose(ib, ia, rtol=2e-6, atol=1e-8)         # solaraz went through degrees and back
        assert ia.shape == cam['ray_mu'].shape and ia.max() > 0.01
        pix = b.average_subpixel_rays(b.integrate_to_sensor(cam))
        assert pix.shape == (1, int(np.prod(cam['image_shape'])))
    assert a.adaptive_fluxes.shape == (2, a._solved.npts) and np.all(a.adaptive_fluxes >= 0)
    np.testing.assert_array_equal(a.fluxes.reshape(2, -1), a.adaptive_fluxes[:, :a.fluxes[0].size])
    # a tighter tolerance, continuing from the solution at hand
    first = a.num_iterations
    a.set_solution_accuracy(1e-6)
    a.solve(maxiter=60, init_solution=False)
    assert a.solution_accuracy == 1e-6 and a.check_solved(verbose=False) and a.num_iterations > first
    a.close(); b.close()
    # the ocean factory lays SFCPARMS out as the facade test above does by hand
    nxs, nys = 2, 2
    wind = 4.0 + 3.0 * np.arange(nxs)[:, None] + 1.0 * np.arange(nys)[None, :]
    pig = 0.1 + 0.05 * np.arange(nys)[None, :] + 0.0 * wind
    ds = SFC.ocean_unpolarized(wind, pig, ground_temperature=290.0, delx=0.2, dely=0.15)
    sp = np.asarray(ds['sfcparms']).reshape((3, nxs + 1, nys + 1), order='F')
    np.testing.assert_array_equal(sp[1, :nxs, :nys], wind.astype(np.float32))
    np.testing.assert_array_equal(sp[:, nxs, :], sp[:, 0, :]); np.testing.assert_array_equal(sp[:, :, nys], sp[:, :, 0])
    params, medium, source, _ = make_inputs(7, 6, 9, 'periodic', 1, False)
    rte = RTE(params, medium, source, ds)
    rte.solve(maxiter=60)
    assert rte.check_solved(verbose=False) and ds['sfctype'] == 'VO' and float(ds['gndtemp']) == 290.0
    rte.close()
