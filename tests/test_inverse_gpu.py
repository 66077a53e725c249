"""A reference-style inversion script, end to end through the at3d-shaped API of at3d_b200 (containers.SensorsDict /
SolversDict / UnknownScatterers, gradient.LevisApproxGradientUncorrelated, optimize.ObjectiveFunction / Optimizer,
GridStateGenerator): two wavelengths, forward measurements, cost + gradient for extinction and single-scattering-albedo
unknowns against the oracle's LEVISAPPROX_GRADIENT on the same states, and a few L-BFGS-B iterations that reduce the cost."""
from collections import OrderedDict
import numpy as np
import pytest
import oracle_lib as O
from test_rte_gpu import make_inputs, make_sensor

pytestmark = pytest.mark.gpu


def build(scale=1.0, wavelengths=(0.66, 0.86)):
    from at3d_b200.rte import RTE
    from at3d_b200.containers import SolversDict
    solvers, mediums, sources, surfaces, params, nst = SolversDict(), OrderedDict(), {}, {}, {}, {}
    for k, wl in enumerate(wavelengths):
        p, medium, source, surface = make_inputs(7, 6, 9, 'open', 1, True, seed=k)
        p['solution_accuracy'] = 1e-5
        source = dict(source, wavelength=wl)
        medium = OrderedDict((n, dict(s)) for n, s in medium.items())
        medium['cloud']['extinction'] = (medium['cloud']['extinction'] * (1.0 + 0.2 * k) * scale).astype(np.float32)
        mediums[wl], sources[wl], surfaces[wl], params[wl], nst[wl] = medium, source, surface, p, 1
        solvers.add_solver(wl, RTE(p, medium, source, surface, num_stokes=1))
    return solvers, mediums, sources, surfaces, params, nst


def sensors_for(wavelengths):
    from at3d_b200.containers import SensorsDict
    sd = SensorsDict()
    for wl in wavelengths:
        for seed in (0, 1):
            s = make_sensor(0.05 * 7, 0.05 * 6, seed=seed)
            s['stokes'] = np.array([True, False, False, False])
            s['wavelength'] = wl
            sd.add_sensor('cam', s)
    return sd


def test_reference_style_inversion_script():
    from at3d_b200.containers import UnknownScatterers
    from at3d_b200.gradient import LevisApproxGradientUncorrelated
    from at3d_b200.optimize import ObjectiveFunction, Optimizer, GridStateGenerator
    from at3d_b200 import gradsetup
    from at3d_b200.state import Rays
    wls = (0.66, 0.86)
    # ---- forward: synthetic measurements of the true medium ----
    truth, *_ = build(1.0, wls)
    measurements = sensors_for(wls)
    measurements.get_measurements(truth, maxiter=100, verbose=False)
    assert measurements.npixels == 4 * 144 and measurements.nmeasurements == 4 * 144
    for s in measurements['cam']['sensor_list']:
        assert s['I'].shape == (144,) and np.all(s['I'] > 0)
        # 3 % radiometric uncertainty: the inverse error covariance per pixel (uncertainties.py of the reference)
        s['uncertainties'] = np.full((1, 1, 144), 1.0 / (0.03 * s['I'].max()) ** 2)
    # the same observables straight from the RTE
    rte0 = truth[0.66]
    s0 = dict(measurements['cam']['sensor_list'][0])
    direct = rte0.average_subpixel_rays(rte0.integrate_to_sensor({k: v for k, v in s0.items() if k not in ('I', 'uncertainties')}))
    np.testing.assert_array_equal(direct[0], measurements['cam']['sensor_list'][0]['I'])
    for r in truth.values():
        r.close()
    # ---- inverse: cost and gradient at a first guess (80 % of the extinction), unknowns cloud extinction + ssalb ----
    solvers, mediums, sources, surfaces, params, nst = build(0.8, wls)
    unknowns = UnknownScatterers()
    unknowns.add_unknowns('cloud', ['extinction', 'ssalb'])
    forward = measurements.make_forward_sensors()
    grad_fn = LevisApproxGradientUncorrelated(measurements, solvers, forward, unknowns,
                                              dict(n_jobs=1, mpi_comm=None, verbose=False, maxiter=100, init_solution=True),
                                              dict(cost_function='L2', exact_single_scatter=True), dict(add_noise=False))
    loss, gds, jac = grad_fn()
    assert jac is None and gds['gradient'].shape == (7, 6, 9, 2)
    # an uncertainty model of at3d_b200/uncertainties.py attached to the instrument fills `uncertainties` itself:
    # NullUncertainty(scaling 2) doubles the unweighted cost and gradient
    import copy
    from at3d_b200 import uncertainties as UNC
    plain = copy.deepcopy(measurements)
    for sensor in plain['cam']['sensor_list']:
        del sensor['uncertainties']
    scaled = copy.deepcopy(plain)
    scaled.add_uncertainty_model('cam', UNC.NullUncertainty('L2', 2.0))
    with pytest.raises(ValueError, match='inconsistent'):
        LevisApproxGradientUncorrelated(scaled, solvers, forward, unknowns, dict(verbose=False, maxiter=100, init_solution=True),
                                        dict(cost_function='LL', exact_single_scatter=True), dict(add_noise=False))
    pair = []
    for meas in (plain, scaled):
        fn = LevisApproxGradientUncorrelated(meas, solvers, meas.make_forward_sensors(), unknowns,
                                             dict(verbose=False, maxiter=100, init_solution=True),
                                             dict(cost_function='L2', exact_single_scatter=True), dict(add_noise=False))
        pair.append(fn()[:2])
    assert scaled['cam']['sensor_list'][0]['uncertainties'].shape == (4, 4, 144)
    assert pair[1][0] == pytest.approx(2.0 * pair[0][0], rel=1e-12)
    np.testing.assert_allclose(pair[1][1]['gradient'], 2.0 * pair[0][1]['gradient'], rtol=1e-12)
    assert gds['derivative_index'] == [('cloud', 'extinction'), ('cloud', 'ssalb')]
    # the oracle on the same solved states and derivative tables
    rte_sensors, _ = forward.sort_sensors(solvers, measurements)
    lref, gref = 0.0, 0.0
    for wl, rte in solvers.items():
        d = rte._deriv
        gi = gradsetup.optical_gradient_inputs(rte._solved, rte._pg, O, d['partder'], d['doexact'], d['dext'], d['dalb'],
                                               d['diphasep'], d['dphasewtp'], d['dleg'], d['dphasetab'], rte._t['extmin'],
                                               rte._t['scatmin'])
        s = rte_sensors[wl]
        rays = Rays(s['ray_x'], s['ray_y'], s['ray_z'], s['ray_mu'], s['ray_phi'])
        pix = gradsetup.PixelData(s['measurement_data'], s['uncertainties'], s['rays_per_pixel'], s['ray_weight'],
                                  s['stokes_weights'])
        g, c, so = O.levisapprox_gradient(rte._solved, rays, gradsetup.with_pixels(gi, pix))[:3]
        lref += c; gref = gref + g
    lref /= forward.nmeasurements; gref = gref / forward.nmeasurements
    assert abs(loss - lref) <= 1e-4 * abs(lref) and loss > 0
    got = gds['gradient'].reshape(-1, 2)
    for k in range(2):
        np.testing.assert_allclose(got[:, k], gref[:, k], rtol=1e-4, atol=1e-4 * np.abs(gref[:, k]).max())
    assert np.abs(gref[:, 1]).max() > 0                         # the ssalb unknown has a signal
    # the forward sensors hold the modelled observables of the guess
    assert all('I' in s and s['I'].shape == (144,) for s in forward['cam']['sensor_list'])
    # ---- a few L-BFGS-B iterations on the cloud extinction (reference workflow: optimize.Optimizer) ----
    unknown_ext = UnknownScatterers()
    unknown_ext.add_unknowns('cloud', ['extinction'])
    mask = mediums[wls[0]]['cloud']['extinction'] > 0
    gen = GridStateGenerator(solvers, unknown_ext, mediums, sources, surfaces, params, nst, mask=mask)
    x0 = gen.get_state()
    obj = ObjectiveFunction.LevisApproxUncorrelatedL2(
        measurements, solvers, forward, unknown_ext, gen, gen.project_gradient_to_state,
        parallel_solve_kwargs=dict(verbose=False, maxiter=100, init_solution=True),
        gradient_kwargs=dict(cost_function='L2', exact_single_scatter=True), uncertainty_kwargs=dict(add_noise=False),
        min_bounds=1e-3, max_bounds=200.0)
    opt = Optimizer(obj, options=dict(maxiter=4, maxls=8, gtol=1e-16, ftol=1e-16))
    res = opt.minimize(x0)
    hist = opt.loss_history
    assert len(hist) >= 3 and hist[-1] < 0.5 * hist[0] and res.fun <= hist[0]
    for r in solvers.values():
        r.close()
