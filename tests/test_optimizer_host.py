"""Optimizer / CallbackFn host logic (at3d/optimize.py:154-250, at3d/callback.py:32-75) on an analytic objective: priors
are called with (state, iteration) and added, callbacks see the optimizer, bounds are broadcast."""
import numpy as np
from at3d_b200.optimize import ObjectiveFunction, Optimizer, CallbackFn


def test_optimizer_with_prior_callback_and_bounds():
    target = np.array([1.0, -2.0, 3.0, 0.5])

    def loss_fn(state, measurements):
        d = state - target
        return 0.5 * float(d @ d), d.copy()

    seen = []

    def prior(state, iteration):
        seen.append(iteration)
        return 0.5 * 0.1 * float(state @ state), 0.1 * state

    cb = CallbackFn(lambda optimizer: {'iteration': optimizer.iteration, 'state': optimizer.state.copy(),
                                       'loss': optimizer.loss_history[-1]})
    obj = ObjectiveFunction(None, loss_fn, min_bounds=-1.0, max_bounds=10.0)
    opt = Optimizer(obj, prior_fn=prior, callback_fn=cb, options=dict(maxiter=50, gtol=1e-12, ftol=1e-14))
    assert opt.method == 'L-BFGS-B' and opt.options['maxiter'] == 50 and opt.objective_fn is obj
    res = opt.minimize(np.zeros(4), iteration_step=7)
    expect = np.clip(target / 1.1, -1.0, 10.0)                      # the bound is active for the second unknown
    np.testing.assert_allclose(res.x, expect, atol=1e-6)
    assert seen[0] == 7 and 7 < max(seen) <= opt.iteration and seen == sorted(seen)          # priors get the running iteration number
    assert cb.output['iteration'] == list(range(8, opt.iteration + 1))
    assert len(cb.output['state']) == len(cb.output['loss']) == opt.iteration - 7
    assert opt.loss_history[-1] <= opt.loss_history[0]
    # a callback period longer than the run: never called
    quiet = CallbackFn(lambda optimizer: {'n': 1}, ckpt_period=3600.0)
    Optimizer(obj, callback_fn=quiet).minimize(np.zeros(4))
    assert quiet.output == {}
