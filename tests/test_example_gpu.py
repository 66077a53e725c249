"""examples/retrieve_extinction.py end to end: factories -> sensors -> measurements -> save / load -> transforms ->
optimizer.  The cost and the extinction error must fall."""
import os
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_retrieval_example_runs_and_converges(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples'))
    import retrieve_extinction as ex
    out = ex.run(maxiter=6, save=str(tmp_path / 'forward_model.nc'), verbose=False)
    assert os.path.exists(out['path']) and out['nrays'] == 4 * 14 * 14 * 4
    hist, rms = out['history'], out['output']['rms_error']
    assert len(rms) >= 3 and hist[-1] < 0.2 * hist[0]
    assert rms[-1] < 0.8 * rms[0]
    np.testing.assert_array_equal(out['rte_grid']['z'], np.linspace(0.0, 0.5, 11))
