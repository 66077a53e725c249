"""Host-side sharding logic (at3d_b200/parallel.py) and the gradient all-reduce over a 2-rank gloo group.

* ``subdivide_raytrace_jobs`` reproduces the reference's golden splits (reference tests/test_shdom.py:357-372);
* two CPU ranks each evaluate their pixel shard (the oracle stands in for the GPU evaluation) and
  all-reduce: the result equals the single-process evaluation -- the N>1 path of bench.py / config 4.
"""
import os
import sys
from collections import OrderedDict
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_subdivide_raytrace_jobs_reference_golden():
    from at3d_b200.parallel import subdivide_raytrace_jobs
    # merged sensors of reference tests/test_shdom.py:279-355, described by their rays per pixel
    sensors = OrderedDict([
        (3.4, dict(rays_per_pixel=np.ones(7400, int))),
        (0.8, dict(rays_per_pixel=np.array([4] * 360 + [9] * 1370))),
        (1.6, dict(rays_per_pixel=np.full(360, 4))),
        (2.0, dict(rays_per_pixel=np.ones(360, int))),
    ])
    keys, rays, pixels = subdivide_raytrace_jobs(sensors, 4)
    assert rays == [(0, 3700), (3700, 7400), (0, 4590), (4590, 9180), (9180, 13770), (0, 1440), (0, 360)]
    assert pixels == [(0, 3700), (3700, 7400), (0, 710), (710, 1220), (1220, 1730), (0, 360), (0, 360)]
    assert keys == [3.4, 3.4, 0.8, 0.8, 0.8, 1.6, 2.0]


@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_shards_partition_pixels_and_rays(world):
    from at3d_b200.parallel import shard_for_rank
    rng = np.random.default_rng(world)
    rpp = rng.integers(1, 6, 997)
    got = [shard_for_rank(rpp, r, world) for r in range(world)]
    assert got[0][0] == 0 and got[0][2] == 0 and got[-1][1] == rpp.sum() and got[-1][3] == rpp.size
    starts = np.concatenate([[0], np.cumsum(rpp)])
    for a, b in zip(got[:-1], got[1:]):
        assert a[1] == b[0] and a[3] == b[2]
    for lo, hi, p0, p1 in got:
        assert starts[p0] == lo and starts[p1] == hi          # pixel aligned


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import oracle_lib as O
    import scenes
    from at3d_b200 import gradsetup, parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sc = scenes.make('scalar_periodic_split', O)
    rays = scenes.ray_set(sc, n_persp=6, res=0.04)
    gi = gradsetup.make_gradient_inputs(sc, O, seed=11, numder=2)
    rad = O.render(sc.state, rays)
    pix = gradsetup.make_pixels(1, rays.nrays, rad, seed=5, rays_per_pixel=2)

    def compute(r, p):
        g, c, so = O.levisapprox_gradient(sc.state, r, gradsetup.with_pixels(gi, p))
        return g, np.array([c]), so
    g, c, so, (p0, p1) = parallel.sharded_gradient(compute, rays, pix, rank, world)
    if rank == 0:
        gfull, cfull, sofull = O.levisapprox_gradient(sc.state, rays, gradsetup.with_pixels(gi, pix))
        q.put((float(np.abs(g - gfull).max() / np.abs(gfull).max()), float(abs(c[0] - cfull) / abs(cfull)),
               bool(np.array_equal(so, sofull[:, p0:p1]))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gerr, cerr, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert gerr < 1e-12 and cerr < 1e-12 and same
