"""Device-side sharding of rays / pixels and the one collective of the path.

Replaces the thread / MPI fan-out of the reference (at3d/parallel.py:13-174, at3d/containers.py:187-233):
the concatenated ray list of a merged sensor is cut into contiguous, pixel-aligned ranges -- the
same rule as ``subdivide_raytrace_jobs`` -- one range per GPU (rank) instead of one per joblib thread;
every rank holds a replica of the solved state and marches only its rays; the per-worker
``np.stack(...).sum(axis=-1)`` of the gradient (at3d/parallel.py:100, gradient.py:437-438) becomes one
all-reduce(sum) of ``float64[maxpg*numder]`` plus the scalar cost over NCCL (gloo in the CPU tests).
All rays of a pixel stay on one rank because the adjoint weights need complete pixels
(LEVISAPPROX_GRADIENT phase 2, shdomsub4.f:700-709).
"""
from collections import OrderedDict
import numpy as np


def _aligned_chunks(nrays, rays_per_pixel, nchunks):
    """Cut [0,nrays) into ``nchunks`` contiguous ranges whose ends fall on pixel boundaries; the rule of
    at3d/parallel.py:146-167 (np.array_split of arange(nrays+1), ends snapped to the nearest pixel end)."""
    rpp = np.asarray(rays_per_pixel, dtype=np.int64)
    pixel_inds = np.concatenate([[0], np.cumsum(rpp)]).astype(np.int64)
    ends = pixel_inds[1:]                      # ray index one past the last ray of every pixel
    split = np.array_split(np.arange(nrays + 1), nchunks)
    ray_start_end, pixel_start_end = [], []
    new_start = 0
    for chunk in split:
        end = int(chunk.max())
        new_end = int(ends[np.abs(ends - end).argmin()])
        ray_start_end.append((new_start, new_end))
        new_start = new_end
    for start, end in ray_start_end:
        pixel_start_end.append((int(np.where(pixel_inds == start)[0][0]), int(np.where(pixel_inds == end)[0][0])))
    assert ray_start_end[-1][1] == nrays and pixel_start_end[-1][1] == rpp.size
    return ray_start_end, pixel_start_end


def subdivide_raytrace_jobs(rte_sensors, n_jobs, job_factor=1):
    """Same contract as at3d.parallel.subdivide_raytrace_jobs (at3d/parallel.py:114-174).

    ``rte_sensors``: OrderedDict key -> merged sensor; a merged sensor is anything with a
    ``rays_per_pixel`` integer array (xarray Dataset, dict or object).  Returns
    (keys, ray_start_end, pixel_start_end)."""
    def rpp_of(s):
        r = s['rays_per_pixel'] if isinstance(s, dict) or hasattr(s, 'keys') else s.rays_per_pixel
        return np.asarray(getattr(r, 'data', r), dtype=np.int64)
    counts = OrderedDict((k, int(rpp_of(s).sum())) for k, s in rte_sensors.items())
    ray_count = sum(counts.values())
    keys, ray_start_end, pixel_start_end = [], [], []
    for key, s in rte_sensors.items():
        njob = max(int(np.ceil(counts[key] / ray_count * n_jobs * job_factor)), 1)
        r, p = _aligned_chunks(counts[key], rpp_of(s), njob)
        ray_start_end.extend(r)
        pixel_start_end.extend(p)
        keys.extend([key] * len(r))
    return keys, ray_start_end, pixel_start_end


def shard_for_rank(rays_per_pixel, rank, world):
    """(ray_lo, ray_hi, pixel_lo, pixel_hi) of this rank's contiguous pixel-aligned share."""
    rpp = np.asarray(rays_per_pixel, dtype=np.int64)
    r, p = _aligned_chunks(int(rpp.sum()), rpp, world)
    return r[rank] + p[rank]


def allreduce_gradient(gradient, cost, group=None):
    """Sum the per-rank gradient and cost in place (torch tensors on the rank's device, or numpy
    arrays, which go through a CPU tensor).  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return gradient, cost
    nccl = dist.get_backend(group) == 'nccl'
    def red(x):
        if isinstance(x, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(x))
            if nccl:                            # NCCL reduces device tensors only: stage through the rank's GPU
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            x[...] = t.cpu().numpy().reshape(x.shape)
            return x
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x
    return red(gradient), red(cost)


def sharded_gradient(compute, rays, pix, rank, world, group=None):
    """One cost+gradient evaluation with the pixels sharded over ``world`` ranks.

    ``compute(rays_shard, pix_shard) -> (gradient[maxpg,numder] f64, cost[1] f64, stokesout)`` is the
    per-rank evaluation (``DeviceState.gradient`` on a GPU).  Returns the all-reduced (gradient, cost)
    and this rank's pixel Stokes vectors with its pixel range."""
    lo, hi, p0, p1 = shard_for_rank(pix.rays_per_pixel, rank, world)
    psh, r0, r1 = pix.slice_pixels(p0, p1)
    assert (r0, r1) == (lo, hi)
    g, c, so = compute(rays.slice(lo, hi), psh)[:3]
    g, c = allreduce_gradient(g, c, group)
    return g, c, so, (p0, p1)
