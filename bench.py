#!/usr/bin/env python
"""bench.py -- radiance+gradient rays/s of the SHDOM hot path on N B200s (one process per GPU).

A step is one full LEVISAPPROX_GRADIENT evaluation (forward radiances, adjoint weights, adjoint ray
pass, direct-beam pass) over all rays of the workload: BASELINE.json configs[1] -- LES-like cloud
field 32x37x27, NMU=16/NPHI=32 (NLM=256), 9 AirMSPI-like perspective views of 200x200 pixels, scalar
radiance + Levis gradient w.r.t. extinction.  Synthetic state (at3d_b200/synthetic.py), no solver.

  python bench.py [--gpus N --steps K --warmup W]          our CUDA path
  python bench.py --impl reference [...]                    the reference algorithm on the host CPU cores

Multi-GPU (torchrun, one rank per GPU): weak scaling.  Every rank holds its own replica of the
workload (BASELINE.json config 5: one wavelength / state per GPU, rays never cross GPUs) and the only
collective is the NCCL all-reduce of the per-voxel gradient and the cost; value = N x rays / max-rank time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VIEW_ZENITHS = [70.5, 60.0, 45.6, 26.1, 0.0, -26.1, -45.6, -60.0, -70.5]   # AirMSPI-like along-track


def build_scene(args):
    from at3d_b200 import synthetic as S
    if args.workload == 'cfg2':
        kw = dict(nx=32, ny=37, nz=27, nmu=16, nphi=32, nstokes=1, bc='open', dx=0.02, dy=0.02, dz=0.04,
                  cloud='les', ext_max=90.0, numphase=18, nsplits=1700, seed=0, truncate=True)
        npix_side = 200
    elif args.workload == 'cfg3':
        kw = dict(nx=32, ny=37, nz=27, nmu=16, nphi=32, nstokes=3, bc='open', dx=0.02, dy=0.02, dz=0.04,
                  cloud='les', ext_max=90.0, numphase=18, nsplits=1700, seed=0, truncate=True)
        npix_side = 200
    elif args.workload in ('cfg4', 'cfg4s'):
        # BASELINE.json configs[3]: 256x256x100 cloud + Rayleigh (NPART=2), 9 views x 512x512 (cfg4s: 128x128x64,
        # 9 x 256x256); Lambertian surface instead of the ocean BRDF (not on the GPU path yet)
        big = args.workload == 'cfg4'
        kw = dict(nx=256 if big else 128, ny=256 if big else 128, nz=100 if big else 64, nmu=16, nphi=32, nstokes=1,
                  bc='open', dx=0.02, dy=0.02, dz=0.04, cloud='les', ext_max=60.0, numphase=18, nsplits=0, seed=1,
                  truncate=True, rayleigh=True, mix_fraction=0.0)
        npix_side = 512 if big else 256
    else:   # 'small': CI-sized
        kw = dict(nx=16, ny=16, nz=14, nmu=8, nphi=16, nstokes=1, bc='open', dx=0.03, dy=0.03, dz=0.04,
                  cloud='les', ext_max=60.0, numphase=6, nsplits=60, seed=0)
        npix_side = 48
    if args.pixels:
        npix_side = args.pixels
    sc = S.make_scene(**kw)
    m = sc.meta
    cx, cy, cz = 0.5 * m['xmax'], 0.5 * m['ymax'], 0.5 * m['zmax']
    alt = 20.0
    views = []
    for zen in VIEW_ZENITHS:
        t = np.tan(np.deg2rad(zen))
        pos = (cx + (alt - cz) * t, cy, alt)
        dist = (alt - cz) / np.cos(np.deg2rad(zen))
        fov = 2.0 * np.rad2deg(np.arctan(0.45 * max(m['xmax'], m['ymax']) / dist))
        views.append(S.perspective_rays(pos, (cx, cy, cz), fov, npix_side, npix_side)[0])
    rays = S.concat_rays(views)
    return sc, rays, dict(kw, views=len(views), pixels_per_view=npix_side * npix_side)


def workload_text(name, st):
    return name + (': LES-like %dx%dx%d open BC, NLM=%d, NSTOKES=%d, NPART=%d, 9 perspective views, radiance + Levis '
                   'gradient (NUMDER=1); one replica of the workload per GPU, gradient all-reduced'
                   % (st.nx, st.ny, st.nz, st.nlm, st.nstokes, st.npart))


def compute_source_leg(B, st, steps, warmup, oracle=None):
    """COMPUTE_SOURCE (a1) on the workload's state: ms per call (CUDA events around the three kernels inside the C-ABI
    call, and wall time of the whole host-buffer call), algorithmic bytes per SURVEY 8(d) and the HBM fraction."""
    npts, nst = st.npts, st.nstokes
    tot = int(st.shptr[npts])
    if tot * nst * 4 > 6 * (1 << 30):
        return dict(skipped='host staging of SOURCE/DELSOURCE would need %.1f GB per array' % (tot * nst * 4 / 2**30))
    big = st.nlm * npts * nst * 4 > (1 << 31)
    fixsh = big                                    # adaptive truncation may grow NS up to NLM: needs MAXIV = NLM*NPTS
    maxiv = tot if big else st.nlm * npts
    source = np.zeros((nst, maxiv), np.float32, order='F')
    source[:, :tot] = st.source[:, :tot]
    delsource = np.zeros((nst, maxiv), np.float32, order='F')
    shptr = st.shptr[:npts + 1].copy()
    kms, wms, res = [], [], None
    for i in range(warmup + steps):
        t = time.perf_counter()
        res = B.compute_source(st, shptr, source, shptr.copy(), delsource, fixsh=fixsh, maxiv=maxiv, timing=True)
        if res[0] != 0:
            return dict(error='COMPUTE_SOURCE returned %d' % res[0])
        if i >= warmup:
            kms.append(res[-1]); wms.append(1e3 * (time.perf_counter() - t))
    ns_new = int(res[1][npts])
    nr = int(st.rshptr[npts])
    P = 28 + st.npart * (8 + 64 * st.maxnmicro)
    # 4*NSTOKES*(NR + NS_old + NS_new + 2*NS [acceleration: DELSOURCE read and written]) + P + 4*NPART + 8 per point
    b = 4 * nst * (nr + tot + ns_new + 2 * tot) + npts * (P + 4 * st.npart + 8)
    out = dict(kernel_ms=float(np.mean(kms)), call_ms_host_buffers=float(np.mean(wms)), algorithmic_bytes=int(b),
               achieved_gbs=b / (np.mean(kms) * 1e-3) / 1e9, npts=int(npts), sum_ns=tot, sum_nr=nr, fixsh=bool(fixsh))
    if oracle is not None:
        out['cpu_ms'] = oracle.compute_source(st, shptr, source, shptr.copy(), delsource, fixsh=fixsh, maxiv=maxiv,
                                              timing=True)[-1]
    return out


def compute_source_cfg4_leg(B, steps, warmup, peak, nx=256, ny=256, nz=100):
    """COMPUTE_SOURCE at the size of BASELINE.json configs[3] (256x256x100 = 6.55 M grid points, NLM=256, NSTOKES=1,
    cloud + Rayleigh) on DEVICE-RESIDENT arrays through at3d_compute_source_device: synthetic fields generated on the GPU
    (random SH with power-law decay, 35 % of the points truncated to a random shell, seed 0).  Two variants: the general
    call (adaptive truncation: norms kernel, SHPTR scan, write kernel) and FIXSH (one fused pass).  ms per call = CUDA
    events around the launches inside the call; bytes per SURVEY 8(d)."""
    import torch
    from at3d_b200 import grid as G
    try:
        g = torch.Generator(device='cuda'); g.manual_seed(0)
        npts, nmu, nphi, npart, nq = nx * ny * nz, 16, 32, 2, 8
        ml, mm, nlm = G.sh_sizes(nmu, nphi)
        nleg, numphase = ml, 19
        lj = torch.from_numpy(G.lofj(ml, mm).astype(np.int64)).cuda()

        def sh_field(scale):
            ltr = torch.where(torch.rand(npts, generator=g, device='cuda') < 0.35,
                              torch.randint(0, ml + 1, (npts,), generator=g, device='cuda'), torch.full((npts,), ml, device='cuda'))
            ns = torch.where(ltr <= mm, ltr * (ltr + 1) + ltr + 1, (2 * mm + 1) * ltr - mm * mm + mm + 1).to(torch.int32)
            ptr = torch.zeros(npts + 1, dtype=torch.int32, device='cuda')
            ptr[1:] = torch.cumsum(ns, 0).to(torch.int32)
            tot = int(ptr[npts])
            arr = torch.empty(tot, dtype=torch.float32, device='cuda')
            chunk = 1 << 26
            for a in range(0, tot, chunk):                       # bounded temporaries
                b = min(tot, a + chunk)
                arr[a:b] = scale * torch.randn(b - a, generator=g, device='cuda')
            return ptr, arr, tot
        rshptr, radiance, nr = sh_field(0.05)
        shptr, source, ns_tot = sh_field(0.1)
        delsource = 0.01 * source
        ext = torch.zeros((npart, npts), dtype=torch.float32, device='cuda')      # Fortran [npts, npart]
        ext[0] = 30.0 * torch.clamp(torch.rand(npts, generator=g, device='cuda') - 0.5, min=0.0)
        ext[1] = 0.02
        alb = torch.ones((npart, npts), dtype=torch.float32, device='cuda')
        total_ext = ext.sum(0)
        gs = np.linspace(0.80, 0.87, numphase - 1)
        legen = np.zeros((numphase, nleg + 1), np.float32)                         # Fortran [1, nleg+1, numphase]
        for k, gg in enumerate(gs):
            legen[k] = gg ** np.arange(nleg + 1) * (1.0 - gg ** (nleg + 1))        # delta-M-like scaling, values only matter as floats
        legen[-1, 0], legen[-1, 2] = 1.0, 0.5
        iphase = torch.ones((npart, npts, nq), dtype=torch.int32, device='cuda')   # Fortran [nq, npts, npart]
        iphase[0, :, 0] = torch.randint(1, numphase, (npts,), generator=g, device='cuda').to(torch.int32)
        iphase[1, :, 0] = numphase
        pwt = torch.zeros((npart, npts, nq), dtype=torch.float32, device='cuda')
        pwt[:, :, 0] = 1.0
        ylmsun = np.asfortranarray(B.ylmall(True, np.float32(-0.5), np.float32(0.3), ml, mm, 1, nlm).reshape(1, nlm))
        meta = dict(npts=npts, nstokes=1, nstleg=1, nlm=nlm, ml=ml, mm=mm, nleg=nleg, npart=npart, maxnmicro=1,
                    numphase=numphase, deltam=1, interp_new=1, srctype='S', phasemax=0.999, solarmu=-0.5)
        dev = B.DeviceSourceState(meta=meta, extinct=ext, albedo=alb, total_ext=total_ext,
                                  legen=torch.from_numpy(legen).cuda(), iphase=iphase, phaseinterpwt=pwt,
                                  dirflux=torch.rand(npts, generator=g, device='cuda'), rshptr=rshptr, radiance=radiance,
                                  ylmsun=torch.from_numpy(np.ascontiguousarray(ylmsun.ravel(order='F'))).cuda())
        cap = npts * nlm                                              # adaptive truncation may grow every point to NLM
        source_new = torch.empty(cap, dtype=torch.float32, device='cuda')
        shptr_new = torch.empty_like(shptr)
        out = dict(npts=npts, nlm=nlm, npart=npart, sum_nr=nr, sum_ns=ns_tot,
                   hbm_bytes=int(sum(t.numel() * t.element_size() for t in (radiance, source, delsource, delsource, source_new, iphase, pwt, ext, alb))))
        P = 28 + npart * (8 + 64 * 1)
        dl_new = torch.empty_like(delsource)
        for name, fixsh in (('adaptive_truncation', False), ('fixsh_fused', True)):
            kms = []
            tot_new = ns_tot
            for i in range(warmup + steps):
                dl = delsource.clone()
                # adaptive truncation: DELSOURCE double-buffered like SOURCE, so that the routine is one pass (cs_adapt_kernel)
                rc, tot_new, sums, ms = B.compute_source_device(dev, shptr, source, shptr, dl, shptr_new, source_new, fixsh=fixsh,
                                                                shacc=0.0 if fixsh else 3e-3, maxiv=cap, timing=True,
                                                                delsource_new=None if fixsh else dl_new)
                if rc != 0:
                    return dict(error='COMPUTE_SOURCE returned %d' % rc)
                if i >= warmup:
                    kms.append(ms)
            b = 4 * (nr + ns_tot + tot_new + 2 * ns_tot) + npts * (P + 4 * npart + 8)
            k = float(np.mean(kms))
            out[name] = dict(kernel_ms=k, algorithmic_bytes=int(b), achieved_gbs=b / (k * 1e-3) / 1e9,
                             frac=b / (k * 1e-3) / 1e9 / peak, sum_ns_new=int(tot_new))
        return out
    except Exception as e:                                            # e.g. not enough device memory on a shared GPU
        return dict(skipped='%s: %s' % (type(e).__name__, e))


def orthographic_leg(dev, sc, npix_side, steps, warmup):
    """RENDER of 9 orthographic views (all rays of a view share their direction): the library evaluates the source of
    every grid point once per view (view_source_kernel) instead of once per (ray, corner).  Host ray arrays through the
    C ABI; kernel ms = CUDA events around the launches inside the call."""
    from at3d_b200 import synthetic as S
    m = sc.meta
    res = max(m['xmax'], m['ymax']) / npix_side
    views = [S.orthographic_rays(sc, abs(z), 0.0 if z >= 0 else 180.0, res)[0] for z in VIEW_ZENITHS]
    rays = S.concat_rays(views)
    ms, wall = [], []
    for i in range(warmup + steps):
        t = time.perf_counter()
        o = dev.render(rays, timing=True)
        if i >= warmup:
            ms.append(o[-1]); wall.append(time.perf_counter() - t)
    return dict(rays=int(rays.nrays), kernel_ms=float(np.mean(ms)), rays_per_s=rays.nrays / (np.mean(ms) * 1e-3),
                e2e_rays_per_s=rays.nrays / float(np.mean(wall)))


def solver_leg(steps, warmup, oracle=None, n=96, nz=48):
    """One SHDOM solution iteration (PATH_INTEGRATION + COMPUTE_SOURCE) on an independent-pixel grid
    of n x n columns x nz levels, NLM=256: kernel ms on the GPU (CUDA events inside the two C-ABI calls) and, when the
    oracle is given, the same iteration of the reference algorithm on one host core."""
    from at3d_b200 import synthetic as S, backend as B, solver
    sc = S.make_scene(nx=n, ny=n, nz=nz, nmu=16, nphi=32, nstokes=1, bc='periodic', dx=0.02, dy=0.02, dz=0.04,
                      cloud='les', ext_max=40.0, numphase=8, nsplits=0, seed=3, truncate=False, ipflag=3, mix_fraction=0.0)
    B.finalize_scene(sc)
    st = sc.state
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    wtmu = (st.wtdo[:, 0] / delphi).astype(np.float32)
    npts = st.npts
    maxiv = st.nlm * npts
    source = np.zeros((1, maxiv), np.float32, order='F')
    tot = int(st.shptr[npts])
    source[:, :tot] = st.source[:, :tot]
    delsource = np.zeros((1, maxiv), np.float32, order='F')
    rshptr = solver.radiance_truncation(st, st.shptr, st.radiance, st.rshptr, False, 0.0, False, maxiv + npts)
    pms, cms = [], []
    for i in range(warmup + steps):
        rad, fluxes, bcrad, ms = solver.path_integration_ip(st, wtmu, st.shptr, source, rshptr, timing=True)
        st2 = st.copy(); st2.rshptr, st2.radiance = rshptr, rad
        res = B.compute_source(st2, st.shptr, source, st.shptr.copy(), delsource, maxiv=maxiv, timing=True)
        if i >= warmup:
            pms.append(ms); cms.append(res[-1])
    out = dict(grid='%dx%dx%d independent-pixel columns' % (n, n, nz), npts=int(npts), nang=int(st.nphi0.sum()),
               path_integration_ms=float(np.mean(pms)), compute_source_ms=float(np.mean(cms)),
               iteration_ms=float(np.mean(pms) + np.mean(cms)))
    if oracle is not None:
        # the oracle's solve runs the same two routines per iteration; time one iteration
        t = time.perf_counter()
        oracle.solve_fixed_grid(st, wtmu, maxiter=1, solacc=1e-9)
        out['cpu_iteration_ms'] = 1e3 * (time.perf_counter() - t)
        out['cpu_cores'] = 1
        out['cpu_note'] = 'oracle solve with maxiter=1: first-guess COMPUTE_SOURCE + one full iteration'
    return out


def sweep3d_leg(st, steps, warmup, oracle=None):
    """PATH_INTEGRATION on the workload's own 3-D grid (BACK_INT_GRID3D data-flow sweep, all ordinates of a hemisphere per
    launch): kernel ms per call (CUDA events inside the C-ABI call, SH_TO_DO and DO_TO_SH included) and
    (ordinate, grid point) updates per second; with the oracle, the reference's serial sweep on one host core."""
    from at3d_b200 import solver
    if st.ipflag & 2:
        return dict(skipped='independent-pixel grid')
    nang = int(st.nphi0.sum())
    if st.npts * st.nstokes * nang * 8 > 60 * (1 << 30):
        return dict(skipped='two discrete-ordinate fields would need %.0f GB' % (st.npts * st.nstokes * nang * 8 / 2**30))
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    wtmu = (st.wtdo[:, 0] / delphi).astype(np.float32)
    t0 = time.perf_counter()
    sv = solver.SweepSolver(st, wtmu)
    setup_ms = 1e3 * (time.perf_counter() - t0)
    ms = []
    for i in range(warmup + steps):
        out = sv.path_integration(st.shptr, st.source, st.rshptr, timing=True)
        if i >= warmup:
            ms.append(out[3])
    sv.close()
    res = dict(npts=int(st.npts), nang=nang, path_integration_ms=float(np.mean(ms)), setup_ms=setup_ms,
               point_updates_per_s=st.npts * nang / (np.mean(ms) * 1e-3))
    # the whole fixed-grid solve (RTE.solve with split_accuracy=0), device-resident loop, host buffers in and out
    maxit = 8
    t0 = time.perf_counter()
    sol, iters, solcrit, tm = solver.solve_fixed_grid(st, wtmu, solacc=1e-4, maxiter=maxit)
    res['solve'] = dict(iterations=iters, solcrit=solcrit, loop_ms=tm['loop_ms'], path_integration_ms=tm['path_integration_ms'],
                        compute_source_ms=tm['compute_source_ms'], e2e_ms=1e3 * (time.perf_counter() - t0), maxiter=maxit)
    if oracle is not None:
        t = time.perf_counter()
        oracle.path_integration(st, wtmu, st.shptr, st.source, st.rshptr)
        res['cpu_path_integration_ms'] = 1e3 * (time.perf_counter() - t)
        t = time.perf_counter()
        _, it_cpu, _ = oracle.solve_fixed_grid(st, wtmu, solacc=1e-4, maxiter=maxit)
        res['solve']['cpu_ms'] = 1e3 * (time.perf_counter() - t)
        res['solve']['cpu_iterations'] = it_cpu
        res['cpu_cores'] = 1
    return res


def inversion_step_leg(sc, rays, gi, pix, DeviceState):
    """One cost-function + gradient evaluation of an inversion (BASELINE.json configs[4], per wavelength / GPU): the
    fixed-grid solve of the workload's medium from scratch (device-resident iterations), the upload of the solved state,
    the derivative tables, and the radiance + gradient pass with host buffers; wall-clock ms of each part.  `warm` is the
    next evaluation of the loop: a slightly different medium on the same grid given to the LIVE solver object
    (at3d_solver_update_medium: sweep order, dependency levels and sorted plan kept), solved, uploaded, differentiated."""
    from at3d_b200 import solver, backend, gradsetup
    was = backend.memory_reuse(True)      # what Optimizer.minimize does: freed device memory stays with the library
    # what RTE.levis_approx_gradient attaches: no dense DPATH/DPTR lists (185 MB to upload per evaluation here), the beam
    # walks run inside the gradient call
    if gi.dpath is not None and gi.exact_single_scatter:
        gi = gradsetup.with_streaming_beam(gi, sc.state, sc.pg, backend)
    st = sc.state
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    wtmu = (st.wtdo[:, 0] / delphi).astype(np.float32)
    t0 = time.perf_counter()
    sv = solver.SweepSolver(st.copy().normalize(), wtmu)
    tc = time.perf_counter()
    sol, iters, solcrit, tm = sv.solve(solacc=1e-4, maxiter=60)
    t1 = time.perf_counter()
    dev = DeviceState(sol)
    dev.attach_gradient(gi)
    t2 = time.perf_counter()
    g, cost, so = dev.gradient(rays, pix)
    t3 = time.perf_counter()
    dev.close()
    out = dict(solve_iterations=iters, solcrit=solcrit, solver_create_ms=1e3 * (tc - t0), solve_ms=1e3 * (t1 - t0),
               solve_loop_ms=tm.get('loop_ms'), state_upload_ms=1e3 * (t2 - t1), gradient_ms=1e3 * (t3 - t2),
               total_ms=1e3 * (t3 - t0), rays=int(rays.nrays), cost=float(cost[0]))
    # the next evaluation: the optimiser moved the extinction a little
    st2 = st.copy()
    st2.extinct = np.asfortranarray(st.extinct * np.float32(1.02))
    st2.total_ext = (st.total_ext * np.float32(1.02)).astype(np.float32)
    st2.normalize()
    w0 = time.perf_counter()
    sv.update_medium(st2)
    w1 = time.perf_counter()
    sol2, iters2, solcrit2, tm2 = sv.solve(solacc=1e-4, maxiter=60, initial=sol)   # continued from the previous solution
    w2 = time.perf_counter()
    dev = DeviceState(sol2)
    dev.attach_gradient(gi)
    w3 = time.perf_counter()
    g2, cost2, so2 = dev.gradient(rays, pix)
    w4 = time.perf_counter()
    dev.close()
    sv.close()
    out['warm'] = dict(update_medium_ms=1e3 * (w1 - w0), solve_ms=1e3 * (w2 - w1), solve_loop_ms=tm2.get('loop_ms'),
                       solve_iterations=iters2, state_upload_ms=1e3 * (w3 - w2), gradient_ms=1e3 * (w4 - w3),
                       total_ms=1e3 * (w4 - w0), cost=float(cost2[0]),
                       note='live solver object (at3d_solver_update_medium), iterations continued from the previous solution (at3d_solver_solve_from), memory reuse (at3d_set_memory_reuse)')
    backend.memory_reuse(was)
    if not was:
        backend.trim_memory()
    return out


def tensor_peak_tf32():
    """Dense TF32 tensor peak in TFLOP/s: half of the measured dense bf16 figure of MEASURED_PEAKS.json (the TF32 path of
    tcgen05 runs at half the bf16 rate: ncu's sm__ops_path_tensor_op_utchmma peaks, 4096 vs 8192 per cycle and SM), else
    half of the 2250 nominal."""
    try:
        d = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(d.get('bf16_tflops', 2250.0)) / 2.0
    except Exception:
        return 1125.0


def transform_leg(B, st, steps, warmup):
    """SH_TO_DO / DO_TO_SH on the workload's SOURCE / RADIANCE: kernel ms (CUDA events inside the
    C-ABI call), FP32 FMA rate against the CUDA-core peak and bytes against HBM."""
    npts, nst = st.npts, st.nstokes
    nang = int(st.nphi0.sum())
    if npts * nst * nang * 4 > 6 * (1 << 30):
        return dict(skipped='host staging of DOFIELD would need %.1f GB' % (npts * nst * nang * 4 / 2**30))
    delphi = np.float32(2.0 * np.pi) / st.nphi0.astype(np.float32)
    wtmu = (st.wtdo[:, 0] / delphi).astype(np.float32)
    ms1, ms2, do = [], [], None
    for i in range(warmup + steps):
        do, t = B.sh_to_do(st, wtmu, st.shptr, st.source, timing=True)
        if i >= warmup:
            ms1.append(t)
    for i in range(warmup + steps):
        _, t = B.do_to_sh(st, wtmu, st.rshptr, do, timing=True)
        if i >= warmup:
            ms2.append(t)
    me = np.clip(st.nphi0 // 2 - 1, 0, st.mm)
    azmacs = int(np.sum(st.nphi0 * (2 * me + 1)))                      # azimuthal stage, per point and Stokes plane
    planes = 1 if nst == 1 else 5                                      # I; Q,U each from two coupled terms
    tot_s, tot_r = int(st.shptr[npts]), int(st.rshptr[npts])
    fl1 = 2.0 * (planes * tot_s * st.nmu + nst * npts * azmacs)
    fl2 = 2.0 * (planes * tot_r * st.nmu + nst * npts * azmacs)
    by1 = 4.0 * nst * (tot_s + npts * nang)
    by2 = 4.0 * nst * (tot_r + npts * nang)
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12                         # TFLOP/s, CUDA-core FMA at the boost clock
    t1, t2 = float(np.mean(ms1)) * 1e-3, float(np.mean(ms2)) * 1e-3
    # the tensor-core variant of SH_TO_DO (NSTOKES=1): tcgen05.mma kind::tf32, 3xTF32 split, dense [NPTS x NLM].[NLM x NANG]
    tc = None
    if nst == 1:
        old = os.environ.get('AT3D_B200_TRANSFORM')
        os.environ['AT3D_B200_TRANSFORM'] = 'tc'
        try:
            ms3 = []
            for i in range(warmup + steps):
                do_tc, t = B.sh_to_do(st, wtmu, st.shptr, st.source, timing=True)
                if i >= warmup:
                    ms3.append(t)
            t3 = float(np.mean(ms3)) * 1e-3
            ms4 = []
            for i in range(warmup + steps):
                sh_tc, t = B.do_to_sh(st, wtmu, st.rshptr, do, timing=True)
                if i >= warmup:
                    ms4.append(t)
            t4 = float(np.mean(ms4)) * 1e-3
            ntot = (nang + 15) // 16 * 16
            tensor_flops = 3 * 2.0 * (-(-npts // 128) * 128) * (-(-st.nlm // 32) * 32) * ntot     # what the MMAs execute
            tf32_peak = tensor_peak_tf32()
            back_flops = 3 * 2.0 * (-(-npts // 128) * 128) * (-(-nang // 32) * 32) * (-(-st.nlm // 16) * 16)
            tc = dict(ms=1e3 * t3, speedup_vs_fp32=t1 / t3, tensor_tflops=tensor_flops / t3 / 1e12,
                      tensor_pipe_frac=tensor_flops / t3 / 1e12 / tf32_peak, tf32_peak_tflops=tf32_peak,
                      do_to_sh=dict(ms=1e3 * t4, speedup_vs_fp32=t2 / t4, tensor_tflops=back_flops / t4 / 1e12,
                                    tensor_pipe_frac=back_flops / t4 / 1e12 / tf32_peak),
                      max_rel_diff_vs_fp32=float(np.abs(do_tc - do).max() / np.abs(do).max()),
                      note='3xTF32: hi.hi + lo.hi + hi.lo, FP32 accumulation in TMEM; opt-in (AT3D_B200_TRANSFORM=tc)')
        finally:
            if old is None:
                del os.environ['AT3D_B200_TRANSFORM']
            else:
                os.environ['AT3D_B200_TRANSFORM'] = old
    return dict(sh_to_do_ms=1e3 * t1, do_to_sh_ms=1e3 * t2, nang=nang, sh_to_do_tensor_core=tc,
                sh_to_do=dict(tflops=fl1 / t1 / 1e12, fma_pipe_frac=fl1 / t1 / 1e12 / fp32_peak, gbs=by1 / t1 / 1e9),
                do_to_sh=dict(tflops=fl2 / t2 / 1e12, fma_pipe_frac=fl2 / t2 / 1e12 / fp32_peak, gbs=by2 / t2 / 1e9),
                fp32_peak_tflops=fp32_peak)


def algorithmic_bytes(st, gi, cnt, gradient=True):
    """bytes the reference algorithm must touch for the work actually done
    (cells / evaluated grid points / SH lengths / sub-intervals counted by the kernel)."""
    nst = st.nstokes
    P = 28 + st.npart * (8 + 64 * st.maxnmicro)          # per-point scalar payload
    Ccell = 66                                           # GRIDPTR 32 + NEIGHPTR 24 + TREEPTR 8 + CELLFLAGS 2
    b = 4 * nst * cnt['sum_ns'] + P * cnt['points'] + Ccell * cnt['cells']
    if gradient:
        nd = gi.numder
        b += 4 * nst * cnt['sum_nr']                     # radiance expansion, read once per new point
        b += cnt['points'] * (8 * nd * 28 + 8 * 8)       # derivative tables + INTERPPTR/OPTINTERPWT
        b += 2 * cnt['subintervals'] * 8 * 8 * nd * 8    # gradient scatter: source term and radiance term
    return int(b)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self._p, self._t = index, [], None, None

    def _run(self):
        for line in self._p.stdout:
            f = [x.strip() for x in line.strip().split(',')]
            if len(f) >= 6:
                self.samples.append(f)

    def __enter__(self):
        # ONE nvidia-smi process looping every 200 ms (the recipe's clocks line): started before the timed region -- its
        # start-up (driver / NVML initialisation) is over when the first sample arrives -- and stopped after it.  Spawning
        # a new nvidia-smi per sample stalls CUDA calls of the timed process on some boxes (tens of ms per spawn).
        try:
            self._p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                        '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                       stderr=subprocess.DEVNULL, text=True)
            first = self._p.stdout.readline()
            f = [x.strip() for x in first.strip().split(',')]
            if len(f) >= 6:
                self.samples.append(f)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self._p = None
        return self

    def __exit__(self, *a):
        if self._p is not None:
            time.sleep(0.25)                      # one more sample under load
            self._p.terminate()                   # the exact process started above
            try:
                self._p.wait(timeout=5)
            except Exception:
                self._p.kill()
            if self._t is not None:
                self._t.join(timeout=5)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons)


def measured_traffic(workload):
    """The ncu --set full capture of one gradient call of `workload` (profiles/traffic.json, written by
    profiles/make_traffic.py from the .ncu-rep of the same round): DRAM bytes of the step, issue-slot utilisation and the
    per-kernel shares, or None when the workload was not captured."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))[workload]
    except Exception:
        return None


def measured_peak():
    """HBM peak for the roofline: the driver-written MEASURED_PEAKS.json when present (the sustained figure is preferred:
    the kernels are timed inside a step), else the fallback of B200_PROFILING.md.  The file's schema is not fixed here:
    every numeric entry whose key path mentions HBM / copy bandwidth is considered; values below 50 are read as TB/s."""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            found = []

            def walk(x, path):
                if isinstance(x, dict):
                    for k, v in x.items():
                        walk(v, path + [str(k).lower()])
                elif isinstance(x, (list, tuple)):
                    for i, v in enumerate(x):
                        walk(v, path + [str(i)])
                elif isinstance(x, (int, float)) and not isinstance(x, bool):
                    key = '/'.join(path)
                    if any(t in key for t in ('hbm', 'copy', 'dram', 'mem_bw', 'membw', 'bandwidth')) and \
                            not any(t in key for t in ('flop', 'tf', 'nvlink', 'pcie', 'bytes', 'size', 'ms', 'time')):
                        v = float(x) * (1000.0 if 0 < float(x) < 50 else 1.0)
                        if 1000.0 <= v <= 9000.0:
                            found.append((key, v))
            walk(json.load(open(p)), [])
            if found:
                pref = [f for f in found if 'sustain' in f[0]] or found
                key, v = pref[0]
                return v, 'measured (MEASURED_PEAKS.json: %s)' % key
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def cpu_reference_rate(sc, rays, gi, pix, target_s, nthreads, steps=1, warmup=0):
    """Times the CPU restatement of the reference algorithm (oracle/, test infrastructure) on a bounded
    sample of the workload's pixels with `nthreads` OpenMP threads.  Returns (rays/s, sample text, ms list)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_lib as O
    from at3d_b200 import gradsetup
    from at3d_b200.state import Rays
    O.build()
    rng = np.random.default_rng(0)
    npix = pix.npix
    # calibration on a small random sample of pixels spread over all views

    def run(nsample):
        idx = np.sort(rng.choice(npix, size=min(nsample, npix), replace=False))
        starts = np.concatenate([[0], np.cumsum(pix.rays_per_pixel)])
        ridx = np.concatenate([np.arange(starts[p], starts[p + 1]) for p in idx])
        r = Rays(rays.camx[ridx], rays.camy[ridx], rays.camz[ridx], rays.cammu[ridx], rays.camphi[ridx])
        p = gradsetup.PixelData(pix.measurements[:, idx], pix.uncertainties[:, :, idx], pix.rays_per_pixel[idx],
                                pix.ray_weights[ridx], pix.stokes_weights[:, idx])
        g = gradsetup.with_pixels(gi, p)
        t = time.perf_counter()
        O.levisapprox_gradient(sc.state, r, g, nthreads=nthreads)
        return r.nrays, time.perf_counter() - t
    n0, t0 = run(64 * nthreads)
    rate0 = n0 / max(t0, 1e-6)
    nsample = int(min(max(rate0 * target_s, 64 * nthreads), npix))
    times = []
    nr = 0
    for i in range(warmup + steps):
        nr, t = run(nsample)
        if i >= warmup:
            times.append(t)
    tm = float(np.mean(times))
    return nr / tm, '%d of %d rays (random pixels over all views), %d OpenMP threads, %.1f s/step' % (
        nr, rays.nrays, nthreads, tm), [1e3 * t for t in times]


def render_only(args, dev, st, sc, rays, dr, l2flush, stream, barrier, world, rank, local, ncores, B, DeviceState):
    """BASELINE.json configs[3]: RENDER of all rays (Lambertian and ocean surface), COMPUTE_SOURCE; no gradient."""
    import torch
    import torch.distributed as dist
    from at3d_b200 import synthetic as S
    nrays = rays.nrays
    rsout = torch.zeros((nrays, st.nstokes), dtype=torch.float32, device='cuda')
    devo = DeviceState(S.with_brdf_surface(st, 'O' if st.nstokes == 1 else 'W', seed=2, wavelen=0.66))

    def timed(d):
        for _ in range(args.warmup):
            l2flush.zero_()
            d.render(dr, out=rsout, stream=stream)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kms = []
        for i in range(args.steps):
            l2flush.zero_()
            ev[i][0].record()
            o = d.render(dr, out=rsout, stream=stream, timing=True)
            ev[i][1].record()
            kms.append(o[-1])
        barrier()
        t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t, float(np.mean(kms)), d.counts()
    with ClockSampler(local) as cs:
        t0 = time.perf_counter()
        t_ocean, ms_ocean, c_ocean = timed(devo)
        t_lamb, ms_lamb, c_lamb = timed(dev)
        wall = time.perf_counter() - t0
    # e2e: host ray arrays in, host Stokes out, through the C-ABI call (ocean surface)
    devo.render(rays)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_h = devo.render(rays)
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    peak, peak_src = measured_peak()
    rbytes = algorithmic_bytes(st, None, c_lamb, gradient=False)
    csrc = compute_source_leg(B, st, args.steps, args.warmup, None)
    if 'kernel_ms' in csrc:
        csrc['frac'] = csrc['achieved_gbs'] / peak
    hbm = dev.hbm_bytes
    devo.close()
    if rank == 0:
        line = dict(
            metric='radiance rays/s', value=world * nrays * args.steps / t_ocean, unit='rays/s', n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * t_ocean / args.steps, higher_is_better=True,
            scaling='weak', vs_baseline=None, dtype='f32 optics / f64 geometry+accumulators', data='synthetic',
            config=dict(workload=args.workload + ': LES-like %dx%dx%d open BC + Rayleigh (NPART=%d), NLM=%d, ocean BRDF surface, '
                        '9 perspective views, RENDER only; one replica of the ray set per GPU'
                        % (st.nx, st.ny, st.nz, st.npart, st.nlm), rays=int(nrays), npts=int(st.npts),
                        ncells=int(st.ncells), nlm=int(st.nlm), l2='flushed between steps (256 MiB memset)',
                        hbm_state_bytes=hbm),
            e2e=dict(value=world * nrays * args.steps / t_e2e, unit='rays/s', h2d_bytes_per_step=int(nrays * 96),
                     d2h_bytes_per_step=int(nrays * st.nstokes * 4)),
            gpu_launches=int(args.steps * 2),
            roofline=dict(kernel='forward_kernel_t (RENDER march, Lambertian run)', bound='hbm',
                          achieved=rbytes / (ms_lamb * 1e-3) / 1e9, peak=peak, unit='GB/s',
                          frac=rbytes / (ms_lamb * 1e-3) / 1e9 / peak, traffic=None, peak_source=peak_src,
                          algorithmic_bytes_per_launch=rbytes, kernel_ms=ms_lamb, counts=c_lamb),
            render=dict(rays_per_s=world * nrays * args.steps / t_lamb, kernel_ms=ms_lamb),
            render_ocean=dict(rays_per_s=world * nrays * args.steps / t_ocean, kernel_ms=ms_ocean,
                              surface_ms=ms_ocean - ms_lamb, surface_hits=c_ocean['surface_hits'],
                              brdf_evals=c_ocean['surface_hits'] * 4 * (st.nang // 2 + 1)),
            compute_source=csrc, compute_source_cfg4=compute_source_cfg4_leg(B, max(3, args.steps // 2), 2, peak),
            transforms=transform_leg(B, st, args.steps, args.warmup),
            render_orthographic=orthographic_leg(dev, sc, int(round((nrays / 9) ** 0.5)), args.steps, args.warmup),
            clocks=cs.summary(), cpu_baseline=None, wall_s=wall)
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg4s', 'small'])
    ap.add_argument('--pixels', type=int, default=0, help='pixels per view side (default: per workload)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--render-only', action='store_true', help='RENDER only')
    ap.add_argument('--stream-beam', action='store_true',
                    help='streaming direct-beam derivative: no dense DPATH/DPTR lists (forced for cfg4, where they would '
                         'need LONGEST_PATH_PTS x NPTS x 8 B = ~125 GB at 256x256x100, in the reference as well)')
    ap.add_argument('--lean', action='store_true', help='gradient step only: skip the side legs (forced for cfg4)')
    args = ap.parse_args()
    if args.workload == 'cfg4':
        args.stream_beam = True
        args.lean = True
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)

    if world > 1 and args.impl != 'reference':
        # torchrun pins OMP_NUM_THREADS=1; the host side of the C-ABI call (per-ray setup with the host libm)
        # is OpenMP-parallel: give every rank its share of the cores
        lw = int(os.environ.get('LOCAL_WORLD_SIZE', world))
        os.environ['OMP_NUM_THREADS'] = str(max(1, ncores // max(lw, 1)))
    if args.impl == 'reference':
        # the reference's own algorithm on the host cores; rank 0 only
        if rank != 0:
            return
        from at3d_b200 import gradsetup
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import oracle_lib as O
        sc, rays, cfg = build_scene(args)
        O.finalize_scene(sc)
        gi = gradsetup.make_gradient_inputs(sc, O, seed=0, numder=1)
        rad = O.render(sc.state, rays, nthreads=ncores)
        pix = gradsetup.make_pixels(sc.state.nstokes, rays.nrays, rad, seed=1)
        rate, sample, ms = cpu_reference_rate(sc, rays, gi, pix, args.cpu_seconds, ncores, args.steps, args.warmup)
        line = dict(metric='radiance+gradient rays/s', value=rate, unit='rays/s', n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=float(np.mean(ms)), higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32 optics / f64 geometry+accumulators', data='synthetic',
                    impl='reference',
                    config=dict(workload=workload_text(args.workload, sc.state), rays=int(rays.nrays),
                                npts=int(sc.state.npts), ncells=int(sc.state.ncells), nlm=int(sc.state.nlm)),
                    cpu_baseline=dict(value=rate, unit='rays/s', cores=ncores, kind='port', sample=sample,
                                      compute_source_ms=compute_source_leg(O, sc.state, 1, 0)['kernel_ms'],
                                      compute_source_cores=1),
                    e2e=dict(value=rate, unit='rays/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL prints while the communicator is created (its version
        # line when NCCL_DEBUG is set on the box) goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            warm = torch.zeros(1, device='cuda')
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from at3d_b200 import backend as B, gradsetup
    from at3d_b200.device import DeviceState
    from at3d_b200 import _lib
    _lib.lib().at3d_set_device(local)

    sc, rays, cfg = build_scene(args)
    B.finalize_scene(sc)
    st = sc.state
    dev = DeviceState(st)
    nrays = rays.nrays
    stream = torch.cuda.current_stream().cuda_stream

    class Bag:
        pass
    dr, dp = Bag(), Bag()
    for k in ('camx', 'camy', 'camz', 'cammu', 'camphi'):
        setattr(dr, k, torch.from_numpy(getattr(rays, k)).cuda())
    # the per-ray setup records evaluated with the host libm (the reference's own), resident in HBM with the rays: the
    # device-timed `value` walks exactly the cells of the host-buffer (`e2e`) path
    dr.packs = dev.make_ray_packs(rays)
    l2flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.render_only:
        render_only(args, dev, st, sc, rays, dr, l2flush, stream, barrier, world, rank, local, ncores, B, DeviceState)
        dev.close()
        if world > 1:
            dist.destroy_process_group()
        return

    gi = gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1, stream_beam=args.stream_beam)
    dev.attach_gradient(gi)
    rad = dev.render(rays)
    pix = gradsetup.make_pixels(st.nstokes, rays.nrays, rad, seed=1)
    npix = pix.npix

    # ---- device-resident inputs (value) ----
    dp.measurements = torch.from_numpy(np.ascontiguousarray(pix.measurements.T)).cuda()
    dp.uncertainties = torch.from_numpy(np.ascontiguousarray(pix.uncertainties.transpose(2, 1, 0))).cuda()
    dp.rays_per_pixel = torch.from_numpy(pix.rays_per_pixel).cuda()
    dp.ray_weights = torch.from_numpy(pix.ray_weights).cuda()
    dp.stokes_weights = torch.from_numpy(np.ascontiguousarray(pix.stokes_weights.T)).cuda()
    gout = torch.zeros((gi.numder, gi.maxpg), dtype=torch.float64, device='cuda')
    sout = torch.zeros((npix, st.nstokes), dtype=torch.float32, device='cuda')
    cout = torch.zeros(1, dtype=torch.float64, device='cuda')

    def step_device(timing=False):
        l2flush.zero_()
        out = dev.gradient(dr, dp, gradout=gout, stokesout=sout, cost=cout, stream=stream, timing=timing)
        if world > 1:
            dist.all_reduce(gout)          # the only collective of the path: per-voxel gradient (+cost)
            dist.all_reduce(cout)
        return out

    for _ in range(args.warmup):
        step_device()
    barrier()
    kms = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as cs:
        t0 = time.perf_counter()
        for i in range(args.steps):
            l2flush.zero_()
            ev[i][0].record()
            out = dev.gradient(dr, dp, gradout=gout, stokesout=sout, cost=cout, stream=stream, timing=True)
            if world > 1:
                dist.all_reduce(gout)
                dist.all_reduce(cout)
            ev[i][1].record()
            kms.append(out[-1])
        barrier()
        wall = time.perf_counter() - t0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_total = sum(step_ms) * 1e-3
    if world > 1:
        tt = torch.tensor([t_total], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total = float(tt.item())
    counts = dev.counts()
    kms = np.array(kms)            # [steps, 4] forward, adjoint(+pixel), beam, total
    value = world * nrays * args.steps / t_total

    # ---- strong scaling (N > 1): the SAME ray set cut into N contiguous pixel-aligned ranges (the rule of
    # at3d/parallel.py:114-174), one per rank on its replica of the state; the pixel values are gathered and the
    # gradient and cost all-reduced over NCCL inside the timed region ----
    strong = None
    if world > 1:
        from at3d_b200.parallel import shard_for_rank
        r0, r1, p0, p1 = shard_for_rank(pix.rays_per_pixel, rank, world)
        drk, dpk = Bag(), Bag()
        for k in ('camx', 'camy', 'camz', 'cammu', 'camphi'):
            setattr(drk, k, getattr(dr, k)[r0:r1])
        dpk.measurements = dp.measurements[p0:p1]
        dpk.uncertainties = dp.uncertainties[p0:p1]
        dpk.rays_per_pixel = dp.rays_per_pixel[p0:p1]
        dpk.ray_weights = dp.ray_weights[r0:r1]
        dpk.stokes_weights = dp.stokes_weights[p0:p1]
        gout_s = torch.zeros_like(gout)
        cout_s = torch.zeros_like(cout)
        sout_all = torch.zeros_like(sout)

        def step_strong():
            l2flush.zero_()
            sout_all.zero_()
            dev.gradient(drk, dpk, gradout=gout_s, stokesout=sout_all[p0:p1], cost=cout_s, stream=stream)
            dist.all_reduce(gout_s)
            dist.all_reduce(cout_s)
            dist.all_reduce(sout_all)          # ranks own disjoint pixel ranges: the sum is the gather of the pixel values
        for _ in range(args.warmup):
            step_strong()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            evs[i][0].record()
            step_strong()
            evs[i][1].record()
        barrier()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) * 1e-3], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_strong = float(tt.item())
        # the sharded evaluation must reproduce the replica's: same pixel values, gradient equal to FP64 rounding
        gerr = float((gout_s - gout / world).abs().max() / (gout / world).abs().max()) if world > 1 else 0.0
        strong = dict(value=nrays * args.steps / t_strong, unit='rays/s', ms_per_step=1e3 * t_strong / args.steps,
                      rays_per_rank=int(r1 - r0), scaling='strong',
                      efficiency_vs_one_gpu_same_run=(t_total / args.steps) / (world * t_strong / args.steps),
                      gradient_max_rel_diff_vs_replica=gerr,
                      note='one_gpu time = this run\'s weak step (every rank marches the whole ray set); collectives: '
                           'all-reduce of GRADOUT f64[%d], cost, and the gathered pixel values f32[%d]'
                           % (gi.maxpg * gi.numder, npix * st.nstokes))

    # ---- e2e: host buffers (pinned) through the C-ABI call, copies inside the timed region ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    for k in ('camx', 'camy', 'camz', 'cammu', 'camphi'):
        setattr(rays, k, pinned(getattr(rays, k)))
    for _ in range(max(1, args.warmup // 2)):
        dev.gradient(rays, pix)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g_h, c_h, s_h = dev.gradient(rays, pix)
        if world > 1:
            gt = torch.from_numpy(g_h).cuda()
            dist.all_reduce(gt)
            g_h = gt.cpu().numpy()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    h2d = nrays * (80 + 16) + pix.measurements.nbytes + pix.uncertainties.nbytes + pix.rays_per_pixel.nbytes + \
        pix.ray_weights.nbytes + pix.stokes_weights.nbytes
    d2h = gi.maxpg * gi.numder * 8 + npix * st.nstokes * 4 + 8

    # ---- RENDER alone and roofline of the dominant kernel ----
    rsout = torch.zeros((nrays, st.nstokes), dtype=torch.float32, device='cuda')
    rms = []
    for i in range(args.warmup + args.steps):
        l2flush.zero_()
        o = dev.render(dr, out=rsout, stream=stream, timing=True)
        if i >= args.warmup:
            rms.append(o[-1])
    rcounts = dev.counts()
    # ---- RENDER over an ocean surface (BASELINE.json configs[3]: ocean BRDF): every ray that reaches the surface
    # costs 4 x (NANG/2 + 1) evaluations of ocean_brdf_sw in surface_kernel ----
    from at3d_b200 import synthetic as S
    oms, ocounts = [], None
    if not args.lean:
        devo = DeviceState(S.with_brdf_surface(st, 'O' if st.nstokes == 1 else 'W', seed=2, wavelen=0.66))
        for i in range(args.warmup + args.steps):
            l2flush.zero_()
            o = devo.render(dr, out=rsout, stream=stream, timing=True)
            if i >= args.warmup:
                oms.append(o[-1])
        ocounts = devo.counts()
        devo.close()
    peak, peak_src = measured_peak()
    adj_ms = float(np.mean(kms[:, 1]))
    step_kernel_ms = float(np.mean(kms[:, 3]))
    # SURVEY 8(d) GRADIENT: the render bytes (SH blocks, point payload, cell topology of the walk) + radiance expansion,
    # derivative tables and gradient scatter, for the work the derivative walk counted; over ALL kernels of the step (the
    # SH sources are contracted once, in the forward pass, and handed to the derivative walk)
    abytes = algorithmic_bytes(st, gi, counts, gradient=True)
    achieved = abytes / (step_kernel_ms * 1e-3) / 1e9
    rbytes = algorithmic_bytes(st, gi, rcounts, gradient=False)
    tr = measured_traffic(args.workload)
    roof_extra = {}
    if tr is not None:
        roof_extra = dict(
            dram_frac=tr['dram_bytes'] / (step_kernel_ms * 1e-3) / 1e9 / peak,
            issue_active_pct=tr['issue_active_pct'], traffic_source=tr['source'],
            per_kernel={k: dict(share=v['share'], dram_bytes=v['dram_bytes'], issue_active_pct=v['issue_active_pct'],
                                warps_active_pct=v['warps_active_pct'], local_memory_instructions=v['local_memory_instructions'])
                        for k, v in list(tr['kernels'].items())[:5]},
            note='the state is L2-resident: frac is the SURVEY 8(d) algorithmic bytes over the HBM peak, dram_frac the DRAM '
                 'bytes ncu counted for the same step over the same peak; the kernels are latency / issue bound')

    cpu = None
    orc = None
    if rank == 0 and world == 1 and not args.no_cpu:
        gi_cpu = gi
        if gi.dpath is None and gi.exact_single_scatter:
            # the oracle reads the dense DPATH/DPTR lists, as the reference does: built for it when they fit
            gi_cpu = (gradsetup.make_gradient_inputs(sc, B, seed=0, numder=1)
                      if 8 * gi.longest_path_pts * st.npts < gradsetup.STREAM_BEAM_BYTES else None)
        if gi_cpu is not None:
            rate, sample, _ = cpu_reference_rate(sc, rays, gi_cpu, pix, args.cpu_seconds, ncores)
            cpu = dict(value=rate, unit='rays/s', cores=ncores, kind='port', sample=sample)
        else:
            cpu = dict(value=None, unit='rays/s', cores=ncores, kind='port',
                       sample='not run: the reference algorithm needs DPATH/DPTR[%d, %d] (%.0f GB) for this gradient'
                              % (gi.longest_path_pts, st.npts, 8e-9 * gi.longest_path_pts * st.npts))
        import oracle_lib as orc
    csrc = compute_source_leg(B, st, args.steps, args.warmup, orc if not args.lean else None)
    if 'kernel_ms' in csrc:
        csrc['frac'] = csrc['achieved_gbs'] / peak
        if cpu is not None and 'cpu_ms' in csrc:
            cpu['compute_source_ms'] = csrc.pop('cpu_ms')
            cpu['compute_source_cores'] = 1
    if world > 1 and 'kernel_ms' in csrc:
        tt = torch.tensor([csrc['kernel_ms']], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        csrc['kernel_ms'] = float(tt.item())
    if rank == 0:
        line = dict(
            metric='radiance+gradient rays/s', value=value, unit='rays/s', n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=1e3 * t_total / args.steps, higher_is_better=True, scaling='weak',
            vs_baseline=None, dtype='f32 optics / f64 geometry+accumulators', data='synthetic',
            config=dict(workload=workload_text(args.workload, st), rays=int(nrays), npts=int(st.npts), ncells=int(st.ncells),
                        nlm=int(st.nlm), l2='flushed between steps (256 MiB memset)', hbm_state_bytes=dev.hbm_bytes,
                        ray_setup='host-libm setup records resident in HBM (at3d_make_ray_packs): bit-exact walk in `value` and `e2e`'),
            e2e=dict(value=world * nrays * args.steps / t_e2e, unit='rays/s', h2d_bytes_per_step=int(h2d),
                     d2h_bytes_per_step=int(d2h)),
            gpu_launches=int(args.steps * (7 + (4 if gi.exact_single_scatter else 0))),   # forward, pixel, cost, derivative walk, apply, pair bounds + sums (+ beam count, beam, bounds, sums); cub scans and sorts not counted
            roofline=dict(kernel='gradient step: forward pass + derivative walk + apply + pair sort/sums + beam (all kernels of '
                                 'at3d_levisapprox_gradient)', bound='hbm', achieved=achieved, peak=peak, unit='GB/s',
                          frac=achieved / peak, traffic=(tr['dram_bytes'] if tr is not None else None),
                          peak_source=peak_src, algorithmic_bytes_per_launch=abytes, kernel_ms=step_kernel_ms, counts=counts,
                          **roof_extra),
            phases_ms=dict(forward=float(np.mean(kms[:, 0])), derivative=adj_ms, weights=float(np.mean(kms[:, 4])),
                           pair_sort_sums=float(np.mean(kms[:, 5])),
                           apply=adj_ms - float(np.mean(kms[:, 4])) - float(np.mean(kms[:, 5])), beam=float(np.mean(kms[:, 2])),
                           total=step_kernel_ms),
            render=dict(rays_per_s=nrays / (np.mean(rms) * 1e-3), kernel_ms=float(np.mean(rms)),
                        achieved_gbs=rbytes / (np.mean(rms) * 1e-3) / 1e9, frac=rbytes / (np.mean(rms) * 1e-3) / 1e9 / peak,
                        algorithmic_bytes=rbytes, counts=rcounts),
            compute_source=csrc, beam_derivative='streaming (no DPATH/DPTR lists)' if gi.dpath is None else 'dense DPATH/DPTR lists',
            clocks=cs.summary(), cpu_baseline=cpu, wall_s=wall)
        if not args.lean:
            line.update(
                compute_source_cfg4=compute_source_cfg4_leg(B, max(3, args.steps // 2), 2, peak),
                transforms=transform_leg(B, st, args.steps, args.warmup),
                render_orthographic=orthographic_leg(dev, sc, cfg['pixels_per_view'] ** 0.5, args.steps, args.warmup),
                solver_iteration=solver_leg(max(1, args.steps // 2), 1, orc if args.workload == 'cfg2' else None),
                path_integration_3d=sweep3d_leg(st, args.steps, 1, orc if args.workload == 'cfg2' else None),
                inversion_step=inversion_step_leg(sc, rays, gi, pix, DeviceState),
                render_ocean=dict(rays_per_s=nrays / (np.mean(oms) * 1e-3), kernel_ms=float(np.mean(oms)),
                                  surface_ms=float(np.mean(oms) - np.mean(rms)), surface_hits=ocounts['surface_hits'],
                                  brdf_evals=ocounts['surface_hits'] * 4 * (st.nang // 2 + 1)))
        if strong is not None:
            line['strong'] = strong
        print(json.dumps(line))
    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
