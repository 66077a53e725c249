// at3d_render.cu -- state preparation kernels and the RENDER kernel (sm_100a).
// Replaces RENDER / INTEGRATE_1RAY / COMPUTE_SOURCE_1CELL[_UNPOL] / FIND_BOUNDARY_RADIANCE
// (src/polarized/shdomsub4.f:93-286, shdomsub2.f:2311-3192 of the AT3D reference).
#include "at3d_mem.h"
#include <cstring>
#include "at3d_tray.cuh"
#include "at3d_host.h"

// ------------------------------------------------------------------------------------------
// State preparation
// ------------------------------------------------------------------------------------------
__global__ void build_cellrec_kernel(int ncells, const int *gridptr, const int *neighptr,
                                     const int *treeptr, const short *cellflags, int4 *cellrec)
{
    int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= ncells) return;
    const int *g = gridptr + 8 * (size_t)ic;
    const int *n = neighptr + 6 * (size_t)ic;
    cellrec[4 * (size_t)ic + 0] = make_int4(g[0], g[1], g[2], g[3]);
    cellrec[4 * (size_t)ic + 1] = make_int4(g[4], g[5], g[6], g[7]);
    cellrec[4 * (size_t)ic + 2] = make_int4(n[0], n[1], n[2], n[3]);
    cellrec[4 * (size_t)ic + 3] = make_int4(n[4], n[5], treeptr[2 * (size_t)ic + 1], (int)cellflags[ic]);
}

__global__ void build_ptrec_kernel(int npts, const float *gridpos, const float *total_ext, float4 *ptrec)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npts) return;
    ptrec[ip] = make_float4(gridpos[3 * (size_t)ip], gridpos[3 * (size_t)ip + 1],
                            gridpos[3 * (size_t)ip + 2], total_ext[ip]);
}

// one 16-byte record per point with everything a new corner needs besides its coordinates:
// SH block offset, NS, single-scatter count and the first single-scatter entry
// Bit 31 of .y marks a DARK point: every end cell it is a corner of has zero extinction at all eight corners.  The
// extinction-weighted source SRCEXT8 = SOURCE*EXT of such a point is exactly 0 and nothing can be absorbed or emitted in
// its cells, so the ray kernels neither contract its SH block nor look up its phase functions (clear air around
// cumulus fields: more than half of the corner visits of BASELINE configs[1]).
__global__ void mark_lit_kernel(int ncells, const int4 *cellrec, const float4 *ptrec, int *lit)
{
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= ncells) return;
    const int4 *p = cellrec + 4 * (size_t)ic;
    const int4 a = p[0], b = p[1], d = p[3];
    if (d.z != 0) return;                                     // not an end cell (TREEPTR(2) > 0)
    const int gp[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float m = 0.0f;
#pragma unroll
    for (int n = 0; n < 8; n++) m = fmaxf(m, fabsf(ptrec[gp[n] - 1].w));
    if (!(m == 0.0f)) {                                       // also lit if the extinction is not finite
#pragma unroll
        for (int n = 0; n < 8; n++) lit[gp[n] - 1] = 1;
    }
}

__global__ void build_ptsrc_kernel(int npts, int kmax, const int2 *srcrec, const int *sscount, const int2 *ssent,
                                   const int *lit, int4 *ptsrc)
{
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npts) return;
    const int2 r = srcrec[ip];
    const int cnt = sscount[ip];
    int2 e = make_int2(1, 0);
    if (cnt > 0) e = ssent[(size_t)ip * kmax];
    ptsrc[ip] = make_int4(r.x, r.y | (cnt << 16) | (lit[ip] ? 0 : (int)0x80000000), e.x, e.y);
}

// FIXED / VARIABLE_LAMBERTIAN_BOUNDARY (shdomsub1.f:2438-2529): bottom BCRAD
__global__ void lambertian_boundary_kernel(DevState S, const float *fluxes, float *bcrad)
{
    int ibc = blockIdx.x * blockDim.x + threadIdx.x;
    if (ibc >= S.nbotpts) return;
    const int i = S.bcptr[S.maxnbc + ibc];
    const float down = fluxes[2 * (size_t)(i - 1)];
    float v = 0.0f;
    if (S.sfctype0 == 'F') {
        const float alb = S.gndalbedo / acosf(-1.0f);
        float gndrad = 0.0f;
        if (S.srctype == 'T' || S.srctype == 'B') {
            gndrad = dev_planck(S.gndtemp, S.units, S.wavelen);
            gndrad = gndrad * (1.0f - S.gndalbedo);
        }
        if (S.srctype == 'S') v = alb * (S.dirflux[i - 1] + down);
        else if (S.srctype == 'T') v = gndrad + alb * down;
        else if (S.srctype == 'B') v = alb * (S.dirflux[i - 1] + down) + gndrad;
    } else {
        const float opi = 1.0f / acosf(-1.0f);
        const float alb = S.sfcgridparms[1 + S.nsfcpar * ibc];
        if (S.srctype == 'S') v = opi * alb * (S.dirflux[i - 1] + down);
        else {
            const float gndrad = S.sfcgridparms[0 + S.nsfcpar * ibc] * (1 - alb);
            if (S.srctype == 'T') v = gndrad + opi * alb * down;
            else if (S.srctype == 'B') v = opi * alb * (S.dirflux[i - 1] + down) + gndrad;
        }
    }
    bcrad[S.nstokes * (size_t)(S.ntoppts + ibc)] = v;
    for (int k = 1; k < S.nstokes; k++) bcrad[k + S.nstokes * (size_t)(S.ntoppts + ibc)] = 0.0f;
}

// Re-layout of one CSR spherical-harmonic array (SOURCE or RADIANCE, [nstokes,*] interleaved) into
// planar, 16-byte aligned per-point blocks.  For SOURCE with a solar source and delta-M the
// truncated single scattering that COMPUTE_SOURCE_1CELL subtracts per ray
// (shdomsub2.f:2979-3003: DA*LEGENT(.,l)*YLMSUN(1,j) for j<=NS) is ray independent, so it is
// folded in here once per state, and the per-point exact single-scatter list
// (shdomsub2.f:3008-3019: iphase, DA*w/(1-F)) is built.  One warp per grid point.
__global__ void prep_sh_kernel(DevState S, int tms, const int *shptr, const float *sh_in,
                               const int2 *rec, float *sh_out, int *sscount, int2 *ssent)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ip = blockIdx.x * (blockDim.x >> 5) + warp;        // 0-based
    float *corr1 = smem + warp * 2 * (S.ml + 2);
    float *corr5 = corr1 + (S.ml + 2);
    if (ip >= S.npts) return;
    const int is = shptr[ip], ns = shptr[ip + 1] - is;
    const int2 r = rec[ip];
    const int nsp = AT3D_SHPAD(ns);
    const int nst = S.nstokes;
    for (int l = lane; l <= S.ml + 1; l += 32) { corr1[l] = 0.0f; corr5[l] = 0.0f; }
    __syncwarp();
    if (tms) {
        const float ext = S.ptrec[ip].w;
        const float secmu0 = (float)(1.0 / fabs((double)S.solarmu));
        const int nlt = S.nstleg * (S.nleg + 1);
        int cnt = 0;
        for (int ipa = 0; ipa < S.npart; ipa++) {
            float w;
            if (ext == 0.0f) w = 1.0f; else w = S.extinct[ip + (size_t)S.npts * ipa] / ext;
            if (w == 0.0f) continue;
            const int *iph = S.iphase + (size_t)S.nq * (ip + (size_t)S.npts * ipa);
            const float *pw = S.phaseinterpwt + (size_t)S.nq * (ip + (size_t)S.npts * ipa);
            const bool single = (!S.interp_new) || (pw[0] >= S.phasemax);
            // F = LEGENT(1,ML+1) of the mixed table
            float f;
            if (single) f = S.legen[(size_t)nlt * (iph[0] - 1) + S.nstleg * (S.ml + 1)];
            else {
                f = 0.0f;
                for (int q = 0; q < S.nq; q++) {
                    if (pw[q] <= 1e-5f) continue;
                    f = f + S.legen[(size_t)nlt * (iph[q] - 1) + S.nstleg * (S.ml + 1)] * pw[q];
                }
            }
            const float da = S.albedo[ip + (size_t)S.npts * ipa] * S.dirflux[ip] * secmu0 * w;
            for (int l = lane; l <= S.ml; l += 32) {
                float l1, l5 = 0.0f;
                if (single) {
                    l1 = S.legen[(size_t)nlt * (iph[0] - 1) + S.nstleg * l];
                    if (S.nstleg > 1) l5 = S.legen[(size_t)nlt * (iph[0] - 1) + S.nstleg * l + 4];
                } else {
                    l1 = 0.0f;
                    for (int q = 0; q < S.nq; q++) {
                        if (pw[q] <= 1e-5f) continue;
                        l1 = l1 + S.legen[(size_t)nlt * (iph[q] - 1) + S.nstleg * l] * pw[q];
                        if (S.nstleg > 1)
                            l5 = l5 + S.legen[(size_t)nlt * (iph[q] - 1) + S.nstleg * l + 4] * pw[q];
                    }
                }
                if (S.interp_new) { l1 = l1 / (1 - f); l5 = l5 / (1 - f); }
                corr1[l] += da * l1;
                corr5[l] += da * l5;
            }
            if (lane == 0) {
                int2 *dst = ssent + (size_t)ip * S.kmax;
                if (pw[0] >= S.phasemax) {
                    dst[cnt++] = make_int2(iph[0], __float_as_int(da / (1 - f)));
                } else {
                    for (int q = 0; q < S.nq; q++) {
                        if (pw[q] <= 1e-5f) continue;
                        dst[cnt++] = make_int2(iph[q], __float_as_int(da * pw[q] / (1 - f)));
                    }
                }
            }
        }
        if (lane == 0) sscount[ip] = cnt;
        __syncwarp();
    }
    for (int j = lane; j < nsp; j += 32) {
        float v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
        if (j < ns) {
            const int l = S.lofj[j];
            const float ys = tms ? S.ylmsun[(size_t)S.nstleg * j] : 0.0f;
            v1 = sh_in[(size_t)nst * (is + j)] - corr1[l] * ys;
            if (nst > 1) {
                v2 = sh_in[(size_t)nst * (is + j) + 1] - corr5[l] * ys;
                v3 = sh_in[(size_t)nst * (is + j) + 2];
            }
        }
        sh_out[r.x + j] = v1;
        if (nst > 1) { sh_out[r.x + nsp + j] = v2; sh_out[r.x + 2 * nsp + j] = v3; }
    }
}

// ------------------------------------------------------------------------------------------
// RENDER / forward pass of the gradient
// ------------------------------------------------------------------------------------------
// Persistent grid; a warp draws four consecutive rays at a time from a global counter (rays differ
// a lot in length: clear sky vs. cloud), one ray per octet.
template <int NST, int MODES, typename OUTA>
__global__ void __launch_bounds__(AT3D_RAY_THREADS, NST == 1 ? AT3D_MINB_FWD1 : AT3D_MINB_FWD3)
forward_kernel(DevState S, int nrays, const float *camx, const float *camy, const float *camz,
               const double *cammu, const double *camphi, const RayPack *packs, OUTA *outA, double *outB,
               int correctinterpolate, int singlescatter, int nosurface, int maxsub,
               int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub, RayErr *err, int *ray_counter,
               int *npt_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Oct o = oct_id();
    const int lane = threadIdx.x & 31;
    float *Ysh = (float *)smem_raw + (size_t)(threadIdx.x >> 3) * S.ny_comp * S.nlmp;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ray_counter, 32 / AT3D_OCT);
        base = __shfl_sync(FULLMASK, base, 0);
        if (base >= nrays) break;
        const int iray = base + (lane >> 3);
        if (iray < nrays) {
            const double mu2 = __ldg(&cammu[iray]), phi2 = __ldg(&camphi[iray]);
            const RayPack pk = dev_get_pack(S, packs, iray, camx, camy, camz, mu2, phi2);
            double radA[NST], radB[NST];
#pragma unroll
            for (int k = 0; k < NST; k++) { radA[k] = 0.0; radB[k] = 0.0; }
            int ntrace = 0, nsubA = 0, nsubB = 0, nptB = 0;
            if (pk.status == 2) { if (o.ol == 0) set_err(err, 2, iray + S.ray_base); }
            else if (pk.status == 0) {
                RayDir rd;
                dev_ray_dir(S, pk, rd); rd.phi2 = (float)phi2;
                rd.hit = S.surfhits ? (SurfHit *)S.surfhits + iray : nullptr;
                __syncwarp(o.m);
                if (!S.viewsrc) group_ylmall(S, (float)mu2, (float)phi2, Ysh, o.ol, AT3D_OCT, o.m);
                const float sky = (-mu2 > 0.0) ? dev_sky_radiance(S, (float)mu2, (float)phi2) : 0.0f;
                const int e = march_forward<NST, MODES>(S, Ysh, rd, mu2, pk.x0, pk.y0, pk.z0, sky,
                                                        correctinterpolate != 0, singlescatter != 0, nosurface != 0,
                                                        maxsub, o, radA, radB,
                                                        trace_cells ? trace_cells + (size_t)trace_cap * iray : nullptr,
                                                        trace_cap, ntrace, nsubA, nsubB, nptB);
                if (e && o.ol == 0) set_err(err, e, iray + S.ray_base);
            }
            if (o.ol == 0) {
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    if (MODES & 1) outA[k + NST * (size_t)iray] = (OUTA)radA[k];
                    if (MODES & 2) outB[k + NST * (size_t)iray] = radB[k];
                }
                if (trace_n) { trace_n[iray] = ntrace; trace_nsub[iray] = (MODES & 1) ? nsubA : nsubB; }
                if (npt_out) npt_out[iray] = nptB;
            }
        }
        __syncwarp();
    }
}

// Thread-per-ray variant for NSTOKES=1 (at3d_tray.cuh): a warp draws 32 consecutive rays at a time.
template <int MODES, typename OUTA>
__global__ void __launch_bounds__(256)
forward_kernel_t(DevState S, int nrays, const float *camx, const float *camy, const float *camz,
                 const double *cammu, const double *camphi, const RayPack *packs, OUTA *outA, double *outB,
                 int correctinterpolate, int singlescatter, int nosurface, int maxsub,
                 int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub, RayErr *err, int *ray_counter,
                 int *npt_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int bt = blockDim.x, lane = threadIdx.x & 31;
    float4 *Y4 = (float4 *)smem_raw + threadIdx.x;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ray_counter, 32);
        base = __shfl_sync(FULLMASK, base, 0);
        if (base >= nrays) break;
        const int iray = base + lane;
        int ntrace = 0, nsubA = 0, nsubB = 0, npt = 0, nsh = 0, marched = 0, nptB = 0;
        if (iray < nrays) {
            const double mu2 = __ldg(&cammu[iray]), phi2 = __ldg(&camphi[iray]);
            const RayPack pk = dev_get_pack(S, packs, iray, camx, camy, camz, mu2, phi2);
            double radA = 0.0, radB = 0.0;
            SrcWriter sw = -1;
            if ((MODES & 2) && S.srcpool) S.srcstart[iray] = -1;
            if (pk.status == 2) set_err(err, 2, iray + S.ray_base);
            else if (pk.status == 0) {
                RayDir rd;
                dev_ray_dir(S, pk, rd); rd.phi2 = (float)phi2;
                rd.hit = S.surfhits ? (SurfHit *)S.surfhits + iray : nullptr;
                if (!S.viewsrc) thread_ylmall_unpol(S, (float)mu2, (float)phi2, (float *)Y4, bt);
                const float sky = (-mu2 > 0.0) ? dev_sky_radiance(S, (float)mu2, (float)phi2) : 0.0f;
                const int e = thread_march_forward<MODES>(S, Y4, bt, rd, mu2, pk.x0, pk.y0, pk.z0, sky,
                                                          correctinterpolate != 0, singlescatter != 0, nosurface != 0,
                                                          maxsub, radA, radB,
                                                          trace_cells ? trace_cells + (size_t)trace_cap * iray : nullptr,
                                                          trace_cap, ntrace, nsubA, nsubB, npt, nsh, nptB,
                                                          sw, S.srcstart + iray, (MODES & 2) && S.srcpool != nullptr);
                if (e) set_err(err, e, iray + S.ray_base);
                else marched = 1;
            }
            if (MODES & 1) outA[iray] = (OUTA)radA;
            if (MODES & 2) outB[iray] = radB;
            if (trace_n) { trace_n[iray] = ntrace; trace_nsub[iray] = (MODES & 1) ? nsubA : nsubB; }
            if (npt_out) npt_out[iray] = nptB;
        }
        __syncwarp();
        if (S.counts) {
            const int c0 = __reduce_add_sync(FULLMASK, marched ? ntrace : 0), c1 = __reduce_add_sync(FULLMASK, marched ? npt : 0);
            const int c2 = __reduce_add_sync(FULLMASK, marched ? nsh : 0);
            const int c4 = __reduce_add_sync(FULLMASK, marched ? ((MODES & 1) ? nsubA : nsubB) : 0);
            const int c5 = __reduce_add_sync(FULLMASK, marched);
            if (lane == 0) {
                atomicAdd(&S.counts[0], (unsigned long long)c0); atomicAdd(&S.counts[1], (unsigned long long)c1);
                atomicAdd(&S.counts[2], (unsigned long long)c2); atomicAdd(&S.counts[4], (unsigned long long)c4);
                atomicAdd(&S.counts[5], (unsigned long long)c5);
            }
        }
    }
}

// SRCEXT of every grid point for one direction (orthographic views): thread = grid point, the same arithmetic as a
// ray's corner evaluation (thread_point_source), YLMDIR once per block in shared memory.  The march of the view's rays
// then gathers 4 bytes per corner instead of contracting an SH block per (ray, corner).
__global__ void __launch_bounds__(256)
view_source_kernel(DevState S, float mu2, float phi2, RayDir rd, int singlescatter, float *viewsrc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *Y4 = (float4 *)smem_raw;
    if (threadIdx.x == 0) thread_ylmall_unpol(S, mu2, phi2, (float *)Y4, 1);
    __syncthreads();
    for (int ip = 1 + blockIdx.x * blockDim.x + threadIdx.x; ip <= S.npts; ip += gridDim.x * blockDim.x) {
        int ns;
        viewsrc[ip - 1] = thread_point_source(S, Y4, 1, rd, singlescatter != 0, ip, __ldg(&S.ptrec[ip - 1]).w, ns);
    }
}

// The same for the octet kernels (NSTOKES=3, or NSTOKES=1 when NLM is too large for the thread-per-ray layout): an
// octet evaluates one grid point exactly as refresh_corners does (load_corner, sh_dot_partial, oct_sum), YLMDIR once
// per block.
template <int NST>
__global__ void __launch_bounds__(AT3D_RAY_THREADS)
view_source_kernel_oct(DevState S, float mu2, float phi2, RayDir rd, int singlescatter, float *viewsrc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Ysh = (float *)smem_raw;
    const Oct o = oct_id();
    if (threadIdx.x < 32) group_ylmall(S, mu2, phi2, Ysh, threadIdx.x, 32, FULLMASK);
    __syncthreads();
    const int octs = (gridDim.x * blockDim.x) >> 3;
    const int first = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    // whole warps iterate together (4 points per warp and iteration) so that the octet shuffles stay converged
    for (int base = first - ((threadIdx.x & 31) >> 3); base < S.npts; base += octs) {
        const int ip0 = base + ((threadIdx.x & 31) >> 3);          // 0-based
        const bool valid = ip0 < S.npts;
        const int ip = valid ? ip0 + 1 : 1;
        float x, y, z, ext, b[NST], a[NST];
        int soff, sns;
        load_corner<NST>(S, ip, rd, x, y, z, ext, soff, sns, b);
#pragma unroll
        for (int k = 0; k < NST; k++) a[k] = 0.0f;
        if (!singlescatter) {
            sh_dot_partial<NST>(S.shsrc + soff, AT3D_SHPAD(sns), Ysh, S.nlmp, o, a);
#pragma unroll
            for (int k = 0; k < NST; k++) a[k] = oct_sum(o.m, a[k]);
        }
        if (valid && o.ol == 0) {
#pragma unroll
            for (int k = 0; k < NST; k++) viewsrc[k + NST * (size_t)ip0] = (a[k] + b[k]) * ext;
        }
    }
}

cudaError_t launch_view_source(const DevState &S, const RayPack &pk, double mu2, double phi2, int singlescatter,
                               float *viewsrc, cudaStream_t stream)
{
    RayDir rd;
    memset(&rd, 0, sizeof(rd));
    rd.f = pk.f; rd.j = pk.j; rd.cos22 = pk.cos22; rd.sin22 = pk.sin22;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (tray_block_threads(S) > 0) {
        const int want = (S.npts + 255) / 256, cap = nsm * 8;
        view_source_kernel<<<want < cap ? want : cap, 256, (size_t)S.nlmp * sizeof(float), stream>>>(
            S, (float)mu2, (float)phi2, rd, singlescatter, viewsrc);
    } else {
        const int per = AT3D_RAY_THREADS / 8;
        const int want = (S.npts + per - 1) / per, cap = nsm * 8;
        const size_t smem = (size_t)S.ny_comp * S.nlmp * sizeof(float);
        if (S.nstokes == 1)
            view_source_kernel_oct<1><<<want < cap ? want : cap, AT3D_RAY_THREADS, smem, stream>>>(
                S, (float)mu2, (float)phi2, rd, singlescatter, viewsrc);
        else
            view_source_kernel_oct<3><<<want < cap ? want : cap, AT3D_RAY_THREADS, smem, stream>>>(
                S, (float)mu2, (float)phi2, rd, singlescatter, viewsrc);
    }
    return cudaGetLastError();
}

// threads per block of the thread-per-ray kernels: as many YLMDIR columns (4*NLMP bytes) as fit
#ifndef AT3D_TRAY_SMEM_KB
#define AT3D_TRAY_SMEM_KB 220
#endif
int tray_block_threads(const DevState &S)
{
#ifdef AT3D_FORCE_OCTET
    return 0;
#endif
    if (S.nstokes != 1) return 0;
    const size_t per = (size_t)S.nlmp * sizeof(float);
    int bt = (int)((AT3D_TRAY_SMEM_KB * 1024) / per) & ~31;
    if (bt > 256) bt = 256;
    return bt >= 64 ? bt : 0;
}

size_t render_smem_bytes(const DevState &S)
{
    return (size_t)AT3D_RAYS_PER_BLOCK * S.ny_comp * S.nlmp * sizeof(float);
}

// persistent launch: one wave, grid = #SMs x resident blocks
template <typename K>
static int persistent_blocks(K kernel, int nrays, size_t smem)
{
    static thread_local KernelFit fit;                  // one per kernel instantiation
    kernel_fit(fit, kernel, AT3D_RAY_THREADS, smem);
    const long want = ((long)nrays + AT3D_RAYS_PER_BLOCK - 1) / AT3D_RAYS_PER_BLOCK, cap = (long)fit.nsm * fit.per_sm;
    return (int)(want < cap ? want : cap);
}

// Launch helper (host).  modes 1: RENDER (out_f32 = STOKES, or out_f64); modes 3: forward pass of the
// gradient (out_f64 = VISRAD in INTEGRATE_1RAY arithmetic, out_tot = totals in the adjoint arithmetic).
cudaError_t launch_forward(const DevState &S, int nrays, const float *camx, const float *camy,
                           const float *camz, const double *cammu, const double *camphi,
                           const RayPack *packs, float *out_f32, double *out_f64, double *out_tot, int modes,
                           int correctinterpolate, int singlescatter, int nosurface, int maxsub,
                           int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub,
                           RayErr *err, int *ray_counter, int *npt_out, cudaStream_t stream)
{
    if (nrays <= 0) return cudaSuccess;
    cudaError_t ce = cudaMemsetAsync(ray_counter, 0, sizeof(int), stream);
    if (ce != cudaSuccess) return ce;
    const int bt = tray_block_threads(S);
    if (bt > 0) {
        const size_t smem_t = (size_t)bt * S.nlmp * sizeof(float);
#define LAUNCHT(MODES, OUTA, outa)                                                                 \
        {                                                                                          \
            static thread_local KernelFit fit;                                                     \
            kernel_fit(fit, forward_kernel_t<MODES, OUTA>, bt, smem_t);                            \
            const long want = ((long)nrays + bt - 1) / bt, cap = (long)fit.nsm * fit.per_sm;       \
            forward_kernel_t<MODES, OUTA><<<(int)(want < cap ? want : cap), bt, smem_t, stream>>>(S, nrays, camx, camy, \
                camz, cammu, camphi, packs, outa, out_tot, correctinterpolate, singlescatter, nosurface, maxsub,  \
                trace_cells, trace_cap, trace_n, trace_nsub, err, ray_counter, npt_out);           \
        }
        if (modes == 3) LAUNCHT(3, double, out_f64)
        else if (out_f64) LAUNCHT(1, double, out_f64)
        else LAUNCHT(1, float, out_f32)
#undef LAUNCHT
        return cudaGetLastError();
    }
    const size_t smem = render_smem_bytes(S);
#define LAUNCH(NST, MODES, OUTA, outa)                                                            \
    {                                                                                             \
        const int nb = persistent_blocks(forward_kernel<NST, MODES, OUTA>, nrays, smem);          \
        forward_kernel<NST, MODES, OUTA><<<nb, AT3D_RAY_THREADS, smem, stream>>>(S, nrays, camx, camy, camz, \
            cammu, camphi, packs, outa, out_tot, correctinterpolate, singlescatter, nosurface, maxsub,        \
            trace_cells, trace_cap, trace_n, trace_nsub, err, ray_counter, npt_out);                          \
    }
    if (S.nstokes == 1) {
        if (modes == 3) LAUNCH(1, 3, double, out_f64)
        else if (out_f64) LAUNCH(1, 1, double, out_f64)
        else LAUNCH(1, 1, float, out_f32)
    } else {
        if (modes == 3) LAUNCH(3, 3, double, out_f64)
        else if (out_f64) LAUNCH(3, 1, double, out_f64)
        else LAUNCH(3, 1, float, out_f32)
    }
#undef LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_build_cellrec(int ncells, const int *gridptr, const int *neighptr, const int *treeptr,
                                 const short *cellflags, int4 *cellrec, cudaStream_t s)
{
    build_cellrec_kernel<<<(ncells + 255) / 256, 256, 0, s>>>(ncells, gridptr, neighptr, treeptr, cellflags, cellrec);
    return cudaGetLastError();
}
cudaError_t launch_build_ptrec(int npts, const float *gridpos, const float *total_ext, float4 *ptrec, cudaStream_t s)
{
    build_ptrec_kernel<<<(npts + 255) / 256, 256, 0, s>>>(npts, gridpos, total_ext, ptrec);
    return cudaGetLastError();
}
cudaError_t launch_build_ptsrc(int npts, int kmax, const int2 *srcrec, const int *sscount, const int2 *ssent,
                               int ncells, const int4 *cellrec, const float4 *ptrec, int4 *ptsrc, cudaStream_t s)
{
    int *lit = nullptr;
    cudaError_t e = at3d_malloc(&lit, sizeof(int) * (size_t)npts);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(lit, 0, sizeof(int) * (size_t)npts, s);
    mark_lit_kernel<<<(ncells + 255) / 256, 256, 0, s>>>(ncells, cellrec, ptrec, lit);
    build_ptsrc_kernel<<<(npts + 255) / 256, 256, 0, s>>>(npts, kmax, srcrec, sscount, ssent, lit, ptsrc);
    e = cudaGetLastError();
    cudaStreamSynchronize(s);
    at3d_free(lit);
    return e;
}
cudaError_t launch_lambertian_boundary(const DevState &S, const float *fluxes, float *bcrad, cudaStream_t s)
{
    if (S.nbotpts <= 0 || S.sfctype1 != 'L') return cudaSuccess;
    lambertian_boundary_kernel<<<(S.nbotpts + 255) / 256, 256, 0, s>>>(S, fluxes, bcrad);
    return cudaGetLastError();
}
cudaError_t launch_prep_sh(const DevState &S, int tms, const int *shptr, const float *sh_in,
                           const int2 *rec, float *sh_out, int *sscount, int2 *ssent, cudaStream_t s)
{
    const int wpb = 8;
    const size_t smem = (size_t)wpb * 2 * (S.ml + 2) * sizeof(float);
    prep_sh_kernel<<<(S.npts + wpb - 1) / wpb, wpb * 32, smem, s>>>(S, tms, shptr, sh_in, rec, sh_out, sscount, ssent);
    return cudaGetLastError();
}
