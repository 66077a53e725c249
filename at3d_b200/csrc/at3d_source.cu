// at3d_source.cu -- COMPUTE_SOURCE on sm_100a.
// Replaces COMPUTE_SOURCE and CALC_SOURCE_PNT[_UNPOL] (src/polarized/shdomsub1.f:823-1611 of the AT3D
// reference): the per-grid-point source-function update in spherical-harmonic space, the series
// acceleration dot products, the adaptive SH truncation and the SHPTR rebuild.
//
// The reference makes three passes over the grid, each recomputing the temporary source of a point.
// Here: one streaming kernel computes the temporary source once for the four norms AND the new
// truncation length, a device-wide exclusive scan rebuilds SHPTR, and a second streaming kernel
// writes DELSOURCE and the re-packed SOURCE.  One warp per grid point, lanes over the SH index j
// (coalesced on the CSR arrays); the mixed Legendre table of the point lives in shared memory.
// HBM-bound: 4*NSTOKES*(2*NR + NS_old [+2*NS accel] + NS_new) bytes per point.
#include "at3d_mem.h"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <cub/cub.cuh>
#include <mutex>
#include "at3d_host.h"

static void set_msg(char *errmsg, const char *fmt, ...)
{
    if (!errmsg) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, AT3D_ERRMSG_LEN, fmt, ap);
    va_end(ap);
}


#define CS_WARPS 8

// Mixed Legendre table of the point (NEWMETHOD=.TRUE.) into shared memory; returns (albedo, planck) to use in
// CALC_SOURCE_PNT.  shdomsub1.f:1089-1141.  Lane t owns table entries t, t+32, ... in registers through the whole
// mixing (the delta-M fraction F of a species travels by shuffle), so a point costs one pass and one warp sync.
// CS_SLOTS = table entries per lane (1, 2, 4 or 8: NSTLEG*(NLEG+1) <= 256), a template parameter of the kernels.
template <int CS_SLOTS>
__device__ __forceinline__ void cs_mix_newmethod(const CsArgs &a, int i, float *legent, float *legent1,
                                                 float &albedo_out, float &planck_out)
{
    const int lane = threadIdx.x & 31;
    const int nlt = a.nstleg * (a.nleg + 1), ndm = a.nstleg * (a.ml + 1), nslots = (nlt + 31) >> 5;
    const float ext = a.total_ext[i];
    double alb = 0.0;
    float total_planck = 0.0f;
    float acc[CS_SLOTS];
#pragma unroll
    for (int u = 0; u < CS_SLOTS; u++) acc[u] = 0.0f;
    for (int ipa = 0; ipa < a.npart; ipa++) {
        const size_t ld = a.ldp ? a.ldp : a.npts;
        const int *iph = a.iphase + (size_t)a.nq * (i + ld * ipa);
        const float *pw = a.phaseinterpwt + (size_t)a.nq * (i + ld * ipa);
        const float e = a.extinct[i + ld * ipa], al = a.albedo[i + ld * ipa];
        const double scat = (double)(e * al);
        alb = alb + scat;
        if (a.planck) total_planck = total_planck + e * a.planck[i + ld * ipa];
        float l1[CS_SLOTS];
        if (!a.interp_new) {
            const float *lg = a.legen + (size_t)nlt * (iph[0] - 1);
#pragma unroll
            for (int u = 0; u < CS_SLOTS; u++) {
                if (u >= nslots) break; const int t = lane + 32 * u; l1[u] = t < nlt ? __ldg(&lg[t]) : 0.0f; }
        } else {
            const bool single = pw[0] >= a.phasemax;
#pragma unroll
            for (int u = 0; u < CS_SLOTS; u++) {
                if (u >= nslots) break;
                const int t = lane + 32 * u;
                float v = 0.0f;
                if (t < nlt) {
                    if (single) v = __ldg(&a.legen[(size_t)nlt * (iph[0] - 1) + t]);
                    else {
                        for (int q = 0; q < a.nq; q++) {
                            if (pw[q] <= 1e-5f) continue;
                            v = v + __ldg(&a.legen[(size_t)nlt * (iph[q] - 1) + t]) * pw[q];
                        }
                    }
                }
                l1[u] = v;
            }
            if (a.deltam) {
                // F = LEGENT1(1, ML+1): entry ndm of the table
                float f = 0.0f;
#pragma unroll
                for (int u = 0; u < CS_SLOTS; u++)
                    if (ndm / 32 == u) f = __shfl_sync(FULLMASK, l1[u], ndm % 32);
#pragma unroll
                for (int u = 0; u < CS_SLOTS; u++) if (u < nslots && lane + 32 * u < ndm) l1[u] = l1[u] / (1 - f);
            }
        }
#pragma unroll
        for (int u = 0; u < CS_SLOTS; u++) if (u < nslots) acc[u] = (float)(acc[u] + scat * l1[u]);
    }
#pragma unroll
    for (int u = 0; u < CS_SLOTS; u++) {
                if (u >= nslots) break;
        if (alb > 1e-10f) acc[u] = (float)(acc[u] / alb);
        else acc[u] = acc[u] / a.npart;
    }
    if (ext > 1e-10f) { alb = alb / ext; total_planck = total_planck / ext; }
    else { alb = 0.0; total_planck = 0.0f; }
    if (lane == 0) acc[0] = 1.0f;
    __syncwarp();                 // the previous point's reads of legent are complete
#pragma unroll
    for (int u = 0; u < CS_SLOTS; u++) if (lane + 32 * u < nlt) legent[lane + 32 * u] = acc[u];
    __syncwarp();
    albedo_out = (float)alb;
    planck_out = total_planck;
}

// The mixing depends on the optical properties only: it runs once per call (once per solve in at3d_solver_solve), one
// warp per point, and leaves the table and (albedo, planck) of every point in HBM; the two source kernels stream the
// rows (68 B per point for NSTOKES=1, NLEG=16) one point ahead instead of chasing IPHASE -> LEGEN per point.
template <int SLOTS>
__global__ void __launch_bounds__(CS_WARPS * 32) cs_mix_kernel(CsArgs a)
{
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlt = a.nstleg * (a.nleg + 1);
    float *legent = smem + (size_t)warp * 2 * nlt, *legent1 = legent + nlt;
    for (int i = blockIdx.x * CS_WARPS + warp; i < a.npts; i += gridDim.x * CS_WARPS) {
        float albedo, planck;
        cs_mix_newmethod<SLOTS>(a, i, legent, legent1, albedo, planck);
        for (int t = lane; t < nlt; t += 32) a.mix_legent[(size_t)nlt * i + t] = legent[t];
        if (lane == 0) a.mix_ap[i] = make_float2(albedo, planck);
    }
}

// a point's row of the mixed table, lane t holding entries t, t+32, ...
template <int SLOTS>
struct CsMixRow {
    float v[SLOTS];
    float2 ap;
    __device__ __forceinline__ void load(const CsArgs &a, int ip, int lane, int nlt)
    {
#pragma unroll
        for (int u = 0; u < SLOTS; u++) { const int t = lane + 32 * u; v[u] = t < nlt ? __ldg(&a.mix_legent[(size_t)nlt * ip + t]) : 0.0f; }
        ap = __ldg(&a.mix_ap[ip]);
    }
    __device__ __forceinline__ void to_shared(float *legent, int lane, int nlt) const
    {
        __syncwarp();                 // the previous point's reads of legent are complete
#pragma unroll
        for (int u = 0; u < SLOTS; u++) if (lane + 32 * u < nlt) legent[lane + 32 * u] = v[u];
        __syncwarp();
    }
};

// CALC_SOURCE_PNT[_UNPOL] for one SH index j (0-based) (shdomsub1.f:858-898, 940-958); r = RADIANCE(:,j) (0 beyond NR)
template <int NST>
__device__ __forceinline__ void cs_calc_j(const CsArgs &a, const float *legen, int j, int l, float ysun, bool inr,
                                          const float (&r)[NST], float flux0, float planck, float albedo, float (&s)[NST])
{
    const bool solar = a.srctype == 'S' || a.srctype == 'B';
    const bool thermal = a.srctype == 'T' || a.srctype == 'B';
    if (NST == 1) {
        float v = 0.0f;
        if (solar) v = flux0 * albedo * legen[l] * ysun;
        if (thermal && j == 0) v = v + 3.544907703f * planck;
        if (inr) v = v + albedo * legen[l] * r[0];
        s[0] = v;
    } else {
        const int ns = a.nstleg;
        float v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
        if (inr) {
            const float r1 = r[0], r2 = r[1], r3 = r[NST - 1];
            v1 = v1 + legen[ns * l] * r1;
            v1 = v1 + legen[4 + ns * l] * r2;
            if (j >= 4) {
                v2 = v2 + legen[4 + ns * l] * r1 + legen[1 + ns * l] * r2;
                v3 = v3 + legen[2 + ns * l] * r3;
            }
        }
        v1 = albedo * v1; v2 = albedo * v2; v3 = albedo * v3;
        if (solar) {
            v1 = v1 + flux0 * albedo * legen[ns * l] * ysun;
            if (j >= 4) v2 = v2 + flux0 * albedo * legen[4 + ns * l] * ysun;
        }
        if (thermal && j == 0) v1 = v1 + 3.544907703f * planck;
        s[0] = v1; s[1] = v2; s[NST - 1] = v3;
    }
}

// ---- new grid points of SPLIT_GRID (INTERPOLATE_POINT, shdomsub1.f:5109-5211) ----
// warp = new point: RADIANCE(:, IR+J) = 0.5*(RAD1 + RAD2) over NR = max(NR1, NR2) terms, then the source function of the
// point from that radiance with the point's mixed Legendre row (cs_mix_points), first NS = max(NS1, NS2) terms.
template <int NST>
__global__ void interp_points_kernel(CsArgs a, const NewPointRec *rec, int count, const int *rshptr, float *radiance, float *source)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= count) return;
    const NewPointRec r = rec[k];
    const int ir1 = rshptr[r.ip1], nr1 = rshptr[r.ip1 + 1] - ir1, ir2 = rshptr[r.ip2], nr2 = rshptr[r.ip2 + 1] - ir2;
    const int nlt = a.nstleg * (a.nleg + 1);
    const float *legen = a.mix_legent + (size_t)nlt * r.ip;
    const float2 ap = a.mix_ap[r.ip];
    const float flux0 = a.dirflux[r.ip] * a.secmu0;
    const int nmax = r.nr > r.ns ? r.nr : r.ns;
    for (int j = lane; j < nmax; j += 32) {
        float rad[NST], s[NST];
#pragma unroll
        for (int q = 0; q < NST; q++) {
            const float r1 = j < nr1 ? radiance[q + (size_t)NST * (ir1 + j)] : 0.0f;
            const float r2 = j < nr2 ? radiance[q + (size_t)NST * (ir2 + j)] : 0.0f;
            rad[q] = 0.5f * (r1 + r2);
        }
        if (j < r.nr) {
#pragma unroll
            for (int q = 0; q < NST; q++) radiance[q + (size_t)NST * (r.ir + j)] = rad[q];
        }
        if (j < r.ns) {
            cs_calc_j<NST>(a, legen, j, a.lofj[j], a.ylmsun[(size_t)a.nstleg * j], j < r.nr, rad, flux0, ap.y, ap.x, s);
#pragma unroll
            for (int q = 0; q < NST; q++) source[q + (size_t)NST * (r.is + j)] = s[q];
        }
    }
}

cudaError_t launch_interp_points(const CsArgs &a, const NewPointRec *rec_d, int count, const int *rshptr_d, float *radiance,
                                 float *source, cudaStream_t s)
{
    if (count < 1) return cudaSuccess;
    const int nb = (int)(((size_t)count * 32 + 127) / 128);
    if (a.nstokes == 1) interp_points_kernel<1><<<nb, 128, 0, s>>>(a, rec_d, count, rshptr_d, radiance, source);
    else interp_points_kernel<3><<<nb, 128, 0, s>>>(a, rec_d, count, rshptr_d, radiance, source);
    return cudaGetLastError();
}

// A point is one warp; its SH index j runs over the lanes in batches of CS_BATCH x 32.  The loads of a batch
// (RADIANCE, old SOURCE, old DELSOURCE: the HBM streams of the routine) are all issued before anything is computed, and
// those of the first batch even before the Legendre table of the point is mixed, so that a warp keeps
// 3 x CS_BATCH x 128 B x NSTOKES in flight instead of one line per array.
template <int NST> struct CsBatch { static const int N = NST == 1 ? 4 : 2; };

// The SH indices a lane meets are the same for every point: j = lane + 32*slot.  Their degree l(j) and YLMSUN(1,j)
// are kept in registers for the first CS_JSLOTS slots (NLM <= 256), instead of two table loads per element.
#define CS_JSLOTS 8
struct CsLaneTab {
    int l[CS_JSLOTS];
    float ys[CS_JSLOTS];
    __device__ __forceinline__ void init(const CsArgs &a, int lane)
    {
#pragma unroll
        for (int q = 0; q < CS_JSLOTS; q++) {
            const int j = lane + 32 * q;
            l[q] = j < a.nlm ? a.lofj[j] : 0;
            ys[q] = j < a.nlm ? a.ylmsun[(size_t)a.nstleg * j] : 0.0f;
        }
    }
};

// Kernel A: norms (passes 1 of the reference) and the new truncation length (first half of pass 3)
#ifndef AT3D_CS_MINB
#define AT3D_CS_MINB 3
#endif
template <int NST, int SLOTS>
__global__ void __launch_bounds__(CS_WARPS * 32, (NST == 1 ? AT3D_CS_MINB : 1)) cs_norms_kernel(CsArgs a)
{
    extern __shared__ float smem[];
    constexpr int NB = CsBatch<NST>::N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlt = a.nstleg * (a.nleg + 1);
    float *legent = smem + (size_t)warp * 2 * nlt, *legent1 = legent + nlt;
    __shared__ double red[CS_WARPS][4];
    double sdot = 0.0, sold = 0.0, snew = 0.0, snorm = 0.0;
    const bool donorm = !a.first, doacc = a.accelflag && !a.first;
    CsLaneTab tab;
    tab.init(a, lane);
    // software pipeline over the warp's points: the block pointers of the next point are requested at the top of an
    // iteration and its first batch of SH values at the bottom, so their HBM latency overlaps this point's arithmetic
    const int stride = gridDim.x * CS_WARPS;
    int i = blockIdx.x * CS_WARPS + warp;
    int ir = 0, nr = 0, is = 0, ns = 0, iso = 0;
    float dfl = 0.0f;
    float r[NB][NST], so[NB][NST], ds[NB][NST];
    auto load_ptrs = [&](int ip, int &ir_, int &nr_, int &is_, int &ns_, int &iso_, float &dfl_) {
        dfl_ = __ldg(&a.dirflux[ip]);
        ir_ = a.rshptr[ip]; nr_ = a.rshptr[ip + 1] - ir_;
        is_ = a.shptr_old[ip]; ns_ = a.shptr_old[ip + 1] - is_;
        iso_ = 0;
        if (doacc) {
            iso_ = a.oshptr_old[ip];
            const int nso = a.oshptr_old[ip + 1] - iso_;
            if (nso < ns_) ns_ = nso;
        }
    };
    auto load_batch = [&](int j0, int ir_, int nr_, int is_, int ns_, int iso_) {
        const float *rad = a.radiance + (size_t)NST * ir_;
        const float *sop = a.source_old + (size_t)NST * is_, *dsp = a.delsource_old + (size_t)NST * iso_;
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int j = j0 + 32 * u + lane;
#pragma unroll
            for (int k = 0; k < NST; k++) {
                r[u][k] = (j < nr_) ? __ldg(&rad[(size_t)NST * j + k]) : 0.0f;
                so[u][k] = (donorm && j < ns_) ? __ldg(&sop[(size_t)NST * j + k]) : 0.0f;
                ds[u][k] = (doacc && j < ns_) ? __ldg(&dsp[(size_t)NST * j + k]) : 0.0f;
            }
        }
    };
    CsMixRow<SLOTS> mix;
    (void)legent1;
    if (i < a.npts) { load_ptrs(i, ir, nr, is, ns, iso, dfl); mix.load(a, i, lane, nlt); load_batch(0, ir, nr, is, ns, iso); }
    for (; i < a.npts; i += stride) {
        const int inext = i + stride;
        int ir2 = 0, nr2 = 0, is2 = 0, ns2 = 0, iso2 = 0;
        float dfl2 = 0.0f;
        if (inext < a.npts) load_ptrs(inext, ir2, nr2, is2, ns2, iso2, dfl2);
        if (nr > a.nlm) {
            if (lane == 0) atomicCAS(a.bad, 0, i + 1);
            ir = ir2; nr = nr2; is = is2; ns = ns2; iso = iso2; dfl = dfl2;
            if (inext < a.npts) { mix.load(a, inext, lane, nlt); load_batch(0, ir, nr, is, ns, iso); }
            continue;
        }
        const float flux0 = dfl * a.secmu0;
        mix.to_shared(legent, lane, nlt);
        const float albedo = mix.ap.x, planck = mix.ap.y;
        if (inext < a.npts) mix.load(a, inext, lane, nlt);
        int jlast = -1;                       // last j with |SOURCET| > SRCMIN
        // the four norms of this point are summed in REAL per lane (at most NLM/32 terms) and added to the DOUBLE
        // accumulators once per point
        float pdot = 0.0f, pold = 0.0f, pnew = 0.0f, pnorm = 0.0f;
#pragma unroll
        for (int b = 0; b < CS_JSLOTS / NB; b++) {
            const int j0 = 32 * NB * b;
            if (j0 >= a.nlm) break;
            if (b > 0) load_batch(j0, ir, nr, is, ns, iso);
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int j = j0 + 32 * u + lane;
                if (j >= a.nlm) continue;
                float s[NST];
                cs_calc_j<NST>(a, legent, j, tab.l[b * NB + u], tab.ys[b * NB + u], j < nr, r[u], flux0, planck, albedo, s);
#pragma unroll
                for (int k = 0; k < NST; k++) if (fabsf(s[k]) > a.srcmin) jlast = j;
                if (donorm && j < ns) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        const float d = s[k] - so[u][k];
                        if (a.accelflag) {
                            pdot = pdot + d * ds[u][k];
                            pold = pold + ds[u][k] * ds[u][k];
                        }
                        pnew = pnew + d * d;
                        pnorm = pnorm + so[u][k] * so[u][k];
                    }
                }
            }
        }
        for (int j0 = 32 * CS_JSLOTS; j0 < a.nlm; j0 += 32 * NB) {        // NLM > 256: table loads
            load_batch(j0, ir, nr, is, ns, iso);
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int j = j0 + 32 * u + lane;
                if (j >= a.nlm) continue;
                float s[NST];
                cs_calc_j<NST>(a, legent, j, a.lofj[j], a.ylmsun[(size_t)a.nstleg * j], j < nr, r[u], flux0, planck, albedo, s);
#pragma unroll
                for (int k = 0; k < NST; k++) if (fabsf(s[k]) > a.srcmin) jlast = j;
                if (donorm && j < ns) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        const float d = s[k] - so[u][k];
                        if (a.accelflag) {
                            pdot = pdot + d * ds[u][k];
                            pold = pold + ds[u][k] * ds[u][k];
                        }
                        pnew = pnew + d * d;
                        pnorm = pnorm + so[u][k] * so[u][k];
                    }
                }
            }
        }
        sdot += (double)pdot; sold += (double)pold; snew += (double)pnew; snorm += (double)pnorm;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jlast = max(jlast, __shfl_xor_sync(FULLMASK, jlast, o));
        if (lane == 0) {
            int nsn;
            if (a.fixsh) nsn = a.shptr_old[i + 1] - a.shptr_old[i];
            else {
                int js = jlast + 1;                            // 1-based; 0 = none
                if (js == 0 && a.srctype != 'S') js = 1;
                if (js == 0) nsn = 0;
                else {
                    const int ls = a.lofj[js - 1], mm = a.mm;
                    if (ls <= mm) nsn = ls * (ls + 1) + ls + 1;
                    else nsn = (2 * mm + 1) * ls - (mm * (1 + (mm - 1))) + mm + 1;
                }
            }
            a.ns_new[i] = nsn;
        }
        __syncwarp();
        ir = ir2; nr = nr2; is = is2; ns = ns2; iso = iso2; dfl = dfl2;
        if (inext < a.npts) load_batch(0, ir, nr, is, ns, iso);
    }
    sdot = warp_sum_d(sdot); sold = warp_sum_d(sold); snew = warp_sum_d(snew); snorm = warp_sum_d(snorm);
    if (lane == 0) { red[warp][0] = sdot; red[warp][1] = sold; red[warp][2] = snew; red[warp][3] = snorm; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < CS_WARPS; w++) t += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// Kernel B: DELSOURCE at the old offsets (pass 2) and the re-packed SOURCE at the new offsets (pass 3)
template <int NST, int SLOTS>
__global__ void __launch_bounds__(CS_WARPS * 32) cs_write_kernel(CsArgs a)
{
    extern __shared__ float smem[];
    constexpr int NB = CsBatch<NST>::N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlt = a.nstleg * (a.nleg + 1);
    float *legent = smem + (size_t)warp * 2 * nlt, *legent1 = legent + nlt;
    const bool dodel = !a.first && a.accelflag;
    CsLaneTab tab;
    tab.init(a, lane);
    // same software pipeline over the warp's points as in cs_norms_kernel
    const int stride = gridDim.x * CS_WARPS;
    int i = blockIdx.x * CS_WARPS + warp;
    int ir = 0, nr = 0, is_old = 0, ns_old = 0, is_new = 0, ns_new = 0;
    float dfl = 0.0f;
    float r[NB][NST], so[NB][NST];
    auto load_ptrs = [&](int ip, int &ir_, int &nr_, int &is_, int &ns_, int &isn_, int &nsn_, float &dfl_) {
        dfl_ = __ldg(&a.dirflux[ip]);
        ir_ = a.rshptr[ip]; nr_ = a.rshptr[ip + 1] - ir_;
        is_ = a.shptr_old[ip]; ns_ = a.shptr_old[ip + 1] - is_;
        isn_ = a.shptr_new[ip]; nsn_ = a.shptr_new[ip + 1] - isn_;
    };
    auto load_batch = [&](int j0, int ir_, int nr_, int is_, int ns_) {
        const float *rad = a.radiance + (size_t)NST * ir_;
        const float *sop = a.source_old + (size_t)NST * is_;
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int j = j0 + 32 * u + lane;
#pragma unroll
            for (int k = 0; k < NST; k++) {
                r[u][k] = (j < nr_) ? __ldg(&rad[(size_t)NST * j + k]) : 0.0f;
                so[u][k] = (dodel && j < ns_) ? __ldg(&sop[(size_t)NST * j + k]) : 0.0f;
            }
        }
    };
    CsMixRow<SLOTS> mix;
    (void)legent1;
    if (i < a.npts) { load_ptrs(i, ir, nr, is_old, ns_old, is_new, ns_new, dfl); mix.load(a, i, lane, nlt); load_batch(0, ir, nr, is_old, ns_old); }
    for (; i < a.npts; i += stride) {
        const int inext = i + stride;
        int ir2 = 0, nr2 = 0, is2 = 0, ns2 = 0, isn2 = 0, nsn2 = 0;
        float dfl2 = 0.0f;
        if (inext < a.npts) load_ptrs(inext, ir2, nr2, is2, ns2, isn2, nsn2, dfl2);
        const int nmax = ns_old > ns_new ? ns_old : ns_new;
        const float flux0 = dfl * a.secmu0;
        mix.to_shared(legent, lane, nlt);
        const float albedo = mix.ap.x, planck = mix.ap.y;
        if (inext < a.npts) mix.load(a, inext, lane, nlt);
#pragma unroll
        for (int b = 0; b < CS_JSLOTS / NB; b++) {
            const int j0 = 32 * NB * b;
            if (j0 >= nmax) break;
            if (b > 0) load_batch(j0, ir, nr, is_old, ns_old);
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int j = j0 + 32 * u + lane;
                if (j >= nmax) continue;
                float s[NST];
                cs_calc_j<NST>(a, legent, j, tab.l[b * NB + u], tab.ys[b * NB + u], j < nr, r[u], flux0, planck, albedo, s);
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    if (dodel && j < ns_old) a.delsource_new[(size_t)NST * (is_old + j) + k] = s[k] - so[u][k];
                    if (j < ns_new) a.source_new[(size_t)NST * (is_new + j) + k] = s[k];
                }
            }
        }
        for (int j0 = 32 * CS_JSLOTS; j0 < nmax; j0 += 32 * NB) {          // NLM > 256: table loads
            load_batch(j0, ir, nr, is_old, ns_old);
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int j = j0 + 32 * u + lane;
                if (j >= nmax) continue;
                float s[NST];
                cs_calc_j<NST>(a, legent, j, a.lofj[j], a.ylmsun[(size_t)a.nstleg * j], j < nr, r[u], flux0, planck, albedo, s);
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    if (dodel && j < ns_old) a.delsource_new[(size_t)NST * (is_old + j) + k] = s[k] - so[u][k];
                    if (j < ns_new) a.source_new[(size_t)NST * (is_new + j) + k] = s[k];
                }
            }
        }
        __syncwarp();
        ir = ir2; nr = nr2; is_old = is2; ns_old = ns2; is_new = isn2; ns_new = nsn2; dfl = dfl2;
        if (inext < a.npts) load_batch(0, ir, nr, is_old, ns_old);
    }
}

// Kernel F (FIXSH): the truncation lengths stay (NS_new = NS_old), so the offsets of the new SOURCE are known and the
// routine is ONE streaming pass: norms, DELSOURCE and SOURCE from a single evaluation of the temporary source -- RADIANCE
// is read once instead of twice and nothing beyond NS_old is evaluated.  Same per-j arithmetic and the same per-lane
// summation order as cs_norms_kernel + cs_write_kernel (identical SOURCE, DELSOURCE and norms).  source_new may alias
// source_old and delsource_new delsource_old (every element is read and then written by the same thread).
template <int NST, int SLOTS>
__global__ void __launch_bounds__(CS_WARPS * 32, (NST == 1 ? AT3D_CS_MINB : 1)) cs_fused_kernel(CsArgs a)
{
    extern __shared__ float smem[];
    constexpr int NB = CsBatch<NST>::N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlt = a.nstleg * (a.nleg + 1);
    float *legent = smem + (size_t)warp * 2 * nlt;
    __shared__ double red[CS_WARPS][4];
    double sdot = 0.0, sold = 0.0, snew = 0.0, snorm = 0.0;
    const bool donorm = !a.first, doacc = a.accelflag && !a.first;
    CsLaneTab tab;
    tab.init(a, lane);
    const int stride = gridDim.x * CS_WARPS;
    int i = blockIdx.x * CS_WARPS + warp;
    int ir = 0, nr = 0, is = 0, ns = 0, iso = 0, nsc = 0;      // ns: NS_old; nsc: terms common with the old DELSOURCE
    float dfl = 0.0f;
    float r[NB][NST], so[NB][NST], ds[NB][NST];
    auto load_ptrs = [&](int ip, int &ir_, int &nr_, int &is_, int &ns_, int &iso_, int &nsc_, float &dfl_) {
        dfl_ = __ldg(&a.dirflux[ip]);
        ir_ = a.rshptr[ip]; nr_ = a.rshptr[ip + 1] - ir_;
        is_ = a.shptr_old[ip]; ns_ = a.shptr_old[ip + 1] - is_;
        iso_ = 0; nsc_ = ns_;
        if (doacc) {
            iso_ = a.oshptr_old[ip];
            const int nso = a.oshptr_old[ip + 1] - iso_;
            if (nso < nsc_) nsc_ = nso;
        }
    };
    auto load_batch = [&](int j0, int ir_, int nr_, int is_, int ns_, int iso_, int nsc_) {
        const float *rad = a.radiance + (size_t)NST * ir_;
        const float *sop = a.source_old + (size_t)NST * is_, *dsp = a.delsource_old + (size_t)NST * iso_;
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int j = j0 + 32 * u + lane;
#pragma unroll
            for (int k = 0; k < NST; k++) {
                r[u][k] = (j < nr_) ? __ldg(&rad[(size_t)NST * j + k]) : 0.0f;
                so[u][k] = (donorm && j < ns_) ? __ldg(&sop[(size_t)NST * j + k]) : 0.0f;
                ds[u][k] = (doacc && j < nsc_) ? __ldg(&dsp[(size_t)NST * j + k]) : 0.0f;
            }
        }
    };
    auto body = [&](int j, int l, float ysun, const float (&rr)[NST], const float (&soo)[NST], const float (&dss)[NST],
                    float flux0, float planck, float albedo, float &pdot, float &pold, float &pnew, float &pnorm) {
        float sv[NST];
        cs_calc_j<NST>(a, legent, j, l, ysun, j < nr, rr, flux0, planck, albedo, sv);
        if (donorm && j < nsc) {
#pragma unroll
            for (int k = 0; k < NST; k++) {
                const float d = sv[k] - soo[k];
                if (a.accelflag) {
                    pdot = pdot + d * dss[k];
                    pold = pold + dss[k] * dss[k];
                }
                pnew = pnew + d * d;
                pnorm = pnorm + soo[k] * soo[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NST; k++) {
            if (doacc) a.delsource_new[(size_t)NST * (is + j) + k] = sv[k] - soo[k];
            a.source_new[(size_t)NST * (is + j) + k] = sv[k];
        }
    };
    CsMixRow<SLOTS> mix;
    if (i < a.npts) { load_ptrs(i, ir, nr, is, ns, iso, nsc, dfl); mix.load(a, i, lane, nlt); load_batch(0, ir, nr, is, ns, iso, nsc); }
    for (; i < a.npts; i += stride) {
        const int inext = i + stride;
        int ir2 = 0, nr2 = 0, is2 = 0, ns2 = 0, iso2 = 0, nsc2 = 0;
        float dfl2 = 0.0f;
        if (inext < a.npts) load_ptrs(inext, ir2, nr2, is2, ns2, iso2, nsc2, dfl2);
        if (nr > a.nlm) {
            if (lane == 0) atomicCAS(a.bad, 0, i + 1);
            ir = ir2; nr = nr2; is = is2; ns = ns2; iso = iso2; nsc = nsc2; dfl = dfl2;
            if (inext < a.npts) { mix.load(a, inext, lane, nlt); load_batch(0, ir, nr, is, ns, iso, nsc); }
            continue;
        }
        const float flux0 = dfl * a.secmu0;
        mix.to_shared(legent, lane, nlt);
        const float albedo = mix.ap.x, planck = mix.ap.y;
        if (inext < a.npts) mix.load(a, inext, lane, nlt);
        float pdot = 0.0f, pold = 0.0f, pnew = 0.0f, pnorm = 0.0f;
#pragma unroll
        for (int b = 0; b < CS_JSLOTS / NB; b++) {
            const int j0 = 32 * NB * b;
            if (j0 >= ns) break;
            if (b > 0) load_batch(j0, ir, nr, is, ns, iso, nsc);
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int j = j0 + 32 * u + lane;
                if (j >= ns) continue;
                body(j, tab.l[b * NB + u], tab.ys[b * NB + u], r[u], so[u], ds[u], flux0, planck, albedo, pdot, pold, pnew, pnorm);
            }
        }
        for (int j0 = 32 * CS_JSLOTS; j0 < ns; j0 += 32 * NB) {            // NLM > 256: table loads
            load_batch(j0, ir, nr, is, ns, iso, nsc);
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int j = j0 + 32 * u + lane;
                if (j >= ns) continue;
                body(j, a.lofj[j], a.ylmsun[(size_t)a.nstleg * j], r[u], so[u], ds[u], flux0, planck, albedo, pdot, pold, pnew, pnorm);
            }
        }
        sdot += (double)pdot; sold += (double)pold; snew += (double)pnew; snorm += (double)pnorm;
        if (lane == 0) a.ns_new[i] = ns;
        __syncwarp();
        ir = ir2; nr = nr2; is = is2; ns = ns2; iso = iso2; nsc = nsc2; dfl = dfl2;
        if (inext < a.npts) load_batch(0, ir, nr, is, ns, iso, nsc);
    }
    sdot = warp_sum_d(sdot); sold = warp_sum_d(sold); snew = warp_sum_d(snew); snorm = warp_sum_d(snorm);
    if (lane == 0) { red[warp][0] = sdot; red[warp][1] = sold; red[warp][2] = snew; red[warp][3] = snorm; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < CS_WARPS; w++) t += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// Kernel G (adaptive truncation in ONE pass): the new truncation length of a point is only known after its temporary source
// has been evaluated, and the offset of its SOURCE block after those of all earlier points -- which is why the reference
// (and cs_norms_kernel + scan + cs_write_kernel) evaluate the source twice.  Here a warp takes a chunk of P consecutive
// points, keeps their temporary sources in shared memory, and obtains the offset of the chunk with a decoupled look-back
// over the chunk totals (chunks are handed out by a ticket counter, so every predecessor of a chunk is already running):
// RADIANCE, SOURCE_old and DELSOURCE_old are read once.  Per-j arithmetic, per-lane summation order and the norms of
// cs_norms_kernel (the partial sums are kept per chunk and reduced in chunk order: deterministic).  delsource_new must
// not alias delsource_old (their offsets differ: OSHPTR vs. SHPTR_old).
struct CsAdapt {
    int P, nchunks, stride;                  // points per chunk, chunks, floats per staged point (NLM * NSTOKES)
    unsigned long long *tile;                // [nchunks] look-back status: flag << 62 | value
    int *ticket;
    double *chunk_sums;                      // [nchunks,4]
    int *shptr_new;                          // [npts+1]
    unsigned long long cap;                  // capacity of source_new in SH terms
    int *overflow;
};
#define CS_TILE_AGG  (1ull << 62)
#define CS_TILE_PFX  (2ull << 62)
#define CS_TILE_MASK ((1ull << 62) - 1)

template <int NST, int SLOTS>
__global__ void __launch_bounds__(CS_WARPS * 32, (NST == 1 ? 3 : 1)) cs_adapt_kernel(CsArgs a, CsAdapt q)
{
    extern __shared__ float smem[];
    constexpr int NB = CsBatch<NST>::N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlt = a.nstleg * (a.nleg + 1);
    float *legent = smem + (size_t)warp * 2 * nlt;
    float *stage = smem + (size_t)CS_WARPS * 2 * nlt + (size_t)warp * q.P * q.stride;
    const bool donorm = !a.first, doacc = a.accelflag && !a.first;
    CsLaneTab tab;
    tab.init(a, lane);
    float r[NB][NST], so[NB][NST], ds[NB][NST];
    int ir = 0, nr = 0, is = 0, ns = 0, iso = 0, nsc = 0;      // ns: NS_old; nsc: terms common with the old DELSOURCE
    float dfl = 0.0f;
    auto load_ptrs = [&](int ip, int &ir_, int &nr_, int &is_, int &ns_, int &iso_, int &nsc_, float &dfl_) {
        dfl_ = __ldg(&a.dirflux[ip]);
        ir_ = a.rshptr[ip]; nr_ = a.rshptr[ip + 1] - ir_;
        is_ = a.shptr_old[ip]; ns_ = a.shptr_old[ip + 1] - is_;
        iso_ = 0; nsc_ = ns_;
        if (doacc) {
            iso_ = a.oshptr_old[ip];
            const int nso = a.oshptr_old[ip + 1] - iso_;
            if (nso < nsc_) nsc_ = nso;
        }
    };
    auto load_batch = [&](int j0, int ir_, int nr_, int is_, int ns_, int iso_, int nsc_) {
        const float *rad = a.radiance + (size_t)NST * ir_;
        const float *sop = a.source_old + (size_t)NST * is_, *dsp = a.delsource_old + (size_t)NST * iso_;
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int j = j0 + 32 * u + lane;
#pragma unroll
            for (int k = 0; k < NST; k++) {
                r[u][k] = (j < nr_) ? __ldg(&rad[(size_t)NST * j + k]) : 0.0f;
                so[u][k] = (donorm && j < ns_) ? __ldg(&sop[(size_t)NST * j + k]) : 0.0f;
                ds[u][k] = (doacc && j < nsc_) ? __ldg(&dsp[(size_t)NST * j + k]) : 0.0f;
            }
        }
    };
    // a tile = CS_WARPS consecutive chunks, one per warp, handed to the block by a ticket counter; ONE look-back per tile
    __shared__ int s_tile[2];
    __shared__ unsigned long long s_base;
    __shared__ int s_wtot[CS_WARPS];
    __shared__ double s_red[CS_WARPS][4];
    const int ntiles = (q.nchunks + CS_WARPS - 1) / CS_WARPS;
    int tsel = 0;
    auto draw = [&](int sel) {                                   // called by all threads of the block
        if (threadIdx.x == 0) s_tile[sel] = atomicAdd(q.ticket, 1);
        __syncthreads();
        return s_tile[sel];
    };
    CsMixRow<SLOTS> mix;
    int tile = draw(tsel);
    int chunk = tile < ntiles ? tile * CS_WARPS + warp : q.nchunks;
    // software pipeline over the warp's points (consecutive within a chunk, then the first point of the next chunk, whose
    // ticket is drawn one chunk ahead): block pointers at the top of an iteration, mixed row after this point's row went to
    // shared memory, first batch of SH values at the bottom
    if (chunk < q.nchunks) {
        const int i = chunk * q.P;
        load_ptrs(i, ir, nr, is, ns, iso, nsc, dfl); mix.load(a, i, lane, nlt); load_batch(0, ir, nr, is, ns, iso, nsc);
    }
    while (tile < ntiles) {
        tsel ^= 1;
        const int next_tile = draw(tsel);
        const int next_chunk = next_tile < ntiles ? next_tile * CS_WARPS + warp : q.nchunks;
        const int i0 = chunk * q.P;
        const int np = chunk < q.nchunks ? min(q.P, a.npts - i0) : 0;
        double sdot = 0.0, sold = 0.0, snew = 0.0, snorm = 0.0;
        int my_ns = 0;                                   // lane p: NS_new of point i0 + p
        for (int p = 0; p < np; p++) {
            const int i = i0 + p;
            const int inext = (p + 1 < np) ? i + 1 : (next_chunk < q.nchunks ? next_chunk * q.P : -1);
            int ir2 = 0, nr2 = 0, is2 = 0, ns2 = 0, iso2 = 0, nsc2 = 0;
            float dfl2 = 0.0f;
            if (inext >= 0) load_ptrs(inext, ir2, nr2, is2, ns2, iso2, nsc2, dfl2);
            if (nr > a.nlm) {
                if (lane == 0) atomicCAS(a.bad, 0, i + 1);
                ir = ir2; nr = nr2; is = is2; ns = ns2; iso = iso2; nsc = nsc2; dfl = dfl2;
                if (inext >= 0) { mix.load(a, inext, lane, nlt); load_batch(0, ir, nr, is, ns, iso, nsc); }
                continue;
            }
            const float flux0 = dfl * a.secmu0;
            mix.to_shared(legent, lane, nlt);
            const float albedo = mix.ap.x, planck = mix.ap.y;
            if (inext >= 0) mix.load(a, inext, lane, nlt);
            float *st = stage + (size_t)p * q.stride;
            int jlast = -1;
            float pdot = 0.0f, pold = 0.0f, pnew = 0.0f, pnorm = 0.0f;
            auto body = [&](int j, int l, float ysun, const float (&rr)[NST], const float (&soo)[NST], const float (&dss)[NST]) {
                float sv[NST];
                cs_calc_j<NST>(a, legent, j, l, ysun, j < nr, rr, flux0, planck, albedo, sv);
#pragma unroll
                for (int k = 0; k < NST; k++) if (fabsf(sv[k]) > a.srcmin) jlast = j;
                if (donorm && j < nsc) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        const float d = sv[k] - soo[k];
                        if (a.accelflag) {
                            pdot = pdot + d * dss[k];
                            pold = pold + dss[k] * dss[k];
                        }
                        pnew = pnew + d * d;
                        pnorm = pnorm + soo[k] * soo[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    if (doacc && j < ns) a.delsource_new[(size_t)NST * (is + j) + k] = sv[k] - soo[k];
                    st[(size_t)NST * j + k] = sv[k];
                }
            };
#pragma unroll
            for (int b = 0; b < CS_JSLOTS / NB; b++) {
                const int j0 = 32 * NB * b;
                if (j0 >= a.nlm) break;
                if (b > 0) load_batch(j0, ir, nr, is, ns, iso, nsc);
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int j = j0 + 32 * u + lane;
                    if (j >= a.nlm) continue;
                    body(j, tab.l[b * NB + u], tab.ys[b * NB + u], r[u], so[u], ds[u]);
                }
            }
            for (int j0 = 32 * CS_JSLOTS; j0 < a.nlm; j0 += 32 * NB) {        // NLM > 256: table loads
                load_batch(j0, ir, nr, is, ns, iso, nsc);
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int j = j0 + 32 * u + lane;
                    if (j >= a.nlm) continue;
                    body(j, a.lofj[j], a.ylmsun[(size_t)a.nstleg * j], r[u], so[u], ds[u]);
                }
            }
            sdot += (double)pdot; sold += (double)pold; snew += (double)pnew; snorm += (double)pnorm;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) jlast = max(jlast, __shfl_xor_sync(FULLMASK, jlast, o));
            int nsn;
            {
                int js = jlast + 1;                            // 1-based; 0 = none
                if (js == 0 && a.srctype != 'S') js = 1;
                if (js == 0) nsn = 0;
                else {
                    const int ls = a.lofj[js - 1], mm = a.mm;
                    if (ls <= mm) nsn = ls * (ls + 1) + ls + 1;
                    else nsn = (2 * mm + 1) * ls - (mm * (1 + (mm - 1))) + mm + 1;
                }
            }
            if (lane == p) my_ns = nsn;
            if (lane == 0) a.ns_new[i] = nsn;
            __syncwarp();
            ir = ir2; nr = nr2; is = is2; ns = ns2; iso = iso2; nsc = nsc2; dfl = dfl2;
            if (inext >= 0) load_batch(0, ir, nr, is, ns, iso, nsc);
        }
        // offsets: exclusive scan of the chunk's NS_new over the lanes, of the chunk totals over the warps, tile base by
        // decoupled look-back over the tile totals (warp 0)
        int incl = my_ns;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
        const int wtotal = __shfl_sync(FULLMASK, incl, 31);
        const int excl = incl - my_ns;
        if (lane == 0) s_wtot[warp] = wtotal;
        __syncthreads();
        if (warp == 0) {
            unsigned long long total = 0;
#pragma unroll
            for (int w = 0; w < CS_WARPS; w++) total += (unsigned long long)s_wtot[w];
            unsigned long long base = 0;
            if (tile == 0) {
                if (lane == 0) { __threadfence(); *(volatile unsigned long long *)&q.tile[0] = CS_TILE_PFX | total; }
            } else {
                if (lane == 0) { __threadfence(); *(volatile unsigned long long *)&q.tile[tile] = CS_TILE_AGG | total; }
                int look = tile - 1;
                for (;;) {
                    const int c = look - lane;
                    unsigned long long v = CS_TILE_PFX;                 // lanes beyond tile 0 read as "prefix 0"
                    if (c >= 0) {
                        do { v = *(volatile unsigned long long *)&q.tile[c]; if (!(v >> 62)) __nanosleep(40); } while (!(v >> 62));
                    }
                    const unsigned pm = __ballot_sync(FULLMASK, (v >> 62) == 2);
                    const int stop = pm ? __ffs(pm) - 1 : 32;           // nearest predecessor that already has its prefix
                    unsigned long long part = (lane <= stop) ? (v & CS_TILE_MASK) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULLMASK, part, o);
                    base += part;
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) { __threadfence(); *(volatile unsigned long long *)&q.tile[tile] = CS_TILE_PFX | (base + total); }
            }
            if (lane == 0) {
                s_base = base;
                if (tile == ntiles - 1) q.shptr_new[a.npts] = (int)(base + total);
            }
        }
        __syncthreads();
        unsigned long long base = s_base;
        for (int w = 0; w < warp; w++) base += (unsigned long long)s_wtot[w];
        // SHPTR_new and the re-packed SOURCE of the chunk's points from shared memory
        if (lane < np) q.shptr_new[i0 + lane] = (int)(base + (unsigned long long)excl);
        __syncwarp();
        for (int p = 0; p < np; p++) {
            const int nsn = __shfl_sync(FULLMASK, my_ns, p);
            const unsigned long long off = base + (unsigned long long)__shfl_sync(FULLMASK, excl, p);
            if (off + (unsigned long long)nsn > q.cap) { if (lane == 0) *q.overflow = 1; continue; }
            const float *st = stage + (size_t)p * q.stride;
            float *dst = a.source_new + (size_t)NST * off;
            for (int t = lane; t < NST * nsn; t += 32) dst[t] = st[t];
        }
        __syncwarp();
        sdot = warp_sum_d(sdot); sold = warp_sum_d(sold); snew = warp_sum_d(snew); snorm = warp_sum_d(snorm);
        if (lane == 0) { s_red[warp][0] = sdot; s_red[warp][1] = sold; s_red[warp][2] = snew; s_red[warp][3] = snorm; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double t = 0.0;
            for (int w = 0; w < CS_WARPS; w++) t += s_red[w][threadIdx.x];
            q.chunk_sums[(size_t)tile * 4 + threadIdx.x] = t;          // per tile, reduced in tile order afterwards
        }
        tile = next_tile;
        chunk = next_chunk;
    }
}

// deterministic final reduction of the per-block partial sums
__global__ void cs_reduce_kernel(int nblocks, const double *partials, double *out)
{
    __shared__ double sh[4][256];
    const int q = threadIdx.x >> 8 & 3, t = threadIdx.x & 255;
    double acc = 0.0;
    for (int b = t; b < nblocks; b += 256) acc += partials[(size_t)b * 4 + q];
    sh[q][t] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) sh[q][t] += sh[q][t + o];
        __syncthreads();
    }
    if (t == 0) out[q] = sh[q][0];
}

namespace {
struct Arena {
    std::vector<void *> p;
    ~Arena() { for (void *q : p) at3d_free(q); }
    template <typename T> T *alloc(size_t n)
    {
        void *q = nullptr;
        if (at3d_malloc(&q, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
        p.push_back(q);
        return (T *)q;
    }
    template <typename T> const T *up(const T *h, size_t n)
    {
        if (!h) return nullptr;
        T *d = alloc<T>(n);
        if (!d) return nullptr;
        if (cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return d;
    }
};
}

// The three launches of one COMPUTE_SOURCE on device-resident arrays (every pointer of `a` except shptr_new /
// source_new set by the caller; scan_tmp sized by cs_scan_bytes): norms + new truncation lengths, SHPTR scan, write.
// Returns 0, 1 (NR>NLM) or 2 (MAXIV exceeded); *total_new = SHPTR(NPTS+1) of the new source.
size_t cs_scan_bytes(int npts)
{
    size_t tmpb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpb, (int *)nullptr, (int *)nullptr, npts + 1);
    return tmpb;
}

int cs_grid_blocks(int npts)
{
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int want = (npts + CS_WARPS - 1) / CS_WARPS;
    return want < nsm * 8 ? want : nsm * 8;          // persistent: a multiple of the SM count
}

// the mixed Legendre rows of the points [first, first+count) (new grid points of SPLIT_GRID): cs_mix_kernel on shifted
// base pointers; a.ldp stays the leading dimension of the species arrays
cudaError_t cs_mix_points(CsArgs a, int first, int count, cudaStream_t st)
{
    if (count < 1) return cudaSuccess;
    const size_t nlt = (size_t)a.nstleg * (a.nleg + 1);
    const int slots = nlt <= 32 ? 1 : nlt <= 64 ? 2 : nlt <= 128 ? 4 : 8;
    const size_t smem = (size_t)CS_WARPS * 2 * nlt * sizeof(float);
    if (!a.ldp) a.ldp = a.npts;
    a.npts = count;
    a.total_ext += first; a.extinct += first; a.albedo += first; if (a.planck) a.planck += first;
    a.iphase += (size_t)a.nq * first; a.phaseinterpwt += (size_t)a.nq * first;
    a.mix_legent += nlt * first; a.mix_ap += first;
    const int nmix = (count + CS_WARPS - 1) / CS_WARPS;
    if (slots == 1) cs_mix_kernel<1><<<nmix, CS_WARPS * 32, smem, st>>>(a);
    else if (slots == 2) cs_mix_kernel<2><<<nmix, CS_WARPS * 32, smem, st>>>(a);
    else if (slots == 4) cs_mix_kernel<4><<<nmix, CS_WARPS * 32, smem, st>>>(a);
    else cs_mix_kernel<8><<<nmix, CS_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

int cs_device_step(CsArgs &a, int nblk, void *scan_tmp, size_t tmpb, int *shptr_new, double *sums, int maxiv, size_t cap_new,
                   float *source_new, int *total_new_out, bool mix_ready, char *errmsg)
{
    const int nst = a.nstokes, npts = a.npts;
    const size_t nlt = (size_t)a.nstleg * (a.nleg + 1);
    const int slots = nlt <= 32 ? 1 : nlt <= 64 ? 2 : nlt <= 128 ? 4 : 8;
    const size_t smem = (size_t)CS_WARPS * 2 * nlt * sizeof(float);
#define CS_LAUNCH(K)                                                                              \
    {                                                                                             \
        if (nst == 1) {                                                                           \
            if (slots == 1) K<1, 1><<<nblk, CS_WARPS * 32, smem>>>(a);                            \
            else if (slots == 2) K<1, 2><<<nblk, CS_WARPS * 32, smem>>>(a);                       \
            else if (slots == 4) K<1, 4><<<nblk, CS_WARPS * 32, smem>>>(a);                       \
            else K<1, 8><<<nblk, CS_WARPS * 32, smem>>>(a);                                       \
        } else {                                                                                  \
            if (slots == 1) K<3, 1><<<nblk, CS_WARPS * 32, smem>>>(a);                            \
            else if (slots == 2) K<3, 2><<<nblk, CS_WARPS * 32, smem>>>(a);                       \
            else if (slots == 4) K<3, 4><<<nblk, CS_WARPS * 32, smem>>>(a);                       \
            else K<3, 8><<<nblk, CS_WARPS * 32, smem>>>(a);                                       \
        }                                                                                         \
    }
    cudaMemsetAsync(a.bad, 0, sizeof(int), 0);
    cudaMemsetAsync(a.ns_new + npts, 0, sizeof(int), 0);
    a.shptr_new = nullptr; a.source_new = nullptr;
    if (!mix_ready) {
        const int nmix = (npts + CS_WARPS - 1) / CS_WARPS;
        if (slots == 1) cs_mix_kernel<1><<<nmix, CS_WARPS * 32, smem>>>(a);
        else if (slots == 2) cs_mix_kernel<2><<<nmix, CS_WARPS * 32, smem>>>(a);
        else if (slots == 4) cs_mix_kernel<4><<<nmix, CS_WARPS * 32, smem>>>(a);
        else cs_mix_kernel<8><<<nmix, CS_WARPS * 32, smem>>>(a);
    }
    if (a.fixsh && source_new) {
        // FIXSH: one pass (cs_fused_kernel); SHPTR is unchanged
        int total_new = 0, hbad = 0;
        cudaMemcpy(&total_new, a.shptr_old + npts, sizeof(int), cudaMemcpyDeviceToHost);
        if (total_new_out) *total_new_out = total_new;
        if ((size_t)total_new > cap_new) { set_msg(errmsg, "COMPUTE_SOURCE: the new SOURCE buffer is too small for FIXSH"); return 2; }
        a.source_new = source_new;
        if (shptr_new != a.shptr_old)
            cudaMemcpyAsync(shptr_new, a.shptr_old, sizeof(int) * ((size_t)npts + 1), cudaMemcpyDeviceToDevice, 0);
        CS_LAUNCH(cs_fused_kernel)
        cs_reduce_kernel<<<1, 1024>>>(nblk, a.partials, sums);
        cudaMemcpy(&hbad, a.bad, sizeof(int), cudaMemcpyDeviceToHost);
        a.shptr_new = shptr_new;
        if (hbad) { set_msg(errmsg, "COMPUTE_SOURCE: NR>NLM 3 %d", hbad); return 1; }
        return 0;
    }
    // adaptive truncation in one pass (cs_adapt_kernel), opt-in with AT3D_B200_CS_ADAPT=one, when the new DELSOURCE does not
    // overwrite the old one and the temporary sources of a chunk fit shared memory.  Measured at 6.55 M points x NLM 256
    // (profiles/r2zj_csadapt_ncu_summary.txt): DRAM traffic 25.1 GB instead of ~35 GB, but 11.4 ms against 10.8 ms for the two
    // passes -- the block-wide wait for the tile offset (look-back) and 24 resident warps per SM (64 KB of staging per
    // block) cost more than the second read of RADIANCE saves; the two-pass route stays the default.
    {
        const bool doacc = a.accelflag && !a.first;
        const size_t stride = (size_t)a.nlm * nst;
        const size_t room = 96 * 1024 > smem ? 96 * 1024 - smem : 0;
        int P = (int)(room / (CS_WARPS * stride * sizeof(float)));
        if (P > 8) P = 8;
        const char *env = getenv("AT3D_B200_CS_ADAPT");
        const bool one = env && !strcmp(env, "one");
        if (one && source_new && P >= 1 && (!doacc || a.delsource_new != a.delsource_old)) {
            CsAdapt q;
            q.P = P; q.nchunks = (npts + P - 1) / P; q.stride = (int)stride;
            const size_t b_tile = ((size_t)q.nchunks * sizeof(unsigned long long) + 255) & ~(size_t)255;
            const size_t b_sums = ((size_t)q.nchunks * 4 * sizeof(double) + 255) & ~(size_t)255;
            static DevBuf scratch;                                  // per process; calls on one stream at a time (as g_cswork)
            if (scratch.reserve(b_tile + b_sums + 256) != cudaSuccess) { set_msg(errmsg, "COMPUTE_SOURCE: device allocation failure"); return 4; }
            unsigned char *sb = (unsigned char *)scratch.p;
            q.tile = (unsigned long long *)sb;
            q.chunk_sums = (double *)(sb + b_tile);
            q.ticket = (int *)(sb + b_tile + b_sums);
            q.overflow = q.ticket + 1;
            q.shptr_new = shptr_new;
            q.cap = (unsigned long long)(cap_new < (size_t)maxiv ? cap_new : (size_t)maxiv);
            cudaMemsetAsync(q.tile, 0, b_tile, 0);
            cudaMemsetAsync(q.ticket, 0, 2 * sizeof(int), 0);
            a.source_new = source_new;
            const size_t smem2 = smem + (size_t)CS_WARPS * P * stride * sizeof(float);
            int dev = 0, nsm = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            const int want = (q.nchunks + CS_WARPS - 1) / CS_WARPS;            // tiles
            const int nb2 = want < nsm * 3 ? want : nsm * 3;
#define CS_LAUNCH_A(NSTV, SL)                                                                                     \
            {                                                                                                     \
                cudaFuncSetAttribute(cs_adapt_kernel<NSTV, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2); \
                cs_adapt_kernel<NSTV, SL><<<nb2, CS_WARPS * 32, smem2>>>(a, q);                                   \
            }
            if (nst == 1) {
                if (slots == 1) CS_LAUNCH_A(1, 1) else if (slots == 2) CS_LAUNCH_A(1, 2) else if (slots == 4) CS_LAUNCH_A(1, 4) else CS_LAUNCH_A(1, 8)
            } else {
                if (slots == 1) CS_LAUNCH_A(3, 1) else if (slots == 2) CS_LAUNCH_A(3, 2) else if (slots == 4) CS_LAUNCH_A(3, 4) else CS_LAUNCH_A(3, 8)
            }
#undef CS_LAUNCH_A
            cs_reduce_kernel<<<1, 1024>>>((q.nchunks + CS_WARPS - 1) / CS_WARPS, q.chunk_sums, sums);
            int total_new = 0, hbad = 0, hov[2] = {0, 0};
            cudaMemcpy(&total_new, shptr_new + npts, sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(&hbad, a.bad, sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(hov, q.ticket, 2 * sizeof(int), cudaMemcpyDeviceToHost);
            if (total_new_out) *total_new_out = total_new;
            a.shptr_new = shptr_new;
            if (hbad) { set_msg(errmsg, "COMPUTE_SOURCE: NR>NLM 3 %d", hbad); return 1; }
            if (hov[1] || total_new > maxiv || (size_t)total_new > cap_new) {
                set_msg(errmsg, "COMPUTE_SOURCE: MAXIV exceeded %d Out of memory for more spherical harmonic terms.", maxiv);
                return 2;
            }
            return 0;
        }
    }
    CS_LAUNCH(cs_norms_kernel)
    cs_reduce_kernel<<<1, 1024>>>(nblk, a.partials, sums);
    cub::DeviceScan::ExclusiveSum(scan_tmp, tmpb, a.ns_new, shptr_new, npts + 1);
    int total_new = 0, hbad = 0;
    cudaMemcpy(&total_new, shptr_new + npts, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(&hbad, a.bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (total_new_out) *total_new_out = total_new;
    if (hbad) { set_msg(errmsg, "COMPUTE_SOURCE: NR>NLM 3 %d", hbad); return 1; }
    if (total_new > maxiv || (size_t)total_new > cap_new) {
        set_msg(errmsg, "COMPUTE_SOURCE: MAXIV exceeded %d Out of memory for more spherical harmonic terms.", maxiv);
        return 2;
    }
    a.shptr_new = shptr_new;
    a.source_new = source_new;
    CS_LAUNCH(cs_write_kernel)
#undef CS_LAUNCH
    return 0;
}

extern "C" int at3d_compute_source(const at3d_state_desc *d, int fixsh, float shacc, int maxiv,
                                   int first, int accelflag, int newmethod,
                                   int32_t *shptr, float *source, int32_t *oshptr, float *delsource,
                                   float *deljdot, float *deljold, float *deljnew, float *jnorm,
                                   double *kernel_ms, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !shptr || !source || !oshptr || !delsource || !deljdot || !deljold || !deljnew || !jnorm) { set_msg(errmsg, "null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if (!(d->nstokes == 1 || d->nstokes == 3)) { set_msg(errmsg, "NSTOKES must be 1 or 3"); return 3; }
    if (!newmethod) { set_msg(errmsg, "COMPUTE_SOURCE: only NEWMETHOD=.TRUE. (the at3d default, solver.py:178) is implemented"); return 3; }
    if (!d->rshptr || !d->radiance) { set_msg(errmsg, "COMPUTE_SOURCE needs RSHPTR/RADIANCE"); return 1; }
    const size_t npts = d->npts;
    const int nst = d->nstokes;
    const size_t nlt = (size_t)d->nstleg * (d->nleg + 1);
    const int nq = 8 * d->maxnmicro;
    Arena A;
    CsArgs a;
    memset(&a, 0, sizeof(a));
    a.npts = d->npts; a.nstokes = nst; a.nstleg = d->nstleg; a.nlm = d->nlm; a.ml = d->ml; a.mm = d->mm; a.nleg = d->nleg;
    a.npart = d->npart; a.nq = nq; a.srctype = d->srctype; a.deltam = d->deltam; a.interp_new = d->interp_new;
    a.newmethod = newmethod; a.first = first; a.accelflag = accelflag; a.fixsh = fixsh;
    a.phasemax = d->phasemax; a.secmu0 = 1.0f / fabsf(d->solarmu); a.srcmin = shacc;
    std::vector<int> lofj(d->nlm);
    {
        int j = 0;
        for (int l = 0; l <= d->ml; l++) {
            const int me = l < d->mm ? l : d->mm;
            for (int m = -me; m <= me; m++) { if (j < d->nlm) lofj[j] = l; j++; }
        }
        if (j != d->nlm) { set_msg(errmsg, "NLM inconsistent with ML, MM"); return 1; }
    }
    const size_t nrad = (size_t)nst * d->rshptr[npts];
    const size_t nsrc_old = (size_t)nst * shptr[npts];
    const size_t ndel_old = (size_t)nst * (accelflag && !first ? (size_t)oshptr[npts] : 0);
    a.extinct = A.up(d->extinct, npts * d->npart); a.albedo = A.up(d->albedo, npts * d->npart);
    a.total_ext = A.up(d->total_ext, npts);
    a.legen = A.up(d->legen, nlt * d->numphase);
    a.iphase = A.up(d->iphase, (size_t)nq * npts * d->npart);
    a.phaseinterpwt = A.up(d->phaseinterpwt, (size_t)nq * npts * d->npart);
    a.dirflux = A.up(d->dirflux, npts);
    a.rshptr = A.up(d->rshptr, npts + 1);
    a.radiance = A.up(d->radiance, nrad ? nrad : 1);
    a.ylmsun = A.up(d->ylmsun, (size_t)d->nstleg * d->nlm);
    a.planck = d->planck ? A.up(d->planck, npts * d->npart) : nullptr;
    a.lofj = A.up(lofj.data(), lofj.size());
    a.shptr_old = A.up(shptr, npts + 1);
    a.oshptr_old = A.up(oshptr, npts + 1);
    a.source_old = A.up(source, nsrc_old ? nsrc_old : 1);
    a.delsource_old = A.up(delsource, ndel_old ? ndel_old : 1);
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nlt > 256) { set_msg(errmsg, "COMPUTE_SOURCE: Legendre table longer than 256 entries"); return 3; }
    const int slots = nlt <= 32 ? 1 : nlt <= 64 ? 2 : nlt <= 128 ? 4 : 8;
    const size_t smem = (size_t)CS_WARPS * 2 * nlt * sizeof(float);
    const int want = (int)((npts + CS_WARPS - 1) / CS_WARPS);
    const int nblk = want < nsm * 8 ? want : nsm * 8;          // persistent: a multiple of the SM count
    a.ns_new = A.alloc<int>(npts + 1);
    a.partials = A.alloc<double>((size_t)nblk * 4);
    a.bad = A.alloc<int>(1);
    int *shptr_new = A.alloc<int>(npts + 1);
    double *sums = A.alloc<double>(4);
    a.mix_legent = A.alloc<float>(npts * nlt);
    a.mix_ap = A.alloc<float2>(npts);
    if (!a.mix_legent || !a.mix_ap) { set_msg(errmsg, "device allocation failure"); return 4; }
    if (!a.extinct || !a.albedo || !a.total_ext || !a.legen || !a.iphase || !a.phaseinterpwt || !a.dirflux ||
        !a.rshptr || !a.radiance || !a.ylmsun || !a.lofj || !a.shptr_old || !a.oshptr_old || !a.source_old ||
        !a.delsource_old || !a.ns_new || !a.partials || !a.bad || !shptr_new || !sums) {
        set_msg(errmsg, "at3d_compute_source: NULL input array or device allocation failure");
        return 4;
    }
    cudaMemset(a.bad, 0, sizeof(int));
    cudaMemset(a.ns_new + npts, 0, sizeof(int));
    // every allocation happens before the timed region: the new SOURCE at its largest possible size
    // (FIXSH keeps the old truncation; otherwise at most NLM terms per point, capped by MAXIV), the scan scratch,
    // and a DELSOURCE that covers the OLD SHPTR offsets it is rewritten at (pass 2)
    const size_t cap_new = fixsh ? (size_t)shptr[npts]
                                 : ((size_t)maxiv < npts * (size_t)d->nlm ? (size_t)maxiv : npts * (size_t)d->nlm);
    float *source_new = A.alloc<float>((size_t)nst * (cap_new ? cap_new : 1));
    size_t tmpb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpb, a.ns_new, shptr_new, (int)npts + 1);
    void *tmp = A.alloc<unsigned char>(tmpb);
    // a separate new DELSOURCE: its offsets (SHPTR_old) differ from the old one's (OSHPTR), and the one-pass adaptive kernel
    // writes it while other warps still read the old values
    a.delsource_new = (float *)a.delsource_old;
    if (accelflag && !first) a.delsource_new = A.alloc<float>((size_t)nst * shptr[npts] + 1);
    if (!source_new || !tmp || !a.delsource_new) { set_msg(errmsg, "device allocation failure"); return 4; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    int total_new = 0;
    int rc = cs_device_step(a, nblk, tmp, tmpb, shptr_new, sums, maxiv, cap_new, source_new, &total_new, false, errmsg);
    float ms = 0.0f;
    if (!rc) {
        cudaEventRecord(e1, 0);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_compute_source", cudaGetErrorString(e)); rc = 4; }
        else cudaEventElapsedTime(&ms, e0, e1);
    }
    if (!rc) {
        double hs[4];
        cudaMemcpy(hs, sums, sizeof(hs), cudaMemcpyDeviceToHost);
        *deljdot = (float)hs[0]; *deljold = (float)hs[1]; *deljnew = (float)hs[2]; *jnorm = (float)hs[3];
        cudaMemcpy(source, a.source_new, (size_t)nst * total_new * sizeof(float), cudaMemcpyDeviceToHost);
        if (accelflag && !first) {
            cudaMemcpy(delsource, a.delsource_new, (size_t)nst * shptr[npts] * sizeof(float), cudaMemcpyDeviceToHost);
            memcpy(oshptr, shptr, (npts + 1) * sizeof(int32_t));          // OSHPTR(I)=SHPTR(I) (old)
        }
        if (cudaMemcpy(shptr, shptr_new, (npts + 1) * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_msg(errmsg, "CUDA error copying SHPTR back"); rc = 4;
        }
    }
    if (kernel_ms) *kernel_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}


// ---- a1 on device-resident arrays: every array pointer is a DEVICE pointer, nothing is staged or copied back except
//      the four norms, the new SHPTR(NPTS+1) and the error flag (at3d_b200.h) ----
namespace {
struct CsWork {                      // scratch of at3d_compute_source_device, kept between calls (one per process)
    std::mutex mu;
    DevBuf ns_new, partials, bad, sums, mix_legent, mix_ap, scan, lofj;
    int lofj_key[3] = {-1, -1, -1};
    const void *mix_key[4] = {nullptr, nullptr, nullptr, nullptr};
    int mix_npts = -1;
};
CsWork g_cswork;
}

extern "C" int at3d_compute_source_device(const at3d_cs_device_desc *d, int fixsh, float shacc, int64_t maxiv, int first,
                                          int accelflag, const int32_t *shptr_old, const float *source_old,
                                          const int32_t *oshptr_old, const float *delsource_old, float *delsource_new,
                                          int32_t *shptr_new, float *source_new, int64_t source_new_capacity,
                                          int properties_changed, float *norms /*host [4]*/, int32_t *total_new /*host*/,
                                          double *kernel_ms, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !shptr_old || !source_old || !shptr_new || !source_new || !norms || !total_new) { set_msg(errmsg, "null argument"); return 1; }
    if (accelflag && !first && (!oshptr_old || !delsource_old || !delsource_new)) { set_msg(errmsg, "the acceleration needs OSHPTR and DELSOURCE"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if (!(d->nstokes == 1 || d->nstokes == 3)) { set_msg(errmsg, "NSTOKES must be 1 or 3"); return 3; }
    if (!d->rshptr || !d->radiance || !d->extinct || !d->albedo || !d->total_ext || !d->legen || !d->iphase ||
        !d->phaseinterpwt || !d->dirflux || !d->ylmsun) { set_msg(errmsg, "a required device array is NULL"); return 1; }
    std::lock_guard<std::mutex> lock(g_cswork.mu);
    CsWork &W = g_cswork;
    const size_t npts = d->npts;
    const size_t nlt = (size_t)d->nstleg * (d->nleg + 1);
    if (nlt > 256) { set_msg(errmsg, "COMPUTE_SOURCE: Legendre table longer than 256 entries"); return 3; }
    CsArgs a;
    memset(&a, 0, sizeof(a));
    a.npts = d->npts; a.nstokes = d->nstokes; a.nstleg = d->nstleg; a.nlm = d->nlm; a.ml = d->ml; a.mm = d->mm; a.nleg = d->nleg;
    a.npart = d->npart; a.nq = 8 * d->maxnmicro; a.srctype = d->srctype; a.deltam = d->deltam; a.interp_new = d->interp_new;
    a.newmethod = 1; a.first = first; a.accelflag = accelflag; a.fixsh = fixsh;
    a.phasemax = d->phasemax; a.secmu0 = 1.0f / fabsf(d->solarmu); a.srcmin = shacc;
    a.extinct = d->extinct; a.albedo = d->albedo; a.total_ext = d->total_ext; a.legen = d->legen; a.iphase = d->iphase;
    a.phaseinterpwt = d->phaseinterpwt; a.dirflux = d->dirflux; a.rshptr = d->rshptr; a.radiance = d->radiance;
    a.ylmsun = d->ylmsun; a.planck = d->planck;
    a.shptr_old = shptr_old; a.oshptr_old = oshptr_old ? oshptr_old : shptr_old; a.source_old = source_old;
    a.delsource_old = delsource_old ? delsource_old : source_old; a.delsource_new = delsource_new;
    const int nblk = cs_grid_blocks(d->npts);
    const size_t tmpb = cs_scan_bytes(d->npts);
    cudaError_t ce = cudaSuccess;
    if (ce == cudaSuccess) ce = W.ns_new.reserve(sizeof(int) * (npts + 1));
    if (ce == cudaSuccess) ce = W.partials.reserve(sizeof(double) * 4 * (size_t)nblk);
    if (ce == cudaSuccess) ce = W.bad.reserve(256);
    if (ce == cudaSuccess) ce = W.sums.reserve(256);
    if (ce == cudaSuccess) ce = W.scan.reserve(tmpb + 256);
    if (ce == cudaSuccess) ce = W.lofj.reserve(sizeof(int) * (size_t)d->nlm);
    const bool mix_realloc = W.mix_legent.cap < sizeof(float) * npts * nlt || W.mix_ap.cap < sizeof(float2) * npts;
    if (ce == cudaSuccess) ce = W.mix_legent.reserve(sizeof(float) * npts * nlt);
    if (ce == cudaSuccess) ce = W.mix_ap.reserve(sizeof(float2) * npts);
    if (ce != cudaSuccess) { set_msg(errmsg, "device allocation failure (%s)", cudaGetErrorString(ce)); return 4; }
    if (W.lofj_key[0] != d->ml || W.lofj_key[1] != d->mm || W.lofj_key[2] != d->nlm) {
        std::vector<int> lofj(d->nlm);
        int j = 0;
        for (int l = 0; l <= d->ml; l++) {
            const int me = l < d->mm ? l : d->mm;
            for (int m = -me; m <= me; m++) { if (j < d->nlm) lofj[j] = l; j++; }
        }
        if (j != d->nlm) { set_msg(errmsg, "NLM inconsistent with ML, MM"); return 1; }
        cudaMemcpy(W.lofj.p, lofj.data(), sizeof(int) * lofj.size(), cudaMemcpyHostToDevice);
        W.lofj_key[0] = d->ml; W.lofj_key[1] = d->mm; W.lofj_key[2] = d->nlm;
    }
    a.lofj = (const int *)W.lofj.p;
    a.ns_new = (int *)W.ns_new.p; a.partials = (double *)W.partials.p; a.bad = (int *)W.bad.p;
    a.mix_legent = (float *)W.mix_legent.p; a.mix_ap = (float2 *)W.mix_ap.p;
    // the mixed Legendre rows depend on the optical properties only: they are kept between calls on the same arrays
    // unless the caller says the properties changed
    const bool mix_ready = !properties_changed && !mix_realloc && W.mix_npts == d->npts && W.mix_key[0] == d->extinct &&
                           W.mix_key[1] == d->albedo && W.mix_key[2] == d->iphase && W.mix_key[3] == d->legen;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    int tot = 0;
    const int maxiv_i = maxiv > 0x7FFFFFFFll ? 0x7FFFFFFF : (int)maxiv;
    int rc = cs_device_step(a, nblk, W.scan.p, tmpb, shptr_new, (double *)W.sums.p, maxiv_i, (size_t)source_new_capacity, source_new,
                            &tot, mix_ready, errmsg);
    float ms = 0.0f;
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess && !rc) { set_msg(errmsg, "CUDA error %s in at3d_compute_source_device", cudaGetErrorString(e)); rc = 4; }
    else cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (!rc) {
        W.mix_npts = d->npts; W.mix_key[0] = d->extinct; W.mix_key[1] = d->albedo; W.mix_key[2] = d->iphase; W.mix_key[3] = d->legen;
        double hs[4];
        cudaMemcpy(hs, W.sums.p, sizeof(hs), cudaMemcpyDeviceToHost);
        for (int q = 0; q < 4; q++) norms[q] = (float)hs[q];
        *total_new = tot;
    } else W.mix_npts = -1;
    if (kernel_ms) *kernel_ms = ms;
    return rc;
}
