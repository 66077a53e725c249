// at3d_grad.cu -- LEVISAPPROX_GRADIENT, default adjoint ("double sweep") path, on sm_100a.
// Replaces (src/polarized/shdomsub4.f of the AT3D reference)
//   LEVISAPPROX_GRADIENT :288-809, ADJOINT_INTEGRATE_1RAY :3223-3967, COMPUTE_SOURCE_GRAD_1CELL
//   :1546-2042, COMPUTE_SOURCE_DIRECTION :2836-2914, FIND_BOUNDARY_RADIANCE_GRAD :2151-2347,
//   COMPUTE_ADJOINT_WEIGHTS :3969-4034, COMPUTE_RADIANCE_DERIVATIVE_ADJOINT :4037-4114,
//   COMPUTE_DIRECT_BEAM_DERIV_ADJOINT :4117-4143.
//
// Design (DESIGN.md "Gradient kernels"): one octet (8 lanes) per ray, same bit-exact FP64 walk as RENDER.
//  * Everything in COMPUTE_SOURCE_GRAD_1CELL that does not depend on the ray -- the mixed Legendre
//    tables LEGENT/LEGENP/DLEGP and DLEGT(l) per (grid point, property corner, unknown) -- is
//    evaluated once per cost-function evaluation by grad_prep_kernel (at attach time) and kept in HBM.
//  * The reference evaluates the radiance SH contraction COMPUTE_SOURCE_DIRECTION 8*NUMDER times per
//    new grid point.  It is linear in the Legendre table, so the octet forms the per-degree shell sums
//    sum_m RADIANCE(.,j)*YLMDIR(.,j) once (radiance SH read once per (ray, point)) and every
//    (corner, unknown) then costs one dot product of length (ML+1)*{1|4}; lane = property corner.
//  * GRAD8 is contracted with the per-ray adjoint weight at once, so a cell corner keeps 8*NUMDER
//    scalars (a shared-memory row) instead of the reference's five NSTOKES*8*8*NUMDER arrays; rows
//    follow their grid point from cell to cell.
//  * The backward cumulative sum over saved sub-intervals (PASSEDRAD) becomes "total - running", the
//    total coming from the forward pass (same arithmetic); nothing is saved per sub-interval.
//  * Sub-interval weights accumulate in registers (lane n owns corner n); global memory is touched
//    once per cell and corner with red.global.add.f64, never inside the sub-interval loop.
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <cub/cub.cuh>
#include "at3d_host.h"
#include "at3d_ray.cuh"

#define AT3D_GRAD_ROWS 16       // shared-memory rows of contracted GRAD8 per octet (8 live + 8 spare)

// ------------------------------------------------------------------------------------------
// per-octet shared scratch of the adjoint kernel (byte offsets)
// ------------------------------------------------------------------------------------------
struct GradLayout {
    int y, stage, shell, tc, vsh, rowd, rowx, rowib, wacc, total;
};

__host__ __device__ inline GradLayout grad_layout(int nstokes, int ny_comp, int nlmp, int ml, int ncomp,
                                                  int ntup, int numder)
{
    GradLayout L;
    int o = 0;
    L.y = o;      o += ny_comp * nlmp * 4;
    L.stage = o;  o += nstokes * nlmp * 4;                        // RADIANCE*YLMDIR products (scalar) / RADIANCE planes
    L.tc = o;     o += ntup * 8;                                  // double Tc[ntup]
    L.rowd = o;   o += AT3D_GRAD_ROWS * 8 * numder * 8;           // double D[row][nb][numder]
    L.wacc = o;   o += 2 * 8 * 8;                                 // double W[8], G[8]
    L.rowx = o;   o += AT3D_GRAD_ROWS * 8 * numder * 4;           // float  XG[row][nb][numder]
    L.rowib = o;  o += AT3D_GRAD_ROWS * 8 * 4 + 8 * 4;            // int    IB[row][nb], ROWOF[8]
    L.shell = o;  o += (nstokes == 1 ? 1 : 8) * (ml + 1) * 4;     // float shell sums (NPART>1 only)
    L.vsh = o;    o += 3 * (ml + 1) * 4;                          // float V1, V5, V6 (no delta-M only)
    L.total = (o + 15) & ~15;
    return L;
}

// "unscaling" of a delta-M scaled tabulated Legendre entry (shdomsub4.f:1905-1925)
__device__ __forceinline__ float unscale_leg(float x, int k /*0-based component*/, int l, int ml, bool deltam,
                                             bool interp_new, float ftemp, int nstleg)
{
    if (!deltam || l > ml) return x;
    if (k == 0) { if (!interp_new) x = x * (1 - ftemp); return x + ftemp; }
    if (nstleg > 1 && k <= 3) return x + ftemp;
    return x;
}

// compact index of Legendre component k (0-based) among those that reach I,Q,U: 0,1,2,4 -> 0,1,2,3
__device__ __forceinline__ int comp_slot(int k) { return k == 4 ? 3 : k; }

// ------------------------------------------------------------------------------------------
// Ray-independent part of COMPUTE_SOURCE_GRAD_1CELL (shdomsub4.f:1801-1981), once per attach:
// per (grid point, unknown): SCATTERJ, F and -- for NPART>1 -- the mixed table used for SOURCET;
// per (grid point, unknown, property corner): the packed scalars and DLEGT(l).  One warp per point,
// lane = nb + 8*g4 (g4 splits the table entries).
// ------------------------------------------------------------------------------------------
__global__ void grad_prep_kernel(DevState S, DevGrad G, float *grec, float *dlegt, float2 *gpnt, float *legs)
{
    extern __shared__ float prep_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ipz = blockIdx.x * (blockDim.x >> 5) + warp;
    const int nstleg = S.nstleg, ml = S.ml;
    const int nlt = nstleg * (S.nleg + 1);
    if (ipz >= S.npts) return;
    const int ip = ipz + 1;
    float *legent = prep_smem + (size_t)warp * nlt;
    const bool deltam = S.deltam != 0, interp_new = S.interp_new != 0;
    const int nb = lane & 7, g4 = lane >> 3;
    const int ib = __ldg(&G.interpptr[nb + 8 * (size_t)ipz]);
    const float xi = __ldg(&G.optinterpwt[nb + 8 * (size_t)ipz]);
    const int pm = G.pmaxnmicro, nd = G.numder, ncomp = G.ncomp, ntup = G.ntup;
    int last_ipa = -1;
    float scatterj = 0.0f, f = 0.0f;
    for (int idr = 0; idr < nd; idr++) {
        const int ipa = __ldg(&G.partder[idr]);       // 1-based species
        const float albp = __ldg(&G.albedop[(ib - 1) + (size_t)G.maxpg * (ipa - 1)]);
        const float extp = __ldg(&G.extinctp[(ib - 1) + (size_t)G.maxpg * (ipa - 1)]);
        const int *iphp = G.iphasep + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
        const float *pwp = G.phasewtp + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
        const float alb_ip = __ldg(&S.albedo[ipz + (size_t)S.npts * (ipa - 1)]);
        float *legs_row = legs ? legs + ((size_t)ipz * nd + idr) * ntup : nullptr;
        if (ipa != last_ipa) {
            last_ipa = ipa;
            __syncwarp();
            const float sw = xi * albp * extp;       // SPATIAL_WEIGHT of property corner nb
            scatterj = 0.0f;
            if (deltam) for (int n = 0; n < 8; n++) scatterj = scatterj + __shfl_sync(FULLMASK, sw, n);
            if (S.npart == 1) {
                // LEGENT left by the forward part: the PHASEINTERPWT mix at this grid point
                const int *iph = S.iphase + (size_t)S.nq * (ipz + (size_t)S.npts * (ipa - 1));
                const float *pw = S.phaseinterpwt + (size_t)S.nq * (ipz + (size_t)S.npts * (ipa - 1));
                const bool single = (!interp_new) || (__ldg(&pw[0]) >= S.phasemax);
                for (int t = lane; t < nlt; t += 32) {
                    float v;
                    if (single) v = __ldg(&S.legen[(size_t)nlt * (__ldg(&iph[0]) - 1) + t]);
                    else {
                        v = 0.0f;
                        for (int q = 0; q < S.nq; q++) {
                            const float w = __ldg(&pw[q]);
                            if (w <= 1e-5f) continue;
                            v = v + __ldg(&S.legen[(size_t)nlt * (__ldg(&iph[q]) - 1) + t]) * w;
                        }
                    }
                    legent[t] = v;
                }
                __syncwarp();
                if (deltam) f = legent[nstleg * (ml + 1)];
                __syncwarp();
                if (deltam) {
                    for (int t = lane; t < nstleg * (ml + 1); t += 32) {
                        const int k = t % nstleg;
                        float v = legent[t] / (1 - f);
                        v = v * (1 - f);
                        if (k == 0 || (nstleg > 1 && k <= 3)) v = v + f;
                        legent[t] = v;
                    }
                }
            } else {
                // property-grid mix of the Legendre table for this species (shdomsub4.f:1835-1880)
                float swn[8]; int ibn[8];
#pragma unroll
                for (int n = 0; n < 8; n++) { swn[n] = __shfl_sync(FULLMASK, sw, n); ibn[n] = __shfl_sync(FULLMASK, ib, n); }
                for (int t = lane; t < nlt; t += 32) {
                    float v = 0.0f;
#pragma unroll
                    for (int n = 0; n < 8; n++) {
                        if (swn[n] <= 1e-6f) continue;
                        const int *iq = G.iphasep + (size_t)pm * ((ibn[n] - 1) + (size_t)G.maxpg * (ipa - 1));
                        const float *wq = G.phasewtp + (size_t)pm * ((ibn[n] - 1) + (size_t)G.maxpg * (ipa - 1));
                        for (int q = 0; q < pm; q++) {
                            const float w = __ldg(&wq[q]);
                            if (w <= 1e-6f) continue;
                            v = v + swn[n] * w * __ldg(&S.legen[(size_t)nlt * (__ldg(&iq[q]) - 1) + t]);
                        }
                    }
                    if (scatterj > G.scatmin) v = v / scatterj; else v = (float)(v / G.scatmin);
                    legent[t] = v;
                }
                __syncwarp();
                f = 0.0f;
                if (deltam) f = legent[nstleg * (ml + 1)];
                __syncwarp();
                if (deltam && interp_new)
                    for (int t = lane; t < nstleg * (ml + 1); t += 32) legent[t] = legent[t] / (1 - f);
                __syncwarp();
                // the table COMPUTE_SOURCE_DIRECTION contracts with the radiance for SOURCET
                for (int t = lane; t < nstleg * (ml + 1); t += 32) {
                    const int k = t % nstleg, l = t / nstleg;
                    if (k == 3 || k == 5) continue;
                    legs_row[comp_slot(k) + ncomp * l] = legent[t];
                }
                __syncwarp();
                if (deltam) {
                    for (int t = lane; t < nstleg * (ml + 1); t += 32) {
                        const int k = t % nstleg;
                        float v = legent[t] * (1 - f);
                        if (k == 0 || (nstleg > 1 && k <= 3)) v = v + f;
                        legent[t] = v;
                    }
                }
            }
            __syncwarp();
        } else if (legs_row && idr > 0) {
            // same species as the previous unknown: same table
            const float *prev = legs + ((size_t)ipz * nd + idr - 1) * ntup;
            for (int t = lane; t < ntup; t += 32) legs_row[t] = prev[t];
        }
        if (lane == 0) gpnt[(size_t)ipz * nd + idr] = make_float2(scatterj, f);
        // ---- per property corner nb ----
        const float dext_v = __ldg(&G.dext[(ib - 1) + (size_t)G.maxpg * idr]);
        const float dalb_v = __ldg(&G.dalb[(ib - 1) + (size_t)G.maxpg * idr]);
        const float dextm_v = __ldg(&G.dextm[(ib - 1) + (size_t)G.maxpg * idr]);
        const float dalbm_v = __ldg(&G.dalbm[nb + 8 * ((size_t)ipz + (size_t)S.npts * idr)]);
        const float dfj_v = __ldg(&G.dfj[nb + 8 * ((size_t)ipz + (size_t)S.npts * idr)]);
        const int doex = __ldg(&G.doexact[idr]);
        const int *dip = G.diphasep + (size_t)G.deriv_maxnmicro * ((ib - 1) + (size_t)G.maxpg * idr);
        const float *dpw = G.dphasewtp + (size_t)G.deriv_maxnmicro * ((ib - 1) + (size_t)G.maxpg * idr);
        const size_t row = ((size_t)ipz * nd + idr) * 8 + nb;
        if (g4 == 0) {
            float4 *gr = (float4 *)(grec + row * 8);
            gr[0] = make_float4(dext_v, dalb_v, dextm_v, dalbm_v);
            gr[1] = make_float4(dfj_v, albp, extp, alb_ip);
        }
        if (xi >= 1e-7f) {
            float *drow = dlegt + row * ntup;
            for (int t = g4; t < nlt; t += 4) {
                const int k = t % nstleg, l = t / nstleg;
                if (k == 3 || k == 5 || l > ml) continue;       // components that never reach I,Q,U
                float legenp = 0.0f, dlegp = 0.0f;
                for (int q = 0; q < pm; q++) {
                    const float *lg = S.legen + (size_t)nlt * (__ldg(&iphp[q]) - 1);
                    const float ftemp = deltam ? __ldg(&lg[nstleg * (ml + 1)]) : 0.0f;
                    const float un = unscale_leg(__ldg(&lg[t]), k, l, ml, deltam, interp_new, ftemp, nstleg);
                    legenp = legenp + __ldg(&pwp[q]) * un;
                    if (doex == 0 && q < G.deriv_maxnmicro) dlegp = dlegp + __ldg(&dpw[q]) * un;
                }
                if (doex == 1)
                    for (int q = 0; q < G.deriv_maxnmicro; q++)
                        dlegp = dlegp + __ldg(&pwp[q]) * __ldg(&G.dleg[(size_t)nlt * (__ldg(&dip[q]) - 1) + t]);
                const float lt = legent[t];
                const float leg_diff = legenp - lt;
                drow[comp_slot(k) + ncomp * l] = dext_v * leg_diff * albp + dalb_v * leg_diff * extp
                                                 + dlegp * extp * albp + (lt - 1) * dfj_v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// COMPUTE_SOURCE_GRAD_1CELL for one new grid point (the 8 lanes of the octet cooperate): forward
// values of the point and its contracted GRAD8 row (written to shared-memory row `row`).
// ------------------------------------------------------------------------------------------
template <int NST>
__device__ void eval_point_grad(const DevState &S, const DevGrad &G, int ip, unsigned char *sm,
                                const GradLayout &L, const RayDir &rd, const double (&adj)[NST], const Oct &o,
                                int row, float &ext_out, float (&src_out)[NST], float (&ss_out)[NST],
                                int &ns_out, int &nr_out)
{
    const float *Ysh = (const float *)(sm + L.y);
    float *stage = (float *)(sm + L.stage);
    float *shell = (float *)(sm + L.shell);
    double *Tc = (double *)(sm + L.tc);
    const float *Vsh = (const float *)(sm + L.vsh);
    const int nlmp = S.nlmp, nstleg = S.nstleg, ml = S.ml, mm = S.mm;
    const int nlt = nstleg * (S.nleg + 1);
    const int ncomp = G.ncomp, ntup = G.ntup, nd = G.numder;
    const bool deltam = S.deltam != 0, interp_new = S.interp_new != 0;
    const float secmu0 = (float)(1.0 / fabs((double)S.solarmu));
    const int ipz = ip - 1;
    // ---------------- forward part (shdomsub4.f:1660-1785) ----------------
    float ext, a[NST], b[NST];
    eval_point<NST>(S, ip, Ysh, rd, false, o, ext, ns_out, a, b);
    const float dirflux = __ldg(&S.dirflux[ipz]);
    if (!deltam) {
        // without delta-M SINGSCAT8 is the truncated single scattering (shdomsub4.f:1723-1760,1781):
        // sum over species of DA*LEGENT(.,l)*sum_m YLMDIR*YLMSUN for the shells present in SOURCE
        float t[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) t[k] = 0.0f;
        for (int ipa = 0; ipa < S.npart; ipa++) {
            float w;
            if (ext == 0.0f) w = 1.0f; else w = __ldg(&S.extinct[ipz + (size_t)S.npts * ipa]) / ext;
            if (w == 0.0f) continue;
            const int *iph = S.iphase + (size_t)S.nq * (ipz + (size_t)S.npts * ipa);
            const float *pw = S.phaseinterpwt + (size_t)S.nq * (ipz + (size_t)S.npts * ipa);
            const bool single = (!interp_new) || (__ldg(&pw[0]) >= S.phasemax);
            const float da = __ldg(&S.albedo[ipz + (size_t)S.npts * ipa]) * dirflux * secmu0 * w;
            for (int l = o.ol; l <= ml; l += 8) {
                const int me = l < mm ? l : mm;
                if (sh_index(l, -me, mm) >= ns_out) continue;
                float l1, l5 = 0.0f;
                if (single) {
                    l1 = __ldg(&S.legen[(size_t)nlt * (__ldg(&iph[0]) - 1) + nstleg * l]);
                    if (nstleg > 1) l5 = __ldg(&S.legen[(size_t)nlt * (__ldg(&iph[0]) - 1) + nstleg * l + 4]);
                } else {
                    l1 = 0.0f;
                    for (int q = 0; q < S.nq; q++) {
                        const float wq = __ldg(&pw[q]);
                        if (wq <= 1e-5f) continue;
                        l1 = l1 + __ldg(&S.legen[(size_t)nlt * (__ldg(&iph[q]) - 1) + nstleg * l]) * wq;
                        if (nstleg > 1) l5 = l5 + __ldg(&S.legen[(size_t)nlt * (__ldg(&iph[q]) - 1) + nstleg * l + 4]) * wq;
                    }
                }
                t[0] = t[0] + da * l1 * Vsh[l];
                if (NST > 1) {
                    t[1] = t[1] + da * l5 * Vsh[(ml + 1) + l];
                    t[NST - 1] = t[NST - 1] + da * l5 * Vsh[2 * (ml + 1) + l];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NST; k++) b[k] = oct_sum(o.m, t[k]);
    }
    float srcfull[NST];      // SRCEXT8 before the multiplication by the extinction
#pragma unroll
    for (int k = 0; k < NST; k++) {
        srcfull[k] = deltam ? a[k] + b[k] : a[k];
        ss_out[k] = b[k] * ext;
        src_out[k] = G.singlescatter ? ss_out[k] : srcfull[k] * ext;
    }
    ext_out = ext;
    // ---------------- radiance SH block: shell sums over m ----------------
    const int2 rr = __ldg(&S.radrec[ipz]);
    const int rns = rr.y, nrp = AT3D_SHPAD(rr.y);
    nr_out = rns;
    {
        const float *rb = S.shrad + rr.x + o.ol * 4;
        if (NST == 1) {
#pragma unroll 4
            for (int j = 0; j < nrp; j += 32) {
                const float4 r = __ldg((const float4 *)(rb + j));
                const float4 y = *(const float4 *)(Ysh + o.ol * 4 + j);
                *(float4 *)(stage + o.ol * 4 + j) = make_float4(r.x * y.x, r.y * y.y, r.z * y.z, r.w * y.w);
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < nrp; j += 32) {
#pragma unroll
                for (int k = 0; k < NST; k++)
                    *(float4 *)(stage + k * nlmp + o.ol * 4 + j) = __ldg((const float4 *)(rb + k * nrp + j));
            }
        }
    }
    for (int t = o.ol; t < ntup; t += 8) Tc[t] = 0.0;
    __syncwarp(o.m);
    // T(l) = adj . sum_m RADIANCE(.,j) YLMDIR(.,j); degrees are paired (l, ML-l) to balance the lanes
    const bool keep_shell = S.npart > 1;
    for (int q = o.ol; 2 * q <= ml; q += 8) {
        for (int half = 0; half < 2; half++) {
            const int l = half ? ml - q : q;
            if (half && l == q) break;
            const int me = l < mm ? l : mm;
            const int jlo = sh_index(l, -me, mm);
            int cnt = 2 * me + 1;
            if (jlo + cnt > rns) cnt = rns - jlo;
            if (NST == 1) {
                float A = 0.0f;
                for (int i = 0; i < cnt; i++) A = A + stage[jlo + i];
                double t1 = adj[0] * A;
                if (!deltam) t1 += adj[0] * (double)(dirflux * secmu0 * Vsh[l]);
                Tc[l] = t1;
                if (keep_shell) shell[l] = A;
            } else {
                float A = 0, B = 0, C = 0, D = 0, E = 0, F = 0, Gq = 0, H = 0;
                for (int i = 0; i < cnt; i++) {
                    const int j = jlo + i;
                    const float r1 = stage[j], y1 = Ysh[j];
                    const float r2 = stage[nlmp + j], r3 = stage[2 * nlmp + j];
                    const float y2 = Ysh[nlmp + j], y5 = Ysh[2 * nlmp + j], y6 = Ysh[3 * nlmp + j], y3 = Ysh[4 * nlmp + j];
                    A = A + r1 * y1;
                    B = B + r2 * y1; C = C + r1 * y2; D = D + r2 * y2; E = E + r3 * y5;
                    F = F + r1 * y6; Gq = Gq + r2 * y6; H = H + r3 * y3;
                }
                double t1 = adj[0] * A;
                if (!deltam) t1 += adj[0] * (double)(dirflux * secmu0 * Vsh[l]);
                double t5 = adj[0] * B + adj[1] * C + adj[NST - 1] * F;
                if (!deltam) t5 += adj[1] * (double)(dirflux * secmu0 * Vsh[(ml + 1) + l]);
                Tc[0 + ncomp * l] = t1;
                Tc[1 + ncomp * l] = adj[1] * D + adj[NST - 1] * Gq;
                Tc[2 + ncomp * l] = adj[1] * E + adj[NST - 1] * H;
                Tc[3 + ncomp * l] = t5;
                if (keep_shell) {
                    const int s = ml + 1;
                    shell[l] = A; shell[s + l] = B; shell[2 * s + l] = C; shell[3 * s + l] = D;
                    shell[4 * s + l] = E; shell[5 * s + l] = F; shell[6 * s + l] = Gq; shell[7 * s + l] = H;
                }
            }
        }
    }
    __syncwarp(o.m);
    // ---------------- gradient part (shdomsub4.f:1786-2019): lane = property corner nb ----------------
    const int nb = o.ol;
    const int ib = __ldg(&G.interpptr[nb + 8 * (size_t)ipz]);
    const float xi = __ldg(&G.optinterpwt[nb + 8 * (size_t)ipz]);
    double *Drow = (double *)(sm + L.rowd) + (size_t)row * 8 * nd;
    float *XGrow = (float *)(sm + L.rowx) + (size_t)row * 8 * nd;
    int *IBrow = (int *)(sm + L.rowib) + row * 8;
    IBrow[nb] = ib;
    int last_ipa = -1;
    float scatterj = 0.0f, f = 0.0f, singscatj[NST], sourcet[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { singscatj[k] = 0.0f; sourcet[k] = 0.0f; }
    const int pm = G.pmaxnmicro;
    for (int idr = 0; idr < nd; idr++) {
        const int ipa = __ldg(&G.partder[idr]);       // 1-based species
        const size_t prow = ((size_t)ipz * nd + idr) * 8 + nb;
        const float4 g0 = __ldg((const float4 *)(G.grec + prow * 8));
        const float4 g1 = __ldg((const float4 *)(G.grec + prow * 8) + 1);
        const float dext_v = g0.x, dalb_v = g0.y, dextm_v = g0.z, dalbm_v = g0.w;
        const float dfj_v = g1.x, albp = g1.y, extp = g1.z, alb_ip = g1.w;
        const int *iphp = G.iphasep + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
        const float *pwp = G.phasewtp + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
        if (ipa != last_ipa) {
            last_ipa = ipa;
            const float2 pn = __ldg(&G.gpnt[(size_t)ipz * nd + idr]);
            scatterj = pn.x; f = pn.y;
            const float sw = xi * albp * extp;       // SPATIAL_WEIGHT of property corner nb
#pragma unroll
            for (int k = 0; k < NST; k++) singscatj[k] = 0.0f;
            if (deltam) {
                float part[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) part[k] = 0.0f;
                if (sw > 1e-6f) {
                    for (int q = 0; q < pm; q++) {
                        const float w = __ldg(&pwp[q]);
                        if (w <= 1e-6f) continue;
                        float sv[NST];
                        ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, __ldg(&iphp[q]), rd, sv);
#pragma unroll
                        for (int k = 0; k < NST; k++) part[k] = part[k] + sw * w * sv[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    float tot = 0.0f;
#pragma unroll
                    for (int n = 0; n < 8; n++) tot = tot + __shfl_sync(o.m, part[k], n, 8);
                    if (scatterj > G.scatmin) singscatj[k] = tot / scatterj;
                    else singscatj[k] = (float)(tot / G.scatmin);
                }
            }
            if (S.npart == 1) {
#pragma unroll
                for (int k = 0; k < NST; k++) sourcet[k] = (alb_ip > 1e-8f) ? srcfull[k] / alb_ip : 0.0f;
            } else {
#pragma unroll
                for (int k = 0; k < NST; k++) sourcet[k] = 0.0f;
                if (scatterj > G.scatmin) {
                    // COMPUTE_SOURCE_DIRECTION with the mixed table, from the shell sums
                    const float *lg = G.legs + ((size_t)ipz * nd + idr) * ntup;
                    float acc[NST];
#pragma unroll
                    for (int k = 0; k < NST; k++) acc[k] = 0.0f;
                    const int s = ml + 1;
                    for (int l = o.ol; l <= ml; l += 8) {
                        const float l1 = __ldg(&lg[ncomp * l]);
                        acc[0] = acc[0] + l1 * shell[l];
                        if (!deltam) acc[0] = acc[0] + dirflux * secmu0 * l1 * Vsh[l];
                        if (NST > 1) {
                            const float l2 = __ldg(&lg[1 + ncomp * l]), l3 = __ldg(&lg[2 + ncomp * l]);
                            const float l5 = __ldg(&lg[3 + ncomp * l]);
                            acc[0] = acc[0] + l5 * shell[s + l];
                            acc[1] = acc[1] + l5 * shell[2 * s + l] + l2 * shell[3 * s + l] + l3 * shell[4 * s + l];
                            acc[NST - 1] = acc[NST - 1] + l5 * shell[5 * s + l] + l2 * shell[6 * s + l]
                                           + l3 * shell[7 * s + l];
                            if (!deltam) acc[1] = acc[1] + dirflux * secmu0 * l5 * Vsh[s + l];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NST; k++) sourcet[k] = oct_sum(o.m, acc[k]);
                    if (deltam) {
#pragma unroll
                        for (int k = 0; k < NST; k++) sourcet[k] = sourcet[k] + dirflux * singscatj[k] * secmu0 / (1 - f);
                    }
                }
            }
            sourcet[0] = fmaxf(0.0f, sourcet[0]);
        }
        double d = 0.0;
        if (xi >= 1e-7f) {
            // DSOURCE contracted with the adjoint weight: DLEGT(l) . T(l)
            const float4 *dl = (const float4 *)(G.dlegt + prow * ntup);
            double dot = 0.0;
            for (int t4 = 0; t4 < ntup / 4; t4++) {
                const float4 v = __ldg(&dl[t4]);
                dot += (double)v.x * Tc[4 * t4];
                dot += (double)v.y * Tc[4 * t4 + 1];
                dot += (double)v.z * Tc[4 * t4 + 2];
                dot += (double)v.w * Tc[4 * t4 + 3];
            }
            const int doex = __ldg(&G.doexact[idr]);
            const int *dip = G.diphasep + (size_t)G.deriv_maxnmicro * ((ib - 1) + (size_t)G.maxpg * idr);
            const float *dpw = G.dphasewtp + (size_t)G.deriv_maxnmicro * ((ib - 1) + (size_t)G.maxpg * idr);
            float singscatp[NST], dsingscatp[NST];
#pragma unroll
            for (int k = 0; k < NST; k++) { singscatp[k] = 0.0f; dsingscatp[k] = 0.0f; }
            if (deltam) {
                for (int q = 0; q < pm; q++) {
                    float sv[NST];
                    ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, __ldg(&iphp[q]), rd, sv);
                    const float w = __ldg(&pwp[q]);
#pragma unroll
                    for (int k = 0; k < NST; k++) singscatp[k] = singscatp[k] + w * sv[k];
                    if (doex == 0 && q < G.deriv_maxnmicro) {
                        const float dw = __ldg(&dpw[q]);
#pragma unroll
                        for (int k = 0; k < NST; k++) dsingscatp[k] = dsingscatp[k] + dw * sv[k];
                    }
                }
                if (doex == 1) {
                    for (int q = 0; q < G.deriv_maxnmicro; q++) {
                        float sv[NST];
                        ray_singscat<NST>(G.dphasetab, S.nstphase, G.dnumphase, __ldg(&dip[q]), rd, sv);
                        const float w = __ldg(&pwp[q]);
#pragma unroll
                        for (int k = 0; k < NST; k++) dsingscatp[k] = dsingscatp[k] + w * sv[k];
                    }
                }
            }
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < NST; k++) {
                float g8 = xi * (sourcet[k] * (alb_ip * dextm_v + dalbm_v));
                if (deltam)
                    g8 = g8 + dirflux * secmu0 * xi * (singscatj[k] * dfj_v + dsingscatp[k] * extp * albp
                              + dalb_v * (singscatp[k] - singscatj[k]) * extp
                              + dext_v * (singscatp[k] - singscatj[k]) * albp);
                sum += adj[k] * (double)g8;
            }
            d = sum + (double)xi * dot;
        }
        Drow[nb * nd + idr] = d;
        XGrow[nb * nd + idr] = dextm_v * xi;
    }
    __syncwarp(o.m);
}

// Corner refresh for the adjoint walk: values and the shared-memory row (contracted GRAD8) of points
// shared with the previous cell are carried over (OLDIPTS/DONEFACE logic of shdomsub4.f:1645-1659).
template <int NST>
__device__ __forceinline__ void refresh_corners_grad(const DevState &S, const DevGrad &G, const CellRec &c,
                                                     unsigned char *sm, const GradLayout &L, const RayDir &rd,
                                                     const double (&adj)[NST], bool first, const Oct &o,
                                                     int &cpt, int &crow, float &cext, float (&csrc)[NST],
                                                     float (&css)[NST], int &npt_eval, int &nsh_eval, int &nrh_eval)
{
    const int myp = own_corner(c, o.ol);
    int hit = -1;
    if (!first) {
#pragma unroll
        for (int k = 0; k < 8; k++) { const int pk = __shfl_sync(o.m, cpt, k, 8); if (pk == myp) hit = k; }
    }
    const int from = hit < 0 ? o.ol : hit;
    cext = __shfl_sync(o.m, cext, from, 8);
    crow = __shfl_sync(o.m, crow, from, 8);
#pragma unroll
    for (int k = 0; k < NST; k++) {
        csrc[k] = __shfl_sync(o.m, csrc[k], from, 8);
        css[k] = __shfl_sync(o.m, css[k], from, 8);
    }
    cpt = myp;
    unsigned used = oct_or(o, hit >= 0 ? (1u << crow) : 0u);
    unsigned need = oct_ballot(o, hit < 0);
    while (need) {
        const int n = __ffs(need) - 1;
        const int ip = __shfl_sync(o.m, myp, n, 8);
        const int row = __ffs(~used) - 1;
        used |= 1u << row;
        float ext, src[NST], ss[NST];
        int ns, nr;
        eval_point_grad<NST>(S, G, ip, sm, L, rd, adj, o, row, ext, src, ss, ns, nr);
        npt_eval++; nsh_eval += ns; nrh_eval += nr;
        const bool mine = (myp == ip);
        if (mine) {
            cext = ext; crow = row;
#pragma unroll
            for (int k = 0; k < NST; k++) { csrc[k] = src[k]; css[k] = ss[k]; }
        }
        need &= ~oct_ballot(o, mine);
    }
}

// ADJOINT_INTEGRATE_1RAY for one ray (one octet).
template <int NST>
__device__ int march_ray_adjoint(const DevState &S, const DevGrad &G, unsigned char *sm, const GradLayout &L,
                                 const RayDir &rd, double mu2, double x0, double y0, double z0, float sky,
                                 const double (&adj)[NST], const double (&total)[NST], const Oct &o,
                                 double *gradout, double *beam_weight,
                                 int *trace_cells, int trace_cap, int &ntrace, int &nsub)
{
    const double *Dall = (const double *)(sm + L.rowd);
    const float *XGall = (const float *)(sm + L.rowx);
    const int *IBall = (const int *)(sm + L.rowib);
    int *rowof = (int *)(sm + L.rowib) + AT3D_GRAD_ROWS * 8;
    double *Wacc = (double *)(sm + L.wacc);
    const int nd = G.numder;
    double xe = x0, ye = y0, ze = z0, transmit = 1.0;
    double radout[NST];
    float ext1 = 0.0f, srcext1[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { radout[k] = 0.0; srcext1[k] = 0.0f; }
    const int p1c = cell_gp(S, 1, 1), p8c = cell_gp(S, 1, 8);
    const double eps = (double)(1.0e-5f * (pt_coord(S, p8c, 3) - pt_coord(S, p1c, 3)));
    const bool exact_ss = G.exact_single_scatter != 0;
    int icell = dev_locate_grid_cell(S, xe, ye, ze);
    int iface = 0, npassed = 1;
    bool done = false, first = true;
    int npt_eval = 0, nsh_eval = 0, nrh_eval = 0;
    int cpt = 0, crow = 0; float cext = 0.0f, csrc[NST], css[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { csrc[k] = 0.0f; css[k] = 0.0f; }
    ntrace = 0; nsub = 0;
    while (!done && icell > 0) {
        if (trace_cells && o.ol == 0 && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        const CellRec c = load_cell(S, icell);
        refresh_corners_grad<NST>(S, G, c, sm, L, rd, adj, first, o, cpt, crow, cext, csrc, css,
                                  npt_eval, nsh_eval, nrh_eval);
        first = false;
        float e8[8], s8[NST][8];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            e8[n] = __shfl_sync(o.m, cext, n, 8);
#pragma unroll
            for (int k = 0; k < NST; k++) s8[k][n] = __shfl_sync(o.m, csrc[k], n, 8);
        }
        const float4 q1 = __ldg(&S.ptrec[c.gp[0] - 1]);
        const float4 q8 = __ldg(&S.ptrec[c.gp[7] - 1]);
        const double delx = (double)(q8.x - q1.x), dely = (double)(q8.y - q1.y), delz = (double)(q8.z - q1.z);
        const double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        const double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        const double invdelz = 1.0 / delz;
        double u = (xe - q1.x) * invdelx, v = (ye - q1.y) * invdely, w = (ze - q1.z) * invdelz;
        double fc[8];
        interp_kernel(u, v, w, fc);
        double fown = interp_kernel_own(u, v, w, o.ol);
#pragma unroll
        for (int k = 0; k < NST; k++) srcext1[k] = (float)fcsum(fc, s8[k]);
        srcext1[0] = fmaxf(0.0f, srcext1[0]);
        double ext1d = fcsum(fc, e8);
        ext1 = (float)ext1d;
        const bool ipinx = DBTEST(c.flags, 0) &&
            !(DBTEST(S.bcflag, 0) && ((rd.cx > 0 && xe < rd.xm) || (rd.cx < 0 && xe > rd.xm)));
        const bool ipiny = DBTEST(c.flags, 1) &&
            !(DBTEST(S.bcflag, 1) && ((rd.cy > 0 && ye < rd.ym) || (rd.cy < 0 && ye > rd.ym)));
        int iopp = c.gp[0];
#pragma unroll
        for (int n = 1; n < 8; n++) if (8 - rd.ioct == n) iopp = c.gp[n];
        const float4 qo = __ldg(&S.ptrec[iopp - 1]);
        const double sox = ipinx ? (double)1.0e20f : (qo.x - xe) * rd.cxinv;
        const double soy = ipiny ? (double)1.0e20f : (qo.y - ye) * rd.cyinv;
        const double soz = (qo.z - ze) * rd.czinv;
        const double so = fmin(fmin(sox, soy), soz);
        if (so < -eps) return 1;
        double xn = xe + so * rd.cx, yn = ye + so * rd.cy, zn = ze + so * rd.cz;
        u = (xn - q1.x) * invdelx; v = (yn - q1.y) * invdely; w = (zn - q1.z) * invdelz;
        float extn;
        { double fcn[8]; interp_kernel(u, v, w, fcn); extn = (float)fcsum(fcn, e8); }
        const double taugrid = so * 0.5f * (ext1 + extn);
        int ntau = 1 + (int)(taugrid / S.tautol);
        if (ntau < 1) ntau = 1;
        const double dels = so / ntau;
        // per-corner accumulators of this cell (lane n owns corner n)
        double Wn = 0.0, Gn = 0.0;
        float bw[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) bw[k] = 0.0f;
        for (int it = 1; it <= ntau; it++) {
            const double f1 = fown;                  // previous interpolation weight of the own corner
            const double s = it * dels;
            const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
            u = (xi - q1.x) * invdelx; v = (yi - q1.y) * invdely; w = (zi - q1.z) * invdelz;
            interp_kernel(u, v, w, fc);
            fown = interp_kernel_own(u, v, w, o.ol);
            const double f0 = fown;
            float ext0, srcext0[NST];
#pragma unroll
            for (int k = 0; k < NST; k++) srcext0[k] = (float)fcsum(fc, s8[k]);
            const double ext0d = fcsum(fc, e8);
            ext0 = (it != ntau) ? (float)ext0d : extn;
            srcext0[0] = fmaxf(0.0f, srcext0[0]);
            const double ext = (double)(0.5f * (ext0 + ext1));
            if (ext != 0.0) {
                const double tau = ext * dels;
                const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                const double transcell = 1.0f - abscell;
                const double corr = dels * (1.0f - 0.05f * (ext1 - ext0) * dels);
                double rcur = 0.0, rnext = 0.0;      // adj . PASSEDRAD(kk), adj . PASSEDRAD(kk+1)
#pragma unroll
                for (int k = 0; k < NST; k++) rcur += adj[k] * (total[k] - radout[k]);
                rcur = rcur / transmit;
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    const double src = (0.5f * (srcext0[k] + srcext1[k])
                        + 0.08333333333f * (ext0 * srcext1[k] - ext1 * srcext0[k]) * dels
                          * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext;
                    radout[k] = radout[k] + transmit * src * abscell;
                }
                const double tnext = transmit * transcell;
#pragma unroll
                for (int k = 0; k < NST; k++) rnext += adj[k] * (total[k] - radout[k]);
                rnext = rnext / tnext;
                // lane-private corner weights
                Wn += transmit * abscell * ((0.5f * (f0 + f1) + 0.08333333333f * (ext0 * f1 - ext1 * f0) * corr) / ext);
                if (exact_ss) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        float ss0 = (float)(f0 * css[k]), ss1 = (float)(f1 * css[k]);
                        if (k == 0) { ss0 = fmaxf(0.0f, ss0); ss1 = fmaxf(0.0f, ss1); }
                        bw[k] = (float)(bw[k] + transmit * abscell *
                                (0.5f * (ss0 + ss1) + 0.08333333333f * (ext0 * ss1 - ext1 * ss0) * corr) / ext);
                    }
                }
                // radiance term (COMPUTE_RADIANCE_DERIVATIVE_ADJOINT): extinctions re-interpolated in double
                const double aext = 0.5f * (ext0d + ext1d);
                if (aext != 0.0) {
                    const double g0 = -rnext * f0, g1 = -rcur * f1;
                    const double ag = (0.5f * (g0 + g1) + 0.08333333333f * (ext0d * g1 - ext1d * g0) * dels
                                       * (1.0f - 0.05f * (ext1d - ext0d) * dels)) / aext;
                    Gn += ag * transmit * abscell;
                }
                transmit = tnext;
                npassed++;
                nsub++;
                if (npassed > G.maxsub) return 4;
            } else {
#pragma unroll
                for (int k = 0; k < NST; k++) bw[k] = 0.0f;
            }
            ext1 = ext0; ext1d = ext0d;
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1[k] = srcext0[k];
        }
        // ---- flush this cell's contributions (once per cell and corner, no atomics above) ----
        Wacc[o.ol] = Wn; Wacc[8 + o.ol] = Gn; rowof[o.ol] = crow;
        if (exact_ss) {
            double bsum = 0.0;
#pragma unroll
            for (int k = 0; k < NST; k++) bsum += adj[k] * (double)bw[k];
            if (bsum != 0.0) atomicAdd(&beam_weight[cpt - 1], bsum);
        }
        __syncwarp(o.m);
        {
            const int nb = o.ol;
            for (int slot = 0; slot < 8; slot++) {
                const int r = rowof[slot];
                const double Ws = Wacc[slot], Gs = Wacc[8 + slot];
                const int ibp = IBall[r * 8 + nb];
                for (int idr = 0; idr < nd; idr++) {
                    const int e = (r * 8 + nb) * nd + idr;
                    const double val = Ws * Dall[e] + (double)XGall[e] * Gs;
                    if (val != 0.0) atomicAdd(&gradout[(size_t)(ibp - 1) + (size_t)G.maxpg * idr], val);
                }
            }
        }
        __syncwarp(o.m);
        int jface;
        bool openbcface;
        if (sox <= soz && sox <= soy) { iface = 2 - rd.bitx; jface = 1; openbcface = DBTEST(c.flags, 0) && DBTEST(S.bcflag, 0); }
        else if (soy <= soz) { iface = 4 - rd.bity; jface = 2; openbcface = DBTEST(c.flags, 1) && DBTEST(S.bcflag, 1); }
        else { iface = 6 - rd.bitz; jface = 3; openbcface = false; }
        int nbr = c.nb[0];
#pragma unroll
        for (int n = 1; n < 6; n++) if (iface - 1 == n) nbr = c.nb[n];
        int inextcell = nbr;
        if (inextcell < 0) inextcell = dev_next_cell(S, xn, yn, zn, iface, jface, inextcell);
        int kface, ic;
        if (nbr >= 0 && !openbcface) { kface = iface; ic = icell; }
        else { kface = ((iface - 1) ^ 1) + 1; ic = inextcell; iface = 0; }
        if (inextcell > 0) {
            const int pn = cell_gp(S, inextcell, rd.ioct);
            if (jface == 1) xn = (double)pt_coord(S, pn, 1);
            else if (jface == 2) yn = (double)pt_coord(S, pn, 2);
            else zn = (double)pt_coord(S, pn, 3);
        }
        if (transmit < S.transcut) {
            done = true;
        } else if (inextcell == 0 && iface >= 5) {
            done = true;
            float radbnd[NST];
            int boundpts[4]; double boundinterp[4], dirrad1[4];
            const int e = boundary_radiance<NST, true>(S, xn, yn, (float)mu2, sky, ic, kface, radbnd,
                                                       boundpts, boundinterp, dirrad1);
            if (e) return e;
            if (exact_ss && o.ol < 4) {
                int bp = boundpts[0]; double bi = boundinterp[0], dr = dirrad1[0];
#pragma unroll
                for (int n = 1; n < 4; n++) if (o.ol == n) { bp = boundpts[n]; bi = boundinterp[n]; dr = dirrad1[n]; }
                const double val = adj[0] * transmit * bi * dr;
                if (val != 0.0) atomicAdd(&beam_weight[bp - 1], val);
            }
        } else {
            icell = inextcell;
        }
        xe = xn; ye = yn; ze = zn;
    }
    if (S.counts && o.ol == 0) {
        atomicAdd(&S.counts[0], (unsigned long long)ntrace);
        atomicAdd(&S.counts[1], (unsigned long long)npt_eval);
        atomicAdd(&S.counts[2], (unsigned long long)nsh_eval);
        atomicAdd(&S.counts[3], (unsigned long long)nrh_eval);
        atomicAdd(&S.counts[4], (unsigned long long)nsub);
        atomicAdd(&S.counts[5], 1ull);
    }
    return 0;
}

template <int NST>
__global__ void __launch_bounds__(AT3D_RAY_THREADS)
adjoint_kernel(DevState S, DevGrad G, GradLayout L, int nrays, const float *camx, const float *camy,
               const float *camz, const double *cammu, const double *camphi, const RayPack *packs,
               const int *raypix, const double *adjw /*[NST,npix]*/, const double *ray_weights,
               const double *stokes_weights, const double *total /*[NST,nrays]*/,
               double *gradout, double *beam_weight,
               int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub, RayErr *err, int *ray_counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Oct o = oct_id();
    const int lane = threadIdx.x & 31;
    unsigned char *sm = smem_raw + (size_t)(threadIdx.x >> 3) * L.total;
    float *Ysh = (float *)(sm + L.y);
    float *Vsh = (float *)(sm + L.vsh);
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ray_counter, 32 / AT3D_OCT);
        base = __shfl_sync(FULLMASK, base, 0);
        if (base >= nrays) break;
        const int iray = base + (lane >> 3);
        if (iray < nrays) {
            const double mu2 = __ldg(&cammu[iray]), phi2 = __ldg(&camphi[iray]);
            const RayPack pk = dev_get_pack(S, packs, iray, camx, camy, camz, mu2, phi2);
            int ntrace = 0, nsub = 0;
            if (pk.status == 2) { if (o.ol == 0) set_err(err, 2, iray); }
            else if (pk.status == 0) {
                const int pix = __ldg(&raypix[iray]);
                double adj[NST], tot[NST];
                const double rw = __ldg(&ray_weights[iray]);
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    adj[k] = __ldg(&adjw[k + NST * (size_t)pix]) * rw * __ldg(&stokes_weights[k + NST * (size_t)pix]);
                    tot[k] = __ldg(&total[k + NST * (size_t)iray]);
                }
                RayDir rd;
                dev_ray_dir(S, pk, rd);
                __syncwarp(o.m);
                group_ylmall(S, (float)mu2, (float)phi2, Ysh, o.ol, AT3D_OCT, o.m);
                if (!S.deltam) {
                    // shell sums of YLMSUN*YLMDIR for the untruncated solar term of COMPUTE_SOURCE_DIRECTION
                    for (int l = o.ol; l <= S.ml; l += 8) {
                        const int me = l < S.mm ? l : S.mm;
                        const int jlo = sh_index(l, -me, S.mm);
                        float v1 = 0.0f, v5 = 0.0f, v6 = 0.0f;
                        for (int i = 0; i < 2 * me + 1; i++) {
                            const float ys = __ldg(&S.ylmsun[(size_t)S.nstleg * (jlo + i)]);
                            v1 = v1 + ys * Ysh[jlo + i];
                            if (NST > 1) { v5 = v5 + ys * Ysh[S.nlmp + jlo + i]; v6 = v6 + ys * Ysh[3 * S.nlmp + jlo + i]; }
                        }
                        Vsh[l] = v1; Vsh[(S.ml + 1) + l] = v5; Vsh[2 * (S.ml + 1) + l] = v6;
                    }
                    __syncwarp(o.m);
                }
                const float sky = (-mu2 > 0.0) ? dev_sky_radiance(S, (float)mu2, (float)phi2) : 0.0f;
                const int e = march_ray_adjoint<NST>(S, G, sm, L, rd, mu2, pk.x0, pk.y0, pk.z0, sky, adj, tot, o,
                                                     gradout, beam_weight,
                                                     trace_cells ? trace_cells + (size_t)trace_cap * iray : nullptr,
                                                     trace_cap, ntrace, nsub);
                if (e && o.ol == 0) set_err(err, e, iray);
            }
            if (o.ol == 0 && trace_n) { trace_n[iray] = ntrace; trace_nsub[iray] = nsub; }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Phase 1 tail / Phase 2: ray -> pixel accumulation (shdomsub4.f:693-696), COMPUTE_ADJOINT_WEIGHTS
// ------------------------------------------------------------------------------------------
template <int NST>
__global__ void pixel_kernel(int npix, const int *pixstart, const int *rays_per_pixel, const double *visrad,
                             const double *ray_weights, const double *stokes_weights, const float *measurements,
                             const double *unc, int nunc, int costfunc_ll, float *stokesout, double *adjw,
                             double *costp, int *raypix)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int r0 = pixstart[p], n = rays_per_pixel[p];
    float so[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) so[k] = 0.0f;
    for (int r = r0; r < r0 + n; r++) {
        raypix[r] = p;
        const double rw = ray_weights[r];
#pragma unroll
        for (int k = 0; k < NST; k++)
            so[k] = (float)(so[k] + visrad[k + NST * (size_t)r] * rw * stokes_weights[k + NST * (size_t)p]);
    }
    double s[NST], m[NST], aw[NST], cost = 0.0;
#pragma unroll
    for (int k = 0; k < NST; k++) {
        stokesout[k + NST * (size_t)p] = so[k];
        s[k] = (double)so[k]; m[k] = (double)measurements[k + NST * (size_t)p]; aw[k] = 0.0;
    }
    const double *U = unc + (size_t)nunc * nunc * p;
#define UNC(a, b) U[((a) - 1) + nunc * ((b) - 1)]
    if (!costfunc_ll) {
        for (int i = 1; i <= NST; i++) {
            const double pe = s[i - 1] - m[i - 1];
            for (int j = 1; j <= NST; j++) {
                cost = cost + 0.5 * UNC(i, j) * (pe * pe);
                aw[i - 1] = aw[i - 1] + UNC(i, j) * pe;
            }
        }
    } else {
        const double raderror = log(s[0]) - log(m[0]);
        cost = cost + 0.5 * (raderror * raderror * UNC(1, 1));
        aw[0] = aw[0] + raderror * UNC(1, 1) / s[0];
        if (NST > 1) {
            const double dolp1 = sqrt(s[1] * s[1] + s[NST - 1] * s[NST - 1]) / s[0];
            const double dolp2 = sqrt(m[1] * m[1] + m[NST - 1] * m[NST - 1]) / m[0];
            const double dolperr = log(dolp1) - log(dolp2);
            cost = cost + 0.5 * (dolperr * dolperr * UNC(2, 2));
            aw[1] = aw[1] + dolperr * UNC(2, 2) * s[1] / (s[1] * s[1] + s[NST - 1] * s[NST - 1]);
            aw[NST - 1] = aw[NST - 1] + dolperr * UNC(2, 2) * s[NST - 1] / (s[1] * s[1] + s[NST - 1] * s[NST - 1]);
        }
    }
#undef UNC
#pragma unroll
    for (int k = 0; k < NST; k++) adjw[k + NST * (size_t)p] = aw[k];
    costp[p] = cost;
}

// deterministic sum of the per-pixel costs (one block, fixed tree)
__global__ void cost_reduce_kernel(int n, const double *costp, double *cost)
{
    __shared__ double sh[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += costp[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) cost[0] = sh[0];
}

// Phase 4: COMPUTE_DIRECT_BEAM_DERIV_ADJOINT (shdomsub4.f:792-802,4117-4143): one warp per grid point
// with a non-zero beam weight walks its zero-terminated DPTR/DPATH list.
__global__ void beam_kernel(DevGrad G, int npts, const double *beam_weight, double *gradout)
{
    const int ip = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ip >= npts) return;
    const double bwt = beam_weight[ip];
    if (bwt == 0.0) return;
    const float *dpath = G.dpath + (size_t)G.longest_path_pts * ip;
    const int *dptr = G.dptr + (size_t)G.longest_path_pts * ip;
    for (int base = 0; base < G.longest_path_pts; base += 32) {
        const int ii = base + lane;
        const int ib = ii < G.longest_path_pts ? __ldg(&dptr[ii]) : 0;
        // the list ends at the first entry <= 0
        const unsigned stop = __ballot_sync(FULLMASK, ib <= 0);
        const int nvalid = stop ? __ffs(stop) - 1 : 32;
        if (lane < nvalid) {
            const double pb = (double)__ldg(&dpath[ii]) * bwt;
            for (int idr = 0; idr < G.numder; idr++) {
                const float dm = __ldg(&G.dextm[(ib - 1) + (size_t)G.maxpg * idr]);
                const double val = dm * pb;
                if (val != 0.0) atomicAdd(&gradout[(ib - 1) + (size_t)G.maxpg * idr], -val);
            }
        }
        if (stop) break;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static void set_msg(char *errmsg, const char *fmt, ...)
{
    if (!errmsg) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, AT3D_ERRMSG_LEN, fmt, ap);
    va_end(ap);
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_msg(errmsg, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__,   \
                    __LINE__, #expr);                                                          \
            return 4;                                                                          \
        }                                                                                      \
    } while (0)

int check_ray_err(at3d_state *st, cudaStream_t stream, char *errmsg);
int stage_rays(at3d_state *st, const at3d_rays *rays, cudaStream_t stream, const float **camx,
               const float **camy, const float **camz, const double **cammu, const double **camphi,
               const RayPack **packs, char *errmsg);

template <typename T>
static int gupload(at3d_state *st, std::vector<void *> &owned, const T *host, size_t n, const T **dev, char *errmsg)
{
    *dev = nullptr;
    if (!host || n == 0) return 0;
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, n * sizeof(T)));
    owned.push_back(p);
    st->bytes += n * sizeof(T);
    CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T *)p;
    return 0;
}

extern "C" int at3d_state_attach_gradient(at3d_state *st, const at3d_grad_desc *g, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!st || !g) { set_msg(errmsg, "null argument"); return 1; }
    if (!st->S.radrec) { set_msg(errmsg, "the state was created without RADIANCE/RSHPTR: the gradient needs them"); return 1; }
    if (g->numder < 1) { set_msg(errmsg, "NUMDER must be >= 1"); return 1; }
    if (g->deriv_maxnmicro > st->S.maxnmicro) { set_msg(errmsg, "DERIV_MAXNMICRO > MAXNMICRO is not supported"); return 3; }
    // drop a previous attachment
    for (void *p : st->grad_owned) cudaFree(p);
    st->grad_owned.clear();
    st->grad_attached = 0;
    DevGrad &G = st->G;
    memset(&G, 0, sizeof(G));
    const DevState &S = st->S;
    G.maxpg = g->maxpg; G.numder = g->numder; G.dnumphase = g->dnumphase;
    G.deriv_maxnmicro = g->deriv_maxnmicro; G.pmaxnmicro = S.maxnmicro;
    G.longest_path_pts = g->longest_path_pts;
    G.exact_single_scatter = g->exact_single_scatter; G.singlescatter = g->singlescatter;
    const int mx = S.nx > S.ny ? (S.nx > S.nz ? S.nx : S.nz) : (S.ny > S.nz ? S.ny : S.nz);
    G.maxsub = g->maxsubgridints > 50 * mx ? g->maxsubgridints : 50 * mx;
    G.scatmin = g->scatmin;
    const size_t mp = (size_t)g->maxpg, nd = (size_t)g->numder, np = (size_t)S.npts;
    const size_t nlt = (size_t)S.nstleg * (S.nleg + 1);
    int rc = 0;
    std::vector<void *> &own = st->grad_owned;
#define GUP(field, count) if (!rc) rc = gupload(st, own, g->field, (size_t)(count), &G.field, errmsg)
    GUP(partder, nd); GUP(doexact, nd);
    GUP(dext, mp * nd); GUP(dalb, mp * nd); GUP(dextm, mp * nd);
    GUP(dalbm, 8 * np * nd); GUP(dfj, 8 * np * nd);
    GUP(optinterpwt, 8 * np); GUP(interpptr, 8 * np);
    GUP(dleg, nlt * g->dnumphase);
    GUP(dphasetab, (size_t)S.nstphase * g->dnumphase * S.nscatangle);
    GUP(diphasep, (size_t)g->deriv_maxnmicro * mp * nd); GUP(dphasewtp, (size_t)g->deriv_maxnmicro * mp * nd);
    GUP(iphasep, (size_t)S.maxnmicro * mp * S.npart); GUP(phasewtp, (size_t)S.maxnmicro * mp * S.npart);
    GUP(extinctp, mp * S.npart); GUP(albedop, mp * S.npart);
    if (g->exact_single_scatter) { GUP(dpath, (size_t)g->longest_path_pts * np); GUP(dptr, (size_t)g->longest_path_pts * np); }
#undef GUP
    if (rc) return rc;
    if (!G.partder || !G.doexact || !G.dext || !G.dalb || !G.dextm || !G.dalbm || !G.dfj || !G.optinterpwt ||
        !G.interpptr || !G.iphasep || !G.phasewtp || !G.extinctp || !G.albedop || !G.diphasep || !G.dphasewtp ||
        !G.dleg || !G.dphasetab) {
        set_msg(errmsg, "at3d_state_attach_gradient: a required derivative array is NULL");
        return 1;
    }
    if (g->exact_single_scatter && (!G.dpath || !G.dptr)) { set_msg(errmsg, "EXACT_SINGLE_SCATTER needs DPATH/DPTR"); return 1; }
    // ray-independent tables of COMPUTE_SOURCE_GRAD_1CELL (grad_prep_kernel)
    {
        G.ncomp = S.nstleg == 1 ? 1 : 4;
        G.ntup = (G.ncomp * (S.ml + 1) + 3) & ~3;
        const size_t nrow = np * nd * 8;
        float *grec = nullptr, *dlegt = nullptr, *legs = nullptr; float2 *gpnt = nullptr;
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, nrow * 8 * sizeof(float))); own.push_back(p); grec = (float *)p; st->bytes += nrow * 8 * sizeof(float);
        CUDA_TRY(cudaMalloc(&p, nrow * G.ntup * sizeof(float))); own.push_back(p); dlegt = (float *)p; st->bytes += nrow * G.ntup * sizeof(float);
        CUDA_TRY(cudaMalloc(&p, np * nd * sizeof(float2))); own.push_back(p); gpnt = (float2 *)p; st->bytes += np * nd * sizeof(float2);
        CUDA_TRY(cudaMemset(dlegt, 0, nrow * G.ntup * sizeof(float)));
        if (S.npart > 1) {
            CUDA_TRY(cudaMalloc(&p, np * nd * G.ntup * sizeof(float))); own.push_back(p); legs = (float *)p;
            st->bytes += np * nd * G.ntup * sizeof(float);
            CUDA_TRY(cudaMemset(legs, 0, np * nd * G.ntup * sizeof(float)));
        }
        const int wpb = 4;
        const size_t smem = (size_t)wpb * nlt * sizeof(float);
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(grad_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grad_prep_kernel<<<(S.npts + wpb - 1) / wpb, wpb * 32, smem>>>(S, G, grec, dlegt, gpnt, legs);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        G.grec = grec; G.dlegt = dlegt; G.gpnt = gpnt; G.legs = legs;
    }
    st->grad_attached = 1;
    return 0;
}

template <typename T>
static int stage_in(const T *src, size_t n, int host, void *dst, const T **out, cudaStream_t s, char *errmsg)
{
    if (!host) { *out = src; return 0; }
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    *out = (const T *)dst;
    return 0;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int at3d_levisapprox_gradient(at3d_state *st, const at3d_rays *rays, const at3d_grad_desc *g,
                                         double *gradout, double *cost, float *stokesout,
                                         const at3d_trace *trace, void *cuda_stream, double *kernel_ms,
                                         char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!st || !rays || !g || !gradout || !cost || !stokesout) { set_msg(errmsg, "null argument"); return 1; }
    if (!st->grad_attached) { set_msg(errmsg, "at3d_state_attach_gradient must be called first"); return 1; }
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const DevState &S = st->S;
    DevGrad G = st->G;
    const int nst = S.nstokes;
    const size_t n = rays->nrays, npix = g->npix;
    const size_t ngrad = (size_t)G.maxpg * G.numder;
    const bool host = rays->memspace == AT3D_MEM_HOST;
    const int nunc = g->nuncertainty;
    if (nunc < nst) { set_msg(errmsg, "NUNCERTAINTY must be >= NSTOKES"); return 1; }
    // ---- stage inputs ----
    const float *camx, *camy, *camz; const double *cammu, *camphi; const RayPack *packs;
    int rc = stage_rays(st, rays, stream, &camx, &camy, &camz, &cammu, &camphi, &packs, errmsg);
    if (rc) return rc;
    size_t o = 0;
    const size_t o_meas = o; o += al256(sizeof(float) * nst * npix);
    const size_t o_unc = o; o += al256(sizeof(double) * nunc * nunc * npix);
    const size_t o_rpp = o; o += al256(sizeof(int) * npix);
    const size_t o_rw = o; o += al256(sizeof(double) * n);
    const size_t o_sw = o; o += al256(sizeof(double) * nst * npix);
    CUDA_TRY(st->pix.reserve(o + 256));
    unsigned char *pb = (unsigned char *)st->pix.p;
    const float *meas; const double *unc, *rw, *sw; const int *rpp;
    if ((rc = stage_in(g->measurements, (size_t)nst * npix, host, pb + o_meas, &meas, stream, errmsg))) return rc;
    if ((rc = stage_in(g->uncertainties, (size_t)nunc * nunc * npix, host, pb + o_unc, &unc, stream, errmsg))) return rc;
    if ((rc = stage_in(g->rays_per_pixel, npix, host, pb + o_rpp, &rpp, stream, errmsg))) return rc;
    if ((rc = stage_in(g->ray_weights, n, host, pb + o_rw, &rw, stream, errmsg))) return rc;
    if ((rc = stage_in(g->stokes_weights, (size_t)nst * npix, host, pb + o_sw, &sw, stream, errmsg))) return rc;
    // ---- work buffers ----
    size_t cubtmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cubtmp, (const int *)nullptr, (int *)nullptr, (int)npix, stream);
    o = 0;
    const size_t w_vis = o; o += al256(sizeof(double) * nst * n);
    const size_t w_tot = o; o += al256(sizeof(double) * nst * n);
    const size_t w_adj = o; o += al256(sizeof(double) * nst * npix);
    const size_t w_costp = o; o += al256(sizeof(double) * npix);
    const size_t w_pixstart = o; o += al256(sizeof(int) * (npix + 1));
    const size_t w_raypix = o; o += al256(sizeof(int) * n);
    const size_t w_beam = o; o += al256(sizeof(double) * S.npts);
    const size_t w_cub = o; o += al256(cubtmp);
    const size_t w_grad = o; o += al256(sizeof(double) * ngrad);
    const size_t w_so = o; o += al256(sizeof(float) * nst * npix);
    const size_t w_cost = o; o += 256;
    CUDA_TRY(st->work.reserve(o + 256));
    unsigned char *wb = (unsigned char *)st->work.p;
    double *visrad = (double *)(wb + w_vis), *total = (double *)(wb + w_tot), *adjw = (double *)(wb + w_adj);
    double *costp = (double *)(wb + w_costp), *beam = (double *)(wb + w_beam);
    int *pixstart = (int *)(wb + w_pixstart), *raypix = (int *)(wb + w_raypix);
    double *grad_d = host ? (double *)(wb + w_grad) : gradout;
    float *so_d = host ? (float *)(wb + w_so) : stokesout;
    double *cost_d = host ? (double *)(wb + w_cost) : cost;
    int *tc = nullptr, *tn = nullptr, *ts = nullptr; int tcap = 0;
    if (trace && trace->cells) {
        tcap = trace->max_per_ray;
        if (host) {
            CUDA_TRY(st->trace.reserve(((size_t)tcap * n + 2 * n) * sizeof(int)));
            tc = (int *)st->trace.p; tn = tc + (size_t)tcap * n; ts = tn + n;
            CUDA_TRY(cudaMemsetAsync(tc, 0, ((size_t)tcap * n + 2 * n) * sizeof(int), stream));
        } else { tc = trace->cells; tn = trace->ncells; ts = trace->nsub; }
    }
    CUDA_TRY(st->err.reserve(sizeof(RayErr)));
    CUDA_TRY(cudaMemsetAsync(st->err.p, 0, sizeof(RayErr), stream));
    CUDA_TRY(cudaMemsetAsync(grad_d, 0, sizeof(double) * ngrad, stream));
    CUDA_TRY(cudaMemsetAsync(beam, 0, sizeof(double) * S.npts, stream));
    CUDA_TRY(cudaMemsetAsync(so_d, 0, sizeof(float) * nst * npix, stream));
    CUDA_TRY(cudaMemsetAsync(cost_d, 0, sizeof(double), stream));
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (kernel_ms) { for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]); cudaEventRecord(ev[0], stream); }
    if (n > 0 && npix > 0) {
        // ---- Phase 1: forward radiances (INTEGRATE_1RAY arithmetic for the pixel values; the
        //      ADJOINT_INTEGRATE_1RAY arithmetic for the totals the derivative pass needs) ----
        DevState Sf = S;          // the work counters describe the adjoint pass only
        Sf.counts = nullptr;
        CUDA_TRY(cudaMemsetAsync(st->counts_dev, 0, 8 * sizeof(unsigned long long), stream));
        CUDA_TRY(launch_forward(Sf, (int)n, camx, camy, camz, cammu, camphi, packs, nullptr, visrad, total, 3, 1,
                                G.singlescatter, 0, G.maxsub, nullptr, 0, nullptr, nullptr, (RayErr *)st->err.p,
                                st->ray_counter, stream));
        if (kernel_ms) cudaEventRecord(ev[1], stream);
        // ---- Phase 2 ----
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(wb + w_cub, cubtmp, rpp, pixstart, (int)npix, stream));
        const int nb = (int)((npix + 127) / 128);
        if (nst == 1)
            pixel_kernel<1><<<nb, 128, 0, stream>>>((int)npix, pixstart, rpp, visrad, rw, sw, meas, unc, nunc,
                                                      g->costfunc_ll, so_d, adjw, costp, raypix);
        else
            pixel_kernel<3><<<nb, 128, 0, stream>>>((int)npix, pixstart, rpp, visrad, rw, sw, meas, unc, nunc,
                                                      g->costfunc_ll, so_d, adjw, costp, raypix);
        CUDA_TRY(cudaGetLastError());
        cost_reduce_kernel<<<1, 1024, 0, stream>>>((int)npix, costp, cost_d);
        CUDA_TRY(cudaGetLastError());
        // ---- Phase 3 ----
        const GradLayout L = grad_layout(nst, S.ny_comp, S.nlmp, S.ml, G.ncomp, G.ntup, G.numder);
        const size_t smem = (size_t)AT3D_RAYS_PER_BLOCK * L.total;
        if (smem > 227 * 1024) { set_msg(errmsg, "gradient kernel needs %zu bytes of shared memory (NUMDER too large)", smem); return 3; }
        const void *fn = nst == 1 ? (const void *)adjoint_kernel<1> : (const void *)adjoint_kernel<3>;
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0, nsm = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, AT3D_RAY_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        long want = ((long)n + AT3D_RAYS_PER_BLOCK - 1) / AT3D_RAYS_PER_BLOCK;
        long cap = (long)nsm * per_sm;
        const int nblk = (int)(want < cap ? want : cap);
        CUDA_TRY(cudaMemsetAsync(st->ray_counter, 0, sizeof(int), stream));
        if (nst == 1)
            adjoint_kernel<1><<<nblk, AT3D_RAY_THREADS, smem, stream>>>(S, G, L, (int)n, camx, camy, camz, cammu,
                camphi, packs, raypix, adjw, rw, sw, total, grad_d, beam, tc, tcap, tn, ts, (RayErr *)st->err.p,
                st->ray_counter);
        else
            adjoint_kernel<3><<<nblk, AT3D_RAY_THREADS, smem, stream>>>(S, G, L, (int)n, camx, camy, camz, cammu,
                camphi, packs, raypix, adjw, rw, sw, total, grad_d, beam, tc, tcap, tn, ts, (RayErr *)st->err.p,
                st->ray_counter);
        CUDA_TRY(cudaGetLastError());
        if (kernel_ms) cudaEventRecord(ev[2], stream);
        // ---- Phase 4 ----
        if (G.exact_single_scatter) {
            const int wpb = 8;
            beam_kernel<<<(S.npts + wpb - 1) / wpb, wpb * 32, 0, stream>>>(G, S.npts, beam, grad_d);
            CUDA_TRY(cudaGetLastError());
        }
    } else if (kernel_ms) { cudaEventRecord(ev[1], stream); cudaEventRecord(ev[2], stream); }
    if (kernel_ms) cudaEventRecord(ev[3], stream);
    if (host) {
        CUDA_TRY(cudaMemcpyAsync(gradout, grad_d, sizeof(double) * ngrad, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(stokesout, so_d, sizeof(float) * nst * npix, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(cost, cost_d, sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (tc) {
            CUDA_TRY(cudaMemcpyAsync(trace->cells, tc, (size_t)tcap * n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(trace->ncells, tn, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(trace->nsub, ts, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
    }
    rc = check_ray_err(st, stream, errmsg);
    if (kernel_ms) {
        float ms;
        cudaEventSynchronize(ev[3]);
        cudaEventElapsedTime(&ms, ev[0], ev[1]); kernel_ms[0] = ms;
        cudaEventElapsedTime(&ms, ev[1], ev[2]); kernel_ms[1] = ms;
        cudaEventElapsedTime(&ms, ev[2], ev[3]); kernel_ms[2] = ms;
        cudaEventElapsedTime(&ms, ev[0], ev[3]); kernel_ms[3] = ms;
        for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
    }
    return rc;
}
