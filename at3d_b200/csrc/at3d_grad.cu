// at3d_grad.cu -- LEVISAPPROX_GRADIENT, default adjoint ("double sweep") path, on sm_100a.
// Replaces (src/polarized/shdomsub4.f of the AT3D reference)
//   LEVISAPPROX_GRADIENT :288-809, ADJOINT_INTEGRATE_1RAY :3223-3967, COMPUTE_SOURCE_GRAD_1CELL
//   :1546-2042, COMPUTE_SOURCE_DIRECTION :2836-2914, FIND_BOUNDARY_RADIANCE_GRAD :2151-2347,
//   COMPUTE_ADJOINT_WEIGHTS :3969-4034, COMPUTE_RADIANCE_DERIVATIVE_ADJOINT :4037-4114,
//   COMPUTE_DIRECT_BEAM_DERIV_ADJOINT :4117-4143.
//
// Design (DESIGN.md "Gradient kernels"): the derivative pass is split into two small kernels.
//  * weights_kernel (phase A): the bit-exact ADJOINT_INTEGRATE_1RAY walk, one octet per ray.  Per visit
//    of a grid point (from the cell where it becomes a corner to the cell where it stops being one) it
//    accumulates in registers the two scalars the gradient is linear in -- W (source term,
//    shdomsub4.f:3692-3764) and G (radiance term, COMPUTE_RADIANCE_DERIVATIVE_ADJOINT, with the
//    backward cumulative sum PASSEDRAD rewritten as "total - running") -- and writes one 32-byte
//    record (point, SRCEXT8/EXT, W, G) per visit.  The forward pass counted the visits per ray, so the
//    records of a ray are contiguous.  BEAM_WEIGHT is accumulated here as well.
//  * apply_kernel (phase B): one octet per ray contracts, for every record, the point's precomputed
//    gradient SH rows with YLMDIR, adds the scalar single-scatter terms and issues one
//    red.global.add.f64 per (property corner, unknown).  No geometry, no FP64 chains, small code.
//  * Everything in COMPUTE_SOURCE_GRAD_1CELL that does not depend on the ray (LEGENT/LEGENP/DLEGP,
//    DLEGT(l), the factors of the single-scatter terms) is evaluated once per cost-function evaluation
//    by grad_prep_kernel; DLEGT is folded with the radiance into SH rows XI*DLEGT(l_j)*RADIANCE(.,j),
//    so the reference's 8*NUMDER COMPUTE_SOURCE_DIRECTION calls per new point become one SH dot
//    product per non-zero property corner.
#include "at3d_mem.h"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <vector>
#include <cub/cub.cuh>
#include "at3d_host.h"
#include "at3d_ray.cuh"

#define AT3D_DTAB_FLAG 0x40000000   // list entry refers to DPHASETAB instead of PHASETAB

// one record per visit of a grid point by a ray (phase A -> phase B)
struct __align__(16) VisitRec {
    int ip;                 // grid point (1-based)
    float srcfull[3];       // SRCEXT8 before the multiplication by the extinction
    double W, G;            // accumulated source-term / radiance-term weights
    double B;               // adjoint-weighted SRCSINGSCAT of the visit: the point's BEAM_WEIGHT contribution
    double pad;
};

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
// Ray-independent part of COMPUTE_SOURCE_GRAD_1CELL (shdomsub4.f:1801-2007), once per attach.
// One warp per grid point.  Per (point, unknown): SCATTERJ, F and the (phase index, weight) list of
// SINGSCATJ.  Per row = (point, unknown, property corner with XI>=1e-7): DLEGT(l) (shdomsub4.f:1978),
// folded with the radiance into SH rows  XI*DLEGT(l_j)*RADIANCE(.,j)  so that the ray only contracts
// them with YLMDIR like a source block; the scalar factors of the GRAD8 terms; the (phase index,
// coefficient) list of the single-scatter terms (shdomsub4.f:1996-2007).
// ------------------------------------------------------------------------------------------
__global__ void grad_prep_kernel(DevState S, DevGrad G, int4 *rowrec, int4 *sprec, float *dsh, float *dlegt_out,
                                 float *legs_out)
{
    extern __shared__ float prep_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ipz = blockIdx.x * (blockDim.x >> 5) + warp;
    const int nstleg = S.nstleg, ml = S.ml, nst = S.nstokes;
    const int nlt = nstleg * (S.nleg + 1);
    if (ipz >= S.npts) return;
    const int pm = G.pmaxnmicro, dm = G.deriv_maxnmicro, nd = G.numder, ncomp = G.ncomp, ntup = G.ntup;
    float *legent = prep_smem + (size_t)warp * (nlt + ntup);
    float *dl = legent + nlt;                     // compact DLEGT (or SOURCET table) of the current row
    const bool deltam = S.deltam != 0, interp_new = S.interp_new != 0;
    // every solar term carries DIRFLUX*SECMU0 (shdomsub4.f:1996-2007, 2899-2911): none for SRCTYPE='T'
    const bool solar = S.srctype != 'T', thermal = S.srctype != 'S';
    const float secmu0 = solar ? (float)(1.0 / fabs((double)S.solarmu)) : 0.0f;
    const float dirflux = solar ? __ldg(&S.dirflux[ipz]) : 0.0f;
    // PLANCK / DPLANCK of the grid point (shdomsub4.f:1792-1799; PLANCK_DERIVATIVE :3171-3221, UNITS 'T' | 'R')
    float planck = 0.0f, dplanck = 0.0f;
    if (thermal) {
        const float tk = __ldg(&S.temp[ipz]);
        planck = dev_planck(tk, S.units, S.wavelen);
        if (S.units == 'T') dplanck = 1.0f;
        else if (tk > 0.0f) {
            const float e = expf(1.4388e4f / (S.wavelen * tk));
            const float w3 = S.wavelen * S.wavelen * S.wavelen;
            dplanck = (1.1911e8f * 1.4388e4f) / (w3 * w3) * e / (tk * tk * ((e - 1) * (e - 1)));
        }
        if (planck < 1e-7f) dplanck = 0.0f;
    }
    const int nb_l = lane & 7;
    const int ib_l = __ldg(&G.interpptr[nb_l + 8 * (size_t)ipz]);
    const float xi_l = __ldg(&G.optinterpwt[nb_l + 8 * (size_t)ipz]);
    const unsigned active = __ballot_sync(FULLMASK, lane < 8 && xi_l >= 1e-7f) & 0xFFu;
    const int4 gp = __ldg(&G.gptrec[ipz]);
    const int nnz = gp.y >> 16, nrp = (gp.w & 0xFF) * 32;
    const size_t rowbase = (size_t)gp.x;
    float *dshp = dsh + (size_t)(unsigned)gp.z * 32;
    const int2 rr = __ldg(&S.radrec[ipz]);
    const float *rad = S.shrad + rr.x;
    const int rns = rr.y;
    int last_ipa = -1;
    float scatterj = 0.0f, f = 0.0f;
    for (int idr = 0; idr < nd; idr++) {
        const int ipa = __ldg(&G.partder[idr]);       // 1-based species
        const float albp_l = __ldg(&G.albedop[(ib_l - 1) + (size_t)G.maxpg * (ipa - 1)]);
        const float extp_l = __ldg(&G.extinctp[(ib_l - 1) + (size_t)G.maxpg * (ipa - 1)]);
        const float alb_ip = __ldg(&S.albedo[ipz + (size_t)S.npts * (ipa - 1)]);
        const float sw_l = xi_l * albp_l * extp_l;       // SPATIAL_WEIGHT of property corner nb
        if (ipa != last_ipa) {
            last_ipa = ipa;
            __syncwarp();
            scatterj = 0.0f;
            if (deltam) for (int n = 0; n < 8; n++) scatterj = scatterj + __shfl_sync(FULLMASK, sw_l, n);
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
                for (int n = 0; n < 8; n++) { swn[n] = __shfl_sync(FULLMASK, sw_l, n); ibn[n] = __shfl_sync(FULLMASK, ib_l, n); }
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
                // the table COMPUTE_SOURCE_DIRECTION contracts with the radiance for SOURCET -> dl
                for (int t = lane; t < ntup; t += 32) dl[t] = 0.0f;
                __syncwarp();
                for (int t = lane; t < nstleg * (ml + 1); t += 32) {
                    const int k = t % nstleg, l = t / nstleg;
                    if (k == 3 || k == 5) continue;
                    dl[comp_slot(k) + ncomp * l] = legent[t];
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
        }
        if (S.npart > 1) {
            // SOURCET row of this unknown: table(l_j)*RADIANCE(.,j) (same species as the previous unknown
            // leaves dl untouched: rebuilt identically)
            if (ipa == last_ipa && idr > 0 && __ldg(&G.partder[idr - 1]) == ipa) {
                // dl was overwritten by the rows of the previous unknown: copy its SOURCET row instead
                const float *prev = dshp + (size_t)(nnz * nd + idr - 1) * nst * nrp;
                float *cur = dshp + (size_t)(nnz * nd + idr) * nst * nrp;
                for (int j = lane; j < nst * nrp; j += 32) cur[j] = prev[j];
                if (legs_out) for (int t = lane; t < ntup; t += 32)
                    legs_out[((size_t)ipz * nd + idr) * ntup + t] = legs_out[((size_t)ipz * nd + idr - 1) * ntup + t];
            } else {
                float *cur = dshp + (size_t)(nnz * nd + idr) * nst * nrp;
                for (int j = lane; j < nrp; j += 32) {
                    float p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
                    if (j < rns) {
                        const int l = __ldg(&S.lofj[j]);
                        const float r1 = rad[j];
                        p1 = dl[ncomp * l] * r1;
                        if (nst > 1) {
                            const float r2 = rad[nrp + j], r3 = rad[2 * nrp + j];
                            p1 = p1 + dl[3 + ncomp * l] * r2;
                            p2 = dl[3 + ncomp * l] * r1 + dl[1 + ncomp * l] * r2;
                            p3 = dl[2 + ncomp * l] * r3;
                        }
                    }
                    cur[j] = p1;
                    if (nst > 1) { cur[nrp + j] = p2; cur[2 * nrp + j] = p3; }
                }
                if (legs_out) for (int t = lane; t < ntup; t += 32) legs_out[((size_t)ipz * nd + idr) * ntup + t] = dl[t];
            }
            __syncwarp();
        }
        // ---- species record: alb, F, SCATTERJ and the SINGSCATJ list (shdomsub4.f:1809-1832) ----
        if (lane == 0) {
            int4 *sp = sprec + ((size_t)ipz * nd + idr) * G.sp_stride;
            int2 *list = (int2 *)(sp + 1);
            int cnt = 0;
            if (deltam && solar) {
                const float sdiv = (scatterj > G.scatmin) ? scatterj : (float)G.scatmin;
                for (int n = 0; n < 8; n++) {
                    const int ibn = __ldg(&G.interpptr[n + 8 * (size_t)ipz]);
                    const float swn = __ldg(&G.optinterpwt[n + 8 * (size_t)ipz]) *
                                      __ldg(&G.albedop[(ibn - 1) + (size_t)G.maxpg * (ipa - 1)]) *
                                      __ldg(&G.extinctp[(ibn - 1) + (size_t)G.maxpg * (ipa - 1)]);
                    if (swn <= 1e-6f) continue;
                    const int *iq = G.iphasep + (size_t)pm * ((ibn - 1) + (size_t)G.maxpg * (ipa - 1));
                    const float *wq = G.phasewtp + (size_t)pm * ((ibn - 1) + (size_t)G.maxpg * (ipa - 1));
                    for (int q = 0; q < pm; q++) {
                        const float w = __ldg(&wq[q]);
                        if (w <= 1e-6f) continue;
                        list[cnt++] = make_int2(__ldg(&iq[q]), __float_as_int(swn * w / sdiv));
                    }
                }
            }
            sp[0] = make_int4(__float_as_int(alb_ip), __float_as_int(f), __float_as_int(scatterj), cnt);
        }
        // ---- rows: one per property corner with a non-zero interpolation weight ----
        unsigned todo = active;
        int rank = 0;
        const int doex = __ldg(&G.doexact[idr]);
        while (todo) {
            const int nb = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ib = __shfl_sync(FULLMASK, ib_l, nb);
            const float xi = __shfl_sync(FULLMASK, xi_l, nb);
            const float albp = __shfl_sync(FULLMASK, albp_l, nb), extp = __shfl_sync(FULLMASK, extp_l, nb);
            const int *iphp = G.iphasep + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
            const float *pwp = G.phasewtp + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
            const float dext_v = __ldg(&G.dext[(ib - 1) + (size_t)G.maxpg * idr]);
            const float dalb_v = __ldg(&G.dalb[(ib - 1) + (size_t)G.maxpg * idr]);
            const float dextm_v = __ldg(&G.dextm[(ib - 1) + (size_t)G.maxpg * idr]);
            const float dalbm_v = __ldg(&G.dalbm[nb + 8 * ((size_t)ipz + (size_t)S.npts * idr)]);
            const float dfj_v = __ldg(&G.dfj[nb + 8 * ((size_t)ipz + (size_t)S.npts * idr)]);
            const int *dip = G.diphasep + (size_t)dm * ((ib - 1) + (size_t)G.maxpg * idr);
            const float *dpw = G.dphasewtp + (size_t)dm * ((ib - 1) + (size_t)G.maxpg * idr);
            const size_t R = rowbase + (size_t)idr * nnz + rank;
            __syncwarp();
            for (int t = lane; t < ntup; t += 32) dl[t] = 0.0f;
            __syncwarp();
            for (int t = lane; t < nlt; t += 32) {
                const int k = t % nstleg, l = t / nstleg;
                if (k == 3 || k == 5 || l > ml) continue;       // components that never reach I,Q,U
                float legenp = 0.0f, dlegp = 0.0f;
                for (int q = 0; q < pm; q++) {
                    const float *lg = S.legen + (size_t)nlt * (__ldg(&iphp[q]) - 1);
                    const float ftemp = deltam ? __ldg(&lg[nstleg * (ml + 1)]) : 0.0f;
                    const float un = unscale_leg(__ldg(&lg[t]), k, l, ml, deltam, interp_new, ftemp, nstleg);
                    legenp = legenp + __ldg(&pwp[q]) * un;
                    if (doex == 0 && q < dm) dlegp = dlegp + __ldg(&dpw[q]) * un;
                }
                if (doex == 1)
                    for (int q = 0; q < dm; q++)
                        dlegp = dlegp + __ldg(&pwp[q]) * __ldg(&G.dleg[(size_t)nlt * (__ldg(&dip[q]) - 1) + t]);
                const float lt = legent[t];
                const float leg_diff = legenp - lt;
                dl[comp_slot(k) + ncomp * l] = dext_v * leg_diff * albp + dalb_v * leg_diff * extp
                                               + dlegp * extp * albp + (lt - 1) * dfj_v;
            }
            __syncwarp();
            if (dlegt_out) for (int t = lane; t < ntup; t += 32) dlegt_out[R * ntup + t] = dl[t];
            // SH row: XI * DLEGT(l_j) (x) RADIANCE(.,j), Stokes coupling of COMPUTE_SOURCE_DIRECTION
            {
                float *cur = dshp + (size_t)(idr * nnz + rank) * nst * nrp;
                for (int j = lane; j < nrp; j += 32) {
                    float p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
                    if (j < rns) {
                        const int l = __ldg(&S.lofj[j]);
                        const float r1 = rad[j];
                        p1 = dl[ncomp * l] * r1;
                        if (nst > 1) {
                            const float r2 = rad[nrp + j], r3 = rad[2 * nrp + j];
                            p1 = p1 + dl[3 + ncomp * l] * r2;
                            p2 = dl[3 + ncomp * l] * r1 + dl[1 + ncomp * l] * r2;
                            p3 = dl[2 + ncomp * l] * r3;
                        }
                    }
                    cur[j] = xi * p1;
                    if (nst > 1) { cur[nrp + j] = xi * p2; cur[2 * nrp + j] = xi * p3; }
                }
            }
            if (lane == 0) {
                int4 *rw = rowrec + R * G.prow_stride;
                int2 *list = (int2 *)(rw + 2);
                int cnt = 0;
                float cj = 0.0f;
                if (deltam) {
                    const float K = dirflux * secmu0 * xi;
                    cj = K * (dfj_v - dalb_v * extp - dext_v * albp);
                    const float cp = K * (dalb_v * extp + dext_v * albp);
                    const float cd = K * extp * albp;
                    for (int q = 0; q < pm; q++) {
                        const float c = cp * __ldg(&pwp[q]);
                        if (c != 0.0f) list[cnt++] = make_int2(__ldg(&iphp[q]), __float_as_int(c));
                    }
                    for (int q = 0; q < dm; q++) {
                        if (doex == 0) {
                            const float c = cd * __ldg(&dpw[q]);
                            if (c != 0.0f) list[cnt++] = make_int2(__ldg(&iphp[q]), __float_as_int(c));
                        } else if (doex == 1) {
                            const float c = cd * __ldg(&pwp[q]);
                            if (c != 0.0f) list[cnt++] = make_int2(__ldg(&dip[q]) | AT3D_DTAB_FLAG, __float_as_int(c));
                        }
                    }
                }
                // thermal component of GRAD8(1,...) (shdomsub4.f:2009-2016): the same for every ray
                float therm = 0.0f;
                if (thermal) {
                    const float dtemp_v = G.dtemp ? __ldg(&G.dtemp[(ib - 1) + (size_t)G.maxpg * idr]) : 0.0f;
                    therm = xi * (__ldg(&S.extinct[ipz + (size_t)S.npts * (ipa - 1)]) * (1.0f - alb_ip) * dplanck * dtemp_v
                                  - planck * dalbm_v + planck * (1.0f - alb_ip) * dextm_v);
                }
                rw[0] = make_int4(__float_as_int(xi * (alb_ip * dextm_v + dalbm_v)), __float_as_int(cj),
                                  __float_as_int(dextm_v * xi), ib);
                rw[1] = make_int4(nb | (cnt << 8), __float_as_int(therm), 0, 0);
            }
            rank++;
        }
    }
}

// Truncated single scattering of one grid point without delta-M (shdomsub4.f:1723-1760,1781): sum over
// species of DA*LEGENT(.,l)*sum_m YLMDIR*YLMSUN for the shells present in SOURCE (octet-cooperative).
template <int NST>
__device__ __forceinline__ void nodeltam_singscat(const DevState &S, int ip, int sns, float ext, const float *Vsh,
                                                  const Oct &o, float (&b)[NST])
{
    const int nstleg = S.nstleg, ml = S.ml, mm = S.mm, ipz = ip - 1;
    const int nlt = nstleg * (S.nleg + 1);
    const bool interp_new = S.interp_new != 0;
    const float dirflux = __ldg(&S.dirflux[ipz]);
    const float secmu0 = S.srctype != 'T' ? (float)(1.0 / fabs((double)S.solarmu)) : 0.0f;
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
            if (sh_index(l, -me, mm) >= sns) continue;
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

// The values lane n keeps for corner n of the current cell in the adjoint walk.
template <int NST>
struct GCorner {
    int pt;
    float x, y, z, ext;
    float src[NST], ss[NST], srcfull[NST];
    double W, G, B;         // weights accumulated since the point became a corner
};

template <int NST>
__device__ __forceinline__ void write_visit(VisitRec *dst, const GCorner<NST> &K)
{
    int4 a;
    a.x = K.pt;
    a.y = __float_as_int(K.srcfull[0]);
    a.z = NST > 1 ? __float_as_int(K.srcfull[NST > 1 ? 1 : 0]) : 0;
    a.w = NST > 2 ? __float_as_int(K.srcfull[NST > 2 ? 2 : 0]) : 0;
    *(int4 *)dst = a;
    *((double2 *)dst + 1) = make_double2(K.W, K.G);
    *((double2 *)dst + 2) = make_double2(K.B, 0.0);
}

// rows of GRADOUT a visit of grid point ip contributes to (+1: its BEAM_WEIGHT entry) = pair slots of its record
__device__ __forceinline__ int visit_pairs(const DevGrad &G, int ip) { return (__ldg(&G.gptrec[ip - 1]).y & 0xFFFF) + 1; }

#include "at3d_gwalk.cuh"
#ifndef GW_BT
#define GW_BT 256
#endif
#ifndef GW_MINB
#define GW_MINB 1
#endif

// ------------------------------------------------------------------------------------------
// Phase A for NSTOKES=1 with delta-M: the thread-per-ray walk of at3d_gwalk.cuh (sources from the forward pass's stream)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GW_BT, GW_MINB)
weights_kernel_t(DevState S, DevGrad G, int nrays, int ray0, const float *camx, const float *camy, const float *camz,
                 const double *cammu, const double *camphi, const RayPack *packs, const int *raypix, const double *adjw,
                 const double *ray_weights, const double *stokes_weights, const double *total,
                 const long long *recoff /*[nrays+1]*/, long long rec_base, VisitRec *recs, int *nrec_out, int *npairs_out,
                 int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub, RayErr *err, int *ray_counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int bt = blockDim.x, lane = threadIdx.x & 31, tid = threadIdx.x;
    GwShared M;
    M.bt = bt;
    {
        unsigned char *q = smem_raw;
        M.W = (double *)q + tid; q += (size_t)8 * 8 * bt;
        M.Gr = (double *)q + tid; q += (size_t)8 * 8 * bt;
        M.B = (double *)q + tid; q += (size_t)8 * 8 * bt;
        M.sf = (float *)q + tid; q += (size_t)8 * 4 * bt;
        M.ss = (float *)q + tid; q += (size_t)8 * 4 * bt;
        M.bw = (float *)q + tid; q += (size_t)8 * 4 * bt;
        M.fa = (double *)q + tid; q += (size_t)8 * 8 * bt;
        M.fb = (double *)q + tid;
    }
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ray_counter, 32);
        base = __shfl_sync(FULLMASK, base, 0);
        if (ray0 + base >= nrays) break;
        const int iray = ray0 + base + lane;
        int ntrace = 0, nsub = 0, npt = 0, nsh = 0, marched = 0, nrec = 0, npairs = 0;
        if (iray < nrays) {
            const double mu2 = __ldg(&cammu[iray]), phi2 = __ldg(&camphi[iray]);
            const RayPack pk = dev_get_pack(S, packs, iray, camx, camy, camz, mu2, phi2);
            if (pk.status == 2) set_err(err, 2, iray);
            else if (pk.status == 0) {
                const int pix = __ldg(&raypix[iray]);
                const double adj = __ldg(&adjw[pix]) * __ldg(&ray_weights[iray]) * __ldg(&stokes_weights[pix]);
                RayDir rd;
                dev_ray_dir(S, pk, rd); rd.phi2 = (float)phi2;
                rd.hit = nullptr;
                const float sky = (-mu2 > 0.0) ? dev_sky_radiance(S, (float)mu2, (float)phi2) : 0.0f;
                const long long r0 = __ldg(&recoff[iray]), r1 = __ldg(&recoff[iray + 1]);
                SrcReader sr;
                const long long s0 = __ldg(&S.srcstart[iray]);
                sr.pool = S.srcpool; sr.p = S.srcpool + (s0 < 0 ? 0 : s0); sr.left = AT3D_SRC_CHUNK - 1;
                const int e = thread_march_weights(S, G, M, rd, mu2, pk.x0, pk.y0, pk.z0, sky, adj, __ldg(&total[iray]), sr,
                                                   recs + (r0 - rec_base), (int)(r1 - r0),
                                                   trace_cells ? trace_cells + (size_t)trace_cap * iray : nullptr, trace_cap,
                                                   ntrace, nsub, npt, nsh, nrec, npairs);
                if (e) { set_err(err, e, iray); nrec = 0; npairs = 0; }
                else marched = 1;
            }
            nrec_out[iray] = nrec;
            npairs_out[iray] = npairs;
            if (trace_n) { trace_n[iray] = ntrace; trace_nsub[iray] = nsub; }
        }
        __syncwarp();
        if (S.counts) {
            const int c0 = __reduce_add_sync(FULLMASK, marched ? ntrace : 0), c1 = __reduce_add_sync(FULLMASK, marched ? npt : 0);
            const int c4 = __reduce_add_sync(FULLMASK, marched ? nsub : 0), c5 = __reduce_add_sync(FULLMASK, marched);
            const int c2 = __reduce_add_sync(FULLMASK, marched ? nsh : 0);
            if (lane == 0) {
                atomicAdd(&S.counts[0], (unsigned long long)c0); atomicAdd(&S.counts[1], (unsigned long long)c1);
                atomicAdd(&S.counts[2], (unsigned long long)c2);
                atomicAdd(&S.counts[4], (unsigned long long)c4); atomicAdd(&S.counts[5], (unsigned long long)c5);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Phase A: ADJOINT_INTEGRATE_1RAY walk of one ray (one octet), forward values only; emits VisitRecs.
// ------------------------------------------------------------------------------------------
template <int NST>
__device__ int march_weights(const DevState &S, const DevGrad &G, const float *Ysh, const float *Vsh,
                             const RayDir &rd, double mu2, double x0, double y0, double z0, float sky,
                             const double (&adj)[NST], const double (&total)[NST], const Oct &o,
                             VisitRec *rec, int cap,
                             int *trace_cells, int trace_cap, int &ntrace, int &nsub, int &nrec, int &npairs)
{
    double xe = x0, ye = y0, ze = z0, transmit = 1.0;
    double radout[NST];
    float ext1 = 0.0f, srcext1[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { radout[k] = 0.0; srcext1[k] = 0.0f; }
    const int p1c = cell_gp(S, 1, 1), p8c = cell_gp(S, 1, 8);
    const double eps = (double)(1.0e-5f * (pt_coord(S, p8c, 3) - pt_coord(S, p1c, 3)));
    const bool exact_ss = G.exact_single_scatter != 0;
    const bool deltam = S.deltam != 0;
    int icell = dev_locate_grid_cell(S, xe, ye, ze);
    int iface = 0, npassed = 1, jf = 0;
    bool done = false;
    int npt_eval = 0, nsh_eval = 0;
    int bnd_pt = 0; double bnd_val = 0.0;            // surface point of lanes 0..3 (ray ends on the ground)
    GCorner<NST> K;
    K.pt = 0; K.x = K.y = K.z = K.ext = 0.0f; K.W = 0.0; K.G = 0.0; K.B = 0.0;
#pragma unroll
    for (int k = 0; k < NST; k++) { K.src[k] = 0.0f; K.ss[k] = 0.0f; K.srcfull[k] = 0.0f; }
    ntrace = 0; nsub = 0; nrec = 0;
    int mypairs = 0;                                 // pair slots of the records this lane wrote
    int err = 0;
    bool any_cell = false;
    CellRec c;
    if (icell > 0) c = load_cell(S, icell);
    while (!done && icell > 0) {
        if (trace_cells && o.ol == 0 && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        // ---- corners: DONEFACE carry-over (values and accumulated weights), records of the points that left ----
        {
            const int myp = own_corner(c, o.ol);
            const int from = o.ol ^ (jf == 1 ? 1 : jf == 2 ? 2 : 4);
            const int cand = __shfl_sync(o.m, K.pt, from, 8);
            const bool hit = (jf != 0) && (cand == myp);
            // the old corner of this lane survives iff lane `from` (its only possible heir) hit it
            const bool heir_hit = __shfl_sync(o.m, (int)hit, from, 8) != 0;
            // (a visit in clear air contributes exactly nothing: no record)
            const bool evict = any_cell && !heir_hit && (K.W != 0.0 || K.G != 0.0 || K.B != 0.0);
            const unsigned ev = oct_ballot(o, evict);
            if (evict) {
                const int k = nrec + __popc(ev & ((1u << o.ol) - 1));
                if (k < cap) { write_visit<NST>(rec + k, K); mypairs += visit_pairs(G, K.pt); } else err = 5;
            }
            nrec += __popc(ev);
            const int src_lane = hit ? from : o.ol;
            K.x = __shfl_sync(o.m, K.x, src_lane, 8); K.y = __shfl_sync(o.m, K.y, src_lane, 8);
            K.z = __shfl_sync(o.m, K.z, src_lane, 8); K.ext = __shfl_sync(o.m, K.ext, src_lane, 8);
            K.W = __shfl_sync(o.m, K.W, src_lane, 8); K.G = __shfl_sync(o.m, K.G, src_lane, 8);
            K.B = __shfl_sync(o.m, K.B, src_lane, 8);
#pragma unroll
            for (int k = 0; k < NST; k++) {
                K.src[k] = __shfl_sync(o.m, K.src[k], src_lane, 8);
                K.ss[k] = __shfl_sync(o.m, K.ss[k], src_lane, 8);
                K.srcfull[k] = __shfl_sync(o.m, K.srcfull[k], src_lane, 8);
            }
            K.pt = myp;
            any_cell = true;
            int soff = 0, sns = 0;
            float b[NST];
#pragma unroll
            for (int k = 0; k < NST; k++) b[k] = 0.0f;
            if (!hit) {
                load_corner<NST>(S, myp, rd, K.x, K.y, K.z, K.ext, soff, sns, b);
                K.W = 0.0; K.G = 0.0; K.B = 0.0;
            }
            unsigned need = oct_ballot(o, !hit);
            npt_eval += __popc(need);
            while (need) {
                const int n = __ffs(need) - 1;
                const int ipn = __shfl_sync(o.m, myp, n, 8);
                const int off = __shfl_sync(o.m, soff, n, 8), ns = __shfl_sync(o.m, sns, n, 8);
                float a[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) a[k] = 0.0f;
                sh_dot_partial<NST>(S.shsrc + off, AT3D_SHPAD(ns), Ysh, S.nlmp, o, a);
#pragma unroll
                for (int k = 0; k < NST; k++) a[k] = oct_sum(o.m, a[k]);
                float bt[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) bt[k] = 0.0f;
                if (!deltam) {
                    const float extn_ = __shfl_sync(o.m, K.ext, n, 8);
                    nodeltam_singscat<NST>(S, ipn, ns, extn_, Vsh, o, bt);
                }
                // new corners that are the same grid point (zero-width open-boundary cells) share the evaluation
                const bool mine = !hit && (myp == ipn);
                const unsigned same = oct_ballot(o, mine);
                nsh_eval += ns * __popc(same);
                need &= ~same;
                if (mine) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        const float bb = deltam ? b[k] : bt[k];
                        K.srcfull[k] = deltam ? a[k] + bb : a[k];
                        K.ss[k] = bb * K.ext;
                        K.src[k] = G.singlescatter ? K.ss[k] : K.srcfull[k] * K.ext;
                    }
                }
            }
        }
        float e8[8], s8[NST][8];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            e8[n] = __shfl_sync(o.m, K.ext, n, 8);
#pragma unroll
            for (int k = 0; k < NST; k++) s8[k][n] = __shfl_sync(o.m, K.src[k], n, 8);
        }
        const float q1x = __shfl_sync(o.m, K.x, 0, 8), q1y = __shfl_sync(o.m, K.y, 0, 8), q1z = __shfl_sync(o.m, K.z, 0, 8);
        const float q8x = __shfl_sync(o.m, K.x, 7, 8), q8y = __shfl_sync(o.m, K.y, 7, 8), q8z = __shfl_sync(o.m, K.z, 7, 8);
        const float qox = __shfl_sync(o.m, K.x, 8 - rd.ioct, 8), qoy = __shfl_sync(o.m, K.y, 8 - rd.ioct, 8),
                    qoz = __shfl_sync(o.m, K.z, 8 - rd.ioct, 8);
        const double delx = (double)(q8x - q1x), dely = (double)(q8y - q1y), delz = (double)(q8z - q1z);
        const double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        const double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        const double invdelz = 1.0 / delz;
        double u = (xe - q1x) * invdelx, v = (ye - q1y) * invdely, w = (ze - q1z) * invdelz;
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
        const double sox = ipinx ? (double)1.0e20f : (qox - xe) * rd.cxinv;
        const double soy = ipiny ? (double)1.0e20f : (qoy - ye) * rd.cyinv;
        const double soz = (qoz - ze) * rd.czinv;
        const double so = fmin(fmin(sox, soy), soz);
        if (so < -eps) { err = 1; break; }
        double xn = xe + so * rd.cx, yn = ye + so * rd.cy, zn = ze + so * rd.cz;
        // ---- exit face and next cell; its record is requested before the sub-interval loop ----
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
        CellRec cn = c;
        float snap = 0.0f;
        if (inextcell > 0) {
            cn = load_cell(S, inextcell);
            int pn = cn.gp[0];
#pragma unroll
            for (int n = 1; n < 8; n++) if (rd.ioct - 1 == n) pn = cn.gp[n];
            snap = pt_coord(S, pn, jface);
        }
        u = (xn - q1x) * invdelx; v = (yn - q1y) * invdely; w = (zn - q1z) * invdelz;
        float extn;
        { double fcn[8]; interp_kernel(u, v, w, fcn); extn = (float)fcsum(fcn, e8); }
        const double taugrid = so * 0.5f * (ext1 + extn);
        int ntau = 1 + (int)(taugrid / S.tautol);
        if (ntau < 1) ntau = 1;
        const double dels = so / ntau;
        float bw[NST];                               // SRCSINGSCAT of the own corner in this cell
#pragma unroll
        for (int k = 0; k < NST; k++) bw[k] = 0.0f;
        for (int it = 1; it <= ntau; it++) {
            const double f1 = fown;                  // previous interpolation weight of the own corner
            const double s = it * dels;
            const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
            u = (xi - q1x) * invdelx; v = (yi - q1y) * invdely; w = (zi - q1z) * invdelz;
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
                K.W += transmit * abscell * ((0.5f * (f0 + f1) + 0.08333333333f * (ext0 * f1 - ext1 * f0) * corr) / ext);
                if (exact_ss) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        float ss0 = (float)(f0 * K.ss[k]), ss1 = (float)(f1 * K.ss[k]);
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
                    K.G += ag * transmit * abscell;
                }
                transmit = tnext;
                npassed++;
                nsub++;
                if (npassed > G.maxsub) { err = 4; break; }
            } else {
#pragma unroll
                for (int k = 0; k < NST; k++) bw[k] = 0.0f;
            }
            ext1 = ext0; ext1d = ext0d;
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1[k] = srcext0[k];
        }
        if (err) break;
        if (exact_ss) {
            double bsum = 0.0;
#pragma unroll
            for (int k = 0; k < NST; k++) bsum += adj[k] * (double)bw[k];
            K.B += bsum;
        }
        if (inextcell > 0) {
            if (jface == 1) xn = (double)snap;
            else if (jface == 2) yn = (double)snap;
            else zn = (double)snap;
        }
        if (transmit < S.transcut) {
            done = true;
        } else if (inextcell == 0 && iface >= 5) {
            done = true;
            float radbnd[NST];
            int boundpts[4]; double boundinterp[4], dirrad1[4];
            const int e = boundary_radiance<NST, true>(S, xn, yn, (float)mu2, rd.phi2, sky, ic, kface, radbnd,
                                                       boundpts, boundinterp, dirrad1);
            if (e) { err = e; break; }
            if (exact_ss && o.ol < 4) {
                int bp = boundpts[0]; double bi = boundinterp[0], dr = dirrad1[0];
#pragma unroll
                for (int n = 1; n < 4; n++) if (o.ol == n) { bp = boundpts[n]; bi = boundinterp[n]; dr = dirrad1[n]; }
                bnd_val = adj[0] * transmit * bi * dr;
                if (bnd_val != 0.0) bnd_pt = bp;
            }
        } else {
            icell = inextcell; c = cn;
        }
        jf = jface;
        xe = xn; ye = yn; ze = zn;
    }
    // the corners of the last cell
    if (any_cell && !err) {
        const bool keep = K.W != 0.0 || K.G != 0.0 || K.B != 0.0;
        const unsigned kb = oct_ballot(o, keep);
        if (keep) {
            const int k = nrec + __popc(kb & ((1u << o.ol) - 1));
            if (k < cap) { write_visit<NST>(rec + k, K); mypairs += visit_pairs(G, K.pt); } else err = 5;
        }
        nrec += __popc(kb);
    }
    // the four surface points of a ray that ends on the ground carry the direct-beam part of the reflected radiance
    // (FIND_BOUNDARY_RADIANCE_GRAD): four records with a beam weight only
    const unsigned bm = oct_ballot(o, bnd_pt > 0);
    if (any_cell && !err && bm) {
        if (bnd_pt > 0) {
            const int k = nrec + __popc(bm & ((1u << o.ol) - 1));
            if (k < cap) {
                GCorner<NST> Z;
                Z.pt = bnd_pt; Z.W = 0.0; Z.G = 0.0; Z.B = bnd_val;
#pragma unroll
                for (int q = 0; q < NST; q++) Z.srcfull[q] = 0.0f;
                write_visit<NST>(rec + k, Z); mypairs += visit_pairs(G, bnd_pt);
            } else err = 5;
        }
        nrec += __popc(bm);
    }
    {
        int t = mypairs;
        t += __shfl_xor_sync(o.m, t, 1, 8); t += __shfl_xor_sync(o.m, t, 2, 8); t += __shfl_xor_sync(o.m, t, 4, 8);
        npairs = t;
    }
    err = __reduce_max_sync(o.m, err);
    if (S.counts && o.ol == 0) {
        atomicAdd(&S.counts[0], (unsigned long long)ntrace);
        atomicAdd(&S.counts[1], (unsigned long long)npt_eval);
        atomicAdd(&S.counts[2], (unsigned long long)nsh_eval);
        atomicAdd(&S.counts[4], (unsigned long long)nsub);
        atomicAdd(&S.counts[5], 1ull);
    }
    return err;
}

// shell sums of YLMSUN*YLMDIR for the untruncated solar terms (no delta-M only)
template <int NST>
__device__ __forceinline__ void ray_vsh(const DevState &S, const float *Ysh, float *Vsh, const Oct &o)
{
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

// per-ray setup shared by the two phases; returns false when the ray contributes nothing
template <int NST>
__device__ __forceinline__ bool ray_setup(const DevState &S, int iray, const float *camx, const float *camy,
                                          const float *camz, const double *cammu, const double *camphi,
                                          const RayPack *packs, const int *raypix, const double *adjw,
                                          const double *ray_weights, const double *stokes_weights, float *Ysh,
                                          float *Vsh, const Oct &o, RayErr *err, RayPack &pk, RayDir &rd,
                                          double &mu2, double &phi2, double (&adj)[NST])
{
    mu2 = __ldg(&cammu[iray]); phi2 = __ldg(&camphi[iray]);
    pk = dev_get_pack(S, packs, iray, camx, camy, camz, mu2, phi2);
    if (pk.status == 2) { if (o.ol == 0) set_err(err, 2, iray); return false; }
    if (pk.status != 0) return false;
    const int pix = __ldg(&raypix[iray]);
    const double rw = __ldg(&ray_weights[iray]);
#pragma unroll
    for (int k = 0; k < NST; k++)
        adj[k] = __ldg(&adjw[k + NST * (size_t)pix]) * rw * __ldg(&stokes_weights[k + NST * (size_t)pix]);
    dev_ray_dir(S, pk, rd); rd.phi2 = (float)phi2;
    __syncwarp(o.m);
    group_ylmall(S, (float)mu2, (float)phi2, Ysh, o.ol, AT3D_OCT, o.m);
    if (!S.deltam) ray_vsh<NST>(S, Ysh, Vsh, o);
    return true;
}

template <int NST>
__global__ void __launch_bounds__(AT3D_RAY_THREADS, NST == 1 ? AT3D_MINB_ADJ1 : AT3D_MINB_ADJ3)
weights_kernel(DevState S, DevGrad G, int nrays, const float *camx, const float *camy,
               const float *camz, const double *cammu, const double *camphi, const RayPack *packs,
               const int *raypix, const double *adjw /*[NST,npix]*/, const double *ray_weights,
               const double *stokes_weights, const double *total /*[NST,nrays]*/,
               const long long *recoff /*[nrays+1]*/, long long rec_base, VisitRec *recs, int *nrec_out,
               int *npairs_out, int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub, RayErr *err,
               int *ray_counter, int ray0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Oct o = oct_id();
    const int lane = threadIdx.x & 31;
    const int ysz = S.ny_comp * S.nlmp, vsz = (3 * (S.ml + 1) + 3) & ~3;
    float *Ysh = (float *)smem_raw + (size_t)(threadIdx.x >> 3) * (ysz + vsz);
    float *Vsh = Ysh + ysz;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ray_counter, 32 / AT3D_OCT);
        base = __shfl_sync(FULLMASK, base, 0);
        if (ray0 + base >= nrays) break;
        const int iray = ray0 + base + (lane >> 3);
        if (iray < nrays) {
            RayPack pk; RayDir rd; double mu2, phi2, adj[NST];
            int ntrace = 0, nsub = 0, nrec = 0, npairs = 0;
            if (ray_setup<NST>(S, iray, camx, camy, camz, cammu, camphi, packs, raypix, adjw, ray_weights,
                               stokes_weights, Ysh, Vsh, o, err, pk, rd, mu2, phi2, adj)) {
                double tot[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) tot[k] = __ldg(&total[k + NST * (size_t)iray]);
                const float sky = (-mu2 > 0.0) ? dev_sky_radiance(S, (float)mu2, (float)phi2) : 0.0f;
                const long long r0 = __ldg(&recoff[iray]), r1 = __ldg(&recoff[iray + 1]);
                const int e = march_weights<NST>(S, G, Ysh, Vsh, rd, mu2, pk.x0, pk.y0, pk.z0, sky, adj, tot, o,
                                                 recs + (r0 - rec_base), (int)(r1 - r0),
                                                 trace_cells ? trace_cells + (size_t)trace_cap * iray : nullptr,
                                                 trace_cap, ntrace, nsub, nrec, npairs);
                if (e && o.ol == 0) set_err(err, e, iray);
                if (e) { nrec = 0; npairs = 0; }
            }
            if (o.ol == 0) {
                nrec_out[iray] = nrec;
                npairs_out[iray] = npairs;
                if (trace_n) { trace_n[iray] = ntrace; trace_nsub[iray] = nsub; }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Phase B: COMPUTE_SOURCE_GRAD_1CELL's ray-dependent remainder for every visit record of a ray
// (shdomsub4.f:1786-2019 with the tables of grad_prep_kernel) and the scatter into GRADOUT
// (shdomsub4.f:3751-3764, 4101-4109).  One octet per ray.
// ------------------------------------------------------------------------------------------
// destination of the contributions: PAIRS: (key, value) slots of the ray (sorted by key and summed per key afterwards:
// no atomics, one owner per GRADOUT entry, deterministic); otherwise red.global.add.f64 (Jacobian path)
struct PairOut {
    unsigned *keys;         // slot -> entry of GRADOUT (< ngrad) or ngrad + grid point (BEAM_WEIGHT)
    double *vals;
    unsigned ngrad;
};

template <int NST, bool PAIRS>
__device__ __forceinline__ void apply_record(const DevState &S, const DevGrad &G, const VisitRec &rc,
                                             const float *Ysh, const float *Vsh, const RayDir &rd,
                                             const double (&adj)[NST], const Oct &o, double *gradout,
                                             double *beam_weight, const PairOut &po, long long &slot,
                                             int &nrh_eval)
{
    const int nlmp = S.nlmp, ml = S.ml, nd = G.numder;
    const bool deltam = S.deltam != 0;
    const int ipz = rc.ip - 1;
    const int4 gp = __ldg(&G.gptrec[ipz]);
    const int nnz = gp.y >> 16, nrows = gp.y & 0xFFFF, nrp = (gp.w & 0xFF) * 32;
    const float *dshp = G.dsh + (size_t)(unsigned)gp.z * 32;
    nrh_eval += gp.w >> 8;
    float dirflux = 0.0f, secmu0 = 0.0f;
    if ((!deltam || S.npart > 1) && S.srctype != 'T') {
        dirflux = __ldg(&S.dirflux[ipz]);
        secmu0 = (float)(1.0 / fabs((double)S.solarmu));
    }
    int last_idr = -1, last_ipa = -1;
    float sourcet[NST], ssj[NST];        // SOURCET; lane-partial SINGSCATJ
#pragma unroll
    for (int k = 0; k < NST; k++) { sourcet[k] = 0.0f; ssj[k] = 0.0f; }
    for (int r = 0; r < nrows; r++) {
        const int idr = r / nnz;
        const int4 *rw = G.rowrec + ((size_t)gp.x + r) * G.prow_stride;
        const int4 h0 = __ldg(rw), h1 = __ldg(rw + 1);
        if (idr != last_idr) {
            last_idr = idr;
            const int ipa = __ldg(&G.partder[idr]);
            if (ipa != last_ipa) {
                last_ipa = ipa;
                const int4 *sp = G.sprec + ((size_t)ipz * nd + idr) * G.sp_stride;
                const int4 s0 = __ldg(sp);
                const float alb_ip = __int_as_float(s0.x), f = __int_as_float(s0.y), scatterj = __int_as_float(s0.z);
#pragma unroll
                for (int k = 0; k < NST; k++) ssj[k] = 0.0f;
                for (int e = o.ol; e < s0.w; e += 8) {
                    const int2 en = __ldg((const int2 *)(sp + 1) + e);
                    float sv[NST];
                    ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, en.x, rd, sv);
                    const float cc = __int_as_float(en.y);
#pragma unroll
                    for (int k = 0; k < NST; k++) ssj[k] = ssj[k] + cc * sv[k];
                }
                if (S.npart == 1) {
#pragma unroll
                    for (int k = 0; k < NST; k++) sourcet[k] = (alb_ip > 1e-8f) ? rc.srcfull[k] / alb_ip : 0.0f;
                } else {
                    // SINGSCATJ is needed on its own; afterwards only lane 0 carries it into the row sums
                    float sj[NST];
#pragma unroll
                    for (int k = 0; k < NST; k++) { sj[k] = oct_sum(o.m, ssj[k]); ssj[k] = (o.ol == 0) ? sj[k] : 0.0f; sourcet[k] = 0.0f; }
                    if (scatterj > G.scatmin) {
                        // COMPUTE_SOURCE_DIRECTION with the mixed table: the SOURCET row of this unknown
                        float acc[NST];
#pragma unroll
                        for (int k = 0; k < NST; k++) acc[k] = 0.0f;
                        sh_dot_partial<NST>(dshp + (size_t)(nnz * nd + idr) * NST * nrp, nrp, Ysh, nlmp, o, acc);
                        if (!deltam) {
                            const float *lg = G.legs + ((size_t)ipz * nd + idr) * G.ntup;
                            for (int l = o.ol; l <= ml; l += 8) {
                                acc[0] = acc[0] + dirflux * secmu0 * __ldg(&lg[G.ncomp * l]) * Vsh[l];
                                if (NST > 1) acc[1] = acc[1] + dirflux * secmu0 * __ldg(&lg[3 + G.ncomp * l]) * Vsh[(ml + 1) + l];
                            }
                        }
#pragma unroll
                        for (int k = 0; k < NST; k++) sourcet[k] = oct_sum(o.m, acc[k]);
                        if (deltam) {
#pragma unroll
                            for (int k = 0; k < NST; k++) sourcet[k] = sourcet[k] + dirflux * sj[k] * secmu0 / (1 - f);
                        }
                    }
                }
                sourcet[0] = fmaxf(0.0f, sourcet[0]);
            }
        }
        const float c_src = __int_as_float(h0.x), cj = __int_as_float(h0.y);
        const int nb = h1.x & 0xFF, npl = h1.x >> 8;
        // DSOURCE: the row's SH block in the ray direction (lane partial)
        float dp[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) dp[k] = cj * ssj[k];
        sh_dot_partial<NST>(dshp + (size_t)r * NST * nrp, nrp, Ysh, nlmp, o, dp);
        // single-scatter terms of the row (SINGSCATP, DSINGSCATP with their factors)
        for (int e = o.ol; e < npl; e += 8) {
            const int2 en = __ldg((const int2 *)(rw + 2) + e);
            float sv[NST];
            if (en.x & AT3D_DTAB_FLAG)
                ray_singscat<NST>(G.dphasetab, S.nstphase, G.dnumphase, en.x & ~AT3D_DTAB_FLAG, rd, sv);
            else
                ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, en.x, rd, sv);
            const float cc = __int_as_float(en.y);
#pragma unroll
            for (int k = 0; k < NST; k++) dp[k] = dp[k] + cc * sv[k];
        }
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NST; k++) v += adj[k] * (double)dp[k];
        if (!deltam) {
            // untruncated solar term of COMPUTE_SOURCE_DIRECTION applied to DLEGT (SOURCET(1), SOURCET(2) only)
            const float xi = __ldg(&G.optinterpwt[nb + 8 * (size_t)ipz]);
            const float *dlg = G.dlegt + ((size_t)gp.x + r) * G.ntup;
            for (int l = o.ol; l <= ml; l += 8) {
                v += adj[0] * (double)(xi * (dirflux * secmu0 * __ldg(&dlg[G.ncomp * l]) * Vsh[l]));
                if (NST > 1) v += adj[1] * (double)(xi * (dirflux * secmu0 * __ldg(&dlg[3 + G.ncomp * l]) * Vsh[(ml + 1) + l]));
            }
        }
        v = oct_sum_d(o.m, v);
#pragma unroll
        for (int k = 0; k < NST; k++) v += adj[k] * (double)(c_src * sourcet[k]);
        v += adj[0] * (double)__int_as_float(h1.y);             // thermal component (0 for solar sources)
        if (o.ol == 0) {
            const double val = rc.W * v + (double)__int_as_float(h0.z) * rc.G;
            const size_t dst = (size_t)(h0.w - 1) + (size_t)G.maxpg * idr;
            if (PAIRS) { po.keys[slot + r] = (unsigned)dst; po.vals[slot + r] = val; }
            else if (val != 0.0) atomicAdd(&gradout[dst], val);
        }
    }
    if (o.ol == 0) {
        if (PAIRS) { po.keys[slot + nrows] = po.ngrad + (unsigned)ipz; po.vals[slot + nrows] = rc.B; }
        else if (rc.B != 0.0) atomicAdd(&beam_weight[ipz], rc.B);
    }
    slot += nrows + 1;
}

#ifndef AT3D_MINB_APPLY
#define AT3D_MINB_APPLY 6
#endif
template <int NST, bool PAIRS>
__global__ void __launch_bounds__(AT3D_RAY_THREADS, (NST == 1 ? AT3D_MINB_APPLY : 1))
apply_kernel(DevState S, DevGrad G, int nrays, const float *camx, const float *camy,
             const float *camz, const double *cammu, const double *camphi, const RayPack *packs,
             const int *raypix, const double *adjw, const double *ray_weights, const double *stokes_weights,
             const long long *recoff, long long rec_base, const VisitRec *recs, const int *nrec_in, double *gradout,
             double *beam_weight, PairOut po, const long long *pairoff /*[nrays - ray0 + 1], PAIRS only*/,
             RayErr *err, int *ray_counter, int ray0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Oct o = oct_id();
    const int lane = threadIdx.x & 31;
    const int ysz = S.ny_comp * S.nlmp, vsz = (3 * (S.ml + 1) + 3) & ~3;
    float *Ysh = (float *)smem_raw + (size_t)(threadIdx.x >> 3) * (ysz + vsz);
    float *Vsh = Ysh + ysz;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ray_counter, 32 / AT3D_OCT);
        base = __shfl_sync(FULLMASK, base, 0);
        if (ray0 + base >= nrays) break;
        const int iray = ray0 + base + (lane >> 3);
        if (iray < nrays) {
            const int nrec = __ldg(&nrec_in[iray]);
            RayPack pk; RayDir rd; double mu2, phi2, adj[NST];
            if (nrec > 0 && ray_setup<NST>(S, iray, camx, camy, camz, cammu, camphi, packs, raypix, adjw, ray_weights,
                                           stokes_weights, Ysh, Vsh, o, err, pk, rd, mu2, phi2, adj)) {
                const VisitRec *rp = recs + (__ldg(&recoff[iray]) - rec_base);
                int nrh = 0;
                long long slot = PAIRS ? __ldg(&pairoff[iray - ray0]) : 0;
                for (int i = 0; i < nrec; i++) {
                    VisitRec rc;
                    const int4 a = __ldg((const int4 *)rp);
                    const double2 b = __ldg((const double2 *)rp + 1);
                    const double2 c = __ldg((const double2 *)rp + 2);
                    rc.ip = a.x; rc.srcfull[0] = __int_as_float(a.y); rc.srcfull[1] = __int_as_float(a.z);
                    rc.srcfull[2] = __int_as_float(a.w); rc.W = b.x; rc.G = b.y; rc.B = c.x;
                    apply_record<NST, PAIRS>(S, G, rc, Ysh, Vsh, rd, adj, o, gradout, beam_weight, po, slot, nrh);
                    rp++;
                }
                if (S.counts && o.ol == 0) atomicAdd(&S.counts[3], (unsigned long long)nrh);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Phase 1 tail / Phase 2: ray -> pixel accumulation (shdomsub4.f:693-696), COMPUTE_ADJOINT_WEIGHTS
// ------------------------------------------------------------------------------------------
template <int NST>
__global__ void pixel_kernel(int pix0, int npix, const int *pixstart, const int *rays_per_pixel, const double *visrad,
                             const double *ray_weights, const double *stokes_weights, const float *measurements,
                             const double *unc, int nunc, int costfunc_ll, float *stokesout, double *adjw,
                             double *costp, int *raypix)
{
    const int p = pix0 + blockIdx.x * blockDim.x + threadIdx.x;      // pixels [pix0, npix)
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
// with a non-zero beam weight walks its zero-terminated DPTR/DPATH list.  PAIRS: the contributions go to the
// point's (key, value) slots (beam_count_kernel sized them) and are summed per GRADOUT entry by pair_sum_kernel.
__global__ void beam_count_kernel(DevGrad G, int npts, const double *beam_weight, int *count)
{
    const int ip = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ip >= npts) return;
    int len = 0;
    if (beam_weight[ip] != 0.0) {
        const int *dptr = G.dptr + (size_t)G.longest_path_pts * ip;
        for (int base = 0; base < G.longest_path_pts; base += 32) {
            const int ii = base + lane;
            const int ib = ii < G.longest_path_pts ? __ldg(&dptr[ii]) : 0;
            const unsigned stop = __ballot_sync(FULLMASK, ib <= 0);
            len += stop ? __ffs(stop) - 1 : 32;
            if (stop) break;
        }
    }
    if (lane == 0) count[ip] = len * G.numder;
}

template <bool PAIRS>
__global__ void beam_kernel(DevGrad G, int npts, const double *beam_weight, double *gradout, const long long *pairoff,
                            unsigned *keys, double *vals)
{
    const int ip = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ip >= npts) return;
    const double bwt = beam_weight[ip];
    if (bwt == 0.0) return;
    const float *dpath = G.dpath + (size_t)G.longest_path_pts * ip;
    const int *dptr = G.dptr + (size_t)G.longest_path_pts * ip;
    const long long slot0 = PAIRS ? pairoff[ip] : 0;
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
                const size_t dst = (size_t)(ib - 1) + (size_t)G.maxpg * idr;
                if (PAIRS) { keys[slot0 + (size_t)ii * G.numder + idr] = (unsigned)dst; vals[slot0 + (size_t)ii * G.numder + idr] = -val; }
                else if (val != 0.0) atomicAdd(&gradout[dst], -val);
            }
        }
        if (stop) break;
    }
}

// Streaming variant of phase 4 (DevGrad::stream_beam): no DPATH/DPTR lists in memory.  A thread per grid point with a
// non-zero beam weight repeats the walk of DIRECT_BEAM_AND_PATHS_PROP (shdomsub5.f:1646-2003; beam_walk<true>) and hands
// every (property point, path integral) entry straight to the accumulation, entry by entry in the order the dense list
// would hold them: the values and their summation order are those of beam_kernel, so the gradient is the same bit for bit.
__global__ void beam_stream_count_kernel(DevGrad G, int npts, const float4 *ptrec, const double *beam_weight, int *count, RayErr *err)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npts) return;
    int n = 0;
    if (beam_weight[ip] != 0.0) {
        BeamCountSink sink{0};
        double path; int npp;
        const float4 pt = __ldg(&ptrec[ip]);
        const int e = beam_walk<true>(G.bg, G.bzl, pt.x, pt.y, pt.z, nullptr, path, npp, sink);
        if (e) { if (atomicCAS(&err->code, 0, 5) == 0) err->ray = ip + 1; }
        n = sink.n;
    }
    count[ip] = n * G.numder;
}

template <bool PAIRS>
struct BeamGradSink {
    const DevGrad &G; double bwt; long long slot; unsigned *keys; double *vals; double *gradout;
    __device__ bool put(const int *p, const float *v)
    {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int ib = p[c];
            const double pb = (double)v[c] * bwt;
            for (int idr = 0; idr < G.numder; idr++) {
                const size_t dst = (size_t)(ib - 1) + (size_t)G.maxpg * idr;
                const float dm = __ldg(&G.dextm[dst]);
                const double val = dm * pb;
                if (PAIRS) { keys[slot] = (unsigned)dst; vals[slot] = -val; slot++; }
                else if (val != 0.0) atomicAdd(&gradout[dst], -val);
            }
        }
        return true;
    }
};

template <bool PAIRS>
__global__ void beam_stream_kernel(DevGrad G, int p0, int p1, const float4 *ptrec, const double *beam_weight, double *gradout,
                                   const long long *pairoff, unsigned *keys, double *vals)
{
    const int ip = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= p1) return;
    const double bwt = beam_weight[ip];
    if (bwt == 0.0) return;
    BeamGradSink<PAIRS> sink{G, bwt, PAIRS ? pairoff[ip] - pairoff[p0] : 0, keys, vals, gradout};
    double path; int npp;
    const float4 pt = __ldg(&ptrec[ip]);
    beam_walk<true>(G.bg, G.bzl, pt.x, pt.y, pt.z, nullptr, path, npp, sink);
}

// ------------------------------------------------------------------------------------------
// Atomics-free accumulation of the (GRADOUT entry, value) pairs of the derivative pass: the pairs are sorted by entry
// (cub radix sort, stable: equal keys keep their slot order, which is fixed by the ray order), pair_bounds_kernel finds
// the run of every entry and pair_sum_kernel gives every entry to ONE warp, which adds the run in a fixed order
// (lane-strided partial sums, then a fixed shuffle tree).  The result does not depend on scheduling: the gradient is
// reproducible bit for bit.
// ------------------------------------------------------------------------------------------
__global__ void pair_bounds_kernel(long long n, const unsigned *keys, unsigned nkeys, long long *lo, long long *hi)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys[i];
    if (k >= nkeys) return;
    if (i == 0 || keys[i - 1] != k) lo[k] = i;
    if (i == n - 1 || keys[i + 1] != k) hi[k] = i + 1;
}

__global__ void pair_sum_kernel(unsigned nkeys, unsigned ngrad, const long long *lo, const long long *hi, const double *vals,
                                double *gradout, double *beam_weight)
{
    const unsigned k = (unsigned)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= nkeys) return;
    const long long a = lo[k], b = hi[k];
    if (b <= a) return;
    double acc = 0.0;
    for (long long i = a + lane; i < b; i += 32) acc += vals[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULLMASK, acc, d);
    if (lane == 0) {
        if (k < ngrad) gradout[k] += acc; else beam_weight[k - ngrad] += acc;
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
    CUDA_TRY(at3d_malloc(&p, n * sizeof(T)));
    owned.push_back(p);
    st->bytes += n * sizeof(T);
    CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T *)p;
    return 0;
}

static inline int nst_of(const DevState &S) { return S.nstokes; }

extern "C" int at3d_state_attach_gradient(at3d_state *st, const at3d_grad_desc *g, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!st || !g) { set_msg(errmsg, "null argument"); return 1; }
    std::lock_guard<std::mutex> lock(st->mu);
    const bool attach_timing = getenv("AT3D_B200_ATTACH_TIMING") != nullptr;
    auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_attach0 = tnow();
    if (!st->S.radrec) { set_msg(errmsg, "the state was created without RADIANCE/RSHPTR: the gradient needs them"); return 1; }
    if (g->numder < 1) { set_msg(errmsg, "NUMDER must be >= 1"); return 1; }
    const bool solar = st->S.srctype != 'T', thermal = st->S.srctype != 'S';
    if (thermal) {
        if (st->S.units == 'B') { set_msg(errmsg, "thermal gradient: band-integrated Planck units (UNITS='B') are not implemented"); return 3; }
        if (!st->S.temp) { set_msg(errmsg, "thermal gradient: the state was created without TEMP"); return 1; }
        // surface emission enters FIND_BOUNDARY_RADIANCE_GRAD through SFCGRIDRAD (shdomsub4.f:2274-2290): zero for the
        // Lambertian surfaces the gradient supports
        if (st->S.sfcgridrad) { set_msg(errmsg, "thermal gradient: SFCGRIDRAD must be zero"); return 3; }
    }
    // the reference itself stops here: SURFACE_BRDF_GRAD has no linearisation for W/D/O/R surfaces (surface.f:395-399)
    if (st->S.sfctype1 != 'L') { set_msg(errmsg, "the gradient needs a Lambertian surface (SFCTYPE 'FL','VL')"); return 3; }
    if (g->deriv_maxnmicro > st->S.maxnmicro) { set_msg(errmsg, "DERIV_MAXNMICRO > MAXNMICRO is not supported"); return 3; }
    // drop a previous attachment
    for (void *p : st->grad_owned) at3d_free(p);
    st->grad_owned.clear();
    st->grad_attached = 0;
    DevGrad &G = st->G;
    memset(&G, 0, sizeof(G));
    const DevState &S = st->S;
    G.maxpg = g->maxpg; G.numder = g->numder; G.dnumphase = g->dnumphase;
    G.deriv_maxnmicro = g->deriv_maxnmicro; G.pmaxnmicro = S.maxnmicro;
    G.longest_path_pts = g->longest_path_pts;
    // the direct-beam derivative exists for solar sources only (shdomsub4.f:793-794)
    G.exact_single_scatter = g->exact_single_scatter && solar; G.singlescatter = g->singlescatter;
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
    const bool stream_beam = G.exact_single_scatter && !g->dpath && !g->dptr;
    if (thermal && g->dtemp) GUP(dtemp, mp * nd);
    if (G.exact_single_scatter && !stream_beam) { GUP(dpath, (size_t)g->longest_path_pts * np); GUP(dptr, (size_t)g->longest_path_pts * np); }
#undef GUP
    if (rc) return rc;
    if (stream_beam) {
        if (!g->beam_d || !g->beam_i || !g->beam_zlevels || g->beam_npx < 1 || g->beam_npy < 1 || g->beam_npz < 2 ||
            (size_t)g->beam_npx * g->beam_npy * g->beam_npz != mp) {
            set_msg(errmsg, "streaming direct-beam derivative: beam_d / beam_i / beam_zlevels and the property-grid size are required");
            return 1;
        }
        BeamGeom &b = G.bg;
        b.bcflag = S.bcflag; b.npx = g->beam_npx; b.npy = g->beam_npy; b.npz = g->beam_npz;
        b.xstart = g->beam_xstart; b.ystart = g->beam_ystart;
        b.cx = g->beam_d[0]; b.cy = g->beam_d[1]; b.cz = g->beam_d[2];
        b.cxinv = g->beam_d[3]; b.cyinv = g->beam_d[4]; b.czinv = g->beam_d[5];
        b.epss = g->beam_d[6]; b.epsz = g->beam_d[7]; b.xdomain = g->beam_d[8]; b.ydomain = g->beam_d[9];
        b.delxd = g->beam_d[11]; b.delyd = g->beam_d[12];
        b.ipdirect = g->beam_i[0]; b.di = g->beam_i[1]; b.dj = g->beam_i[2]; b.dk = g->beam_i[3];
        rc = gupload(st, own, g->beam_zlevels, (size_t)g->beam_npz, &G.bzl, errmsg);
        if (rc) return rc;
        G.stream_beam = 1;
    }
    if (!G.partder || !G.doexact || !G.dext || !G.dalb || !G.dextm || !G.dalbm || !G.dfj || !G.optinterpwt ||
        !G.interpptr || !G.iphasep || !G.phasewtp || !G.extinctp || !G.albedop || !G.diphasep || !G.dphasewtp ||
        !G.dleg || !G.dphasetab) {
        set_msg(errmsg, "at3d_state_attach_gradient: a required derivative array is NULL");
        return 1;
    }
    if (G.exact_single_scatter && !G.stream_beam && (!G.dpath || !G.dptr)) { set_msg(errmsg, "EXACT_SINGLE_SCATTER needs DPATH and DPTR (or neither: streaming)"); return 1; }
    const double t_attach1 = tnow();
    // ray-independent tables of COMPUTE_SOURCE_GRAD_1CELL (grad_prep_kernel)
    {
        const bool deltam = S.deltam != 0;
        G.ncomp = S.nstleg == 1 ? 1 : 4;
        G.ntup = (G.ncomp * (S.ml + 1) + 3) & ~3;
        const int pmaxp = S.maxnmicro + g->deriv_maxnmicro;            // row list: SINGSCATP + DSINGSCATP entries
        G.prow_stride = 2 + (pmaxp + 1) / 2;
        G.sp_stride = 1 + (8 * S.maxnmicro + 1) / 2;                   // species list: SINGSCATJ entries
        if ((int)st->nr_h.size() != S.npts) { set_msg(errmsg, "the state has no RADIANCE/RSHPTR"); return 1; }
        // row / SH-block offsets per grid point (host prefix sums)
        std::vector<int4> gp(np);
        size_t nrows_tot = 0, units = 0;
        const int extra = S.npart > 1 ? (int)nd : 0;                    // SOURCET rows
        for (size_t i = 0; i < np; i++) {
            int nnz = 0;
            for (int nb = 0; nb < 8; nb++) if (g->optinterpwt[nb + 8 * i] >= 1e-7f) nnz++;
            const int nr = st->nr_h[i], nrp32 = AT3D_SHPAD(nr) / 32;
            const size_t nrows = (size_t)nnz * nd;
            if (nnz == 0) nnz = 1;
            if (nrows > 0xFFFF || nrp32 > 0xFF) { set_msg(errmsg, "too many gradient rows per grid point"); return 3; }
            gp[i] = make_int4((int)nrows_tot, (int)nrows | (nnz << 16), (int)(unsigned)units, nrp32 | (nr << 8));
            nrows_tot += nrows;
            units += (nrows + extra) * (size_t)nst_of(S) * nrp32;
            if (nrows_tot >= ((size_t)1 << 31) || units >= ((size_t)1 << 32)) { set_msg(errmsg, "gradient tables too large"); return 2; }
        }
        const int4 *gp_d = nullptr;
        rc = gupload(st, own, gp.data(), np, &gp_d, errmsg); if (rc) return rc;
        G.gptrec = gp_d;
        int4 *rowrec = nullptr, *sprec = nullptr; float *dsh = nullptr, *dlegt = nullptr, *legs = nullptr;
        void *p = nullptr;
        size_t nb_ = (nrows_tot + 1) * G.prow_stride * sizeof(int4);
        CUDA_TRY(at3d_malloc(&p, nb_)); own.push_back(p); rowrec = (int4 *)p; st->bytes += nb_;
        nb_ = np * nd * G.sp_stride * sizeof(int4);
        CUDA_TRY(at3d_malloc(&p, nb_)); own.push_back(p); sprec = (int4 *)p; st->bytes += nb_;
        nb_ = (units + 1) * 32 * sizeof(float);
        CUDA_TRY(at3d_malloc(&p, nb_)); own.push_back(p); dsh = (float *)p; st->bytes += nb_;
        if (!deltam) {
            nb_ = (nrows_tot + 1) * G.ntup * sizeof(float);
            CUDA_TRY(at3d_malloc(&p, nb_)); own.push_back(p); dlegt = (float *)p; st->bytes += nb_;
            if (S.npart > 1) {
                nb_ = np * nd * G.ntup * sizeof(float);
                CUDA_TRY(at3d_malloc(&p, nb_)); own.push_back(p); legs = (float *)p; st->bytes += nb_;
            }
        }
        const int wpb = 4;
        const size_t smem = (size_t)wpb * (nlt + G.ntup) * sizeof(float);
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(grad_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double t_attach2 = tnow();
        grad_prep_kernel<<<(S.npts + wpb - 1) / wpb, wpb * 32, smem>>>(S, G, rowrec, sprec, dsh, dlegt, legs);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        G.rowrec = rowrec; G.sprec = sprec; G.dsh = dsh; G.dlegt = dlegt; G.legs = legs;
        if (attach_timing)
            fprintf(stderr, "[attach] uploads %.2f ms, row tables (host prefix sums, allocations) %.2f ms, grad_prep_kernel %.2f ms\n",
                    t_attach1 - t_attach0, t_attach2 - t_attach1, tnow() - t_attach2);
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

struct IntToLL { __host__ __device__ long long operator()(int v) const { return (long long)v; } };
// visit records of a ray: one per visit + up to four surface records (FIND_BOUNDARY_RADIANCE_GRAD's beam weights)
struct VisitsToLL { __host__ __device__ long long operator()(int v) const { return (long long)v + 4; } };

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// unit adjoint weights of Stokes component k for every pixel (Jacobian path)
__global__ void unit_adj_kernel(int nst, int npix, int k, double *adjw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nst * npix) adjw[i] = (i % nst == k) ? 1.0 : 0.0;
}

// JACOBIAN(k,:,JI,IPIX) = RAYGRAD_PIXEL(k,JACOBIANPTR(JI),:)  (shdomsub4.f:627-630)
__global__ void jac_gather_kernel(int nst, int nd, int njac, int maxpg, int k, int ipix, const int *jacptr,
                                  const double *gtmp, float *jac)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd * njac) return;
    const int idr = i % nd, ji = i / nd;
    jac[k + nst * (idr + nd * ((size_t)ji + (size_t)njac * ipix))] = (float)gtmp[(size_t)(jacptr[ji] - 1) + (size_t)maxpg * idr];
}

// The pair buffers of the derivative pass: keys / values double-buffered for the radix sort, the run bounds per key
// and the sort's scratch, carved out of one allocation.
struct PairBufs {
    unsigned *k0, *k1;
    double *v0, *v1;
    long long *lo, *hi;
    void *tmp;
    size_t tmpb;
};

static int pair_reserve(at3d_state *st, size_t npairs, unsigned nkeys, PairBufs &pb, char *errmsg)
{
    size_t tmpb = 0;
    cub::DoubleBuffer<unsigned> dk((unsigned *)nullptr, (unsigned *)nullptr);
    cub::DoubleBuffer<double> dv((double *)nullptr, (double *)nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, tmpb, dk, dv, (long long)npairs, 0, 32, (cudaStream_t)0);
    size_t o = 0;
    const size_t o_k0 = o; o += al256(sizeof(unsigned) * npairs);
    const size_t o_k1 = o; o += al256(sizeof(unsigned) * npairs);
    const size_t o_v0 = o; o += al256(sizeof(double) * npairs);
    const size_t o_v1 = o; o += al256(sizeof(double) * npairs);
    const size_t o_lo = o; o += al256(sizeof(long long) * nkeys);
    const size_t o_hi = o; o += al256(sizeof(long long) * nkeys);
    const size_t o_tmp = o; o += al256(tmpb);
    CUDA_TRY(st->pairs.reserve(o + 256));
    unsigned char *b = (unsigned char *)st->pairs.p;
    pb.k0 = (unsigned *)(b + o_k0); pb.k1 = (unsigned *)(b + o_k1);
    pb.v0 = (double *)(b + o_v0); pb.v1 = (double *)(b + o_v1);
    pb.lo = (long long *)(b + o_lo); pb.hi = (long long *)(b + o_hi);
    pb.tmp = b + o_tmp; pb.tmpb = tmpb;
    return 0;
}

// GRADOUT(key) += sum of the values with that key (keys >= ngrad: BEAM_WEIGHT(key - ngrad)), in a fixed order
static int pair_accumulate(const PairBufs &pb, size_t npairs, unsigned nkeys, unsigned ngrad, double *grad_d, double *beam,
                           cudaStream_t stream, char *errmsg)
{
    if (npairs == 0) return 0;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)nkeys) bits++;
    cub::DoubleBuffer<unsigned> dk(pb.k0, pb.k1);
    cub::DoubleBuffer<double> dv(pb.v0, pb.v1);
    size_t tmpb = pb.tmpb;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(pb.tmp, tmpb, dk, dv, (long long)npairs, 0, bits, stream));
    CUDA_TRY(cudaMemsetAsync(pb.lo, 0, sizeof(long long) * nkeys, stream));
    CUDA_TRY(cudaMemsetAsync(pb.hi, 0, sizeof(long long) * nkeys, stream));
    pair_bounds_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, stream>>>((long long)npairs, dk.Current(), nkeys, pb.lo, pb.hi);
    pair_sum_kernel<<<(unsigned)(((size_t)nkeys * 32 + 255) / 256), 256, 0, stream>>>(nkeys, ngrad, pb.lo, pb.hi, dv.Current(),
                                                                                     grad_d, beam);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int gradient_impl(at3d_state *st, const at3d_rays *rays, const at3d_grad_desc *g,
                         double *gradout, double *cost, float *stokesout,
                         const at3d_trace *trace, void *cuda_stream, double *kernel_ms,
                         int njac, const int32_t *jacobianptr, float *jacobian, char *errmsg)
{
    std::unique_lock<std::mutex> lock;
    if (st) lock = std::unique_lock<std::mutex>(st->mu);
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
    cub::DeviceScan::ExclusiveSum(nullptr, cubtmp, (const int *)nullptr, (int *)nullptr, (int)(npix > n + 1 ? npix : n + 1), stream);
    {
        size_t t2 = 0;
        cub::TransformInputIterator<long long, IntToLL, const int *> it((const int *)nullptr, IntToLL());
        cub::DeviceScan::ExclusiveSum(nullptr, t2, it, (long long *)nullptr, (int)((n > (size_t)S.npts ? n : (size_t)S.npts) + 1), stream);
        if (t2 > cubtmp) cubtmp = t2;
    }
    o = 0;
    const size_t w_vis = o; o += al256(sizeof(double) * nst * n);
    const size_t w_tot = o; o += al256(sizeof(double) * nst * n);
    const size_t w_adj = o; o += al256(sizeof(double) * nst * npix);
    const size_t w_costp = o; o += al256(sizeof(double) * npix);
    const size_t w_pixstart = o; o += al256(sizeof(int) * (npix + 1));
    const size_t w_raypix = o; o += al256(sizeof(int) * n);
    const size_t w_beam = o; o += al256(sizeof(double) * S.npts);
    const size_t w_cub = o; o += al256(cubtmp);
    const size_t w_npt = o; o += al256(sizeof(int) * (n + 1));
    const size_t w_recoff = o; o += al256(sizeof(long long) * (n + 1));
    const size_t w_nrec = o; o += al256(sizeof(int) * n);
    const size_t w_npairs = o; o += al256(sizeof(int) * (n + 1));
    const size_t w_pairoff = o; o += al256(sizeof(long long) * ((n > (size_t)S.npts ? n : (size_t)S.npts) + 2));
    const size_t w_bcount = o; o += al256(sizeof(int) * ((size_t)S.npts + 1));
    const size_t w_grad = o; o += al256(sizeof(double) * ngrad);
    const size_t w_so = o; o += al256(sizeof(float) * nst * npix);
    const size_t w_cost = o; o += 256;
    CUDA_TRY(st->work.reserve(o + 256));
    unsigned char *wb = (unsigned char *)st->work.p;
    double *visrad = (double *)(wb + w_vis), *total = (double *)(wb + w_tot), *adjw = (double *)(wb + w_adj);
    double *costp = (double *)(wb + w_costp), *beam = (double *)(wb + w_beam);
    int *pixstart = (int *)(wb + w_pixstart), *raypix = (int *)(wb + w_raypix);
    int *npt = (int *)(wb + w_npt), *nrec = (int *)(wb + w_nrec);
    long long *recoff = (long long *)(wb + w_recoff);
    int *npairs = (int *)(wb + w_npairs), *bcount = (int *)(wb + w_bcount);
    long long *pairoff = (long long *)(wb + w_pairoff);
    const size_t nkeys_sz = ngrad + (size_t)S.npts;
    if (nkeys_sz >= 0xFFFFFFF0ull) { set_msg(errmsg, "GRADOUT has too many entries for 32-bit pair keys"); return 2; }
    const unsigned nkeys = (unsigned)nkeys_sz;
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
    if (kernel_ms) { for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]); cudaEventRecord(ev[0], stream); kernel_ms[4] = 0.0; kernel_ms[5] = 0.0; kernel_ms[6] = 0.0; kernel_ms[7] = 0.0; }
    if (n > 0 && npix > 0) {
        // ---- Phase 1: forward radiances (INTEGRATE_1RAY arithmetic for the pixel values; the
        //      ADJOINT_INTEGRATE_1RAY arithmetic for the totals the derivative pass needs) ----
        DevState Sf = S;          // the work counters describe the adjoint pass only
        Sf.counts = nullptr;
        CUDA_TRY(cudaMemsetAsync(st->counts_dev, 0, 8 * sizeof(unsigned long long), stream));
        // NSTOKES=1 with delta-M: the forward pass leaves the corner sources in a stream and the derivative walk is the
        // thread-per-ray one (at3d_gwalk.cuh); AT3D_B200_GRAD_PATH=legacy selects the octet walk, which also serves
        // NSTOKES=3, no delta-M and the Jacobian
        bool use_t = nst == 1 && S.deltam && njac == 0 && tray_block_threads(S) > 0;
        if (const char *e = getenv("AT3D_B200_GRAD_PATH")) { if (!strcmp(e, "legacy")) use_t = false; }
        long long *srcstart = nullptr;
        unsigned *srctop = nullptr;
        size_t src_chunks = 0;
        // orthographic views (runs of rays with one direction) read a per-view source evaluated once per grid point
        std::vector<size_t> seg_start, seg_len;
        std::vector<char> seg_view;
        view_segments(st, rays, seg_start, seg_len, seg_view);
        long long nrec_total = 0;
        for (int attempt = 0;; attempt++) {
            bool at_limit = false;
            if (use_t) {
                // the stream pool is sized by the entries per ray the last call on this state saw
                // (limit: AT3D_B200_SRC_GB, default 45 % of the memory this pool could get, between 4 and 64 GB)
                if (st->src_gb_limit <= 0.0) {              // once per state: cudaMemGetInfo is not a call for every step
                    size_t fr = 0, tot = 0;
                    st->src_gb_limit = 16.0;
                    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
                        const double v = 0.45 * (double)(fr + st->slabs.cap) / 1073741824.0;
                        st->src_gb_limit = v < 4.0 ? 4.0 : (v > 64.0 ? 64.0 : v);
                    }
                }
                double src_gb = st->src_gb_limit;
                if (const char *e = getenv("AT3D_B200_SRC_GB")) { const double v = atof(e); if (v > 0.0) src_gb = v; }
                const double est = st->gw_rec_per_ray > 0.0 ? st->gw_rec_per_ray : (double)(6 * (S.nx + S.ny + S.nz) + 64);
                src_chunks = (size_t)(((double)n * (est * 1.5 + AT3D_SRC_CHUNK) + 4096.0) / AT3D_SRC_CHUNK);
                if (src_chunks < 8 * AT3D_SRC_REGIONS) src_chunks = 8 * AT3D_SRC_REGIONS;
                size_t lim = (size_t)(src_gb * 1073741824.0 / (sizeof(float2) * AT3D_SRC_CHUNK));
                if (lim > 0x7FFFFFF0u) lim = 0x7FFFFFF0u;
                if (src_chunks >= lim) { src_chunks = lim; at_limit = true; }
                CUDA_TRY(st->slabs.reserve(src_chunks * AT3D_SRC_CHUNK * sizeof(float2)));
                const size_t topb = sizeof(unsigned) * AT3D_SRC_TOP_WORDS;
                CUDA_TRY(st->misc.reserve(topb + sizeof(long long) * (n + 1)));
                srctop = (unsigned *)st->misc.p;
                srcstart = (long long *)((unsigned char *)st->misc.p + topb);
                CUDA_TRY(cudaMemsetAsync(srctop, 0, topb, stream));
                Sf.srcpool = (float2 *)st->slabs.p; Sf.srcpool_chunks = (unsigned)src_chunks; Sf.srcpool_top = srctop;
                int dev_ = 0, nsm_ = 148;
                cudaGetDevice(&dev_);
                cudaDeviceGetAttribute(&nsm_, cudaDevAttrMultiProcessorCount, dev_);
                size_t nreg = (n + tray_block_threads(S) - 1) / tray_block_threads(S);     // blocks of the largest launch
                if (nreg > (size_t)nsm_) nreg = (size_t)nsm_;
                if (nreg > AT3D_SRC_REGIONS) nreg = AT3D_SRC_REGIONS;
                Sf.srcpool_nreg = (unsigned)nreg;
                Sf.srcpool_region = (unsigned)(src_chunks / 5 * 4 / nreg);
            } else {
                Sf.srcpool = nullptr;
            }
            CUDA_TRY(cudaMemsetAsync(npt, 0, sizeof(int) * (n + 1), stream));
            for (size_t sgi = 0; sgi < seg_start.size(); sgi++) {
                const size_t s0 = seg_start[sgi], sn = seg_len[sgi];
                DevState Sg = Sf;
                Sg.ray_base = (int)s0;
                if (use_t) Sg.srcstart = srcstart + s0;
                else if (seg_view[sgi]) {
                    CUDA_TRY(st->viewsrc.reserve((size_t)S.npts * nst * sizeof(float)));
                    CUDA_TRY(launch_view_source(Sf, st->packs_h[s0], rays->cammu[s0], rays->camphi[s0], G.singlescatter,
                                                (float *)st->viewsrc.p, stream));
                    Sg.viewsrc = (const float *)st->viewsrc.p;
                }
                CUDA_TRY(launch_forward(Sg, (int)sn, camx ? camx + s0 : nullptr, camy ? camy + s0 : nullptr,
                                        camz ? camz + s0 : nullptr, cammu + s0, camphi + s0, packs ? packs + s0 : nullptr,
                                        nullptr, visrad + (size_t)nst * s0, total + (size_t)nst * s0, 3, 1,
                                        G.singlescatter, 0, G.maxsub, nullptr, 0, nullptr, nullptr, (RayErr *)st->err.p,
                                        st->ray_counter, npt + s0, stream));
            }
            // The rest of phase 1 and phase 2 are queued behind the forward pass without waiting for it: the pool's overflow
            // flag and the total number of visit records come back together, in ONE host round trip per step (pinned
            // buffer, so the copies are asynchronous); an overflow -- rare -- repeats the attempt with a larger pool.
            if (!st->hpin) CUDA_TRY(cudaMallocHost(&st->hpin, sizeof(unsigned) * AT3D_SRC_TOP_WORDS + 64));
            unsigned *htop = (unsigned *)st->hpin;
            long long *htot = (long long *)((unsigned char *)st->hpin + sizeof(unsigned) * AT3D_SRC_TOP_WORDS);
            if (use_t) CUDA_TRY(cudaMemcpyAsync(htop, srctop, sizeof(unsigned) * AT3D_SRC_TOP_WORDS, cudaMemcpyDeviceToHost, stream));
            // visit records of a ray are contiguous: offsets = exclusive scan (64-bit) of the per-ray visit counts
            {
                cub::TransformInputIterator<long long, VisitsToLL, const int *> it(npt, VisitsToLL());
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(wb + w_cub, cubtmp, it, recoff, (int)(n + 1), stream));
            }
            if (kernel_ms) cudaEventRecord(ev[1], stream);
            // ---- Phase 2 ----
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(wb + w_cub, cubtmp, rpp, pixstart, (int)npix, stream));
            const int nb = (int)((npix + 127) / 128);
            if (nst == 1)
                pixel_kernel<1><<<nb, 128, 0, stream>>>(0, (int)npix, pixstart, rpp, visrad, rw, sw, meas, unc, nunc,
                                                          g->costfunc_ll, so_d, adjw, costp, raypix);
            else
                pixel_kernel<3><<<nb, 128, 0, stream>>>(0, (int)npix, pixstart, rpp, visrad, rw, sw, meas, unc, nunc,
                                                          g->costfunc_ll, so_d, adjw, costp, raypix);
            CUDA_TRY(cudaGetLastError());
            cost_reduce_kernel<<<1, 1024, 0, stream>>>((int)npix, costp, cost_d);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(htot, recoff + n, sizeof(long long), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            nrec_total = *htot;
            if (!use_t) break;
            double drawn = (double)htop[0];
            for (int r = 0; r < AT3D_SRC_REGIONS; r++) {
                const double d = (double)htop[32 * (r + 1)];
                drawn += d < (double)Sf.srcpool_region ? d : (double)Sf.srcpool_region;
            }
            const double seen = drawn * AT3D_SRC_CHUNK / (double)n;
            if (getenv("AT3D_B200_DEBUG_POOL")) {
                int full = 0;
                for (int r = 0; r < AT3D_SRC_REGIONS; r++) if ((double)htop[32 * (r + 1)] >= (double)Sf.srcpool_region) full++;
                fprintf(stderr, "[pool] attempt %d chunks %zu region %u nreg %u tail_drawn %u overflow %u full_regions %d seen/ray %.1f est %.1f\n",
                        attempt, src_chunks, Sf.srcpool_region, Sf.srcpool_nreg, htop[0], htop[1], full, seen, st->gw_rec_per_ray);
            }
            if (!htop[1]) { if (seen > st->gw_rec_per_ray) st->gw_rec_per_ray = seen; break; }
            // the pool overflowed: walk again with a pool twice as large; at the memory limit (AT3D_B200_SRC_GB) the octet walk,
            // which evaluates its sources itself, takes over (the radiances of this attempt are complete either way)
            if (at_limit || attempt >= 6) { use_t = false; break; }
            const double est = st->gw_rec_per_ray > 0.0 ? st->gw_rec_per_ray : (double)(6 * (S.nx + S.ny + S.nz) + 64);
            st->gw_rec_per_ray = 2.0 * (seen > est ? seen : est);
        }
        // ---- Phase 3 ----
        // The records of all rays may not fit (cfg4: ~80 GB): the derivative pass runs over chunks of rays whose
        // records fit the budget (AT3D_B200_REC_GB, default 8 GB); chunk boundaries from the offsets on the host, which are
        // only read back when there is more than one chunk (or for the per-pixel passes of the Jacobian).
        double budget_gb = 8.0;
        if (const char *e = getenv("AT3D_B200_REC_GB")) { const double v = atof(e); if (v > 0.0) budget_gb = v; }
        long long budget = (long long)(budget_gb * 1073741824.0 / sizeof(VisitRec));
        {   // the pairs of a chunk are counted with 31 bits (at most 8*NUMDER+1 per record)
            const long long lim = 0x7FFFFFFFll / (8ll * G.numder + 1);
            if (budget > lim) budget = lim;
        }
        const bool one_chunk = nrec_total <= budget && !(njac > 0 && jacobian);
        st->recoff_h.resize(n + 1);
        if (one_chunk) {
            st->recoff_h[0] = 0; st->recoff_h[n] = nrec_total;          // the only entries the single-chunk pass reads
        } else {
            CUDA_TRY(cudaMemcpyAsync(st->recoff_h.data(), recoff, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
        }
        const long long *roff = st->recoff_h.data();
        const size_t smem = (size_t)AT3D_RAYS_PER_BLOCK * (S.ny_comp * S.nlmp + ((3 * (S.ml + 1) + 3) & ~3)) * sizeof(float);
        static thread_local KernelFit fit_w, fit_a, fit_t;
        if (nst == 1) { kernel_fit(fit_w, weights_kernel<1>, AT3D_RAY_THREADS, smem); kernel_fit(fit_a, apply_kernel<1, true>, AT3D_RAY_THREADS, smem); }
        else { kernel_fit(fit_w, weights_kernel<3>, AT3D_RAY_THREADS, smem); kernel_fit(fit_a, apply_kernel<3, true>, AT3D_RAY_THREADS, smem); }
        const int gw_bt = GW_BT;
        const size_t smem_t = gw_smem_per_thread() * gw_bt;
        kernel_fit(fit_t, weights_kernel_t, gw_bt, smem_t);
        const int nsm = fit_t.nsm, per_sm_w = fit_w.per_sm, per_sm_a = fit_a.per_sm, per_sm_t = fit_t.per_sm;
        CUDA_TRY(cudaMemsetAsync(npairs, 0, sizeof(int) * (n + 1), stream));
        float ms_weights = 0.0f, ms_accum = 0.0f;
        size_t r0 = 0;
        while (r0 < n) {
            size_t r1 = r0 + 1;
            if (one_chunk) r1 = n;
            else while (r1 < n && roff[r1 + 1] - roff[r0] <= budget) r1++;
            const long long nrec_chunk = roff[r1] - roff[r0];
            CUDA_TRY(st->recs.reserve(((size_t)nrec_chunk + 8) * sizeof(VisitRec)));
            VisitRec *recs = (VisitRec *)st->recs.p;
            const long want = ((long)(r1 - r0) + AT3D_RAYS_PER_BLOCK - 1) / AT3D_RAYS_PER_BLOCK;
            const long capw = (long)nsm * per_sm_w, capa = (long)nsm * per_sm_a;
            const int nblkw = (int)(want < capw ? want : capw), nblka = (int)(want < capa ? want : capa);
            cudaEvent_t c0 = nullptr, c1 = nullptr, c2 = nullptr, c3 = nullptr;
            if (kernel_ms) { cudaEventCreate(&c0); cudaEventCreate(&c1); cudaEventCreate(&c2); cudaEventCreate(&c3); cudaEventRecord(c0, stream); }
            CUDA_TRY(cudaMemsetAsync(st->ray_counter, 0, sizeof(int), stream));
            if (use_t) {
                DevState St = S;
                St.srcpool = (float2 *)st->slabs.p; St.srcstart = srcstart;
                const long wantt = ((long)(r1 - r0) + gw_bt - 1) / gw_bt, capt = (long)nsm * per_sm_t;
                weights_kernel_t<<<(int)(wantt < capt ? wantt : capt), gw_bt, smem_t, stream>>>(St, G, (int)r1, (int)r0, camx, camy,
                    camz, cammu, camphi, packs, raypix, adjw, rw, sw, total, recoff, roff[r0], recs, nrec, npairs, tc, tcap, tn, ts,
                    (RayErr *)st->err.p, st->ray_counter);
            } else if (nst == 1)
                weights_kernel<1><<<nblkw, AT3D_RAY_THREADS, smem, stream>>>(S, G, (int)r1, camx, camy, camz, cammu, camphi,
                    packs, raypix, adjw, rw, sw, total, recoff, roff[r0], recs, nrec, npairs, tc, tcap, tn, ts,
                    (RayErr *)st->err.p, st->ray_counter, (int)r0);
            else
                weights_kernel<3><<<nblkw, AT3D_RAY_THREADS, smem, stream>>>(S, G, (int)r1, camx, camy, camz, cammu, camphi,
                    packs, raypix, adjw, rw, sw, total, recoff, roff[r0], recs, nrec, npairs, tc, tcap, tn, ts,
                    (RayErr *)st->err.p, st->ray_counter, (int)r0);
            CUDA_TRY(cudaGetLastError());
            if (kernel_ms) cudaEventRecord(c1, stream);
            // pair slots of the chunk's rays: exclusive scan of the per-ray counts (entry r1 is still zero), total to the host
            long long npair_chunk = 0;
            {
                cub::TransformInputIterator<long long, IntToLL, const int *> it(npairs + r0, IntToLL());
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(wb + w_cub, cubtmp, it, pairoff, (int)(r1 - r0 + 1), stream));
                CUDA_TRY(cudaMemcpyAsync(&npair_chunk, pairoff + (r1 - r0), sizeof(long long), cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
            }
            PairBufs pb;
            if ((rc = pair_reserve(st, (size_t)npair_chunk + 1, nkeys, pb, errmsg))) return rc;
            PairOut po; po.keys = pb.k0; po.vals = pb.v0; po.ngrad = (unsigned)ngrad;
            CUDA_TRY(cudaMemsetAsync(st->ray_counter, 0, sizeof(int), stream));
            if (nst == 1)
                apply_kernel<1, true><<<nblka, AT3D_RAY_THREADS, smem, stream>>>(S, G, (int)r1, camx, camy, camz, cammu, camphi,
                    packs, raypix, adjw, rw, sw, recoff, roff[r0], recs, nrec, grad_d, beam, po, pairoff, (RayErr *)st->err.p,
                    st->ray_counter, (int)r0);
            else
                apply_kernel<3, true><<<nblka, AT3D_RAY_THREADS, smem, stream>>>(S, G, (int)r1, camx, camy, camz, cammu, camphi,
                    packs, raypix, adjw, rw, sw, recoff, roff[r0], recs, nrec, grad_d, beam, po, pairoff, (RayErr *)st->err.p,
                    st->ray_counter, (int)r0);
            CUDA_TRY(cudaGetLastError());
            if (kernel_ms) cudaEventRecord(c2, stream);
            if ((rc = pair_accumulate(pb, (size_t)npair_chunk, nkeys, (unsigned)ngrad, grad_d, beam, stream, errmsg))) return rc;
            if (kernel_ms) {
                float t = 0.0f;
                cudaEventRecord(c3, stream);
                cudaEventSynchronize(c3);
                cudaEventElapsedTime(&t, c0, c1); ms_weights += t;
                cudaEventElapsedTime(&t, c2, c3); ms_accum += t;
                cudaEventDestroy(c0); cudaEventDestroy(c1); cudaEventDestroy(c2); cudaEventDestroy(c3);
            }
            r0 = r1;
        }
        if (kernel_ms) { kernel_ms[4] = ms_weights; kernel_ms[5] = ms_accum; }
        CUDA_TRY(cudaGetLastError());
        if (kernel_ms) cudaEventRecord(ev[2], stream);
        // ---- Phase 4 ----
        if (G.exact_single_scatter && G.stream_beam) {
            // streaming: count the entries of every walk, then passes over point ranges that fit the pair buffers
            const int nbt = (S.npts + 127) / 128;
            beam_stream_count_kernel<<<nbt, 128, 0, stream>>>(G, S.npts, S.ptrec, beam, bcount, (RayErr *)st->err.p);
            CUDA_TRY(cudaMemsetAsync(bcount + S.npts, 0, sizeof(int), stream));
            cub::TransformInputIterator<long long, IntToLL, const int *> it(bcount, IntToLL());
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(wb + w_cub, cubtmp, it, pairoff, S.npts + 1, stream));
            std::vector<long long> off_h((size_t)S.npts + 1);
            CUDA_TRY(cudaMemcpyAsync(off_h.data(), pairoff, sizeof(long long) * ((size_t)S.npts + 1), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if ((rc = check_ray_err(st, stream, errmsg))) return rc;
            long long cap = 1ll << 29;                                     // pairs per pass: 12 B each, twice (sort buffers)
            if (const char *e = getenv("AT3D_B200_BEAM_PAIRS")) { const long long v = atoll(e); if (v > 0) cap = v; }
            int p0 = 0;
            while (p0 < S.npts) {
                if (off_h[S.npts] == off_h[p0]) break;                        // nothing left
                int p1 = (int)(std::upper_bound(off_h.begin() + p0, off_h.end(), off_h[p0] + cap) - off_h.begin()) - 1;
                if (p1 <= p0) p1 = p0 + 1;                                    // one point's walk always fits (< 2^31 entries)
                if (p1 > S.npts) p1 = S.npts;
                const long long nbp = off_h[p1] - off_h[p0];
                if (nbp > 0) {
                    if (nbp > 0x7FFFFFF0ll) { set_msg(errmsg, "too many direct-beam derivative terms for one grid point"); return 2; }
                    PairBufs pb;
                    if ((rc = pair_reserve(st, (size_t)nbp + 1, nkeys, pb, errmsg))) return rc;
                    beam_stream_kernel<true><<<(p1 - p0 + 127) / 128, 128, 0, stream>>>(G, p0, p1, S.ptrec, beam, grad_d,
                                                                                         pairoff, pb.k0, pb.v0);
                    CUDA_TRY(cudaGetLastError());
                    if ((rc = pair_accumulate(pb, (size_t)nbp, nkeys, (unsigned)ngrad, grad_d, beam, stream, errmsg))) return rc;
                }
                p0 = p1;
            }
        } else if (G.exact_single_scatter) {
            const int wpb = 8, nbb = (S.npts + wpb - 1) / wpb;
            long long nbp = 0;
            beam_count_kernel<<<nbb, wpb * 32, 0, stream>>>(G, S.npts, beam, bcount);
            CUDA_TRY(cudaMemsetAsync(bcount + S.npts, 0, sizeof(int), stream));
            {
                cub::TransformInputIterator<long long, IntToLL, const int *> it(bcount, IntToLL());
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(wb + w_cub, cubtmp, it, pairoff, S.npts + 1, stream));
                CUDA_TRY(cudaMemcpyAsync(&nbp, pairoff + S.npts, sizeof(long long), cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
            }
            if (nbp > 0x7FFFFFF0ll) { set_msg(errmsg, "too many direct-beam derivative terms for one pass"); return 2; }
            PairBufs pb;
            if ((rc = pair_reserve(st, (size_t)nbp + 1, nkeys, pb, errmsg))) return rc;
            beam_kernel<true><<<nbb, wpb * 32, 0, stream>>>(G, S.npts, beam, grad_d, pairoff, pb.k0, pb.v0);
            CUDA_TRY(cudaGetLastError());
            if ((rc = pair_accumulate(pb, (size_t)nbp, nkeys, (unsigned)ngrad, grad_d, beam, stream, errmsg))) return rc;
        }
        // ---- Jacobian (MAKEJACOBIAN=.TRUE., shdomsub4.f:536-631): RAYGRAD_PIXEL(k,:,:) of a pixel is the
        //      derivative pass over that pixel's rays with the unit adjoint weight e_k (the pass is linear in
        //      the weight).  One pass per pixel and Stokes component, like the reference's slow path. ----
        if (njac > 0 && jacobian) {
            std::vector<int> rpp_h(npix);
            if (host) memcpy(rpp_h.data(), g->rays_per_pixel, sizeof(int) * npix);
            else CUDA_TRY(cudaMemcpy(rpp_h.data(), g->rays_per_pixel, sizeof(int) * npix, cudaMemcpyDeviceToHost));
            const size_t njv = (size_t)nst * G.numder * njac * npix;
            size_t oj = 0;
            const size_t j_adj = oj; oj += al256(sizeof(double) * nst * npix);
            const size_t j_gtmp = oj; oj += al256(sizeof(double) * ngrad);
            const size_t j_beam = oj; oj += al256(sizeof(double) * S.npts);
            const size_t j_ptr = oj; oj += al256(sizeof(int) * njac);
            const size_t j_jac = oj; oj += al256(sizeof(float) * njv);
            CUDA_TRY(st->misc.reserve(oj + 256));
            unsigned char *jb = (unsigned char *)st->misc.p;
            double *adj1 = (double *)(jb + j_adj), *gtmp = (double *)(jb + j_gtmp), *beam1 = (double *)(jb + j_beam);
            int *jptr_d = (int *)(jb + j_ptr);
            float *jac_d = host ? (float *)(jb + j_jac) : jacobian;
            CUDA_TRY(cudaMemcpyAsync(jptr_d, jacobianptr, sizeof(int) * njac,
                                     host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
            CUDA_TRY(cudaMemsetAsync(jac_d, 0, sizeof(float) * njv, stream));
            DevState Sq = S;
            Sq.counts = nullptr;
            const size_t smem = (size_t)AT3D_RAYS_PER_BLOCK * (S.ny_comp * S.nlmp + ((3 * (S.ml + 1) + 3) & ~3)) * sizeof(float);
            const long long *roff = st->recoff_h.data();
            CUDA_TRY(cudaFuncSetAttribute(nst == 1 ? (const void *)apply_kernel<1, false> : (const void *)apply_kernel<3, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            {   // one record buffer large enough for the rays of any single pixel
                long long mx = 0; size_t q0 = 0;
                for (size_t p = 0; p < npix; p++) { const size_t q1 = q0 + rpp_h[p]; if (roff[q1] - roff[q0] > mx) mx = roff[q1] - roff[q0]; q0 = q1; }
                CUDA_TRY(st->recs.reserve(((size_t)mx + 8) * sizeof(VisitRec)));
            }
            for (int k = 0; k < nst; k++) {
                unit_adj_kernel<<<(int)((nst * npix + 255) / 256), 256, 0, stream>>>(nst, (int)npix, k, adj1);
                size_t r0 = 0;
                for (size_t p = 0; p < npix; p++) {
                    const size_t r1 = r0 + rpp_h[p];
                    if (r1 > r0) {
                        VisitRec *recs = (VisitRec *)st->recs.p;
                        const int nblk = (int)((r1 - r0 + AT3D_RAYS_PER_BLOCK - 1) / AT3D_RAYS_PER_BLOCK);
                        CUDA_TRY(cudaMemsetAsync(gtmp, 0, sizeof(double) * ngrad, stream));
                        CUDA_TRY(cudaMemsetAsync(beam1, 0, sizeof(double) * S.npts, stream));
                        CUDA_TRY(cudaMemsetAsync(st->ray_counter, 0, sizeof(int), stream));
                        if (nst == 1)
                            weights_kernel<1><<<nblk, AT3D_RAY_THREADS, smem, stream>>>(Sq, G, (int)r1, camx, camy, camz, cammu,
                                camphi, packs, raypix, adj1, rw, sw, total, recoff, roff[r0], recs, nrec, npairs, nullptr, 0,
                                nullptr, nullptr, (RayErr *)st->err.p, st->ray_counter, (int)r0);
                        else
                            weights_kernel<3><<<nblk, AT3D_RAY_THREADS, smem, stream>>>(Sq, G, (int)r1, camx, camy, camz, cammu,
                                camphi, packs, raypix, adj1, rw, sw, total, recoff, roff[r0], recs, nrec, npairs, nullptr, 0,
                                nullptr, nullptr, (RayErr *)st->err.p, st->ray_counter, (int)r0);
                        CUDA_TRY(cudaMemsetAsync(st->ray_counter, 0, sizeof(int), stream));
                        if (nst == 1)
                            apply_kernel<1, false><<<nblk, AT3D_RAY_THREADS, smem, stream>>>(Sq, G, (int)r1, camx, camy, camz, cammu,
                                camphi, packs, raypix, adj1, rw, sw, recoff, roff[r0], recs, nrec, gtmp, beam1, PairOut(), nullptr,
                                (RayErr *)st->err.p, st->ray_counter, (int)r0);
                        else
                            apply_kernel<3, false><<<nblk, AT3D_RAY_THREADS, smem, stream>>>(Sq, G, (int)r1, camx, camy, camz, cammu,
                                camphi, packs, raypix, adj1, rw, sw, recoff, roff[r0], recs, nrec, gtmp, beam1, PairOut(), nullptr,
                                (RayErr *)st->err.p, st->ray_counter, (int)r0);
                        if (G.exact_single_scatter && G.stream_beam) {
                            beam_stream_kernel<false><<<(S.npts + 127) / 128, 128, 0, stream>>>(G, 0, S.npts, S.ptrec, beam1,
                                                                                                 gtmp, nullptr, nullptr, nullptr);
                        } else if (G.exact_single_scatter) {
                            const int wpb = 8;
                            beam_kernel<false><<<(S.npts + wpb - 1) / wpb, wpb * 32, 0, stream>>>(G, S.npts, beam1, gtmp, nullptr, nullptr, nullptr);
                        }
                        jac_gather_kernel<<<(G.numder * njac + 127) / 128, 128, 0, stream>>>(nst, G.numder, njac, G.maxpg, k,
                                                                                           (int)p, jptr_d, gtmp, jac_d);
                        CUDA_TRY(cudaGetLastError());
                    }
                    r0 = r1;
                }
            }
            if (host) CUDA_TRY(cudaMemcpyAsync(jacobian, jac_d, sizeof(float) * njv, cudaMemcpyDeviceToHost, stream));
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

extern "C" int at3d_levisapprox_gradient(at3d_state *st, const at3d_rays *rays, const at3d_grad_desc *g,
                                         double *gradout, double *cost, float *stokesout,
                                         const at3d_trace *trace, void *cuda_stream, double *kernel_ms,
                                         char *errmsg)
{
    return gradient_impl(st, rays, g, gradout, cost, stokesout, trace, cuda_stream, kernel_ms, 0, nullptr, nullptr, errmsg);
}

extern "C" int at3d_levisapprox_gradient_jacobian(at3d_state *st, const at3d_rays *rays, const at3d_grad_desc *g,
                                                  double *gradout, double *cost, float *stokesout,
                                                  int num_jacobian_pts, const int32_t *jacobianptr, float *jacobian,
                                                  void *cuda_stream, char *errmsg)
{
    if (num_jacobian_pts < 1 || !jacobianptr || !jacobian) {
        if (errmsg) snprintf(errmsg, AT3D_ERRMSG_LEN, "at3d_levisapprox_gradient_jacobian: NUM_JACOBIAN_PTS >= 1 and non-null arrays required");
        return 1;
    }
    return gradient_impl(st, rays, g, gradout, cost, stokesout, nullptr, cuda_stream, nullptr, num_jacobian_pts,
                         jacobianptr, jacobian, errmsg);
}
