// at3d_grad.cu -- LEVISAPPROX_GRADIENT, default adjoint ("double sweep") path, on sm_100a.
// Replaces (src/polarized/shdomsub4.f of the AT3D reference)
//   LEVISAPPROX_GRADIENT :288-809, ADJOINT_INTEGRATE_1RAY :3223-3967, COMPUTE_SOURCE_GRAD_1CELL
//   :1546-2042, COMPUTE_SOURCE_DIRECTION :2836-2914, FIND_BOUNDARY_RADIANCE_GRAD :2151-2347,
//   COMPUTE_ADJOINT_WEIGHTS :3969-4034, COMPUTE_RADIANCE_DERIVATIVE_ADJOINT :4037-4114,
//   COMPUTE_DIRECT_BEAM_DERIV_ADJOINT :4117-4143.
//
// Design (DESIGN.md "Gradient kernels"): one warp per ray, same bit-exact FP64 walk as RENDER.
//  * The reference evaluates the radiance SH contraction 8*NUMDER times per new grid point
//    (COMPUTE_SOURCE_DIRECTION per property corner and unknown).  That contraction is linear in the
//    Legendre table, so the warp forms the per-degree shell sums sum_m RADIANCE(.,j)*YLMDIR(.,j) once
//    (radiance SH read once per (ray, point)) and every (corner, unknown) then costs one
//    (NLEG+1)*NSTLEG dot product.
//  * GRAD8 is contracted with the per-ray adjoint weight at once, so a corner keeps 8*NUMDER scalars
//    instead of the reference's five NSTOKES*8*8*NUMDER arrays.
//  * The backward cumulative sum over saved sub-intervals (PASSEDRAD) becomes "total - running",
//    the total coming from a forward march with identical arithmetic; nothing is saved per
//    sub-interval.
//  * Sub-interval weights accumulate in registers (lane n owns corner n); global memory is touched
//    once per cell and corner with red.global.add.f64, never inside the sub-interval loop.
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <cub/cub.cuh>
#include "at3d_host.h"
#include "at3d_ray.cuh"

// ------------------------------------------------------------------------------------------
// per-warp shared scratch of the adjoint kernel
// ------------------------------------------------------------------------------------------
struct GradLayout {
    int y, rad, cc, dslot, ib, xg, legent, tc, vsh, wacc, total;   // byte offsets
};

__host__ __device__ inline GradLayout grad_layout(int nstokes, int ny_comp, int nlmp, int nstleg, int nleg,
                                                  int ml, int numder)
{
    GradLayout L;
    int o = 0;
    L.y = o;      o += ny_comp * nlmp * 4;
    L.rad = o;    o += nstokes * nlmp * 4;
    L.dslot = o;  o += 8 * 8 * numder * 8;                 // double D[8 corners][8 nb][numder]
    L.wacc = o;   o += 3 * 8 * 8;                          // double W[8], G[8], BW[8]
    L.tc = o;     o += nstleg * (nleg + 1) * 8;            // double Tc[nlt]
    o = (o + 15) & ~15;
    L.cc = o;     o += (nstokes == 1 ? (int)sizeof(CornerCache<1>) : (int)sizeof(CornerCache<3>));
    L.ib = o;     o += 8 * 8 * 4;                          // int IB[8][8]
    L.xg = o;     o += 8 * 8 * numder * 4;                 // float XG[8][8][numder]
    L.legent = o; o += nstleg * (nleg + 1) * 4;            // float legent[nlt]
    L.vsh = o;    o += 3 * (ml + 1) * 4;                   // float V1[ml+1], V2[ml+1], V6[ml+1]
    L.total = (o + 15) & ~15;
    return L;
}

// SINGSCAT(:,iph) of the ray (shdomsub2.f:2425-2441), evaluated on demand
template <int NST>
__device__ __forceinline__ void ray_singscat(const float *tab, int nstphase, int numphase, int iph,
                                             const RayDir &rd, float (&s)[NST])
{
    const float *p0 = tab + (size_t)nstphase * ((iph - 1) + (size_t)numphase * (rd.j - 1));
    const float *p1 = p0 + (size_t)nstphase * numphase;
    s[0] = (1 - rd.f) * __ldg(p0) + rd.f * __ldg(p1);
    if (NST > 1) {
        const float b1 = (1 - rd.f) * __ldg(p0 + 1) + rd.f * __ldg(p1 + 1);
        s[1] = (float)(b1 * rd.cos22);
        s[NST - 1] = (float)(b1 * rd.sin22);
    }
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

// COMPUTE_SOURCE_GRAD_1CELL for one new grid point (all lanes cooperate).
template <int NST>
__device__ void eval_point_grad(const DevState &S, const DevGrad &G, int ip, unsigned char *sm,
                                const GradLayout &L, const RayDir &rd, const double (&adj)[NST],
                                float &ext_out, float (&src_out)[NST], float (&ss_out)[NST],
                                double *Dslot, int *IBslot, float *XGslot, float &fpersist)
{
    const int lane = lane_id();
    const float *Ysh = (const float *)(sm + L.y);
    float *radsh = (float *)(sm + L.rad);
    float *legent = (float *)(sm + L.legent);
    double *Tc = (double *)(sm + L.tc);
    const float *Vsh = (const float *)(sm + L.vsh);
    const int nlmp = S.nlmp, nstleg = S.nstleg, ml = S.ml, mm = S.mm;
    const int nlt = nstleg * (S.nleg + 1);
    const bool deltam = S.deltam != 0, interp_new = S.interp_new != 0;
    const float secmu0 = (float)(1.0 / fabs((double)S.solarmu));
    // ---------------- forward part (shdomsub4.f:1660-1785) ----------------
    const float4 pr = __ldg(&S.ptrec[ip - 1]);
    const float ext = pr.w;
    float a[NST], b[NST];
    {
        const int2 sr = __ldg(&S.srcrec[ip - 1]);
        const int nsp = (sr.y + 3) & ~3;
        const float *base = S.shsrc + sr.x;
#pragma unroll
        for (int k = 0; k < NST; k++) { a[k] = 0.0f; b[k] = 0.0f; }
        for (int j4 = lane * 4; j4 < nsp; j4 += 128) {
            const float4 s = __ldg((const float4 *)(base + j4));
            const float4 y = *(const float4 *)(Ysh + j4);
            a[0] = fmaf(s.x, y.x, a[0]); a[0] = fmaf(s.y, y.y, a[0]);
            a[0] = fmaf(s.z, y.z, a[0]); a[0] = fmaf(s.w, y.w, a[0]);
            if (NST > 1) {
                const float4 q = __ldg((const float4 *)(base + nsp + j4));
                const float4 u = __ldg((const float4 *)(base + 2 * nsp + j4));
                const float4 y2 = *(const float4 *)(Ysh + 1 * nlmp + j4);
                const float4 y5 = *(const float4 *)(Ysh + 2 * nlmp + j4);
                const float4 y6 = *(const float4 *)(Ysh + 3 * nlmp + j4);
                const float4 y3 = *(const float4 *)(Ysh + 4 * nlmp + j4);
                a[1] = fmaf(q.x, y2.x, a[1]); a[1] = fmaf(u.x, y5.x, a[1]);
                a[1] = fmaf(q.y, y2.y, a[1]); a[1] = fmaf(u.y, y5.y, a[1]);
                a[1] = fmaf(q.z, y2.z, a[1]); a[1] = fmaf(u.z, y5.z, a[1]);
                a[1] = fmaf(q.w, y2.w, a[1]); a[1] = fmaf(u.w, y5.w, a[1]);
                a[NST - 1] = fmaf(q.x, y6.x, a[NST - 1]); a[NST - 1] = fmaf(u.x, y3.x, a[NST - 1]);
                a[NST - 1] = fmaf(q.y, y6.y, a[NST - 1]); a[NST - 1] = fmaf(u.y, y3.y, a[NST - 1]);
                a[NST - 1] = fmaf(q.z, y6.z, a[NST - 1]); a[NST - 1] = fmaf(u.z, y3.z, a[NST - 1]);
                a[NST - 1] = fmaf(q.w, y6.w, a[NST - 1]); a[NST - 1] = fmaf(u.w, y3.w, a[NST - 1]);
            }
        }
        const int cnt = __ldg(&S.sscount[ip - 1]);
        for (int k = lane; k < cnt; k += 32) {
            const int2 e = __ldg(&S.ssent[(size_t)(ip - 1) * S.kmax + k]);
            const float coef = __int_as_float(e.y);
            float sv[NST];
            ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, e.x, rd, sv);
#pragma unroll
            for (int kk = 0; kk < NST; kk++) b[kk] = fmaf(coef, sv[kk], b[kk]);
        }
#pragma unroll
        for (int k = 0; k < NST; k++) { a[k] = warp_sum(a[k]); b[k] = warp_sum(b[k]); }
    }
    if (!deltam) {
        // without delta-M SINGSCAT8 is the truncated single scattering (shdomsub4.f:1723-1760,1781):
        // sum over species of DA*LEGENT(.,l)*sum_m YLMDIR*YLMSUN for the shells present in SOURCE
        const int ns = __ldg(&S.srcrec[ip - 1]).y;
        const float dirflux0 = __ldg(&S.dirflux[ip - 1]);
        float t[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) t[k] = 0.0f;
        for (int ipa = 0; ipa < S.npart; ipa++) {
            float w;
            if (ext == 0.0f) w = 1.0f; else w = __ldg(&S.extinct[(ip - 1) + (size_t)S.npts * ipa]) / ext;
            if (w == 0.0f) continue;
            const int *iph = S.iphase + (size_t)S.nq * ((ip - 1) + (size_t)S.npts * ipa);
            const float *pw = S.phaseinterpwt + (size_t)S.nq * ((ip - 1) + (size_t)S.npts * ipa);
            const bool single = (!interp_new) || (__ldg(&pw[0]) >= S.phasemax);
            const float da = __ldg(&S.albedo[(ip - 1) + (size_t)S.npts * ipa]) * dirflux0 * secmu0 * w;
            for (int l = lane; l <= ml; l += 32) {
                const int me = l < mm ? l : mm;
                if (sh_index(l, -me, mm) >= ns) continue;
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
        for (int k = 0; k < NST; k++) b[k] = warp_sum(t[k]);
    }
    float srcfull[NST];      // SRCEXT8 before the multiplication by the extinction
#pragma unroll
    for (int k = 0; k < NST; k++) {
        srcfull[k] = deltam ? a[k] + b[k] : a[k];
        ss_out[k] = b[k] * ext;
        src_out[k] = G.singlescatter ? ss_out[k] : srcfull[k] * ext;
    }
    ext_out = ext;
    // ---------------- radiance SH block -> shared memory ----------------
    const int2 rr = __ldg(&S.radrec[ip - 1]);
    const int rns = rr.y, nrp = (rr.y + 3) & ~3;
    {
        const float *rb = S.shrad + rr.x;
        for (int j4 = lane * 4; j4 < nrp; j4 += 128) {
#pragma unroll
            for (int k = 0; k < NST; k++)
                *(float4 *)(radsh + k * nlmp + j4) = __ldg((const float4 *)(rb + k * nrp + j4));
        }
    }
    for (int t = lane; t < nlt; t += 32) Tc[t] = 0.0;
    __syncwarp();
    // shell sums  T_k(l) = sum_m adj . RADIANCE(.,j) YLMDIR(.,j)   (lane <-> degree l)
    const float dirflux = __ldg(&S.dirflux[ip - 1]);
    for (int l = lane; l <= ml; l += 32) {
        const int me = l < mm ? l : mm;
        const int jlo = sh_index(l, -me, mm), cnt = 2 * me + 1;
        float A = 0, B = 0, C = 0, D = 0, E = 0, F = 0, Gq = 0, H = 0;
        for (int i = 0; i < cnt; i++) {
            const int j = jlo + i;
            if (j >= rns) break;
            const float r1 = radsh[j], y1 = Ysh[j];
            A = A + r1 * y1;
            if (NST > 1) {
                const float r2 = radsh[nlmp + j], r3 = radsh[2 * nlmp + j];
                const float y2 = Ysh[nlmp + j], y5 = Ysh[2 * nlmp + j], y6 = Ysh[3 * nlmp + j], y3 = Ysh[4 * nlmp + j];
                B = B + r2 * y1; C = C + r1 * y2; D = D + r2 * y2; E = E + r3 * y5;
                F = F + r1 * y6; Gq = Gq + r2 * y6; H = H + r3 * y3;
            }
        }
        double t1 = adj[0] * A;
        if (!deltam) t1 += adj[0] * (double)(dirflux * secmu0 * Vsh[l]);
        Tc[0 + nstleg * l] = t1;
        if (NST > 1) {
            double t5 = adj[0] * B + adj[1] * C + adj[NST - 1] * F;
            if (!deltam) t5 += adj[1] * (double)(dirflux * secmu0 * Vsh[(ml + 1) + l]);
            Tc[4 + nstleg * l] = t5;
            Tc[1 + nstleg * l] = adj[1] * D + adj[NST - 1] * Gq;
            Tc[2 + nstleg * l] = adj[1] * E + adj[NST - 1] * H;
        }
    }
    __syncwarp();
    // ---------------- gradient part (shdomsub4.f:1786-2019) ----------------
    const int nb = lane & 7, g4 = lane >> 3;
    const int ib = __ldg(&G.interpptr[nb + 8 * (size_t)(ip - 1)]);
    const float xi = __ldg(&G.optinterpwt[nb + 8 * (size_t)(ip - 1)]);
    if (lane < 8) IBslot[lane] = ib;
    int last_ipa = -1;
    float scatterj = 0.0f, singscatj[NST], sourcet[NST], f = fpersist;
#pragma unroll
    for (int k = 0; k < NST; k++) { singscatj[k] = 0.0f; sourcet[k] = 0.0f; }
    const int pm = G.pmaxnmicro;
    for (int idr = 0; idr < G.numder; idr++) {
        const int ipa = __ldg(&G.partder[idr]);       // 1-based species
        const float albp = __ldg(&G.albedop[(ib - 1) + (size_t)G.maxpg * (ipa - 1)]);
        const float extp = __ldg(&G.extinctp[(ib - 1) + (size_t)G.maxpg * (ipa - 1)]);
        const int *iphp = G.iphasep + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
        const float *pwp = G.phasewtp + (size_t)pm * ((ib - 1) + (size_t)G.maxpg * (ipa - 1));
        const float alb_ip = __ldg(&S.albedo[(ip - 1) + (size_t)S.npts * (ipa - 1)]);
        if (ipa != last_ipa) {
            last_ipa = ipa;
            __syncwarp();
            const float sw = xi * albp * extp;       // SPATIAL_WEIGHT of property corner nb
            scatterj = 0.0f;
#pragma unroll
            for (int k = 0; k < NST; k++) singscatj[k] = 0.0f;
            if (deltam) {
                for (int n = 0; n < 8; n++) scatterj = scatterj + __shfl_sync(FULLMASK, sw, n);
                float part[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) part[k] = 0.0f;
                if (g4 == 0 && sw > 1e-6f) {
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
                    for (int n = 0; n < 8; n++) tot = tot + __shfl_sync(FULLMASK, part[k], n);
                    if (scatterj > G.scatmin) singscatj[k] = tot / scatterj;
                    else singscatj[k] = (float)(tot / G.scatmin);
                }
            }
            if (S.npart == 1) {
                // LEGENT left by the forward part: the PHASEINTERPWT mix at this grid point
                const int *iph = S.iphase + (size_t)S.nq * ((ip - 1) + (size_t)S.npts * (ipa - 1));
                const float *pw = S.phaseinterpwt + (size_t)S.nq * ((ip - 1) + (size_t)S.npts * (ipa - 1));
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
#pragma unroll
                for (int k = 0; k < NST; k++) sourcet[k] = (alb_ip > 1e-8f) ? srcfull[k] / alb_ip : 0.0f;
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
#pragma unroll
                for (int k = 0; k < NST; k++) sourcet[k] = 0.0f;
                if (scatterj > G.scatmin) {
                    // COMPUTE_SOURCE_DIRECTION with the mixed table
                    float acc[NST];
#pragma unroll
                    for (int k = 0; k < NST; k++) acc[k] = 0.0f;
                    for (int j = lane; j < rns; j += 32) {
                        const int l = __ldg(&S.lofj[j]);
                        const float r1 = radsh[j], y1 = Ysh[j];
                        acc[0] = acc[0] + legent[nstleg * l] * r1 * y1;
                        if (NST > 1) {
                            const float r2 = radsh[nlmp + j], r3 = radsh[2 * nlmp + j];
                            const float l5 = legent[4 + nstleg * l], l2 = legent[1 + nstleg * l], l3 = legent[2 + nstleg * l];
                            acc[0] = acc[0] + l5 * r2 * y1;
                            acc[1] = acc[1] + l5 * r1 * Ysh[nlmp + j] + l2 * r2 * Ysh[nlmp + j] + l3 * r3 * Ysh[2 * nlmp + j];
                            acc[NST - 1] = acc[NST - 1] + l5 * r1 * Ysh[3 * nlmp + j] + l2 * r2 * Ysh[3 * nlmp + j]
                                           + l3 * r3 * Ysh[4 * nlmp + j];
                        }
                    }
                    if (!deltam) {
                        for (int l = lane; l <= ml; l += 32) {
                            acc[0] = acc[0] + dirflux * secmu0 * legent[nstleg * l] * Vsh[l];
                            if (NST > 1) acc[1] = acc[1] + dirflux * secmu0 * legent[4 + nstleg * l] * Vsh[(ml + 1) + l];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NST; k++) sourcet[k] = warp_sum(acc[k]);
                    if (deltam) {
#pragma unroll
                        for (int k = 0; k < NST; k++) sourcet[k] = sourcet[k] + dirflux * singscatj[k] * secmu0 / (1 - f);
                    }
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
            sourcet[0] = fmaxf(0.0f, sourcet[0]);
            __syncwarp();
        }
        // ---- per property corner nb (lane = nb + 8*g4, g4 splits the table entries) ----
        const float dext_v = __ldg(&G.dext[(ib - 1) + (size_t)G.maxpg * idr]);
        const float dalb_v = __ldg(&G.dalb[(ib - 1) + (size_t)G.maxpg * idr]);
        const float dextm_v = __ldg(&G.dextm[(ib - 1) + (size_t)G.maxpg * idr]);
        const float dalbm_v = __ldg(&G.dalbm[nb + 8 * ((size_t)(ip - 1) + (size_t)S.npts * idr)]);
        const float dfj_v = __ldg(&G.dfj[nb + 8 * ((size_t)(ip - 1) + (size_t)S.npts * idr)]);
        const int doex = __ldg(&G.doexact[idr]);
        const int *dip = G.diphasep + (size_t)G.deriv_maxnmicro * ((ib - 1) + (size_t)G.maxpg * idr);
        const float *dpw = G.dphasewtp + (size_t)G.deriv_maxnmicro * ((ib - 1) + (size_t)G.maxpg * idr);
        double dot = 0.0;
        if (xi >= 1e-7f) {
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
                const float dlegt = dext_v * leg_diff * albp + dalb_v * leg_diff * extp + dlegp * extp * albp
                                    + (lt - 1) * dfj_v;
                dot += (double)dlegt * Tc[t];
            }
        }
        dot += __shfl_xor_sync(FULLMASK, dot, 8);
        dot += __shfl_xor_sync(FULLMASK, dot, 16);
        if (lane < 8) {
            double d = 0.0;
            if (xi >= 1e-7f) {
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
            Dslot[lane * G.numder + idr] = d;
            XGslot[lane * G.numder + idr] = dextm_v * xi;
        }
    }
    fpersist = f;
    __syncwarp();
}

// Corner refresh for the adjoint walk: values (and the contracted GRAD8 rows) of points shared with
// the previous cell are carried over (OLDIPTS/DONEFACE logic of shdomsub4.f:1645-1659).
template <int NST>
__device__ __forceinline__ void refresh_corners_grad(const DevState &S, const DevGrad &G, const CellRec &c,
                                                     unsigned char *sm, const GradLayout &L, const RayDir &rd,
                                                     const double (&adj)[NST], bool first, float &fpersist,
                                                     int &npt_eval, int &nsh_eval, int &nrh_eval)
{
    const int lane = lane_id();
    CornerCache<NST> *cc = (CornerCache<NST> *)(sm + L.cc);
    double *Dall = (double *)(sm + L.dslot);
    int *IBall = (int *)(sm + L.ib);
    float *XGall = (float *)(sm + L.xg);
    const int nd = G.numder;
    // lane n<8 decides; then every lane helps copying the 8*nd rows of reused corners
    int myp = 0, hit = -1;
    if (lane < 8) {
        myp = c.gp[0];
#pragma unroll
        for (int n = 1; n < 8; n++) if (lane == n) myp = c.gp[n];
        if (!first) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (cc->pt[k] == myp) hit = k;
        }
    }
    float oext = 0.0f, osrc[NST], oss[NST];
    if (lane < 8 && hit >= 0) {
        oext = cc->ext[hit];
#pragma unroll
        for (int k = 0; k < NST; k++) { osrc[k] = cc->src[k][hit]; oss[k] = cc->ss[k][hit]; }
    }
    // row copies: each lane handles (slot = lane&7, part = lane>>3) of the 8*nd doubles/floats + ints
    const int slot = lane & 7, part = lane >> 3;
    const int shit = __shfl_sync(FULLMASK, hit, slot);
    double dtmp[16]; float xtmp[16]; int itmp[2];
    const int per = (8 * nd + 3) / 4;      // entries per part
    int cntd = 0;
    if (shit >= 0 && shit != slot) {
        for (int e = part * per; e < (part + 1) * per && e < 8 * nd && cntd < 16; e++, cntd++) {
            dtmp[cntd] = Dall[shit * 8 * nd + e];
            xtmp[cntd] = XGall[shit * 8 * nd + e];
        }
        itmp[0] = IBall[shit * 8 + 2 * part];
        itmp[1] = IBall[shit * 8 + 2 * part + 1];
    }
    __syncwarp();
    if (lane < 8) {
        cc->pt[lane] = myp;
        if (hit >= 0) {
            cc->ext[lane] = oext;
#pragma unroll
            for (int k = 0; k < NST; k++) { cc->src[k][lane] = osrc[k]; cc->ss[k][lane] = oss[k]; }
        }
    }
    if (shit >= 0 && shit != slot) {
        int i = 0;
        for (int e = part * per; e < (part + 1) * per && e < 8 * nd && i < 16; e++, i++) {
            Dall[slot * 8 * nd + e] = dtmp[i];
            XGall[slot * 8 * nd + e] = xtmp[i];
        }
        IBall[slot * 8 + 2 * part] = itmp[0];
        IBall[slot * 8 + 2 * part + 1] = itmp[1];
    }
    __syncwarp();
    unsigned need = __ballot_sync(FULLMASK, lane < 8 && hit < 0);
    while (need) {
        const int n = __ffs(need) - 1;
        need &= need - 1;
        const int ip = __shfl_sync(FULLMASK, myp, n);
        float ext, src[NST], ss[NST];
        eval_point_grad<NST>(S, G, ip, sm, L, rd, adj, ext, src, ss, Dall + n * 8 * nd, IBall + n * 8,
                             XGall + n * 8 * nd, fpersist);
        npt_eval++; nsh_eval += __ldg(&S.srcrec[ip - 1]).y; nrh_eval += __ldg(&S.radrec[ip - 1]).y;
        // duplicate corners inside one cell share the evaluation
        const unsigned same = __ballot_sync(FULLMASK, lane < 8 && myp == ip);
        if (lane < 8 && myp == ip) {
            cc->ext[lane] = ext;
#pragma unroll
            for (int k = 0; k < NST; k++) { cc->src[k][lane] = src[k]; cc->ss[k][lane] = ss[k]; }
        }
        unsigned dup = same & ~(1u << n);
        while (dup) {
            const int m = __ffs(dup) - 1;
            dup &= dup - 1;
            for (int e = lane; e < 8 * nd; e += 32) {
                Dall[m * 8 * nd + e] = Dall[n * 8 * nd + e];
                XGall[m * 8 * nd + e] = XGall[n * 8 * nd + e];
            }
            if (lane < 8) IBall[m * 8 + lane] = IBall[n * 8 + lane];
        }
        need &= ~same;
        __syncwarp();
    }
    __syncwarp();
}

// ADJOINT_INTEGRATE_1RAY for one ray (one warp).
template <int NST>
__device__ int march_ray_adjoint(const DevState &S, const DevGrad &G, unsigned char *sm, const GradLayout &L,
                                 const RayDir &rd, double mu2, double x0, double y0, double z0, float sky,
                                 const double (&adj)[NST], const double (&total)[NST],
                                 double *gradout, double *beam_weight,
                                 int *trace_cells, int trace_cap, int &ntrace, int &nsub)
{
    const int lane = lane_id();
    CornerCache<NST> *cc = (CornerCache<NST> *)(sm + L.cc);
    const double *Dall = (const double *)(sm + L.dslot);
    const int *IBall = (const int *)(sm + L.ib);
    const float *XGall = (const float *)(sm + L.xg);
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
    float fpersist = 0.0f;
    int npt_eval = 0, nsh_eval = 0, nrh_eval = 0;
    ntrace = 0; nsub = 0;
    while (!done && icell > 0) {
        if (trace_cells && lane == 0 && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        const CellRec c = load_cell(S, icell);
        refresh_corners_grad<NST>(S, G, c, sm, L, rd, adj, first, fpersist, npt_eval, nsh_eval, nrh_eval);
        first = false;
        float e8[8], s8[NST][8];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            e8[n] = cc->ext[n];
#pragma unroll
            for (int k = 0; k < NST; k++) s8[k][n] = cc->src[k][n];
        }
        float myss[NST];           // SINGSCAT8(:,n) of the corner this lane owns
#pragma unroll
        for (int k = 0; k < NST; k++) myss[k] = cc->ss[k][lane & 7];
        const float4 q1 = __ldg(&S.ptrec[c.gp[0] - 1]);
        const float4 q8 = __ldg(&S.ptrec[c.gp[7] - 1]);
        const double delx = (double)(q8.x - q1.x), dely = (double)(q8.y - q1.y), delz = (double)(q8.z - q1.z);
        const double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        const double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        const double invdelz = 1.0 / delz;
        double u = (xe - q1.x) * invdelx, v = (ye - q1.y) * invdely, w = (ze - q1.z) * invdelz;
        double fc[8], fcprev[8];
        interp_kernel(u, v, w, fc);
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
        // per-corner accumulators of this cell (lane n < 8 owns corner n)
        double Wn = 0.0, Gn = 0.0;
        float bw[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) bw[k] = 0.0f;
        for (int it = 1; it <= ntau; it++) {
#pragma unroll
            for (int n = 0; n < 8; n++) fcprev[n] = fc[n];
            const double s = it * dels;
            const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
            u = (xi - q1.x) * invdelx; v = (yi - q1.y) * invdely; w = (zi - q1.z) * invdelz;
            interp_kernel(u, v, w, fc);
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
                double f0 = fc[0], f1 = fcprev[0];
#pragma unroll
                for (int n = 1; n < 8; n++) if ((lane & 7) == n) { f0 = fc[n]; f1 = fcprev[n]; }
                Wn += transmit * abscell * ((0.5f * (f0 + f1) + 0.08333333333f * (ext0 * f1 - ext1 * f0) * corr) / ext);
                if (exact_ss) {
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        float ss0 = (float)(f0 * myss[k]), ss1 = (float)(f1 * myss[k]);
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
        if (lane < 8) {
            Wacc[lane] = Wn; Wacc[8 + lane] = Gn;
            if (exact_ss) {
                double bsum = 0.0;
#pragma unroll
                for (int k = 0; k < NST; k++) bsum += adj[k] * (double)bw[k];
                int ipn = c.gp[0];
#pragma unroll
                for (int n = 1; n < 8; n++) if (lane == n) ipn = c.gp[n];
                if (bsum != 0.0) atomicAdd(&beam_weight[ipn - 1], bsum);
            }
        }
        __syncwarp();
        for (int e = lane; e < 64 * nd; e += 32) {
            const int slot = e / (8 * nd), r = e - slot * 8 * nd, nb = r / nd, idr = r - nb * nd;
            const double val = Wacc[slot] * Dall[e] + (double)XGall[e] * Wacc[8 + slot];
            if (val != 0.0) atomicAdd(&gradout[(size_t)(IBall[slot * 8 + nb] - 1) + (size_t)G.maxpg * idr], val);
        }
        __syncwarp();
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
            if (exact_ss && lane < 4) {
                int bp = boundpts[0]; double bi = boundinterp[0], dr = dirrad1[0];
#pragma unroll
                for (int n = 1; n < 4; n++) if (lane == n) { bp = boundpts[n]; bi = boundinterp[n]; dr = dirrad1[n]; }
                const double val = adj[0] * transmit * bi * dr;
                if (val != 0.0) atomicAdd(&beam_weight[bp - 1], val);
            }
        } else {
            icell = inextcell;
        }
        xe = xn; ye = yn; ze = zn;
    }
    if (S.counts && lane == 0) {
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
__global__ void __launch_bounds__(AT3D_WARPS_PER_BLOCK * 32)
adjoint_kernel(DevState S, DevGrad G, GradLayout L, int nrays, const float *camx, const float *camy,
               const float *camz, const double *cammu, const double *camphi, const RayPack *packs,
               const int *raypix, const double *adjw /*[NST,npix]*/, const double *ray_weights,
               const double *stokes_weights, const double *total /*[NST,nrays]*/,
               double *gradout, double *beam_weight,
               int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub, RayErr *err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *sm = smem_raw + (size_t)warp * L.total;
    float *Ysh = (float *)(sm + L.y);
    float *Vsh = (float *)(sm + L.vsh);
    const int nwarps = gridDim.x * AT3D_WARPS_PER_BLOCK;
    for (int iray = blockIdx.x * AT3D_WARPS_PER_BLOCK + warp; iray < nrays; iray += nwarps) {
        const double mu2 = __ldg(&cammu[iray]), phi2 = __ldg(&camphi[iray]);
        const RayPack pk = dev_get_pack(S, packs, iray, camx, camy, camz, mu2, phi2);
        int ntrace = 0, nsub = 0;
        if (pk.status == 2) { if (lane == 0) set_err(err, 2, iray); }
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
            __syncwarp();
            warp_ylmall(S, (float)mu2, (float)phi2, Ysh);
            if (!S.deltam) {
                // shell sums of YLMSUN*YLMDIR for the untruncated solar term of COMPUTE_SOURCE_DIRECTION
                for (int l = lane; l <= S.ml; l += 32) {
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
                __syncwarp();
            }
            const float sky = (-mu2 > 0.0) ? dev_sky_radiance(S, (float)mu2, (float)phi2) : 0.0f;
            const int e = march_ray_adjoint<NST>(S, G, sm, L, rd, mu2, pk.x0, pk.y0, pk.z0, sky, adj, tot, gradout,
                                                 beam_weight,
                                                 trace_cells ? trace_cells + (size_t)trace_cap * iray : nullptr,
                                                 trace_cap, ntrace, nsub);
            if (e && lane == 0) set_err(err, e, iray);
        }
        if (lane == 0 && trace_n) { trace_n[iray] = ntrace; trace_nsub[iray] = nsub; }
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
        CUDA_TRY(launch_render(Sf, (int)n, camx, camy, camz, cammu, camphi, packs, nullptr, visrad, 0, 1, G.singlescatter, 0, 0,
                               nullptr, 0, nullptr, nullptr, (RayErr *)st->err.p, stream));
        CUDA_TRY(launch_render(Sf, (int)n, camx, camy, camz, cammu, camphi, packs, nullptr, total, 1, 1, G.singlescatter, 0,
                               G.maxsub, nullptr, 0, nullptr, nullptr, (RayErr *)st->err.p, stream));
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
        const GradLayout L = grad_layout(nst, S.ny_comp, S.nlmp, S.nstleg, S.nleg, S.ml, G.numder);
        const size_t smem = (size_t)AT3D_WARPS_PER_BLOCK * L.total;
        if (smem > 227 * 1024) { set_msg(errmsg, "gradient kernel needs %zu bytes of shared memory (NUMDER too large)", smem); return 3; }
        if (8 * G.numder > 64) { set_msg(errmsg, "NUMDER > 8 is not supported by the adjoint kernel"); return 3; }
        const void *fn = nst == 1 ? (const void *)adjoint_kernel<1> : (const void *)adjoint_kernel<3>;
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0, nsm = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, AT3D_WARPS_PER_BLOCK * 32, smem);
        if (per_sm < 1) per_sm = 1;
        long want = ((long)n + AT3D_WARPS_PER_BLOCK - 1) / AT3D_WARPS_PER_BLOCK;
        long cap = (long)nsm * per_sm;
        const int nblk = (int)(want < cap ? want : cap);
        if (nst == 1)
            adjoint_kernel<1><<<nblk, AT3D_WARPS_PER_BLOCK * 32, smem, stream>>>(S, G, L, (int)n, camx, camy, camz, cammu,
                camphi, packs, raypix, adjw, rw, sw, total, grad_d, beam, tc, tcap, tn, ts, (RayErr *)st->err.p);
        else
            adjoint_kernel<3><<<nblk, AT3D_WARPS_PER_BLOCK * 32, smem, stream>>>(S, G, L, (int)n, camx, camy, camz, cammu,
                camphi, packs, raypix, adjw, rw, sw, total, grad_d, beam, tc, tcap, tn, ts, (RayErr *)st->err.p);
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
