/*
 * oracle_grad.c -- TEST INFRASTRUCTURE (see shdom_oracle.h).
 * Levis-approximation cost-function gradient, default adjoint ("double sweep") path:
 *   LEVISAPPROX_GRADIENT            /root/reference/src/polarized/shdomsub4.f:288-809
 *   COMPUTE_SOURCE_GRAD_1CELL       shdomsub4.f:1546-2042
 *   FIND_BOUNDARY_RADIANCE_GRAD     shdomsub4.f:2151-2347 (Lambertian surfaces)
 *   COMPUTE_SOURCE_DIRECTION        shdomsub4.f:2836-2914
 *   PREPARE_DERIV_INTERPS / COMPUTE_INTERP_WEIGHTS  shdomsub4.f:2917-3169
 *   ADJOINT_INTEGRATE_1RAY          shdomsub4.f:3223-3967
 *   COMPUTE_ADJOINT_WEIGHTS         shdomsub4.f:3969-4034
 *   COMPUTE_RADIANCE_DERIVATIVE_ADJOINT  shdomsub4.f:4037-4114
 *   COMPUTE_DIRECT_BEAM_DERIV_ADJOINT    shdomsub4.f:4117-4143
 *   UPDATE_COSTFUNCTION             shdomsub4.f:13-91
 *   GET_INTERP_KERNEL               /root/reference/src/shdomsub5.f:1497-1536
 *   average_subpixel_rays           /root/reference/src/util.f90:484-518
 * Solar, thermal and combined sources (SRCTYPE 'S','T','B'; UNITS 'R' or 'T'): the thermal terms are
 *   PLANCK_DERIVATIVE               shdomsub4.f:3171-3221
 *   PLANCK / DPLANCK per grid point shdomsub4.f:1792-1799, GRAD8 thermal component :2009-2016
 *   surface emission (RADEMIS) in FIND_BOUNDARY_RADIANCE_GRAD :2274-2290
 * With SRCTYPE='T' and delta-M the reference reads SINGSCAT without allocating it (shdomsub4.f:3413-3423 vs
 * :1811-1824); the values only enter terms that are multiplied out for thermal sources, so here they are simply
 * evaluated.  Likewise LEGENT / F for NPART=1 are only set inside the solar block of COMPUTE_SOURCE_GRAD_1CELL
 * (:1693-1722) but read by the gradient part (:1835-1840, :1963-1985): undefined for SRCTYPE='T' in the reference,
 * evaluated as for 'S' / 'B' here (see compute_source_grad_1cell).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "shdom_oracle.h"
#include "oracle_internal.h"

/* GET_INTERP_KERNEL  shdomsub5.f:1497-1536 */
static void get_interp_kernel(const oracle_state *st, int icell, double x, double y, double z, double *f)
{
    int ipt1 = GRIDPTR(st, 1, icell), ipt2 = GRIDPTR(st, 8, icell);
    double delx, dely, delz, invdelx, invdely, invdelz, u, v, w;
    delx = GRIDPOS(st, 1, ipt2) - GRIDPOS(st, 1, ipt1);
    if (delx <= 0.0) invdelx = 1.0; else invdelx = 1.0 / delx;
    dely = GRIDPOS(st, 2, ipt2) - GRIDPOS(st, 2, ipt1);
    if (dely <= 0.0) invdely = 1.0; else invdely = 1.0 / dely;
    delz = GRIDPOS(st, 3, ipt2) - GRIDPOS(st, 3, ipt1);
    invdelz = 1.0 / delz;
    u = (x - GRIDPOS(st, 1, ipt1)) * invdelx;
    v = (y - GRIDPOS(st, 2, ipt1)) * invdely;
    w = (z - GRIDPOS(st, 3, ipt1)) * invdelz;
    f[0] = (1 - w) * (1 - v) * (1 - u);
    f[1] = (1 - w) * (1 - v) * u;
    f[2] = (1 - w) * v * (1 - u);
    f[3] = (1 - w) * v * u;
    f[4] = w * (1 - v) * (1 - u);
    f[5] = w * (1 - v) * u;
    f[6] = w * v * (1 - u);
    f[7] = w * v * u;
}

typedef struct {
    int nstokes, nstleg, nleg, nlm, ml, mm, numder;
    int *lofj;
    /* Legendre work tables [nstleg,0:nleg] */
    float *legent, *legenp, *unscaled, *dlegp, *dlegt, *leg_diff;
    float f;                       /* persists like the Fortran local */
    /* per-ray arrays [nstokes,8(nb),8(n),numder] */
    float *grad8, *ograd8, *grad0, *grad1, *srcgrad;
    /* saved sub-interval records */
    int maxsub;
    int *passedpoints;             /* [8,maxsub] */
    double *passedinterp0, *passedinterp1;  /* [8,maxsub] */
    double *passeddels, *passedrad /*[nstokes,maxsub]*/, *passedabscell, *passedtransmit;
    ray_scratch *sc;
} grad_work;

#define LT(tab, i, l) (tab)[((i) - 1) + gw->nstleg * (l)]
#define G8(arr, ns, nb, n, idr) \
    (arr)[((ns) - 1) + gw->nstokes * (((nb) - 1) + 8 * (((n) - 1) + 8 * (size_t)((idr) - 1)))]

/* COMPUTE_SOURCE_DIRECTION  shdomsub4.f:2836-2914 */
static void compute_source_direction(const oracle_state *st, const grad_work *gw, const float *legen,
                                     float *sourcet, int ris, int rns, const float *ylmdir,
                                     float dirflux, float secmu0)
{
    const int nstokes = gw->nstokes, nstleg = gw->nstleg, nlm = gw->nlm;
    const int *lofj = gw->lofj;
    int j;
#define YD(i, jj) ylmdir[((i) - 1) + nstleg * ((jj) - 1)]
#define YS(i, jj) st->ylmsun[((i) - 1) + nstleg * ((jj) - 1)]
#define LEG(i, l) legen[((i) - 1) + nstleg * (l)]
    for (j = 1; j <= rns; j++)
        sourcet[0] = sourcet[0] + LEG(1, lofj[j - 1]) * RADIANCE(st, 1, ris + j) * YD(1, j);
    if (nstokes > 1) {
        for (j = 1; j <= rns; j++)
            sourcet[0] = sourcet[0] + LEG(5, lofj[j - 1]) * RADIANCE(st, 2, ris + j) * YD(1, j);
        for (j = 5; j <= rns; j++) {
            sourcet[1] = sourcet[1]
                + LEG(5, lofj[j - 1]) * RADIANCE(st, 1, ris + j) * YD(2, j)
                + LEG(2, lofj[j - 1]) * RADIANCE(st, 2, ris + j) * YD(2, j)
                + LEG(3, lofj[j - 1]) * RADIANCE(st, 3, ris + j) * YD(5, j);
            sourcet[2] = sourcet[2]
                + LEG(5, lofj[j - 1]) * RADIANCE(st, 1, ris + j) * YD(6, j)
                + LEG(2, lofj[j - 1]) * RADIANCE(st, 2, ris + j) * YD(6, j)
                + LEG(3, lofj[j - 1]) * RADIANCE(st, 3, ris + j) * YD(3, j);
        }
    }
    if (nstokes == 4) {
        for (j = 1; j <= rns; j++) {
            sourcet[1] = sourcet[1] + LEG(6, lofj[j - 1]) * RADIANCE(st, 4, ris + j) * YD(5, j);
            sourcet[2] = sourcet[2] + LEG(6, lofj[j - 1]) * RADIANCE(st, 4, ris + j) * YD(3, j);
            sourcet[3] = sourcet[3] - LEG(6, lofj[j - 1]) * RADIANCE(st, 3, ris + j) * YD(4, j)
                                    + LEG(4, lofj[j - 1]) * RADIANCE(st, 4, ris + j) * YD(4, j);
        }
    }
    if (!st->deltam && (st->srctype == 'S' || st->srctype == 'B')) {
        for (j = 1; j <= nlm; j++)
            sourcet[0] = sourcet[0] + dirflux * secmu0 * LEG(1, lofj[j - 1]) * YS(1, j) * YD(1, j);
        if (nstokes > 1)
            for (j = 5; j <= nlm; j++)
                sourcet[1] = sourcet[1] + dirflux * secmu0 * LEG(5, lofj[j - 1]) * YS(1, j) * YD(2, j);
    }
#undef YD
#undef YS
#undef LEG
}

/* LEGENT of one species at a grid point: the PHASEINTERPWT mix of the tabulated Legendre series and, with delta-M,
 * F = LEGENT(1,ML+1) and the division by 1-F  (shdomsub4.f:1700-1722) */
static void mix_legent(const oracle_state *st, grad_work *gw, const int *iph, const float *pw, float *legent)
{
    const int nstleg = st->nstleg, ml = st->ml, nlt = nstleg * (st->nleg + 1), nq = 8 * st->maxnmicro;
    int t, q, l, k;
    if (!st->interp_new) {
        const float *lg = &st->legen[(size_t)nlt * (iph[0] - 1)];
        for (t = 0; t < nlt; t++) legent[t] = lg[t];
    } else {
        if (pw[0] >= st->phasemax) {
            const float *lg = &st->legen[(size_t)nlt * (iph[0] - 1)];
            for (t = 0; t < nlt; t++) legent[t] = lg[t];
        } else {
            for (t = 0; t < nlt; t++) legent[t] = 0.0f;
            for (q = 0; q < nq; q++) {
                const float *lg;
                if (pw[q] <= 1e-5f) continue;
                lg = &st->legen[(size_t)nlt * (iph[q] - 1)];
                for (t = 0; t < nlt; t++) legent[t] = legent[t] + lg[t] * pw[q];
            }
        }
    }
    if (st->deltam) {
        gw->f = LT(legent, 1, ml + 1);
        for (l = 0; l <= ml; l++)
            for (k = 1; k <= nstleg; k++)
                LT(legent, k, l) = LT(legent, k, l) / (1 - gw->f);
    }
}

/* PLANCK_DERIVATIVE  shdomsub4.f:3171-3221 (UNITS 'T' and 'R'; the band integration of UNITS='B' is not restated) */
static float planck_derivative(float temp, int units, float wavelen)
{
    if (units == 'T') return 1.0f;
    if (temp > 0.0f) {
        const float e = expf(1.4388e4f / (wavelen * temp));
        const float w3 = wavelen * wavelen * wavelen;
        return (1.1911e8f * 1.4388e4f) / (w3 * w3) * e / (temp * temp * ((e - 1) * (e - 1)));
    }
    return 0.0f;
}

/* COMPUTE_SOURCE_GRAD_1CELL  shdomsub4.f:1546-2042 */
static void compute_source_grad_1cell(const oracle_state *st, const oracle_grad_in *g, grad_work *gw,
                                      int icell, const float *ylmdir, const float *singscat,
                                      const float *dsingscat, const int *donethis, int *oldipts,
                                      const float *oextinct8, const float *osrcext8,
                                      float *extinct8, float *srcext8,
                                      float *singscat8, const float *osingscat8)
{
    const int nstokes = st->nstokes, nstleg = st->nstleg, ml = st->ml, mm = st->mm;
    const int nleg = st->nleg, npart = st->npart, npts = st->npts, maxpg = g->maxpg;
    const int nq = 8 * st->maxnmicro, maxnmicro = st->maxnmicro;
    const int nlt = nstleg * (nleg + 1), numder = g->numder;
    const int solar = (st->srctype == 'S' || st->srctype == 'B');
    const int thermal = (st->srctype == 'T' || st->srctype == 'B');
    float secmu0 = (float)(1.0 / fabs((double)st->solarmu));
    float planck = 0.0f, dplanck = 0.0f;
    int n, k, j, l, m, q, ipa, t, idr, nb;
#define SRC8(kk, nn) srcext8[((kk) - 1) + nstokes * ((nn) - 1)]
#define OSRC8(kk, nn) osrcext8[((kk) - 1) + nstokes * ((nn) - 1)]
#define SS8(kk, nn) singscat8[((kk) - 1) + nstokes * ((nn) - 1)]
#define OSS8(kk, nn) osingscat8[((kk) - 1) + nstokes * ((nn) - 1)]
#define YD(i, jj) ylmdir[((i) - 1) + nstleg * ((jj) - 1)]
#define YS(i, jj) st->ylmsun[((i) - 1) + nstleg * ((jj) - 1)]
#define SSC(kk, ii) singscat[((kk) - 1) + nstokes * ((ii) - 1)]
#define DSSC(kk, ii) dsingscat[((kk) - 1) + nstokes * ((ii) - 1)]
    for (n = 1; n <= 8; n++) {
        int ip = GRIDPTR(st, n, icell);
        int i = donethis[n - 1];
        if (i > 0 && ip == oldipts[n - 1]) {
            extinct8[n - 1] = oextinct8[i - 1];
            for (k = 1; k <= nstokes; k++) { SRC8(k, n) = OSRC8(k, i); SS8(k, n) = OSS8(k, i); }
            for (idr = 1; idr <= numder; idr++)
                for (nb = 1; nb <= 8; nb++)
                    for (k = 1; k <= nstokes; k++)
                        G8(gw->grad8, k, nb, n, idr) = G8(gw->ograd8, k, nb, i, idr);
        } else if (i < 0) {
            extinct8[n - 1] = extinct8[-i - 1];
            for (k = 1; k <= nstokes; k++) { SRC8(k, n) = SRC8(k, -i); SS8(k, n) = SS8(k, -i); }
            for (idr = 1; idr <= numder; idr++)
                for (nb = 1; nb <= 8; nb++)
                    for (k = 1; k <= nstokes; k++)
                        G8(gw->grad8, k, nb, n, idr) = G8(gw->grad8, k, nb, -i, idr);
        } else {
            float ext = st->total_ext[ip - 1];
            int is, ns, ris, rns, last_ipa;
            float truncsingscat[4], sourcet[4], singscatj[4], scatterj = 0.0f;
            float singscatp[4], dsingscatp[4], dsource[4];
            float *legent = gw->legent;
            oldipts[n - 1] = ip;
            is = st->shptr[ip - 1];
            ns = st->shptr[ip] - is;
            for (idr = 1; idr <= numder; idr++)
                for (nb = 1; nb <= 8; nb++)
                    for (k = 1; k <= nstokes; k++) G8(gw->grad8, k, nb, n, idr) = 0.0f;
            ris = st->rshptr[ip - 1];
            rns = st->rshptr[ip] - ris;
            for (k = 1; k <= nstokes; k++) { SRC8(k, n) = 0.0f; SS8(k, n) = 0.0f; }
            for (j = 1; j <= ns; j++)
                SRC8(1, n) = SRC8(1, n) + SOURCE(st, 1, is + j) * YD(1, j);
            if (nstokes > 1) {
                for (j = 1; j <= ns; j++) {
                    SRC8(2, n) = SRC8(2, n) + SOURCE(st, 2, is + j) * YD(2, j)
                                            + SOURCE(st, 3, is + j) * YD(5, j);
                    SRC8(3, n) = SRC8(3, n) + SOURCE(st, 2, is + j) * YD(6, j)
                                            + SOURCE(st, 3, is + j) * YD(3, j);
                }
            }
            if (nstokes == 4)
                for (j = 1; j <= ns; j++)
                    SRC8(4, n) = SRC8(4, n) + SOURCE(st, 4, is + j) * YD(4, j);
            if (solar) {
                for (ipa = 1; ipa <= npart; ipa++) {
                    float w, da;
                    const int *iph = &st->iphase[(size_t)nq * ((ip - 1) + (size_t)npts * (ipa - 1))];
                    const float *pw = &st->phaseinterpwt[(size_t)nq * ((ip - 1) + (size_t)npts * (ipa - 1))];
                    if (ext == 0.0f) w = 1.0f;
                    else w = st->extinct[(ip - 1) + (size_t)npts * (ipa - 1)] / ext;
                    if (w == 0.0f) continue;
                    mix_legent(st, gw, iph, pw, legent);
                    da = st->albedo[(ip - 1) + (size_t)npts * (ipa - 1)] * st->dirflux[ip - 1] * secmu0 * w;
                    j = 1;
                    for (k = 0; k < 4; k++) truncsingscat[k] = 0.0f;
                    for (l = 0; l <= ml; l++) {
                        int me = l < mm ? l : mm;
                        int ms = -me;
                        float a1 = da * LT(legent, 1, l);
                        float b1 = nstleg > 1 ? da * LT(legent, 5, l) : 0.0f;
                        if (j <= ns) {
                            int jt = j;
                            for (m = ms; m <= me; m++) {
                                truncsingscat[0] = truncsingscat[0] + a1 * YD(1, j) * YS(1, j);
                                j = j + 1;
                            }
                            if (nstokes > 1) {
                                j = jt;
                                for (m = ms; m <= me; m++) {
                                    truncsingscat[1] = truncsingscat[1] + b1 * YD(2, j) * YS(1, j);
                                    truncsingscat[2] = truncsingscat[2] + b1 * YD(6, j) * YS(1, j);
                                    j = j + 1;
                                }
                            }
                        }
                    }
                    if (st->deltam) {
                        if (pw[0] >= st->phasemax) {
                            for (k = 1; k <= nstokes; k++)
                                SS8(k, n) = SS8(k, n) + da * SSC(k, iph[0]) / (1 - gw->f);
                        } else {
                            for (q = 0; q < nq; q++) {
                                if (pw[q] <= 1e-5f) continue;
                                for (k = 1; k <= nstokes; k++)
                                    SS8(k, n) = SS8(k, n) + da * SSC(k, iph[q]) * pw[q] / (1 - gw->f);
                            }
                        }
                        for (k = 1; k <= nstokes; k++) SRC8(k, n) = SRC8(k, n) - truncsingscat[k - 1];
                    } else {
                        for (k = 1; k <= nstokes; k++) SS8(k, n) = SS8(k, n) + truncsingscat[k - 1];
                    }
                }
                if (st->deltam)
                    for (k = 1; k <= nstokes; k++) SRC8(k, n) = SRC8(k, n) + SS8(k, n);
            } else if (npart == 1) {
                /* The reference sets LEGENT (and F) inside the solar block above only (shdomsub4.f:1693-1722) and then
                 * uses it for NPART=1 in the gradient part (:1835-1840, :1963-1985): with SRCTYPE='T' it is read
                 * undefined.  The defined value -- the same mix at this grid point -- is used here and on the GPU. */
                mix_legent(st, gw, &st->iphase[(size_t)nq * (ip - 1)], &st->phaseinterpwt[(size_t)nq * (ip - 1)], legent);
            }

            /* ---------------- gradient part (shdomsub4.f:1786-2019) ---------------- */
            if (thermal) {          /* shdomsub4.f:1792-1799 */
                const float wn[2] = {st->waveno0, st->waveno1};
                dplanck = planck_derivative(st->temp[ip - 1], st->units, st->wavelen);
                planck = oracle_planck_function(st->temp[ip - 1], st->units, wn, st->wavelen);
                if (planck < 1e-7f) dplanck = 0.0f;
            }
            last_ipa = -1;
            for (idr = 1; idr <= numder; idr++) {
                ipa = g->partder[idr - 1];
#define ALBP(ib) g->albedop[((ib) - 1) + (size_t)maxpg * (ipa - 1)]
#define EXTP(ib) g->extinctp[((ib) - 1) + (size_t)maxpg * (ipa - 1)]
#define PWP(qq, ib) g->phasewtp[((qq) - 1) + maxnmicro * (((ib) - 1) + (size_t)maxpg * (ipa - 1))]
#define IPHP(qq, ib) g->iphasep[((qq) - 1) + maxnmicro * (((ib) - 1) + (size_t)maxpg * (ipa - 1))]
#define INTERPPTR(nn, ii) g->interpptr[((nn) - 1) + 8 * (size_t)((ii) - 1)]
#define OPTW(nn, ii) g->optinterpwt[((nn) - 1) + 8 * (size_t)((ii) - 1)]
                if (ipa != last_ipa) {
                    last_ipa = ipa;
                    scatterj = 0.0f;
                    for (k = 0; k < 4; k++) singscatj[k] = 0.0f;
                    if (st->deltam) {
                        for (nb = 1; nb <= 8; nb++) {
                            int ib = INTERPPTR(nb, ip);
                            float xi = OPTW(nb, ip);
                            float spatial_weight = xi * ALBP(ib) * EXTP(ib);
                            scatterj = scatterj + spatial_weight;
                            if (spatial_weight <= 1e-6f) continue;
                            for (q = 1; q <= maxnmicro; q++) {
                                if (PWP(q, ib) <= 1e-6f) continue;
                                for (k = 1; k <= nstokes; k++)
                                    singscatj[k - 1] = singscatj[k - 1]
                                        + spatial_weight * PWP(q, ib) * SSC(k, IPHP(q, ib));
                            }
                        }
                        if (scatterj > g->scatmin) {
                            for (k = 0; k < nstokes; k++) singscatj[k] = singscatj[k] / scatterj;
                        } else {
                            for (k = 0; k < nstokes; k++) singscatj[k] = (float)(singscatj[k] / g->scatmin);
                        }
                    }
                    if (npart == 1) {
                        float alb = st->albedo[(ip - 1) + (size_t)npts * (ipa - 1)];
                        if (alb > 1e-8f) { for (k = 1; k <= nstokes; k++) sourcet[k - 1] = SRC8(k, n) / alb; }
                        else { for (k = 0; k < nstokes; k++) sourcet[k] = 0.0f; }
                    } else {
                        for (k = 0; k < 4; k++) sourcet[k] = 0.0f;
                        for (t = 0; t < nlt; t++) legent[t] = 0.0f;
                        gw->f = 0.0f;
                        for (nb = 1; nb <= 8; nb++) {
                            int ib = INTERPPTR(nb, ip);
                            float xi = OPTW(nb, ip);
                            float spatial_weight = xi * ALBP(ib) * EXTP(ib);
                            if (spatial_weight <= 1e-6f) continue;
                            for (q = 1; q <= maxnmicro; q++) {
                                const float *lg;
                                if (PWP(q, ib) <= 1e-6f) continue;
                                lg = &st->legen[(size_t)nlt * (IPHP(q, ib) - 1)];
                                for (t = 0; t < nlt; t++)
                                    legent[t] = legent[t] + spatial_weight * PWP(q, ib) * lg[t];
                            }
                        }
                        if (scatterj > g->scatmin) { for (t = 0; t < nlt; t++) legent[t] = legent[t] / scatterj; }
                        else { for (t = 0; t < nlt; t++) legent[t] = (float)(legent[t] / g->scatmin); }
                        if (st->deltam) {
                            gw->f = LT(legent, 1, ml + 1);
                            if (st->interp_new)
                                for (l = 0; l <= ml; l++)
                                    for (k = 1; k <= nstleg; k++)
                                        LT(legent, k, l) = LT(legent, k, l) / (1 - gw->f);
                        }
                        if (scatterj > g->scatmin) {
                            compute_source_direction(st, gw, legent, sourcet, ris, rns, ylmdir,
                                                     st->dirflux[ip - 1], secmu0);
                            if (st->deltam && solar)
                                for (k = 0; k < nstokes; k++)
                                    sourcet[k] = sourcet[k]
                                        + st->dirflux[ip - 1] * singscatj[k] * secmu0 / (1 - gw->f);
                        }
                    }
                    sourcet[0] = fmaxf(0.0f, sourcet[0]);
                    if (st->deltam) {
                        for (l = 0; l <= ml; l++)
                            for (k = 1; k <= nstleg; k++)
                                LT(legent, k, l) = LT(legent, k, l) * (1 - gw->f);
                        for (l = 0; l <= ml; l++) LT(legent, 1, l) = LT(legent, 1, l) + gw->f;
                        if (nstleg > 1)
                            for (l = 0; l <= ml; l++)
                                for (k = 2; k <= 4; k++) LT(legent, k, l) = LT(legent, k, l) + gw->f;
                    }
                }
                for (nb = 1; nb <= 8; nb++) {
                    int ib = INTERPPTR(nb, ip);
                    float xi = OPTW(nb, ip);
                    float dext_v = g->dext[(ib - 1) + (size_t)maxpg * (idr - 1)];
                    float dalb_v = g->dalb[(ib - 1) + (size_t)maxpg * (idr - 1)];
                    float dextm_v = g->dextm[(ib - 1) + (size_t)maxpg * (idr - 1)];
                    float dalbm_v = g->dalbm[(nb - 1) + 8 * ((ip - 1) + (size_t)npts * (idr - 1))];
                    float dfj_v = g->dfj[(nb - 1) + 8 * ((ip - 1) + (size_t)npts * (idr - 1))];
                    float alb_ip = st->albedo[(ip - 1) + (size_t)npts * (ipa - 1)];
                    if (xi < 1e-7f) continue;
                    for (t = 0; t < nlt; t++) gw->legenp[t] = 0.0f;
                    if (st->deltam) {
                        for (q = 1; q <= maxnmicro; q++) {
                            const float *lg = &st->legen[(size_t)nlt * (IPHP(q, ib) - 1)];
                            float ftemp;
                            for (t = 0; t < nlt; t++) gw->unscaled[t] = lg[t];
                            ftemp = LT(gw->unscaled, 1, ml + 1);
                            if (!st->interp_new)
                                for (l = 0; l <= ml; l++)
                                    LT(gw->unscaled, 1, l) = LT(gw->unscaled, 1, l) * (1 - ftemp);
                            for (l = 0; l <= ml; l++)
                                LT(gw->unscaled, 1, l) = LT(gw->unscaled, 1, l) + ftemp;
                            if (nstleg > 1)
                                for (l = 0; l <= ml; l++)
                                    for (k = 2; k <= 4; k++)
                                        LT(gw->unscaled, k, l) = LT(gw->unscaled, k, l) + ftemp;
                            for (t = 0; t < nlt; t++)
                                gw->legenp[t] = gw->legenp[t] + PWP(q, ib) * gw->unscaled[t];
                        }
                    } else {
                        for (q = 1; q <= maxnmicro; q++) {
                            const float *lg = &st->legen[(size_t)nlt * (IPHP(q, ib) - 1)];
                            for (t = 0; t < nlt; t++) gw->legenp[t] = gw->legenp[t] + PWP(q, ib) * lg[t];
                        }
                    }
                    for (t = 0; t < nlt; t++) gw->leg_diff[t] = gw->legenp[t] - legent[t];
                    if (solar && st->deltam) {
                        for (k = 0; k < 4; k++) singscatp[k] = 0.0f;
                        for (q = 1; q <= maxnmicro; q++)
                            for (k = 1; k <= nstokes; k++)
                                singscatp[k - 1] = singscatp[k - 1] + PWP(q, ib) * SSC(k, IPHP(q, ib));
                    }
                    for (t = 0; t < nlt; t++) gw->dlegp[t] = 0.0f;
                    for (k = 0; k < 4; k++) dsingscatp[k] = 0.0f;
                    for (q = 1; q <= g->deriv_maxnmicro; q++) {
                        int dip = g->diphasep[(q - 1) + g->deriv_maxnmicro * ((ib - 1) + (size_t)maxpg * (idr - 1))];
                        float dpw = g->dphasewtp[(q - 1) + g->deriv_maxnmicro * ((ib - 1) + (size_t)maxpg * (idr - 1))];
                        if (g->doexact[idr - 1] == 1) {
                            const float *dl = &g->dleg[(size_t)nlt * (dip - 1)];
                            if (st->deltam)
                                for (k = 1; k <= nstokes; k++)
                                    dsingscatp[k - 1] = dsingscatp[k - 1] + PWP(q, ib) * DSSC(k, dip);
                            for (t = 0; t < nlt; t++) gw->dlegp[t] = gw->dlegp[t] + PWP(q, ib) * dl[t];
                        } else if (g->doexact[idr - 1] == 0) {
                            const float *lg = &st->legen[(size_t)nlt * (IPHP(q, ib) - 1)];
                            for (t = 0; t < nlt; t++) gw->unscaled[t] = lg[t];
                            if (st->deltam) {
                                float ftemp;
                                for (k = 1; k <= nstokes; k++)
                                    dsingscatp[k - 1] = dsingscatp[k - 1] + dpw * SSC(k, IPHP(q, ib));
                                ftemp = LT(gw->unscaled, 1, ml + 1);
                                if (!st->interp_new)
                                    for (l = 0; l <= ml; l++)
                                        LT(gw->unscaled, 1, l) = LT(gw->unscaled, 1, l) * (1 - ftemp);
                                for (l = 0; l <= ml; l++)
                                    LT(gw->unscaled, 1, l) = LT(gw->unscaled, 1, l) + ftemp;
                                if (nstleg > 1)
                                    for (l = 0; l <= ml; l++)
                                        for (k = 2; k <= 4; k++)
                                            LT(gw->unscaled, k, l) = LT(gw->unscaled, k, l) + ftemp;
                            }
                            for (t = 0; t < nlt; t++) gw->dlegp[t] = gw->dlegp[t] + dpw * gw->unscaled[t];
                        }
                    }
                    for (t = 0; t < nlt; t++)
                        gw->dlegt[t] = dext_v * gw->leg_diff[t] * ALBP(ib)
                                     + dalb_v * gw->leg_diff[t] * EXTP(ib)
                                     + gw->dlegp[t] * EXTP(ib) * ALBP(ib)
                                     + (legent[t] - 1) * dfj_v;
                    for (k = 0; k < 4; k++) dsource[k] = 0.0f;
                    compute_source_direction(st, gw, gw->dlegt, dsource, ris, rns, ylmdir,
                                             st->dirflux[ip - 1], secmu0);
                    for (k = 1; k <= nstokes; k++)
                        G8(gw->grad8, k, nb, n, idr) = G8(gw->grad8, k, nb, n, idr)
                            + xi * (sourcet[k - 1] * (alb_ip * dextm_v + dalbm_v) + dsource[k - 1]);
                    if (solar && st->deltam) {
                        for (k = 1; k <= nstokes; k++)
                            G8(gw->grad8, k, nb, n, idr) = G8(gw->grad8, k, nb, n, idr)
                                + st->dirflux[ip - 1] * secmu0 * xi * (
                                    singscatj[k - 1] * dfj_v
                                  + dsingscatp[k - 1] * EXTP(ib) * ALBP(ib)
                                  + dalb_v * (singscatp[k - 1] - singscatj[k - 1]) * EXTP(ib)
                                  + dext_v * (singscatp[k - 1] - singscatj[k - 1]) * ALBP(ib));
                    }
                    if (thermal) {      /* shdomsub4.f:2009-2016 */
                        const float dtemp_v = g->dtemp ? g->dtemp[(ib - 1) + (size_t)maxpg * (idr - 1)] : 0.0f;
                        G8(gw->grad8, 1, nb, n, idr) = G8(gw->grad8, 1, nb, n, idr)
                            + xi * (st->extinct[(ip - 1) + (size_t)npts * (ipa - 1)] * (1.0f - alb_ip) * dplanck * dtemp_v
                                    - planck * dalbm_v
                                    + planck * (1.0f - alb_ip) * dextm_v);
                    }
                }
            }
            for (k = 1; k <= nstokes; k++) SS8(k, n) = SS8(k, n) * ext;
            if (g->singlescatter) { for (k = 1; k <= nstokes; k++) SRC8(k, n) = SS8(k, n); }
            else { for (k = 1; k <= nstokes; k++) SRC8(k, n) = SRC8(k, n) * ext; }
            extinct8[n - 1] = ext;
        }
    }
#undef SRC8
#undef OSRC8
#undef SS8
#undef OSS8
#undef YD
#undef YS
#undef SSC
#undef DSSC
}

static const int GRIDFACE[6][4] = {{1,3,5,7},{2,4,6,8},{1,2,5,6},{3,4,7,8},{1,2,3,4},{5,6,7,8}};
static const int OPPFACE[6] = {2, 1, 4, 3, 6, 5};

/* FIND_BOUNDARY_RADIANCE_GRAD  shdomsub4.f:2151-2347 (Lambertian; surface derivative outputs,
 * which the caller discards (shdomsub4.f:469-471), are not produced) */
static int find_boundary_radiance_grad(const oracle_state *st, const float *bcrad, double xb, double yb,
                                       float mu2, int icell, int kface, float *radbnd,
                                       int *boundpts, double *boundinterp, double *dirrad /*[nstokes,4]*/,
                                       char *errmsg)
{
    const int nstokes = st->nstokes;
    float x[4], y[4], rad[4][4];
    float opi = 1.0f / acosf(-1.0f);
    double u, v;
    int j, k;
    if (st->sfctype1 != 'L') {
        if (errmsg) snprintf(errmsg, 600, "oracle: only Lambertian surfaces are restated");
        return 3;
    }
    for (j = 0; j < 4 * nstokes; j++) dirrad[j] = 0.0;
    for (j = 0; j < 4; j++) {
        int ip = GRIDPTR(st, GRIDFACE[kface - 1][j], icell);
        int ibc;
        boundpts[j] = ip;
        x[j] = GRIDPOS(st, 1, ip);
        y[j] = GRIDPOS(st, 2, ip);
        if (mu2 < 0.0f) {
            ibc = oracle_bc_search(st->bcptr, st->ntoppts, ip);
            if (!ibc) { if (errmsg) snprintf(errmsg, 600, "FIND_BOUNDARY_RADIANCE: Not at boundary"); return 1; }
            for (k = 0; k < nstokes; k++) rad[j][k] = bcrad[k + nstokes * (ibc - 1)];
        } else {
            ibc = oracle_bc_search(st->bcptr + st->maxnbc, st->nbotpts, ip);
            if (!ibc) { if (errmsg) snprintf(errmsg, 600, "FIND_BOUNDARY_RADIANCE: Not at boundary"); return 1; }
            if (st->srctype == 'S' || st->srctype == 'B') {
                if (st->sfctype0 == 'V')
                    dirrad[nstokes * j] = opi * st->sfcgridparms[1 + st->nsfcpar * (ibc - 1)] * st->dirflux[ip - 1];
                else if (st->sfctype0 == 'F')
                    dirrad[nstokes * j] = opi * st->gndalbedo * st->dirflux[ip - 1];
            }
            /* RADEMIS (COMPUTE_TOP_RADIANCES_GRAD flag 2 on SFCGRIDRAD, shdomsub4.f:2274-2290): SFCGRIDRAD is zero on
             * this path (checked by the caller), so the interpolated value is 0 and RADEMIS = 0 ('S','B') or
             * PLANCK_FUNCTION(0) ('T') */
            for (k = 0; k < nstokes; k++) rad[j][k] = 0.0f;
            if (st->srctype == 'T') {
                const float wn[2] = {st->waveno0, st->waveno1};
                rad[j][0] = oracle_planck_function(0.0f, st->units, wn, st->wavelen);
            }
            for (k = 0; k < nstokes; k++)
                rad[j][k] = rad[j][k] + bcrad[k + nstokes * (st->ntoppts + ibc - 1)];
        }
    }
    if (x[1] - x[0] > 0.0f) u = (xb - x[0]) / (x[1] - x[0]); else u = 0.0;
    if (y[2] - y[0] > 0.0f) v = (yb - y[0]) / (y[2] - y[0]); else v = 0.0;
    boundinterp[0] = (1 - u) * (1 - v);
    boundinterp[1] = u * (1 - v);
    boundinterp[2] = (1 - u) * v;
    boundinterp[3] = u * v;
    for (k = 0; k < nstokes; k++)
        radbnd[k] = (float)((1 - u) * (1 - v) * rad[0][k] + u * (1 - v) * rad[1][k]
                            + (1 - u) * v * rad[2][k] + u * v * rad[3][k]);
    return 0;
}

/* COMPUTE_RADIANCE_DERIVATIVE_ADJOINT  shdomsub4.f:4037-4114 */
static void compute_radiance_derivative_adjoint(const oracle_state *st, const oracle_grad_in *g,
                                                const grad_work *gw, const double *adj_weight,
                                                double *gradout, int npassed)
{
    const int nstokes = st->nstokes, maxpg = g->maxpg;
    int kk, k, nb, idr, ns;
    for (kk = 1; kk <= npassed - 1; kk++) {
        double ext0 = 0.0, ext1 = 0.0, dels = gw->passeddels[kk - 1], ext;
        const int *pp = &gw->passedpoints[8 * (size_t)(kk - 1)];
        const double *i0 = &gw->passedinterp0[8 * (size_t)(kk - 1)];
        const double *i1 = &gw->passedinterp1[8 * (size_t)(kk - 1)];
        for (k = 0; k < 8; k++) {
            ext0 = ext0 + st->total_ext[pp[k] - 1] * i0[k];
            ext1 = ext1 + st->total_ext[pp[k] - 1] * i1[k];
        }
        ext = 0.5f * (ext0 + ext1);
        if (ext != 0.0) {
            for (k = 0; k < 8; k++) {
                int ip = pp[k];
                double gbase0 = 0.0, gbase1 = 0.0, adj_gbase;
                for (ns = 0; ns < nstokes; ns++) {
                    gbase0 = gbase0 - adj_weight[ns] * gw->passedrad[ns + nstokes * (size_t)kk] * i0[k];
                    gbase1 = gbase1 - adj_weight[ns] * gw->passedrad[ns + nstokes * (size_t)(kk - 1)] * i1[k];
                }
                adj_gbase = (0.5f * (gbase0 + gbase1)
                             + 0.08333333333f * (ext0 * gbase1 - ext1 * gbase0) * dels
                               * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext;
                adj_gbase = adj_gbase * gw->passedtransmit[kk - 1] * gw->passedabscell[kk - 1];
                for (idr = 1; idr <= g->numder; idr++) {
                    for (nb = 1; nb <= 8; nb++) {
                        int ib = g->interpptr[(nb - 1) + 8 * (size_t)(ip - 1)];
                        float xi = g->optinterpwt[(nb - 1) + 8 * (size_t)(ip - 1)];
                        float extgrad = g->dextm[(ib - 1) + (size_t)maxpg * (idr - 1)] * xi;
                        gradout[(ib - 1) + (size_t)maxpg * (idr - 1)] += extgrad * adj_gbase;
                    }
                }
            }
        }
    }
}

/* COMPUTE_RADIANCE_DERIVATIVE  shdomsub4.f:2588-2662 (single-sweep / Jacobian path; REAL temporaries) */
static void compute_radiance_derivative(const oracle_state *st, const oracle_grad_in *g,
                                        const grad_work *gw, double *raygrad, int npassed)
{
    const int nstokes = st->nstokes, maxpg = g->maxpg;
    int kk, k, nb, idr, ns;
    for (kk = 1; kk <= npassed - 1; kk++) {
        double ext0 = 0.0, ext1 = 0.0, dels = gw->passeddels[kk - 1], ext;
        const int *pp = &gw->passedpoints[8 * (size_t)(kk - 1)];
        const double *i0 = &gw->passedinterp0[8 * (size_t)(kk - 1)];
        const double *i1 = &gw->passedinterp1[8 * (size_t)(kk - 1)];
        for (k = 0; k < 8; k++) {
            ext0 = ext0 + st->total_ext[pp[k] - 1] * i0[k];
            ext1 = ext1 + st->total_ext[pp[k] - 1] * i1[k];
        }
        ext = 0.5f * (ext0 + ext1);
        if (ext != 0.0) {
            for (k = 0; k < 8; k++) {
                int ip = pp[k];
                for (idr = 1; idr <= g->numder; idr++) {
                    for (nb = 1; nb <= 8; nb++) {
                        int ib = g->interpptr[(nb - 1) + 8 * (size_t)(ip - 1)];
                        float xi = g->optinterpwt[(nb - 1) + 8 * (size_t)(ip - 1)];
                        float extgrad = g->dextm[(ib - 1) + (size_t)maxpg * (idr - 1)] * xi;
                        for (ns = 0; ns < nstokes; ns++) {
                            float radgrad0 = (float)(-1 * gw->passedrad[ns + nstokes * (size_t)kk] * extgrad * i0[k]);
                            float radgrad1 = (float)(-1 * gw->passedrad[ns + nstokes * (size_t)(kk - 1)] * extgrad * i1[k]);
                            float radgrad = (float)((0.5f * (radgrad0 + radgrad1)
                                + 0.08333333333f * (ext0 * radgrad1 - ext1 * radgrad0) * dels
                                  * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext);
                            raygrad[ns + nstokes * ((size_t)(ib - 1) + (size_t)maxpg * (idr - 1))] +=
                                radgrad * gw->passedtransmit[kk - 1] * gw->passedabscell[kk - 1];
                        }
                    }
                }
            }
        }
    }
}

/* COMPUTE_DIRECT_BEAM_DERIV  shdomsub4.f:2778-2834 */
static void compute_direct_beam_deriv(const oracle_grad_in *g, int nstokes, int ip, double transmit,
                                      double abscell, const float *inputweight, double *raygrad)
{
    const float *dpath = &g->dpath[(size_t)g->longest_path_pts * (ip - 1)];
    const int *dptr = &g->dptr[(size_t)g->longest_path_pts * (ip - 1)];
    int ii = 1, idr, ns;
    while (ii <= g->longest_path_pts && dptr[ii - 1] > 0) {
        int ib = dptr[ii - 1];
        for (idr = 1; idr <= g->numder; idr++)
            for (ns = 0; ns < nstokes; ns++)
                raygrad[ns + nstokes * ((size_t)(ib - 1) + (size_t)g->maxpg * (idr - 1))] -=
                    g->dextm[(ib - 1) + (size_t)g->maxpg * (idr - 1)] * dpath[ii - 1] * abscell * transmit
                    * inputweight[ns];
        ii = ii + 1;
    }
}

/* ADJOINT_INTEGRATE_1RAY  shdomsub4.f:3223-3967; with RAYGRAD != NULL the same walk is GRAD_INTEGRATE_1RAY
 * (shdomsub4.f:811-1544, the single-sweep / Jacobian path): the two routines share everything up to
 * where the per-sub-interval source gradient goes (RAYGRAD(NSTOKES,MAXPG,NUMDER) un-contracted vs.
 * GRADOUT contracted with the adjoint weight), how the direct-beam derivative is applied (per cell,
 * COMPUTE_DIRECT_BEAM_DERIV, vs. BEAM_WEIGHT batching) and the radiance-derivative routine. */

static int adjoint_integrate_1ray(const oracle_state *st, const oracle_grad_in *g, grad_work *gw,
                                  const float *bcrad, double mu2, double phi2,
                                  double x0, double y0, double z0, const double *adj_weight,
                                  double *gradout, double *beam_weight, double *raygrad, double *radout_ret,
                                  int *trace_cells, int trace_cap, int *trace_n, int *nsub_out,
                                  char *errmsg)
{
    const int nstokes = st->nstokes, numder = g->numder, maxpg = g->maxpg;
    const size_t ng8 = (size_t)nstokes * 64 * numder;
    ray_dir rd;
    float *ylmdir = gw->sc->ylmdir, *singscat = gw->sc->singscat, *dsingscat = gw->sc->dsingscat;
    int oldipts[8] = {0,0,0,0,0,0,0,0}, donethis[8];
    float oextinct8[8], osrcext8[32], extinct8[8], srcext8[32], singscat8[32], osingscat8[32];
    float singscat0[32], singscat1[32], srcsingscat[32];
    float ext0, ext1, extn, srcext0[4], srcext1[4], radbnd[4];
    double fc[8], fcn[8];
    double xe, ye, ze, xn, yn, zn, xi, yi, zi, so, sox, soy, soz, eps;
    double taugrid, s, dels, ext, tau, transcell, abscell, src[4], radout[4] = {0, 0, 0, 0};
    double transmit = 1.0, dirrad[16], boundinterp[4];
    int boundpts[4];
    int icell, inextcell, iface, jface, kface, ic, iopp, ntau, it, i, k, n, kk, idr, ns, ngrid;
    int ipinx, ipiny, openbcface, validrad, npassed, ntrace = 0, nsub = 0;
    const int exact_ss = g->exact_single_scatter && (st->srctype == 'S' || st->srctype == 'B');
    const double tautol = st->tautol, transcut = st->transcut;
    const int maxsub = gw->maxsub;
    size_t t;

    memset(extinct8, 0, sizeof(extinct8)); memset(srcext8, 0, sizeof(srcext8));
    memset(singscat8, 0, sizeof(singscat8)); memset(singscat1, 0, sizeof(singscat1));
    memset(singscat0, 0, sizeof(singscat0));
    memset(gw->grad8, 0, sizeof(float) * ng8); memset(gw->grad1, 0, sizeof(float) * ng8);
    memset(gw->grad0, 0, sizeof(float) * ng8);
    eps = 1.0e-5f * (GRIDPOS(st, 3, GRIDPTR(st, 8, 1)) - GRIDPOS(st, 3, GRIDPTR(st, 1, 1)));
    npassed = 1;
    oracle_ray_setup(st, mu2, phi2, &rd, ylmdir, singscat, g->dphasetab, g->dnumphase, dsingscat);
    xe = x0; ye = y0; ze = z0;
    icell = oracle_locate_grid_cell(st, &xe, &ye, &ze);
    iface = 0;
    ngrid = 0;
    validrad = 0;
    ext1 = 0.0f;
    for (k = 0; k < 4; k++) srcext1[k] = 0.0f;
    while (!validrad && icell > 0) {
        ngrid = ngrid + 1;
        if (trace_cells && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        oracle_donethis(st, iface, donethis);
        for (i = 0; i < 8; i++) {
            oextinct8[i] = extinct8[i];
            for (k = 0; k < nstokes; k++) {
                osrcext8[k + nstokes * i] = srcext8[k + nstokes * i];
                osingscat8[k + nstokes * i] = singscat8[k + nstokes * i];
            }
        }
        memcpy(gw->ograd8, gw->grad8, sizeof(float) * ng8);
        compute_source_grad_1cell(st, g, gw, icell, ylmdir, singscat, dsingscat, donethis, oldipts,
                                  oextinct8, osrcext8, extinct8, srcext8, singscat8, osingscat8);
        get_interp_kernel(st, icell, xe, ye, ze, fc);
#define FCSUM(F, A) ((F)[0] * A(1) + (F)[1] * A(2) + (F)[2] * A(3) + (F)[3] * A(4) \
                     + (F)[4] * A(5) + (F)[5] * A(6) + (F)[6] * A(7) + (F)[7] * A(8))
#define E8(nn) extinct8[(nn) - 1]
        for (k = 0; k < nstokes; k++) {
#define S8(nn) srcext8[k + nstokes * ((nn) - 1)]
            srcext1[k] = (float)FCSUM(fc, S8);
#undef S8
        }
        srcext1[0] = fmaxf(0.0f, srcext1[0]);
        ext1 = (float)FCSUM(fc, E8);
        for (k = 0; k < 8; k++) gw->passedinterp1[k + 8 * (size_t)(npassed - 1)] = fc[k];
        for (n = 1; n <= 8; n++) {
            for (k = 0; k < nstokes; k++)
                singscat1[k + nstokes * (n - 1)] = (float)(fc[n - 1] * singscat8[k + nstokes * (n - 1)]);
            singscat1[nstokes * (n - 1)] = fmaxf(0.0f, singscat1[nstokes * (n - 1)]);
        }
        for (idr = 1; idr <= numder; idr++)
            for (n = 1; n <= 8; n++)
                for (kk = 1; kk <= 8; kk++)
                    for (k = 1; k <= nstokes; k++)
                        G8(gw->grad1, k, kk, n, idr) = (float)(fc[n - 1] * G8(gw->grad8, k, kk, n, idr));
        ipinx = BTEST(CELLFLAGS(st, icell), 0) &&
                !(BTEST(st->bcflag, 0) && ((rd.cx > 0 && xe < rd.xm) || (rd.cx < 0 && xe > rd.xm)));
        ipiny = BTEST(CELLFLAGS(st, icell), 1) &&
                !(BTEST(st->bcflag, 1) && ((rd.cy > 0 && ye < rd.ym) || (rd.cy < 0 && ye > rd.ym)));
        iopp = GRIDPTR(st, 9 - rd.ioct, icell);
        if (ipinx) sox = 1.0e20f; else sox = (GRIDPOS(st, 1, iopp) - xe) * rd.cxinv;
        if (ipiny) soy = 1.0e20f; else soy = (GRIDPOS(st, 2, iopp) - ye) * rd.cyinv;
        soz = (GRIDPOS(st, 3, iopp) - ze) * rd.czinv;
        so = fmin(fmin(sox, soy), soz);
        if (so < -eps) {
            if (errmsg) snprintf(errmsg, 600, "ADJOINT_INTEGRATE_1RAY: SO<0 %g %g %g %g %g %g %d",
                                 mu2, phi2, xe, ye, ze, so, icell);
            return 1;
        }
        xn = xe + so * rd.cx;
        yn = ye + so * rd.cy;
        zn = ze + so * rd.cz;
        get_interp_kernel(st, icell, xn, yn, zn, fcn);
        extn = (float)FCSUM(fcn, E8);
        taugrid = so * 0.5f * (ext1 + extn);
        ntau = 1 + (int)(taugrid / tautol);
        if (ntau < 1) ntau = 1;
        dels = so / ntau;
        memset(srcsingscat, 0, sizeof(srcsingscat));
        for (it = 1; it <= ntau; it++) {
            gw->passeddels[npassed - 1] = dels;
            for (k = 0; k < 8; k++) {
                gw->passedinterp1[k + 8 * (size_t)(npassed - 1)] = fc[k];
                gw->passedpoints[k + 8 * (size_t)(npassed - 1)] = GRIDPTR(st, k + 1, icell);
            }
            s = it * dels;
            xi = xe + s * rd.cx;
            yi = ye + s * rd.cy;
            zi = ze + s * rd.cz;
            get_interp_kernel(st, icell, xi, yi, zi, fc);
            for (k = 0; k < nstokes; k++) {
#define S8(nn) srcext8[k + nstokes * ((nn) - 1)]
                srcext0[k] = (float)FCSUM(fc, S8);
#undef S8
            }
            if (it != ntau) ext0 = (float)FCSUM(fc, E8);
            else ext0 = extn;
            srcext0[0] = fmaxf(0.0f, srcext0[0]);
            for (k = 0; k < 8; k++) gw->passedinterp0[k + 8 * (size_t)(npassed - 1)] = fc[k];
            for (n = 1; n <= 8; n++) {
                for (k = 0; k < nstokes; k++)
                    singscat0[k + nstokes * (n - 1)] = (float)(fc[n - 1] * singscat8[k + nstokes * (n - 1)]);
                singscat0[nstokes * (n - 1)] = fmaxf(0.0f, singscat0[nstokes * (n - 1)]);
            }
            for (idr = 1; idr <= numder; idr++)
                for (n = 1; n <= 8; n++)
                    for (kk = 1; kk <= 8; kk++)
                        for (k = 1; k <= nstokes; k++)
                            G8(gw->grad0, k, kk, n, idr) = (float)(fc[n - 1] * G8(gw->grad8, k, kk, n, idr));
            ext = 0.5f * (ext0 + ext1);
            if (ext != 0.0) {
                tau = ext * dels;
                abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                transcell = 1.0f - abscell;
                for (k = 0; k < nstokes; k++)
                    src[k] = (0.5f * (srcext0[k] + srcext1[k])
                              + 0.08333333333f * (ext0 * srcext1[k] - ext1 * srcext0[k]) * dels
                                * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext;
                for (t = 0; t < ng8; t++)
                    gw->srcgrad[t] = (float)((0.5f * (gw->grad0[t] + gw->grad1[t])
                              + 0.08333333333f * (ext0 * gw->grad1[t] - ext1 * gw->grad0[t]) * dels
                                * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext);
                if (g->exact_single_scatter && st->srctype != 'T') {
                    for (i = 0; i < 8 * nstokes; i++)
                        srcsingscat[i] = (float)(srcsingscat[i] + transmit * abscell *
                            (0.5f * (singscat0[i] + singscat1[i])
                             + 0.08333333333f * (ext0 * singscat1[i] - ext1 * singscat0[i]) * dels
                               * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext);
                }
                for (k = 0; k < nstokes; k++) radout[k] = radout[k] + transmit * src[k] * abscell;
                gw->passedabscell[npassed - 1] = abscell;
                gw->passedtransmit[npassed - 1] = transmit;
                for (k = 0; k < nstokes; k++)
                    gw->passedrad[k + nstokes * (size_t)(npassed - 1)] = transmit * src[k] * abscell;
                for (kk = 1; kk <= 8; kk++) {
                    int ip = GRIDPTR(st, kk, icell);
                    for (k = 1; k <= 8; k++) {
                        int ib = g->interpptr[(k - 1) + 8 * (size_t)(ip - 1)];
                        for (idr = 1; idr <= numder; idr++) {
                            if (raygrad) {       /* shdomsub4.f:1335-1341 */
                                for (ns = 1; ns <= nstokes; ns++)
                                    raygrad[(ns - 1) + nstokes * ((size_t)(ib - 1) + (size_t)maxpg * (idr - 1))] +=
                                        transmit * G8(gw->srcgrad, ns, k, kk, idr) * abscell;
                            } else {
                                double contrib = 0.0;
                                for (ns = 1; ns <= nstokes; ns++)
                                    contrib = contrib + adj_weight[ns - 1] * G8(gw->srcgrad, ns, k, kk, idr);
                                gradout[(ib - 1) + (size_t)maxpg * (idr - 1)] += transmit * contrib * abscell;
                            }
                        }
                    }
                }
                npassed = npassed + 1;
                nsub++;
                if (npassed > maxsub) {
                    if (errmsg) snprintf(errmsg, 600, "ADJOINT_INTEGRATE_1RAY: The maximum number of "
                        "subgrid intervals for calculation of the radiance along the ray path has been "
                        "exceeded. NPASSED=%d MAXSUBGRIDINTS=%d", npassed, maxsub);
                    return 1;
                }
            } else {
                abscell = 0.0;
                transcell = 1.0;
                for (k = 0; k < nstokes; k++) src[k] = 0.0;
                memset(gw->srcgrad, 0, sizeof(float) * ng8);
                memset(srcsingscat, 0, sizeof(srcsingscat));
            }
            transmit = transmit * transcell;
            ext1 = ext0;
            for (k = 0; k < nstokes; k++) srcext1[k] = srcext0[k];
            memcpy(gw->grad1, gw->grad0, sizeof(float) * ng8);
            memcpy(singscat1, singscat0, sizeof(singscat1));
        }
#undef E8
        if (exact_ss) {
            for (kk = 1; kk <= 8; kk++) {
                int ip = GRIDPTR(st, kk, icell);
                if (raygrad) {           /* shdomsub4.f:1386-1398 */
                    compute_direct_beam_deriv(g, nstokes, ip, 1.0, 1.0, &srcsingscat[nstokes * (kk - 1)], raygrad);
                } else {
                    for (ns = 0; ns < nstokes; ns++)
                        beam_weight[ip - 1] = beam_weight[ip - 1]
                            + adj_weight[ns] * srcsingscat[ns + nstokes * (kk - 1)];
                }
            }
        }
        if (sox <= soz && sox <= soy) {
            iface = 2 - rd.bitx; jface = 1;
            openbcface = BTEST(CELLFLAGS(st, icell), 0) && BTEST(st->bcflag, 0);
        } else if (soy <= soz) {
            iface = 4 - rd.bity; jface = 2;
            openbcface = BTEST(CELLFLAGS(st, icell), 1) && BTEST(st->bcflag, 1);
        } else {
            iface = 6 - rd.bitz; jface = 3;
            openbcface = 0;
        }
        inextcell = NEIGHPTR(st, iface, icell);
        if (inextcell < 0)
            inextcell = oracle_next_cell(st, xn, yn, zn, iface, jface, icell);
        if (NEIGHPTR(st, iface, icell) >= 0 && !openbcface) {
            kface = iface;
            ic = icell;
        } else {
            kface = OPPFACE[iface - 1];
            ic = inextcell;
            iface = 0;
        }
        if (inextcell > 0) {
            if (jface == 1) xn = GRIDPOS(st, 1, GRIDPTR(st, rd.ioct, inextcell));
            else if (jface == 2) yn = GRIDPOS(st, 2, GRIDPTR(st, rd.ioct, inextcell));
            else zn = GRIDPOS(st, 3, GRIDPTR(st, rd.ioct, inextcell));
        }
        if (transmit < transcut) {
            validrad = 1;
            for (k = 0; k < nstokes; k++) gw->passedrad[k + nstokes * (size_t)(npassed - 1)] = 0.0;
            gw->passedtransmit[npassed - 1] = 1.0;
        } else if (inextcell == 0 && iface >= 5) {
            int ierr;
            validrad = 1;
            ierr = find_boundary_radiance_grad(st, bcrad, xn, yn, (float)mu2, ic, kface, radbnd,
                                               boundpts, boundinterp, dirrad, errmsg);
            if (ierr) return ierr;
            for (k = 0; k < nstokes; k++) radout[k] = radout[k] + transmit * radbnd[k];
            gw->passedtransmit[npassed - 1] = transmit;
            gw->passedabscell[npassed - 1] = -1.0;
            for (k = 0; k < nstokes; k++)
                gw->passedrad[k + nstokes * (size_t)(npassed - 1)] = transmit * radbnd[k];
            if (exact_ss) {
                for (kk = 0; kk < 4; kk++) {
                    int ip = boundpts[kk];
                    if (raygrad) {       /* shdomsub4.f:1495-1505 */
                        float wgt[4];
                        for (ns = 0; ns < nstokes; ns++) wgt[ns] = (float)(boundinterp[kk] * dirrad[ns + nstokes * kk]);
                        compute_direct_beam_deriv(g, nstokes, ip, transmit, 1.0, wgt, raygrad);
                    } else {
                        for (ns = 0; ns < nstokes; ns++)
                            beam_weight[ip - 1] = beam_weight[ip - 1]
                                + adj_weight[ns] * transmit * boundinterp[kk] * dirrad[ns + nstokes * kk];
                    }
                }
            }
        } else {
            icell = inextcell;
        }
        xe = xn; ye = yn; ze = zn;
    }
    /* NOTE: if the loop ended because ICELL<=0 (ray left an open-BC side), PASSEDRAD(:,NPASSED) and
     * PASSEDTRANSMIT(NPASSED) are whatever the allocator left there in the reference; we use 0 / 1. */
    if (!validrad) {
        for (k = 0; k < nstokes; k++) gw->passedrad[k + nstokes * (size_t)(npassed - 1)] = 0.0;
        gw->passedtransmit[npassed - 1] = 1.0;
    }
    for (kk = npassed - 1; kk >= 1; kk--)
        for (k = 0; k < nstokes; k++)
            gw->passedrad[k + nstokes * (size_t)(kk - 1)] += gw->passedrad[k + nstokes * (size_t)kk];
    for (kk = 1; kk <= npassed; kk++)
        for (k = 0; k < nstokes; k++)
            gw->passedrad[k + nstokes * (size_t)(kk - 1)] /= gw->passedtransmit[kk - 1];
    if (raygrad) compute_radiance_derivative(st, g, gw, raygrad, npassed);
    else compute_radiance_derivative_adjoint(st, g, gw, adj_weight, gradout, npassed);
    if (radout_ret) for (k = 0; k < nstokes; k++) radout_ret[k] = radout[k];
    if (trace_n) *trace_n = ntrace;
    if (nsub_out) *nsub_out = nsub;
    return 0;
}

/* COMPUTE_ADJOINT_WEIGHTS  shdomsub4.f:3969-4034 */
static int compute_adjoint_weights(const double *stokesout, const double *measurement,
                                   const double *unc, int costfunc_ll, int nstokes, int nunc,
                                   double *adj_weight, double *cost)
{
    int i, j;
#define UNC(a, b) unc[((a) - 1) + nunc * ((b) - 1)]
    for (i = 0; i < nstokes; i++) adj_weight[i] = 0.0;
    if (!costfunc_ll) {
        for (i = 1; i <= nstokes; i++) {
            double pixel_error = stokesout[i - 1] - measurement[i - 1];
            for (j = 1; j <= nstokes; j++) {
                *cost = *cost + 0.5 * UNC(i, j) * (pixel_error * pixel_error);
                adj_weight[i - 1] = adj_weight[i - 1] + UNC(i, j) * pixel_error;
            }
        }
    } else {
        double raderror = log(stokesout[0]) - log(measurement[0]);
        *cost = *cost + 0.5 * (raderror * raderror * UNC(1, 1));
        adj_weight[0] = adj_weight[0] + raderror * UNC(1, 1) / stokesout[0];
        if (nstokes > 1) {
            double dolp1 = sqrt(stokesout[1] * stokesout[1] + stokesout[2] * stokesout[2]) / stokesout[0];
            double dolp2 = sqrt(measurement[1] * measurement[1] + measurement[2] * measurement[2]) / measurement[0];
            double dolperr = log(dolp1) - log(dolp2);
            *cost = *cost + 0.5 * (dolperr * dolperr * UNC(2, 2));
            adj_weight[1] = adj_weight[1] + dolperr * UNC(2, 2) * stokesout[1]
                            / (stokesout[1] * stokesout[1] + stokesout[2] * stokesout[2]);
            adj_weight[2] = adj_weight[2] + dolperr * UNC(2, 2) * stokesout[2]
                            / (stokesout[1] * stokesout[1] + stokesout[2] * stokesout[2]);
        }
    }
#undef UNC
    return 0;
}

/* UPDATE_COSTFUNCTION  shdomsub4.f:13-91 */
int oracle_update_costfunction(const double *stokesout, const double *raygrad_pixel,
                               double *gradout, double *cost, const double *unc,
                               int costfunc_ll, int nstokes, int maxpg, int numder,
                               const double *measurement, int nunc)
{
    size_t n = (size_t)maxpg * numder, t;
    int i, j;
#define UNC(a, b) unc[((a) - 1) + nunc * ((b) - 1)]
#define RG(ns, tt) raygrad_pixel[((ns) - 1) + (size_t)nstokes * (tt)]
    if (!costfunc_ll) {
        for (i = 1; i <= nstokes; i++) {
            double pixel_error = stokesout[i - 1] - measurement[i - 1];
            for (j = 1; j <= nstokes; j++) {
                cost[0] = cost[0] + 0.5 * UNC(i, j) * (pixel_error * pixel_error);
                for (t = 0; t < n; t++)
                    gradout[t] = gradout[t] + UNC(i, j) * pixel_error * RG(i, t);
            }
        }
    } else {
        double raderror = log(stokesout[0]) - log(measurement[0]);
        cost[0] = cost[0] + 0.5 * (raderror * raderror * UNC(1, 1));
        for (t = 0; t < n; t++)
            gradout[t] = gradout[t] + raderror * UNC(1, 1) * RG(1, t) / stokesout[0];
        if (nstokes > 1) {
            double dolp1 = sqrt(stokesout[1] * stokesout[1] + stokesout[2] * stokesout[2]) / stokesout[0];
            double dolp2 = sqrt(measurement[1] * measurement[1] + measurement[2] * measurement[2]) / measurement[0];
            double dolperr = log(dolp1) - log(dolp2);
            cost[0] = cost[0] + 0.5 * (dolperr * dolperr * UNC(2, 2));
            for (t = 0; t < n; t++)
                gradout[t] = gradout[t] + dolperr * UNC(2, 2)
                    * (stokesout[1] * RG(2, t) + stokesout[2] * RG(3, t))
                    / (stokesout[1] * stokesout[1] + stokesout[2] * stokesout[2]);
        }
    }
#undef UNC
#undef RG
    return 0;
}

static grad_work *grad_work_new(const oracle_state *st, const oracle_grad_in *g)
{
    grad_work *gw = (grad_work *)calloc(1, sizeof(grad_work));
    size_t nlt = (size_t)st->nstleg * (st->nleg + 2);
    size_t ng8 = (size_t)st->nstokes * 64 * g->numder;
    int j = 0, l, m;
    int maxcells = 50 * IMAX3(st->nx, st->ny, st->nz);
    gw->nstokes = st->nstokes; gw->nstleg = st->nstleg; gw->nleg = st->nleg; gw->nlm = st->nlm;
    gw->ml = st->ml; gw->mm = st->mm; gw->numder = g->numder;
    gw->lofj = (int *)malloc(sizeof(int) * st->nlm);
    for (l = 0; l <= st->ml; l++) {
        int me = l < st->mm ? l : st->mm;
        for (m = -me; m <= me; m++) gw->lofj[j++] = l;
    }
    gw->legent = (float *)calloc(nlt, sizeof(float));
    gw->legenp = (float *)calloc(nlt, sizeof(float));
    gw->unscaled = (float *)calloc(nlt, sizeof(float));
    gw->dlegp = (float *)calloc(nlt, sizeof(float));
    gw->dlegt = (float *)calloc(nlt, sizeof(float));
    gw->leg_diff = (float *)calloc(nlt, sizeof(float));
    gw->grad8 = (float *)calloc(ng8, sizeof(float));
    gw->ograd8 = (float *)calloc(ng8, sizeof(float));
    gw->grad0 = (float *)calloc(ng8, sizeof(float));
    gw->grad1 = (float *)calloc(ng8, sizeof(float));
    gw->srcgrad = (float *)calloc(ng8, sizeof(float));
    gw->maxsub = g->maxsubgridints > maxcells ? g->maxsubgridints : maxcells;
    gw->passedpoints = (int *)calloc((size_t)8 * (gw->maxsub + 1), sizeof(int));
    gw->passedinterp0 = (double *)calloc((size_t)8 * (gw->maxsub + 1), sizeof(double));
    gw->passedinterp1 = (double *)calloc((size_t)8 * (gw->maxsub + 1), sizeof(double));
    gw->passeddels = (double *)calloc((size_t)gw->maxsub + 1, sizeof(double));
    gw->passedrad = (double *)calloc((size_t)st->nstokes * (gw->maxsub + 1), sizeof(double));
    gw->passedabscell = (double *)calloc((size_t)gw->maxsub + 1, sizeof(double));
    gw->passedtransmit = (double *)calloc((size_t)gw->maxsub + 1, sizeof(double));
    gw->sc = oracle_scratch_new(st, g->dnumphase);
    return gw;
}

static void grad_work_free(grad_work *gw)
{
    free(gw->lofj); free(gw->legent); free(gw->legenp); free(gw->unscaled); free(gw->dlegp);
    free(gw->dlegt); free(gw->leg_diff); free(gw->grad8); free(gw->ograd8); free(gw->grad0);
    free(gw->grad1); free(gw->srcgrad); free(gw->passedpoints); free(gw->passedinterp0);
    free(gw->passedinterp1); free(gw->passeddels); free(gw->passedrad); free(gw->passedabscell);
    free(gw->passedtransmit); oracle_scratch_free(gw->sc); free(gw);
}

static void set_top_bcrad(const oracle_state *st, float *bcrad, double mu2, double phi2)
{   /* shdomsub4.f:662-672 (top boundary radiances for this ray direction) */
    const int nstokes = st->nstokes;
    int itop, k;
    if (-mu2 > 0.0) {
        float sky = oracle_sky_radiance(st, (float)mu2, (float)phi2);
        for (itop = 0; itop < st->ntoppts; itop++) {
            bcrad[nstokes * itop] = sky;
            for (k = 1; k < nstokes; k++) bcrad[k + nstokes * itop] = 0.0f;
        }
    } else {
        for (itop = 0; itop < st->ntoppts * nstokes; itop++) bcrad[itop] = 0.0f;
    }
}

/* LEVISAPPROX_GRADIENT (MAKEJACOBIAN=.FALSE.)  shdomsub4.f:288-809.
 * nthreads>1: pixels are split into contiguous chunks, one private GRADOUT/BEAM_WEIGHT per thread,
 * summed in thread order -- the reference's own thread scheme (at3d/parallel.py:83-100). */
int oracle_levisapprox_gradient(const oracle_state *st, const oracle_rays *rays,
                                const oracle_grad_in *g, double *gradout, double *cost,
                                float *stokesout, oracle_trace *trace, int nthreads, char *errmsg)
{
    const int nstokes = st->nstokes, npix = g->npix, npts = st->npts;
    const size_t ngrad = (size_t)g->maxpg * g->numder;
    double *adj_weights, *beam_weight;
    int *raystart;
    int ierr_all = 0, ipix, ip, idr, ii;
    size_t t;
    if (st->sfctype1 != 'L') {
        /* the reference itself STOPs here: SURFACE_BRDF_GRAD has no linearisation for W/D/O/R (surface.f:395-399) */
        if (errmsg) snprintf(errmsg, 600, "oracle: gradient with a non-Lambertian surface is not restated");
        return 3;
    }
    if (st->srctype != 'S') {
        int i_;
        if (st->units == 'B') { if (errmsg) snprintf(errmsg, 600, "oracle: UNITS='B' is not restated"); return 3; }
        if (!st->temp) { if (errmsg) snprintf(errmsg, 600, "oracle: thermal gradient needs TEMP"); return 1; }
        if (st->sfcgridrad)
            for (i_ = 0; i_ < (st->nang / 2 + 1) * st->nbotpts; i_++)
                if (st->sfcgridrad[i_] != 0.0f) {
                    if (errmsg) snprintf(errmsg, 600, "oracle: gradient with SFCGRIDRAD != 0 is not restated");
                    return 3;
                }
    }
    if (nthreads < 1) nthreads = 1;
    oracle_lambertian_boundary(st, st->bcrad);
    raystart = (int *)malloc(sizeof(int) * (npix + 1));
    raystart[0] = 0;
    for (ipix = 0; ipix < npix; ipix++) raystart[ipix + 1] = raystart[ipix] + g->rays_per_pixel[ipix];
    adj_weights = (double *)calloc((size_t)nstokes * npix, sizeof(double));
    beam_weight = (double *)calloc((size_t)npts, sizeof(double));
    for (t = 0; t < ngrad; t++) gradout[t] = 0.0;
    for (t = 0; t < (size_t)nstokes * npix; t++) stokesout[t] = 0.0f;

    /* ---- Phase 1: forward radiance pass (shdomsub4.f:636-698) ---- */
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        ray_scratch *sc = oracle_scratch_new(st, 0);
        size_t nbc = (size_t)nstokes * (st->ntoppts + st->nbotpts);
        float *bcrad = (float *)malloc(sizeof(float) * (nbc + 1));
        char lmsg[600];
        int jp, iray, k;
        memcpy(bcrad, st->bcrad, sizeof(float) * nbc);
        lmsg[0] = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (jp = 0; jp < npix; jp++) {
            if (ierr_all) continue;
            for (iray = raystart[jp]; iray < raystart[jp + 1]; iray++) {
                double x0 = rays->camx[iray], y0 = rays->camy[iray], z0 = rays->camz[iray];
                double mu2 = rays->cammu[iray], phi2 = rays->camphi[iray];
                double transmit = 1.0, visrad[4] = {0, 0, 0, 0};
                int ierr = 0, dark;
                dark = oracle_ray_start(st, mu2, phi2, &x0, &y0, &z0, &ierr);
                if (ierr) snprintf(lmsg, 600, "LEVISAPPROX_GRADIENT: Level below domain");
                else if (!dark) {
                    set_top_bcrad(st, bcrad, mu2, phi2);
                    ierr = oracle_integrate_1ray(st, bcrad, 0.0f, mu2, phi2, x0, y0, z0, &transmit, visrad,
                                                 1, g->singlescatter, 0, sc, NULL, 0, NULL, NULL, lmsg);
                }
                if (ierr) {
#ifdef _OPENMP
#pragma omp critical
#endif
                    { if (!ierr_all) { ierr_all = ierr; if (errmsg) { strncpy(errmsg, lmsg, 599); errmsg[599] = 0; } } }
                    break;
                }
                for (k = 0; k < nstokes; k++)
                    stokesout[k + nstokes * jp] = (float)(stokesout[k + nstokes * jp]
                        + visrad[k] * g->ray_weights[iray] * g->stokes_weights[k + nstokes * jp]);
            }
        }
        free(bcrad);
        oracle_scratch_free(sc);
    }
    if (ierr_all) goto done;

    /* ---- Phase 2: adjoint weights + cost (shdomsub4.f:700-709) ---- */
    cost[0] = 0.0;
    for (ipix = 0; ipix < npix; ipix++) {
        double so[4], me[4];
        int k;
        for (k = 0; k < nstokes; k++) {
            so[k] = (double)stokesout[k + nstokes * ipix];
            me[k] = (double)g->measurements[k + nstokes * ipix];
        }
        compute_adjoint_weights(so, me, &g->uncertainties[(size_t)g->nuncertainty * g->nuncertainty * ipix],
                                g->costfunc_ll, nstokes, g->nuncertainty,
                                &adj_weights[nstokes * ipix], cost);
    }

    /* ---- Phase 3: adjoint derivative pass (shdomsub4.f:711-790) ---- */
    {
        int nt = nthreads;
        double **pgrad = (double **)calloc(nt, sizeof(double *));
        double **pbeam = (double **)calloc(nt, sizeof(double *));
        int tid;
        pgrad[0] = gradout;
        pbeam[0] = beam_weight;
        for (tid = 1; tid < nt; tid++) {
            pgrad[tid] = (double *)calloc(ngrad, sizeof(double));
            pbeam[tid] = (double *)calloc(npts, sizeof(double));
        }
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
        {
#ifdef _OPENMP
            int me = omp_get_thread_num();
#else
            int me = 0;
#endif
            grad_work *gw = grad_work_new(st, g);
            size_t nbc = (size_t)nstokes * (st->ntoppts + st->nbotpts);
            float *bcrad = (float *)malloc(sizeof(float) * (nbc + 1));
            char lmsg[600];
            int jp, iray, k;
            /* contiguous pixel chunk for this thread */
            int p0 = (int)(((long long)npix * me) / nt), p1 = (int)(((long long)npix * (me + 1)) / nt);
            memcpy(bcrad, st->bcrad, sizeof(float) * nbc);
            lmsg[0] = 0;
            for (jp = p0; jp < p1 && !ierr_all; jp++) {
                for (iray = raystart[jp]; iray < raystart[jp + 1]; iray++) {
                    double x0 = rays->camx[iray], y0 = rays->camy[iray], z0 = rays->camz[iray];
                    double mu2 = rays->cammu[iray], phi2 = rays->camphi[iray];
                    double weight_vec[4];
                    int ierr = 0, dark, ntr = 0, nsub = 0;
                    dark = oracle_ray_start(st, mu2, phi2, &x0, &y0, &z0, &ierr);
                    if (ierr) snprintf(lmsg, 600, "LEVISAPPROX_GRADIENT: Level below domain");
                    else if (!dark) {
                        set_top_bcrad(st, bcrad, mu2, phi2);
                        for (k = 0; k < nstokes; k++)
                            weight_vec[k] = adj_weights[k + nstokes * jp] * g->ray_weights[iray]
                                            * g->stokes_weights[k + nstokes * jp];
                        ierr = adjoint_integrate_1ray(st, g, gw, bcrad, mu2, phi2, x0, y0, z0, weight_vec,
                                    pgrad[me], pbeam[me], NULL, NULL,
                                    trace ? trace->cells + (size_t)trace->max_per_ray * iray : NULL,
                                    trace ? trace->max_per_ray : 0, &ntr, &nsub, lmsg);
                    }
                    if (trace) { trace->ncells[iray] = ntr; trace->nsub[iray] = nsub; }
                    if (ierr) {
#ifdef _OPENMP
#pragma omp critical
#endif
                        { if (!ierr_all) { ierr_all = ierr; if (errmsg) { strncpy(errmsg, lmsg, 599); errmsg[599] = 0; } } }
                        break;
                    }
                }
            }
            free(bcrad);
            grad_work_free(gw);
        }
        for (tid = 1; tid < nt; tid++) {
            for (t = 0; t < ngrad; t++) gradout[t] += pgrad[tid][t];
            for (t = 0; t < (size_t)npts; t++) beam_weight[t] += pbeam[tid][t];
            free(pgrad[tid]); free(pbeam[tid]);
        }
        free(pgrad); free(pbeam);
    }
    if (ierr_all) goto done;

    /* ---- Phase 4: batched direct beam derivatives (shdomsub4.f:792-802, 4117-4143) ---- */
    if (g->exact_single_scatter && (st->srctype == 'S' || st->srctype == 'B')) {
        for (ip = 1; ip <= npts; ip++) {
            if (beam_weight[ip - 1] != 0.0) {
                const float *dpath = &g->dpath[(size_t)g->longest_path_pts * (ip - 1)];
                const int *dptr = &g->dptr[(size_t)g->longest_path_pts * (ip - 1)];
                ii = 1;
                while (ii <= g->longest_path_pts && dptr[ii - 1] > 0) {
                    int ib = dptr[ii - 1];
                    for (idr = 1; idr <= g->numder; idr++)
                        gradout[(ib - 1) + (size_t)g->maxpg * (idr - 1)] -=
                            g->dextm[(ib - 1) + (size_t)g->maxpg * (idr - 1)] * dpath[ii - 1]
                            * beam_weight[ip - 1];
                    ii = ii + 1;
                }
            }
        }
    }
done:
    free(raystart); free(adj_weights); free(beam_weight);
    return ierr_all;
}

/* LEVISAPPROX_GRADIENT with MAKEJACOBIAN=.TRUE. (single sweep)  shdomsub4.f:536-631: per ray
 * GRAD_INTEGRATE_1RAY -> RAYGRAD, per pixel RAYGRAD_PIXEL -> UPDATE_COSTFUNCTION and
 * JACOBIAN(:,:,JI,IPIX) = RAYGRAD_PIXEL(:,JACOBIANPTR(JI),:).  Serial (test sizes only). */
int oracle_levisapprox_jacobian(const oracle_state *st, const oracle_rays *rays,
                                const oracle_grad_in *g, int num_jacobian_pts, const int *jacobianptr,
                                double *gradout, double *cost, float *stokesout, float *jacobian,
                                char *errmsg)
{
    const int nstokes = st->nstokes, npix = g->npix, numder = g->numder, maxpg = g->maxpg;
    const size_t ngrad = (size_t)maxpg * numder, nrg = (size_t)nstokes * ngrad;
    size_t nbc = (size_t)nstokes * (st->ntoppts + st->nbotpts), t;
    double *raygrad, *raygrad_pixel;
    float *bcrad;
    grad_work *gw;
    int ipix, iray = 0, i2, k, ji, idr, ierr = 0;
    if (st->sfctype1 != 'L') {
        /* the reference itself STOPs here: SURFACE_BRDF_GRAD has no linearisation for W/D/O/R (surface.f:395-399) */
        if (errmsg) snprintf(errmsg, 600, "oracle: gradient with a non-Lambertian surface is not restated");
        return 3;
    }
    if (st->srctype != 'S') {
        int i_;
        if (st->units == 'B') { if (errmsg) snprintf(errmsg, 600, "oracle: UNITS='B' is not restated"); return 3; }
        if (!st->temp) { if (errmsg) snprintf(errmsg, 600, "oracle: thermal gradient needs TEMP"); return 1; }
        if (st->sfcgridrad)
            for (i_ = 0; i_ < (st->nang / 2 + 1) * st->nbotpts; i_++)
                if (st->sfcgridrad[i_] != 0.0f) {
                    if (errmsg) snprintf(errmsg, 600, "oracle: gradient with SFCGRIDRAD != 0 is not restated");
                    return 3;
                }
    }
    oracle_lambertian_boundary(st, st->bcrad);
    raygrad = (double *)calloc(nrg, sizeof(double));
    raygrad_pixel = (double *)calloc(nrg, sizeof(double));
    bcrad = (float *)malloc(sizeof(float) * (nbc + 1));
    memcpy(bcrad, st->bcrad, sizeof(float) * nbc);
    gw = grad_work_new(st, g);
    for (t = 0; t < ngrad; t++) gradout[t] = 0.0;
    for (t = 0; t < (size_t)nstokes * npix; t++) stokesout[t] = 0.0f;
    cost[0] = 0.0;
    for (ipix = 0; ipix < npix && !ierr; ipix++) {
        double so[4], me[4];
        for (t = 0; t < nrg; t++) raygrad_pixel[t] = 0.0;
        for (i2 = 0; i2 < g->rays_per_pixel[ipix] && !ierr; i2++, iray++) {
            double x0 = rays->camx[iray], y0 = rays->camy[iray], z0 = rays->camz[iray];
            double mu2 = rays->cammu[iray], phi2 = rays->camphi[iray];
            double visrad[4] = {0, 0, 0, 0};
            int dark = oracle_ray_start(st, mu2, phi2, &x0, &y0, &z0, &ierr);
            if (ierr) { if (errmsg) snprintf(errmsg, 600, "LEVISAPPROX_GRADIENT: Level below domain"); break; }
            for (t = 0; t < nrg; t++) raygrad[t] = 0.0;
            if (!dark) {
                set_top_bcrad(st, bcrad, mu2, phi2);
                ierr = adjoint_integrate_1ray(st, g, gw, bcrad, mu2, phi2, x0, y0, z0, NULL, NULL, NULL,
                                              raygrad, visrad, NULL, 0, NULL, NULL, errmsg);
                if (ierr) break;
            }
            for (k = 0; k < nstokes; k++) {
                const double wgt = g->ray_weights[iray] * g->stokes_weights[k + nstokes * ipix];
                stokesout[k + nstokes * ipix] = (float)(stokesout[k + nstokes * ipix] + visrad[k] * wgt);
                for (t = 0; t < ngrad; t++) raygrad_pixel[k + nstokes * t] += raygrad[k + nstokes * t] * wgt;
            }
        }
        if (ierr) break;
        for (k = 0; k < nstokes; k++) {
            so[k] = (double)stokesout[k + nstokes * ipix];
            me[k] = (double)g->measurements[k + nstokes * ipix];
        }
        oracle_update_costfunction(so, raygrad_pixel, gradout, cost,
                                   &g->uncertainties[(size_t)g->nuncertainty * g->nuncertainty * ipix],
                                   g->costfunc_ll, nstokes, maxpg, numder, me, g->nuncertainty);
        for (ji = 0; ji < num_jacobian_pts; ji++)
            for (idr = 0; idr < numder; idr++)
                for (k = 0; k < nstokes; k++)
                    jacobian[k + nstokes * (idr + numder * (ji + (size_t)num_jacobian_pts * ipix))] =
                        (float)raygrad_pixel[k + nstokes * ((size_t)(jacobianptr[ji] - 1) + (size_t)maxpg * idr)];
    }
    free(raygrad); free(raygrad_pixel); free(bcrad);
    grad_work_free(gw);
    return ierr;
}

/* COMPUTE_INTERP_WEIGHTS  shdomsub4.f:3082-3169 */
static int compute_interp_weights(float x, float y, float z, int npx, int npy, int npz,
                                  float delx, float dely, float xstart, float ystart,
                                  const float *zlevels, int *interpptr, float *optinterpwt, char *errmsg)
{
    int il = 0, iu = npz, im, iz, ix, ixp, iy, iyp, i1, i2, i3, i4;
    double u, v, w;
    while (iu - il > 1) {
        im = (iu + il) / 2;
        if (z >= zlevels[im - 1]) il = im; else iu = im;
    }
    iz = il > 1 ? il : 1;
    w = (double)(z - zlevels[iz - 1]) / (zlevels[iz] - zlevels[iz - 1]);
    w = fmax(fmin(w, 1.0), 0.0);
    ix = (int)((x - xstart) / delx) + 1;
    if (fabsf(x - xstart - npx * delx) < 0.01f * delx) ix = npx;
    if (ix < 1 || ix > npx) {
        if (errmsg) snprintf(errmsg, 600, "TRILIN: Beyond X domain %d %d %g %g", ix, npx, x, xstart);
        return 1;
    }
    ixp = (ix % npx) + 1;
    u = (double)(x - xstart - delx * (ix - 1)) / delx;
    u = fmax(fmin(u, 1.0), 0.0);
    if (u < 1.0e-5) u = 0.0;
    if (u > 1.0 - 1.0e-5) u = 1.0;
    iy = (int)((y - ystart) / dely) + 1;
    if (fabsf(y - ystart - npy * dely) < 0.01f * dely) iy = npy;
    if (iy < 1 || iy > npy) {
        if (errmsg) snprintf(errmsg, 600, "TRILIN: Beyond Y domain %d %d %g %g", iy, npy, y, ystart);
        return 1;
    }
    iyp = (iy % npy) + 1;
    v = (double)(y - ystart - dely * (iy - 1)) / dely;
    v = fmax(fmin(v, 1.0), 0.0);
    if (v < 1.0e-5) v = 0.0;
    if (v > 1.0 - 1.0e-5) v = 1.0;
    optinterpwt[0] = (float)((1 - u) * (1 - v) * (1 - w));
    optinterpwt[1] = (float)(u * (1 - v) * (1 - w));
    optinterpwt[2] = (float)((1 - u) * v * (1 - w));
    optinterpwt[3] = (float)(u * v * (1 - w));
    optinterpwt[4] = (float)((1 - u) * (1 - v) * w);
    optinterpwt[5] = (float)(u * (1 - v) * w);
    optinterpwt[6] = (float)((1 - u) * v * w);
    optinterpwt[7] = (float)(u * v * w);
    i1 = iz + npz * (iy - 1) + npz * npy * (ix - 1);
    i2 = iz + npz * (iy - 1) + npz * npy * (ixp - 1);
    i3 = iz + npz * (iyp - 1) + npz * npy * (ix - 1);
    i4 = iz + npz * (iyp - 1) + npz * npy * (ixp - 1);
    interpptr[0] = i1; interpptr[1] = i2; interpptr[2] = i3; interpptr[3] = i4;
    interpptr[4] = i1 + 1; interpptr[5] = i2 + 1; interpptr[6] = i3 + 1; interpptr[7] = i4 + 1;
    return 0;
}

/* PREPARE_DERIV_INTERPS  shdomsub4.f:2917-3079 */
int oracle_prepare_deriv_interps(const oracle_state *st, int npx, int npy, int npz, int maxpg,
                                 float delx, float dely, float xstart, float ystart,
                                 const float *zlevels, const oracle_grad_in *g,
                                 float *optinterpwt, int *interpptr,
                                 float *dalbm, float *dextm, float *dfj, char *errmsg)
{
    const int npts = st->npts, numder = g->numder, ml = st->ml, nstleg = st->nstleg, nleg = st->nleg;
    const int maxnmicro = st->maxnmicro, nq = 8 * st->maxnmicro;
    const size_t nlt = (size_t)nstleg * (nleg + 1);
    float *fp_buf, *dfp_buf;
    int ip, idr, q, ipa, nb, ib, ierr;
#define LEGEN1(l, iph) st->legen[0 + nstleg * ((l) + (size_t)(nleg + 1) * ((iph) - 1))]
#define DLEG1(l, iph) g->dleg[0 + nstleg * ((l) + (size_t)(nleg + 1) * ((iph) - 1))]
    (void)nlt;
    for (ip = 1; ip <= npts; ip++) {
        ierr = compute_interp_weights(GRIDPOS(st, 1, ip), GRIDPOS(st, 2, ip), GRIDPOS(st, 3, ip),
                                      npx, npy, npz, delx, dely, xstart, ystart, zlevels,
                                      &interpptr[8 * (size_t)(ip - 1)], &optinterpwt[8 * (size_t)(ip - 1)],
                                      errmsg);
        if (ierr) return ierr;
    }
    fp_buf = (float *)malloc(sizeof(float) * (size_t)maxpg * numder);
    dfp_buf = (float *)malloc(sizeof(float) * (size_t)maxpg * numder);
    for (idr = 1; idr <= numder; idr++) {
        ipa = g->partder[idr - 1];
        for (ib = 1; ib <= maxpg; ib++) {
            float fp_val = 0.0f, dfp_val = 0.0f;
            float albp = g->albedop[(ib - 1) + (size_t)maxpg * (ipa - 1)];
            float extp = g->extinctp[(ib - 1) + (size_t)maxpg * (ipa - 1)];
            float dext = g->dext[(ib - 1) + (size_t)maxpg * (idr - 1)];
            float dalb = g->dalb[(ib - 1) + (size_t)maxpg * (idr - 1)];
            if (st->deltam) {
                for (q = 1; q <= g->deriv_maxnmicro; q++) {
                    float pwp = g->phasewtp[(q - 1) + maxnmicro * ((ib - 1) + (size_t)maxpg * (ipa - 1))];
                    int iphp = g->iphasep[(q - 1) + maxnmicro * ((ib - 1) + (size_t)maxpg * (ipa - 1))];
                    fp_val = fp_val + pwp * LEGEN1(ml + 1, iphp);
                    if (g->doexact[idr - 1] == 1) {
                        int dip = g->diphasep[(q - 1) + g->deriv_maxnmicro * ((ib - 1) + (size_t)maxpg * (idr - 1))];
                        dfp_val = dfp_val + pwp * DLEG1(ml + 1, dip);
                    } else if (g->doexact[idr - 1] == 0) {
                        float dpw = g->dphasewtp[(q - 1) + g->deriv_maxnmicro * ((ib - 1) + (size_t)maxpg * (idr - 1))];
                        dfp_val = dfp_val + dpw * LEGEN1(ml + 1, iphp);
                    }
                }
            }
            fp_buf[(ib - 1) + (size_t)maxpg * (idr - 1)] = fp_val;
            dfp_buf[(ib - 1) + (size_t)maxpg * (idr - 1)] = dfp_val;
            dextm[(ib - 1) + (size_t)maxpg * (idr - 1)] =
                dext * (1 - fp_val * albp) - dalb * fp_val * extp - extp * albp * dfp_val;
        }
    }
    for (idr = 1; idr <= numder; idr++) {
        ipa = g->partder[idr - 1];
        for (ip = 1; ip <= npts; ip++) {
            const int *iph = &st->iphase[(size_t)nq * ((ip - 1) + (size_t)npts * (ipa - 1))];
            const float *pw = &st->phaseinterpwt[(size_t)nq * ((ip - 1) + (size_t)npts * (ipa - 1))];
            float alb = st->albedo[(ip - 1) + (size_t)npts * (ipa - 1)];
            float f = 0.0f, albedoj, divide;
            if (st->deltam) {
                if (!st->interp_new) {
                    f = LEGEN1(ml + 1, iph[0]);
                } else {
                    if (pw[0] >= st->phasemax) {
                        f = LEGEN1(ml + 1, iph[0]);
                    } else {
                        for (q = 0; q < nq; q++) {
                            if (pw[q] < 1e-7f) continue;
                            f = f + pw[q] * LEGEN1(ml + 1, iph[q]);
                        }
                    }
                }
                albedoj = alb / (f * (alb - 1) + 1);
            } else {
                albedoj = alb;
            }
            divide = 1.0f / (1.0f - albedoj * f);
            for (nb = 1; nb <= 8; nb++) {
                float albp, extp, dext, dalb, fp, dfp;
                ib = interpptr[(nb - 1) + 8 * (size_t)(ip - 1)];
                albp = g->albedop[(ib - 1) + (size_t)maxpg * (ipa - 1)];
                extp = g->extinctp[(ib - 1) + (size_t)maxpg * (ipa - 1)];
                dext = g->dext[(ib - 1) + (size_t)maxpg * (idr - 1)];
                dalb = g->dalb[(ib - 1) + (size_t)maxpg * (idr - 1)];
                fp = fp_buf[(ib - 1) + (size_t)maxpg * (idr - 1)];
                dfp = dfp_buf[(ib - 1) + (size_t)maxpg * (idr - 1)];
                dalbm[(nb - 1) + 8 * ((ip - 1) + (size_t)npts * (idr - 1))] = divide * (
                    dext * ((1 - f) * (albp - albedoj) + (albedoj - 1) * albp * (fp - f))
                    + dalb * ((1 - f) * extp + (albedoj - 1) * extp * (fp - f))
                    + dfp * (albedoj - 1) * extp * albp);
                dfj[(nb - 1) + 8 * ((ip - 1) + (size_t)npts * (idr - 1))] =
                    (dext * (fp - f) * albp + dalb * (fp - f) * extp + dfp * extp * albp) / (1 - f);
            }
        }
    }
    free(fp_buf); free(dfp_buf);
#undef LEGEN1
#undef DLEG1
    return 0;
}

/* average_subpixel_rays  util.f90:484-518 (pixel_index holds 0-based pixel numbers; the last
 * ray is added to the last pixel separately, exactly as the reference does) */
void oracle_average_subpixel_rays(int npixels, int nrays, int nstokes, const float *weighted_stokes,
                                  const int *pixel_index, float *observables)
{
    double *temp = (double *)calloc(nstokes, sizeof(double));
    int i, k, iray = 1, pixind = 0;
    for (i = 1; i <= npixels; i++) {
        for (k = 0; k < nstokes; k++) temp[k] = 0.0;
        while (pixind + 1 == i) {
            for (k = 0; k < nstokes; k++)
                temp[k] = temp[k] + weighted_stokes[k + nstokes * (size_t)(iray - 1)];
            iray = iray + 1;
            if (iray >= nrays) pixind = i + 100;
            else pixind = pixel_index[iray - 1];
        }
        for (k = 0; k < nstokes; k++) observables[k + nstokes * (size_t)(i - 1)] = (float)temp[k];
    }
    for (k = 0; k < nstokes; k++)
        observables[k + nstokes * (size_t)(npixels - 1)] =
            observables[k + nstokes * (size_t)(npixels - 1)] + weighted_stokes[k + nstokes * (size_t)(nrays - 1)];
    free(temp);
}
