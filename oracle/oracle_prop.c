/*
 * oracle_prop.c -- TEST INFRASTRUCTURE ONLY (see shdom_oracle.h).
 * Property grid -> RTE grid, restated from (paths relative to /root/reference):
 *   src/polarized/shdom90.f90:17-346    TRILIN_INTERP_PROP
 *   src/polarized/shdomsub2.f:313-390   INTERP_GRID
 *   src/polarized/shdomsub2.f:479-612   PREPARE_PROP (delta-M scaling, both INTERPMETHOD(2:2) modes)
 *   src/polarized/shdomsub1.f:19-112    TRANSFER_PA_TO_GRID
 *   src/polarized/shdomsub2.f:4961-5244 SSORT (SLATEC; the tie order decides cell numbering in SPLIT_GRID)
 * REAL = float, DOUBLE PRECISION = double, mixed expressions promoted exactly as Fortran does.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

/* SSORT(X, Y, N, KFLAG)  shdomsub2.f:4961-5244: Singleton's quicksort on X carrying the INTEGER array Y.
 * kflag = 2 / -2 (increasing / decreasing, carry Y), 1 / -1 (X only). Arrays are 1-based inside. */
void oracle_ssort(float *x0, int *y0, int n, int kflag)
{
    float *x = x0 - 1;
    int *y = y0 ? y0 - 1 : NULL;
    float r, t, tt;
    int ty = 0, tty;
    int i, ij, j, k, kk, l, m, nn = n;
    int il[52], iu[52];
    if (nn < 1) return;
    kk = abs(kflag);
    if (kflag <= -1) for (i = 1; i <= nn; i++) x[i] = -x[i];
    (void)kk;
    m = 1; i = 1; j = nn; r = 0.375f;
L110:
    if (i == j) goto L150;
    if (r <= 0.5898437f) r = r + 3.90625e-2f; else r = r - 0.21875f;
L120:
    k = i;
    ij = i + (int)((j - i) * r);
    t = x[ij]; if (y) ty = y[ij];
    if (x[i] > t) {
        x[ij] = x[i]; x[i] = t; t = x[ij];
        if (y) { y[ij] = y[i]; y[i] = ty; ty = y[ij]; }
    }
    l = j;
    if (x[j] < t) {
        x[ij] = x[j]; x[j] = t; t = x[ij];
        if (y) { y[ij] = y[j]; y[j] = ty; ty = y[ij]; }
        if (x[i] > t) {
            x[ij] = x[i]; x[i] = t; t = x[ij];
            if (y) { y[ij] = y[i]; y[i] = ty; ty = y[ij]; }
        }
    }
L130:
    l = l - 1;
    if (x[l] > t) goto L130;
L140:
    k = k + 1;
    if (x[k] < t) goto L140;
    if (k <= l) {
        tt = x[l]; x[l] = x[k]; x[k] = tt;
        if (y) { tty = y[l]; y[l] = y[k]; y[k] = tty; }
        goto L130;
    }
    if (l - i > j - k) { il[m] = i; iu[m] = l; i = k; m = m + 1; }
    else { il[m] = k; iu[m] = j; j = l; m = m + 1; }
    goto L160;
L150:
    m = m - 1;
    if (m == 0) goto L190;
    i = il[m]; j = iu[m];
L160:
    if (j - i >= 1) goto L120;
    if (i == 1) goto L110;
    i = i - 1;
L170:
    i = i + 1;
    if (i == j) goto L150;
    t = x[i + 1]; if (y) ty = y[i + 1];
    if (x[i] <= t) goto L170;
    k = i;
L180:
    x[k + 1] = x[k]; if (y) y[k + 1] = y[k];
    k = k - 1;
    if (t < x[k]) goto L180;
    x[k + 1] = t; if (y) y[k + 1] = ty;
    goto L170;
L190:
    if (kflag <= -1) for (i = 1; i <= nn; i++) x[i] = -x[i];
}

/* EXTMIN, SCATMIN of TRILIN_INTERP_PROP(INIT=.TRUE.)  shdom90.f90:73-75 */
void oracle_prop_extmin(const oracle_prop *pg, double *extmin, double *scatmin)
{
    float e = 1.0e-5f / ((pg->zlevels[pg->npz - 1] - pg->zlevels[0]) / pg->npz);
    *extmin = (double)e;
    *scatmin = 0.1f * *extmin;
}

/* TRILIN_INTERP_PROP (INIT=.FALSE.)  shdom90.f90:78-346 for species ipa (1-based).
 * iphase / phaseinterpwt point at the 8*MAXNMICRO entries of this point and species. */
int oracle_trilin_interp_prop(const oracle_prop *pg, int ipa, float x, float y, float z, int interp_new,
                              double extmin, double scatmin,
                              float *temp, float *extinct, float *albedo, int *iphase, float *phaseinterpwt,
                              float *kg, char *errmsg)
{
    const int npx = pg->npx, npy = pg->npy, npz = pg->npz, mnm = pg->maxnmicro, maxpg = npx * npy * npz;
    const float *zl = pg->zlevels;
    const float *extp = pg->extinctp + (size_t)maxpg * (ipa - 1);
    const float *albp = pg->albedop + (size_t)maxpg * (ipa - 1);
    int *iphp = pg->iphasep + (size_t)mnm * maxpg * (ipa - 1);
    float *pwp = pg->phasewtp + (size_t)mnm * maxpg * (ipa - 1);
    int il, iu, im, ix, ixp, iy, iyp, iz, i, q, q2, c;
    int ic[8];
    double u, v, w, f[8], scat[8], scatter, maxscat;
    il = 0; iu = npz;
    while (iu - il > 1) {
        im = (iu + il) / 2;
        if (z >= zl[im - 1]) il = im; else iu = im;
    }
    iz = il > 1 ? il : 1;
    w = (double)(z - zl[iz - 1]) / (zl[iz] - zl[iz - 1]);
    w = fmax(fmin(w, 1.0), 0.0);
    ix = (int)((x - pg->xstart) / pg->delx) + 1;
    if (fabsf(x - pg->xstart - npx * pg->delx) < 0.01f * pg->delx) ix = npx;
    if (ix < 1 || ix > npx) {
        if (errmsg) snprintf(errmsg, 600, "TRILIN: Beyond X domain %d %d %g %g", ix, npx, x, pg->xstart);
        return 1;
    }
    ixp = ix % npx + 1;
    u = (double)(x - pg->xstart - pg->delx * (ix - 1)) / pg->delx;
    u = fmax(fmin(u, 1.0), 0.0);
    if (u < 1.0e-5) u = 0.0;
    if (u > 1.0 - 1.0e-5) u = 1.0;
    iy = (int)((y - pg->ystart) / pg->dely) + 1;
    if (fabsf(y - pg->ystart - npy * pg->dely) < 0.01f * pg->dely) iy = npy;
    if (iy < 1 || iy > npy) {
        if (errmsg) snprintf(errmsg, 600, "TRILIN: Beyond Y domain %d %d %g %g", iy, npy, y, pg->ystart);
        return 1;
    }
    iyp = iy % npy + 1;
    v = (double)(y - pg->ystart - pg->dely * (iy - 1)) / pg->dely;
    v = fmax(fmin(v, 1.0), 0.0);
    if (v < 1.0e-5) v = 0.0;
    if (v > 1.0 - 1.0e-5) v = 1.0;
    f[0] = (1 - u) * (1 - v) * (1 - w);
    f[1] = u * (1 - v) * (1 - w);
    f[2] = (1 - u) * v * (1 - w);
    f[3] = u * v * (1 - w);
    f[4] = (1 - u) * (1 - v) * w;
    f[5] = u * (1 - v) * w;
    f[6] = (1 - u) * v * w;
    f[7] = u * v * w;
    ic[0] = iz + npz * (iy - 1) + npz * npy * (ix - 1);
    ic[1] = iz + npz * (iy - 1) + npz * npy * (ixp - 1);
    ic[2] = iz + npz * (iyp - 1) + npz * npy * (ix - 1);
    ic[3] = iz + npz * (iyp - 1) + npz * npy * (ixp - 1);
    ic[4] = ic[0] + 1; ic[5] = ic[1] + 1; ic[6] = ic[2] + 1; ic[7] = ic[3] + 1;
    if (pg->tempp) {
        const float *t = pg->tempp;
        *temp = (float)(f[0] * t[ic[0] - 1] + f[1] * t[ic[1] - 1] + f[2] * t[ic[2] - 1] + f[3] * t[ic[3] - 1]
                        + f[4] * t[ic[4] - 1] + f[5] * t[ic[5] - 1] + f[6] * t[ic[6] - 1] + f[7] * t[ic[7] - 1]);
    } else *temp = 0.0f;
    *extinct = (float)(f[0] * extp[ic[0] - 1] + f[1] * extp[ic[1] - 1] + f[2] * extp[ic[2] - 1]
                       + f[3] * extp[ic[3] - 1] + f[4] * extp[ic[4] - 1] + f[5] * extp[ic[5] - 1]
                       + f[6] * extp[ic[6] - 1] + f[7] * extp[ic[7] - 1]);
    for (c = 0; c < 8; c++) scat[c] = f[c] * extp[ic[c] - 1] * albp[ic[c] - 1];
    scatter = scat[0] + scat[1] + scat[2] + scat[3] + scat[4] + scat[5] + scat[6] + scat[7];
    if (*extinct > extmin) *albedo = (float)(scatter / *extinct);
    else *albedo = (float)(scatter / extmin);
    for (c = 0; c < 8; c++)
        for (q = 0; q < mnm; q++) iphase[c * mnm + q] = iphp[q + (size_t)mnm * (ic[c] - 1)];
    if (interp_new) {
        const double den = scatter >= scatmin ? scatter : scatmin;
        for (c = 0; c < 8; c++)
            for (q = 0; q < mnm; q++)
                phaseinterpwt[c * mnm + q] = (float)(pwp[q + (size_t)mnm * (ic[c] - 1)] * scat[c] / den);
        for (q = 1; q <= 8 * mnm; q++) {
            const int cur = iphase[q - 1];
            for (q2 = q + 1; q2 <= 8 * mnm; q2++)
                if (cur == iphase[q2 - 1]) {
                    phaseinterpwt[q - 1] = phaseinterpwt[q - 1] + phaseinterpwt[q2 - 1];
                    phaseinterpwt[q2 - 1] = 0.0f;
                }
        }
        oracle_ssort(phaseinterpwt, iphase, 8 * mnm, -2);
    } else {
        maxscat = -1.0f;
        for (q = 0; q < 8 * mnm; q++) phaseinterpwt[q] = 0.0f;
        phaseinterpwt[0] = 1.0f;
        for (c = 0; c < 8; c++)
            if (scat[c] > maxscat || fabs(f[c] - 1) < 0.001f) {
                oracle_ssort(pwp + (size_t)mnm * (ic[c] - 1), iphp + (size_t)mnm * (ic[c] - 1), mnm, -2);
                maxscat = scat[c];
                iphase[0] = iphp[(size_t)mnm * (ic[c] - 1)];
            }
    }
    if (pg->nzckd > 0) {
        double ff;
        il = 1; iu = pg->nzckd;
        while (iu - il > 1) {
            im = (iu + il) / 2;
            if (z <= pg->zckd[im - 1]) il = im; else iu = im;
        }
        i = il > 1 ? il : 1;
        if (i > pg->nzckd - 1) i = pg->nzckd - 1;
        ff = (z - pg->zckd[i - 1]) / (pg->zckd[i] - pg->zckd[i - 1]);
        ff = fmin(fmax(ff, 0.0), 1.0);
        *kg = (float)((1.0f - ff) * pg->gasabs[i - 1] + ff * pg->gasabs[i]);
    } else *kg = 0.0f;
    return 0;
}

/* F of the delta-M scaling of one point/species (PREPARE_PROP shdomsub2.f:554-565, INTERPOLATE_POINT
 * shdomsub1.f:5063-5077): LEGEN(1,ML+1,.) of the dominant table or the weighted mixture. */
float oracle_deltam_f(const float *legen, int nstleg, int nleg, int ml, const int *iphase, const float *pwt, int nq,
                      int interp_new, float phasemax)
{
#define LEG1(l, iph) legen[nstleg * ((l) + (size_t)(nleg + 1) * ((iph) - 1))]
    float f;
    int q;
    if (!interp_new) return LEG1(ml + 1, iphase[0]);
    if (pwt[0] >= phasemax) return LEG1(ml + 1, iphase[0]);
    f = 0.0f;
    for (q = 0; q < nq; q++) f = f + LEG1(ml + 1, iphase[q]) * pwt[q];
    return f;
#undef LEG1
}

/* TRANSFER_PA_TO_GRID = INTERP_GRID + PREPARE_PROP.  Arrays have leading dimension npts:
 * temp[npts], planck/extinct/albedo[npts,npart], legen[nstleg,0:nleg,numphase],
 * iphase/phaseinterpwt[8*maxnmicro,npts,npart], total_ext[npts]. */
int oracle_transfer_pa_to_grid(const oracle_prop *pg, int npts, const float *gridpos, int ml, int nleg, int deltam,
                               int interp_new, float phasemax, int srctype, int units, const float *waveno,
                               float wavelen, float *temp, float *planck, float *extinct, float *albedo, float *legen,
                               int *iphase, float *phaseinterpwt, float *total_ext, double *extmin_o,
                               double *scatmin_o, float *albmax_o, char *errmsg)
{
    const int nstleg = pg->nstleg, npart = pg->npart, nq = 8 * pg->maxnmicro, numphase = pg->numphase;
    double extmin, scatmin;
    float albmax = 0.0f, kg, f, bb;
    int i, l, j, ip, ipa, iph, ierr;
#define LEGEN(j, l, i) legen[((j) - 1) + nstleg * ((l) + (size_t)(nleg + 1) * ((i) - 1))]
    for (i = 1; i <= numphase; i++)
        for (l = 0; l <= nleg; l++)
            for (j = 1; j <= nstleg; j++)
                LEGEN(j, l, i) = pg->legenp[(j - 1) + nstleg * (l + (size_t)(pg->nlegp + 1) * (i - 1))] / (2 * l + 1);
    oracle_prop_extmin(pg, &extmin, &scatmin);
    for (ip = 0; ip < npts; ip++) total_ext[ip] = 0.0f;
    for (ipa = 1; ipa <= npart; ipa++)
        for (ip = 1; ip <= npts; ip++) {
            const size_t o = (ip - 1) + (size_t)npts * (ipa - 1);
            ierr = oracle_trilin_interp_prop(pg, ipa, gridpos[3 * (size_t)(ip - 1)], gridpos[1 + 3 * (size_t)(ip - 1)],
                                             gridpos[2 + 3 * (size_t)(ip - 1)], interp_new, extmin, scatmin,
                                             &temp[ip - 1], &extinct[o], &albedo[o], &iphase[nq * o],
                                             &phaseinterpwt[nq * o], &kg, errmsg);
            if (ierr) return ierr;
            if (ipa == 1) total_ext[ip - 1] = total_ext[ip - 1] + kg;
            total_ext[ip - 1] = total_ext[ip - 1] + extinct[o];
        }
    /* PREPARE_PROP */
    if (deltam && numphase > 0)
        for (iph = 1; iph <= numphase; iph++) {
            f = LEGEN(1, ml + 1, iph);
            for (l = 0; l <= ml; l++) {
                if (!interp_new) {
                    LEGEN(1, l, iph) = (LEGEN(1, l, iph) - f) / (1 - f);
                    if (nstleg > 1) {
                        LEGEN(2, l, iph) = (LEGEN(2, l, iph) - f) / (1 - f);
                        LEGEN(3, l, iph) = (LEGEN(3, l, iph) - f) / (1 - f);
                        LEGEN(4, l, iph) = (LEGEN(4, l, iph) - f) / (1 - f);
                        LEGEN(5, l, iph) = LEGEN(5, l, iph) / (1 - f);
                        LEGEN(6, l, iph) = LEGEN(6, l, iph) / (1 - f);
                    }
                } else {
                    LEGEN(1, l, iph) = (LEGEN(1, l, iph) - f);
                    if (nstleg > 1) {
                        LEGEN(2, l, iph) = (LEGEN(2, l, iph) - f);
                        LEGEN(3, l, iph) = (LEGEN(3, l, iph) - f);
                        LEGEN(4, l, iph) = (LEGEN(4, l, iph) - f);
                    }
                }
            }
        }
    if (deltam) {
        for (ip = 0; ip < npts; ip++) {
            float s = 0.0f;
            for (ipa = 0; ipa < npart; ipa++) s = s + extinct[ip + (size_t)npts * ipa];
            total_ext[ip] = total_ext[ip] - s;
            if (total_ext[ip] < 0.0f) total_ext[ip] = 0.0f;
        }
        for (ipa = 1; ipa <= npart; ipa++)
            for (ip = 1; ip <= npts; ip++) {
                const size_t o = (ip - 1) + (size_t)npts * (ipa - 1);
                f = oracle_deltam_f(legen, nstleg, nleg, ml, &iphase[nq * o], &phaseinterpwt[nq * o], nq, 1, phasemax);
                extinct[o] = (1.0f - albedo[o] * f) * extinct[o];
                albedo[o] = (1.0f - f) * albedo[o] / (1.0f - albedo[o] * f);
                total_ext[ip - 1] = total_ext[ip - 1] + extinct[o];
            }
    }
    if (srctype != 'S' && planck)
        for (ipa = 1; ipa <= npart; ipa++)
            for (ip = 1; ip <= npts; ip++) {
                const size_t o = (ip - 1) + (size_t)npts * (ipa - 1);
                bb = oracle_planck_function(temp[ip - 1], units, waveno, wavelen);
                planck[o] = (1.0f - albedo[o]) * bb;
            }
    for (i = 0; i < npts * npart; i++) if (albedo[i] > albmax) albmax = albedo[i];
    if (extmin_o) *extmin_o = extmin;
    if (scatmin_o) *scatmin_o = scatmin;
    if (albmax_o) *albmax_o = albmax;
#undef LEGEN
    return 0;
}
