/*
 * oracle_special.c -- TEST INFRASTRUCTURE (see shdom_oracle.h).
 * Real generalized spherical harmonics and Wigner d-functions, phase-function tables.
 * Follows /root/reference/src/polarized/shdomsub2.f:4244-4645 (YLMALL, YLMALL_UNPOL,
 * WIGNERFCT02P2M_NORMALIZED, DMM1_N0, WIGNERFCT_DM0, DM_M10_N0, WIGNERFCT) and
 * /root/reference/src/polarized/shdomsub4.f:2388-2585 (PRECOMPUTE_PHASE_CHECK[_GRAD]).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shdom_oracle.h"

static double ipow(double x, int n)
{   /* gfortran x**n for integer n>=0: repeated multiplication (binary powering) */
    double r = 1.0;
    double b = x;
    int e = n;
    while (e > 0) {
        if (e & 1) r *= b;
        e >>= 1;
        if (e) b *= b;
    }
    return r;
}

/* shdomsub2.f:4452-4486 */
static double dmm1_n0(double x, int m, int m1)
{
    double fact, cmm1, prod, r;
    int p, maxm, minm;
    if (m == m1) {
        fact = ipow((1.0 + x) / 2.0, m);
        return fact;
    }
    if (m1 > m) cmm1 = 1.0;
    else cmm1 = ((m - m1) & 1) ? -1.0 : 1.0;
    maxm = m > m1 ? m : m1;
    minm = m < m1 ? m : m1;
    prod = 1.0;
    for (p = 1; p <= maxm - minm; p++) {
        fact = sqrt((double)(m + m1 + p) / (double)p);
        prod = prod * fact;
    }
    fact = sqrt((1.0 - x) / 2.0);
    r = cmm1 * prod * ipow(fact, maxm - minm);
    fact = sqrt((1.0 + x) / 2.0);
    r = r * ipow(fact, maxm + minm);
    return r;
}

/* shdomsub2.f:4363-4448 */
static void wignerfct02p2m_normalized(double miu, int nrank, int m,
                                      double *dm0, double *dm2p, double *dm2m)
{
    double xp = miu, xm = -miu, fact1, fact2, factp, factm;
    int n, n0;
    for (n = 0; n <= nrank; n++) dm0[n] = 0.0;
    if (m <= nrank) {
        if (m == 0) {
            dm0[0] = 1.0;
            dm0[1] = xp;
            for (n = 1; n <= nrank - 1; n++) {
                fact1 = (double)(2 * n + 1) * xp / (double)(n + 1);
                fact2 = (double)n / (double)(n + 1);
                dm0[n + 1] = fact1 * dm0[n] - fact2 * dm0[n - 1];
            }
        } else {
            dm0[m] = dmm1_n0(xp, m, 0);
            for (n = m; n <= nrank - 1; n++) {
                fact1 = (double)(n * (n + 1)) * xp;
                fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - m * m));
                fact1 = fact1 / (double)(n + 1);
                fact1 = fact1 * (double)(2 * n + 1) / (double)n;
                fact2 = sqrt((double)(n * n - m * m)) * (double)n;
                fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m * m));
                fact2 = fact2 / (double)(n + 1);
                fact2 = fact2 * (double)(n + 1) / (double)n;
                dm0[n + 1] = fact1 * dm0[n] - fact2 * (n - 1 >= 0 ? dm0[n - 1] : 0.0);
            }
        }
        for (n = 0; n <= nrank; n++) dm0[n] = sqrt(n + 0.5) * dm0[n];
    }
    n0 = m > 2 ? m : 2;
    for (n = 0; n <= nrank; n++) { dm2p[n] = 0.0; dm2m[n] = 0.0; }
    if (m <= nrank && nrank >= 2) {
        dm2p[n0] = dmm1_n0(xp, m, 2);
        dm2m[n0] = dmm1_n0(xm, m, 2);
        for (n = n0; n <= nrank - 1; n++) {
            factp = (double)(n * (n + 1)) * xp - (double)(2 * m);
            factm = (double)(n * (n + 1)) * xm - (double)(2 * m);
            fact1 = 1.0 / sqrt((double)((n + 1) * (n + 1) - m * m));
            fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - 4));
            fact1 = fact1 * (double)(2 * n + 1) / (double)n;
            fact2 = sqrt((double)(n * n - m * m)) * sqrt((double)(n * n - 4));
            fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m * m));
            fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - 4));
            fact2 = fact2 * (double)(n + 1) / (double)n;
            dm2p[n + 1] = factp * fact1 * dm2p[n] - fact2 * dm2p[n - 1];
            dm2m[n + 1] = factm * fact1 * dm2m[n] - fact2 * dm2m[n - 1];
        }
        for (n = 0; n <= nrank; n++) {
            dm2m[n] = (((n + m) & 1) ? -1.0 : 1.0) * dm2m[n];
            dm2p[n] = sqrt(n + 0.5) * dm2p[n];
            dm2m[n] = sqrt(n + 0.5) * dm2m[n];
        }
    }
}

/* shdomsub2.f:4580-4600 */
static double dm_m10_n0(double x, int m)
{
    double cm, prod;
    int p;
    if (m == 0) return 1.0;
    cm = (m & 1) ? -1.0 : 1.0;
    prod = 1.0;
    for (p = 1; p <= m; p++) prod = prod * sqrt((m + p) / (double)p);
    return cm * prod * ipow(0.5 * sqrt((1.0 - x) * (1.0 + x)), m);
}

/* shdomsub2.f:4542-4577 */
static void wignerfct_dm0(double x, int nrank, int m, double *dm0)
{
    int n;
    for (n = 0; n <= nrank; n++) dm0[n] = 0.0;
    if (m <= nrank) {
        if (m == 0) {
            dm0[0] = 1.0;
            dm0[1] = x;
            for (n = 1; n <= nrank - 1; n++)
                dm0[n + 1] = ((2 * n + 1) * x * dm0[n] - n * dm0[n - 1]) / (n + 1);
        } else {
            dm0[m] = dm_m10_n0(x, m);
            for (n = m; n <= nrank - 1; n++)
                dm0[n + 1] = ((2 * n + 1) * x * dm0[n]
                              - sqrt((double)(n * n - m * m)) * dm0[n - 1])
                             / sqrt((double)((n + 1) * (n + 1) - m * m));
        }
        for (n = 0; n <= nrank; n++) dm0[n] = sqrt(n + 0.5) * dm0[n];
    }
}

/* shdomsub2.f:4490-4539 */
static void ylmall_unpol(float mu, float phi, int ml, int mm, float *yr)
{
    double *dm0 = (double *)malloc(sizeof(double) * (ml + 2));
    double x = (double)mu;
    double pi = acos(-1.0);
    double fct = 1.0 / sqrt(2.0 * pi);
    double cosm, sinm;
    int m, l, j;
    for (m = 0; m <= mm; m++) {
        wignerfct_dm0(x, ml, m, dm0);
        for (l = 0; l <= ml; l++) dm0[l] = fct * dm0[l];
        if (m > 0) {
            /* COS(M*PHI): INTEGER*REAL -> REAL, REAL intrinsic, then widened */
            cosm = (double)cosf((float)m * phi);
            sinm = (double)sinf((float)m * phi);
        } else {
            cosm = 1.0;
            sinm = 0.0;
        }
        for (l = m; l <= ml; l++) {
            if (l <= mm) j = l * (l + 1) + m + 1;
            else j = (2 * mm + 1) * l - mm * mm + m + 1;
            yr[j - 1] = (float)((cosm - sinm) * dm0[l]);
            j = j - 2 * m;
            yr[j - 1] = (float)((cosm + sinm) * dm0[l]);
        }
    }
    free(dm0);
}

/* shdomsub2.f:4244-4360.  yr is [nstleg, nlm] Fortran order */
void oracle_ylmall(int transpose, float mu, float phi, int ml, int mm, int nstleg, float *yr)
{
    double *dm0, *dm2p, *dm2m;
    double x, pi, fct, p1, p2, p3, cosm, sinm, sign;
    int j, m, mabs, l;
#define YR(i, jj) yr[((i) - 1) + (size_t)nstleg * ((jj) - 1)]
    if (nstleg == 1) {
        ylmall_unpol(mu, phi, ml, mm, yr);
        return;
    }
    dm0 = (double *)malloc(sizeof(double) * (ml + 3));
    dm2p = (double *)malloc(sizeof(double) * (ml + 3));
    dm2m = (double *)malloc(sizeof(double) * (ml + 3));
    x = (double)mu;
    pi = acos(-1.0);
    fct = 1.0 / sqrt(2.0 * pi);
    sign = transpose ? -1.0 : 1.0;

    m = 0;
    wignerfct02p2m_normalized(x, ml, m, dm0, dm2p, dm2m);
    for (l = 0; l <= ml; l++) {
        if (l <= mm) j = l * (l + 1) + m + 1;
        else j = (2 * mm + 1) * l - mm * mm + m + 1;
        p1 = fct * dm0[l];
        p2 = -0.5 * fct * (dm2p[l] + dm2m[l]);
        p3 = -0.5 * fct * (dm2p[l] - dm2m[l]);
        YR(1, j) = (float)p1;
        if (nstleg == 6) {
            YR(2, j) = (float)p2;
            YR(3, j) = (float)p2;
            YR(4, j) = (float)p1;
            YR(5, j) = (float)p3;
            YR(6, j) = (float)p3;
        }
    }
    for (mabs = 1; mabs <= mm; mabs++) {
        wignerfct02p2m_normalized(x, ml, mabs, dm0, dm2p, dm2m);
        cosm = (double)cosf((float)mabs * phi);
        sinm = (double)sinf((float)mabs * phi);
        for (l = mabs; l <= ml; l++) {
            m = mabs;
            if (l <= mm) j = l * (l + 1) + m + 1;
            else j = (2 * mm + 1) * l - mm * mm + m + 1;
            p1 = fct * dm0[l];
            p2 = -0.5 * fct * (dm2p[l] + dm2m[l]);
            p3 = -0.5 * fct * (dm2p[l] - dm2m[l]);
            YR(1, j) = (float)(p1 * cosm - p1 * sinm);
            if (nstleg == 6) {
                YR(2, j) = (float)(p2 * cosm - p2 * sinm);
                YR(3, j) = (float)(p2 * cosm + p2 * sinm);
                YR(4, j) = (float)(p1 * cosm + p1 * sinm);
                YR(5, j) = (float)(p3 * cosm - sign * p3 * sinm);
                YR(6, j) = (float)(p3 * cosm + sign * p3 * sinm);
            }
            m = -mabs;
            if (l <= mm) j = l * (l + 1) + m + 1;
            else j = (2 * mm + 1) * l - mm * mm + m + 1;
            YR(1, j) = (float)(p1 * sinm + p1 * cosm);
            if (nstleg == 6) {
                YR(2, j) = (float)(p2 * sinm + p2 * cosm);
                YR(3, j) = (float)(p2 * sinm - p2 * cosm);
                YR(4, j) = (float)(p1 * sinm - p1 * cosm);
                YR(5, j) = (float)(p3 * sinm + sign * p3 * cosm);
                YR(6, j) = (float)(p3 * sinm - sign * p3 * cosm);
            }
        }
    }
#undef YR
    free(dm0); free(dm2p); free(dm2m);
}

/* shdomsub2.f:4604-4645 */
static void wignerfct(double x, int nrank, int m, int m1, double *dmm1)
{
    int n0, n;
    double fact1, fact2;
    n0 = m > m1 ? m : m1;
    for (n = 0; n <= nrank; n++) dmm1[n] = 0.0;
    if (n0 == 0) {
        dmm1[0] = 1.0;
        if (nrank >= 1) dmm1[1] = x;
        for (n = 1; n <= nrank - 1; n++) {
            fact1 = (double)(2 * n + 1) * x / (double)(n + 1);
            fact2 = (double)n / (double)(n + 1);
            dmm1[n + 1] = fact1 * dmm1[n] - fact2 * dmm1[n - 1];
        }
    } else {
        if (n0 <= nrank) dmm1[n0] = dmm1_n0(x, m, m1);
        for (n = n0; n <= nrank - 1; n++) {
            fact1 = (double)(n * (n + 1)) * x - (double)(m * m1);
            fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - m * m));
            fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - m1 * m1));
            fact1 = fact1 * (double)(2 * n + 1) / (double)n;
            fact2 = sqrt((double)(n * n - m * m)) * sqrt((double)(n * n - m1 * m1));
            fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m * m));
            fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m1 * m1));
            fact2 = fact2 * (double)(n + 1) / (double)n;
            dmm1[n + 1] = fact1 * dmm1[n] - fact2 * dmm1[n - 1];
        }
    }
}

/* shdomsub4.f:2388-2490 (scale_by_2l1=1) and :2493-2585 (scale_by_2l1=0) */
static int phase_check_impl(int nscatangle, int numphase, int nstphase, int nstokes,
                            int nstleg, int nleg, const float *legen, float *phasetab,
                            int negcheck, int scale_by_2l1, const char *who, char *errmsg)
{
    double *unsc1 = (double *)malloc(sizeof(double) * (nleg + 2));
    double *unsc2 = (double *)malloc(sizeof(double) * (nleg + 2));
    double *dmm1 = (double *)malloc(sizeof(double) * (nleg + 2));
    double *dmm2 = (double *)malloc(sizeof(double) * (nleg + 2));
    double pi = acos(-1.0);
    /* OFOURPI = 1.0/(4.0*PI): REAL 4.0 * DOUBLE PI */
    double ofourpi = 1.0 / (4.0 * pi);
    int iph, j, l, ierr = 0;
#define LEGEN(i, ll, ip) legen[((i) - 1) + (size_t)nstleg * ((ll) + (size_t)(nleg + 1) * ((ip) - 1))]
#define PHASETAB(i, ip, jj) phasetab[((i) - 1) + (size_t)nstphase * (((ip) - 1) + (size_t)numphase * ((jj) - 1))]
    for (j = 1; j <= nscatangle && !ierr; j++) {
        double cosscat = cos(pi * (double)(j - 1) / (nscatangle - 1));
        double x = cosscat;
        wignerfct(x, nleg, 0, 0, dmm1);
        if (nstokes > 1) wignerfct(x, nleg, 2, 0, dmm2);
        for (iph = 1; iph <= numphase; iph++) {
            double a1, b1, fct;
            for (l = 0; l <= nleg; l++) {
                /* UNSCLEGEN is DOUBLE; RHS LEGEN/(2*L+1) evaluated in REAL */
                if (scale_by_2l1) unsc1[l] = (double)(LEGEN(1, l, iph) / (float)(2 * l + 1));
                else unsc1[l] = (double)LEGEN(1, l, iph);
                if (nstleg > 1) {
                    if (scale_by_2l1) unsc2[l] = (double)(LEGEN(5, l, iph) / (float)(2 * l + 1));
                    else unsc2[l] = (double)LEGEN(5, l, iph);
                }
            }
            a1 = 0.0;
            for (l = 0; l <= nleg; l++) {
                fct = 2.0 * l + 1.0;
                a1 = a1 + fct * unsc1[l] * dmm1[l];
            }
            if (negcheck && a1 <= 0.0) {
                ierr = 1;
                if (errmsg)
                    snprintf(errmsg, 600, "%s: negative phase function for tabulated phase "
                             "function: IPH %d J %d A1 %g", who, iph, j, a1);
                break;
            }
            PHASETAB(1, iph, j) = (float)(a1 * ofourpi);
            if (nstokes > 1) {
                b1 = 0.0;
                for (l = 0; l <= nleg; l++) {
                    fct = 2.0 * l + 1.0;
                    b1 = b1 - fct * unsc2[l] * dmm2[l];
                }
                PHASETAB(2, iph, j) = (float)(b1 * ofourpi);
            }
        }
    }
#undef LEGEN
#undef PHASETAB
    free(unsc1); free(unsc2); free(dmm1); free(dmm2);
    return ierr;
}

int oracle_precompute_phase_check(int nscatangle, int numphase, int nstphase, int nstokes,
                                  int ml, int nstleg, int nleg, const float *legen,
                                  float *phasetab, int deltam, int negcheck, char *errmsg)
{
    (void)ml; (void)deltam; /* all three DELTAM branches are identical: shdomsub4.f:2438-2445 */
    return phase_check_impl(nscatangle, numphase, nstphase, nstokes, nstleg, nleg, legen,
                            phasetab, negcheck, 1, "PRECOMPUTE_PHASE_CHECK", errmsg);
}

int oracle_precompute_phase_check_grad(int nscatangle, int dnumphase, int nstphase, int nstokes,
                                  int ml, int nstleg, int nleg, const float *dleg,
                                  float *dphasetab, int deltam, int negcheck, char *errmsg)
{
    (void)ml; (void)deltam;
    return phase_check_impl(nscatangle, dnumphase, nstphase, nstokes, nstleg, nleg, dleg,
                            dphasetab, negcheck, 0, "PRECOMPUTE_PHASE_CHECK_GRAD", errmsg);
}

/* PLMALL  shdomsub2.f:4650-4752: the mu-dependent part of the generalised spherical harmonics used by the
 * SH <-> discrete-ordinate transforms.  prc is PRC(6,NLM) in Fortran order. */
void oracle_plmall(int transpose, float mu, int ml, int mm, float *prc)
{
    const double x = (double)mu;
    const double pi = acos(-1.0);
    const double fct = 1.0 / sqrt(2.0 * pi);
    const double sign = transpose ? -1.0 : 1.0;
    double *dm0 = (double *)malloc(sizeof(double) * 3 * (ml + 2));
    double *dm2p = dm0 + (ml + 2), *dm2m = dm2p + (ml + 2);
    int l, m, mabs, j;
    double p1, p2, p3;
#define PRC(q, j) prc[((q) - 1) + 6 * ((j) - 1)]
    m = 0;
    wignerfct02p2m_normalized(x, ml, m, dm0, dm2p, dm2m);
    for (l = 0; l <= ml; l++) {
        if (l <= mm) j = l * (l + 1) + m + 1; else j = (2 * mm + 1) * l - mm * mm + m + 1;
        p1 = fct * dm0[l];
        p2 = -0.5 * fct * (dm2p[l] + dm2m[l]);
        p3 = -0.5 * fct * (dm2p[l] - dm2m[l]);
        PRC(1, j) = (float)p1; PRC(2, j) = (float)p2; PRC(3, j) = (float)p2;
        PRC(4, j) = (float)p1; PRC(5, j) = (float)p3; PRC(6, j) = (float)p3;
    }
    for (mabs = 1; mabs <= mm; mabs++) {
        wignerfct02p2m_normalized(x, ml, mabs, dm0, dm2p, dm2m);
        for (l = mabs; l <= ml; l++) {
            m = mabs;
            if (l <= mm) j = l * (l + 1) + m + 1; else j = (2 * mm + 1) * l - mm * mm + m + 1;
            p1 = fct * dm0[l];
            p2 = -0.5 * fct * (dm2p[l] + dm2m[l]);
            p3 = -0.5 * fct * (dm2p[l] - dm2m[l]);
            PRC(1, j) = (float)p1; PRC(2, j) = (float)p2; PRC(3, j) = (float)p2;
            PRC(4, j) = (float)p1; PRC(5, j) = (float)p3; PRC(6, j) = (float)p3;
            m = -mabs;
            if (l <= mm) j = l * (l + 1) + m + 1; else j = (2 * mm + 1) * l - mm * mm + m + 1;
            PRC(1, j) = (float)p1; PRC(2, j) = (float)p2; PRC(3, j) = (float)(-p2);
            PRC(4, j) = (float)(-p1); PRC(5, j) = (float)(sign * p3); PRC(6, j) = (float)(-sign * p3);
        }
    }
#undef PRC
    free(dm0);
}
