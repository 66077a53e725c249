/*
 * oracle_source.c -- TEST INFRASTRUCTURE (see shdom_oracle.h).
 * COMPUTE_SOURCE and CALC_SOURCE_PNT[_UNPOL]:
 *   /root/reference/src/polarized/shdomsub1.f:823-962 and :967-1611.
 * Volumetric sources (VOLSRC, "UNIMPLEMENTED" upstream, shdomsub1.f:900-906) are not restated.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shdom_oracle.h"
#include "oracle_internal.h"

typedef struct {
    int nstokes, nstleg, nleg, nlm, ml, mm;
    int *lofj;         /* [nlm] */
    float *legent;     /* [nstleg,0:nleg] */
    float *legent1;    /* [nstleg,0:nleg] */
    float *sourcet;    /* [nstokes,nlm] */
    float *sourcet1;   /* [nstokes,nlm] */
} src_work;

#define LG(tab, i, l) (tab)[((i) - 1) + w->nstleg * (l)]
#define ST(arr, k, j) (arr)[((k) - 1) + w->nstokes * ((j) - 1)]

/* CALC_SOURCE_PNT  shdomsub1.f:823-909 */
static void calc_source_pnt(const src_work *w, int srctype, float flux0, const float *ylmsun,
                            float planck, float albedo, const float *legen /*[nstleg,0:nleg]*/,
                            int nr, const float *radiance /*[nstokes,nr]*/, float *sourcet)
{
    const float c = 3.544907703f;
    const int nstokes = w->nstokes, nlm = w->nlm;
    int j, k;
#define RAD(k, j) radiance[((k) - 1) + nstokes * ((j) - 1)]
#define YS(i, j) ylmsun[((i) - 1) + w->nstleg * ((j) - 1)]
    for (j = 0; j < nstokes * nlm; j++) sourcet[j] = 0.0f;
    for (j = 1; j <= nr; j++)
        ST(sourcet, 1, j) = ST(sourcet, 1, j) + LG(legen, 1, w->lofj[j - 1]) * RAD(1, j);
    if (nstokes > 1) {
        for (j = 1; j <= nr; j++)
            ST(sourcet, 1, j) = ST(sourcet, 1, j) + LG(legen, 5, w->lofj[j - 1]) * RAD(2, j);
        for (j = 5; j <= nr; j++) {
            ST(sourcet, 2, j) = ST(sourcet, 2, j) + LG(legen, 5, w->lofj[j - 1]) * RAD(1, j)
                                                  + LG(legen, 2, w->lofj[j - 1]) * RAD(2, j);
            ST(sourcet, 3, j) = ST(sourcet, 3, j) + LG(legen, 3, w->lofj[j - 1]) * RAD(3, j);
        }
    }
    if (nstokes == 4) {
        for (j = 1; j <= nr; j++) {
            ST(sourcet, 3, j) = ST(sourcet, 3, j) + LG(legen, 6, w->lofj[j - 1]) * RAD(4, j);
            ST(sourcet, 4, j) = ST(sourcet, 4, j) - LG(legen, 6, w->lofj[j - 1]) * RAD(3, j)
                                                  + LG(legen, 4, w->lofj[j - 1]) * RAD(4, j);
        }
    }
    for (j = 0; j < nstokes * nlm; j++) sourcet[j] = albedo * sourcet[j];
    if (srctype == 'S' || srctype == 'B') {
        for (j = 1; j <= nlm; j++)
            ST(sourcet, 1, j) = ST(sourcet, 1, j)
                + flux0 * albedo * LG(legen, 1, w->lofj[j - 1]) * YS(1, j);
        if (nstokes > 1)
            for (j = 5; j <= nlm; j++)
                ST(sourcet, 2, j) = ST(sourcet, 2, j)
                    + flux0 * albedo * LG(legen, 5, w->lofj[j - 1]) * YS(1, j);
    }
    if (srctype == 'T' || srctype == 'B')
        ST(sourcet, 1, 1) = ST(sourcet, 1, 1) + c * planck;
#undef RAD
#undef YS
}

/* CALC_SOURCE_PNT_UNPOL  shdomsub1.f:913-962 */
static void calc_source_pnt_unpol(const src_work *w, int srctype, float flux0, const float *ylmsun,
                                  float planck, float albedo, const float *legen /*[0:nleg]*/,
                                  int nr, const float *radiance, float *sourcet)
{
    const float c = 3.544907703f;
    int j;
    for (j = 0; j < w->nlm; j++) sourcet[j] = 0.0f;
    if (srctype == 'S' || srctype == 'B')
        for (j = 1; j <= w->nlm; j++)
            sourcet[j - 1] = flux0 * albedo * legen[w->lofj[j - 1]] * ylmsun[j - 1];
    if (srctype == 'T' || srctype == 'B')
        sourcet[0] = sourcet[0] + c * planck;
    for (j = 1; j <= nr; j++)
        sourcet[j - 1] = sourcet[j - 1] + albedo * legen[w->lofj[j - 1]] * radiance[j - 1];
}

/* the per-point temporary source function (three textual copies in the reference:
 * shdomsub1.f:1089-1219, :1269-1398, :1433-1561) */
static void point_source_ld(const oracle_state *st, const src_work *w, int i, int newmethod,
                            float secmu0, float *sourcet, int npts /* leading dimension of the point arrays */)
{
    const int npart = st->npart, nq = 8 * st->maxnmicro;
    const int nlt = w->nstleg * (w->nleg + 1), ml = w->ml;
    const int nsl = w->nstokes * w->nlm;
    float ext = st->total_ext[i - 1];
    int ir = st->rshptr[i - 1];
    int nr = st->rshptr[i] - ir;
    const float *rad = &st->radiance[(size_t)w->nstokes * ir];
    int ipa, q, t, l, k;
    float f;
    if (newmethod) {
        double alb = 0.0, scat;
        float total_planck = 0.0f;
        for (t = 0; t < nlt; t++) w->legent[t] = 0.0f;
        for (ipa = 1; ipa <= npart; ipa++) {
            const int *iph = &st->iphase[(size_t)nq * ((i - 1) + (size_t)npts * (ipa - 1))];
            const float *pw = &st->phaseinterpwt[(size_t)nq * ((i - 1) + (size_t)npts * (ipa - 1))];
            float e = st->extinct[(i - 1) + (size_t)npts * (ipa - 1)];
            float a = st->albedo[(i - 1) + (size_t)npts * (ipa - 1)];
            scat = (double)(e * a);
            alb = alb + scat;
            if (st->planck)
                total_planck = total_planck + e * st->planck[(i - 1) + (size_t)npts * (ipa - 1)];
            if (!st->interp_new) {
                const float *lg = &st->legen[(size_t)nlt * (iph[0] - 1)];
                for (t = 0; t < nlt; t++) w->legent[t] = (float)(w->legent[t] + scat * lg[t]);
            } else {
                if (pw[0] >= st->phasemax) {
                    const float *lg = &st->legen[(size_t)nlt * (iph[0] - 1)];
                    for (t = 0; t < nlt; t++) w->legent1[t] = lg[t];
                } else {
                    for (t = 0; t < nlt; t++) w->legent1[t] = 0.0f;
                    for (q = 0; q < nq; q++) {
                        const float *lg;
                        if (pw[q] <= 1e-5f) continue;
                        lg = &st->legen[(size_t)nlt * (iph[q] - 1)];
                        for (t = 0; t < nlt; t++) w->legent1[t] = w->legent1[t] + lg[t] * pw[q];
                    }
                }
                if (st->deltam) {
                    f = LG(w->legent1, 1, ml + 1);
                    for (l = 0; l <= ml; l++)
                        for (k = 1; k <= w->nstleg; k++)
                            LG(w->legent1, k, l) = LG(w->legent1, k, l) / (1 - f);
                }
                for (t = 0; t < nlt; t++) w->legent[t] = (float)(w->legent[t] + scat * w->legent1[t]);
            }
        }
        if (alb > 1e-10f) { for (t = 0; t < nlt; t++) w->legent[t] = (float)(w->legent[t] / alb); }
        else { for (t = 0; t < nlt; t++) w->legent[t] = w->legent[t] / npart; }
        if (ext > 1e-10f) {
            alb = alb / ext;
            total_planck = total_planck / ext;
        } else {
            alb = 0.0;
            total_planck = 0.0f;
        }
        LG(w->legent, 1, 0) = 1.0f;
        if (w->nstokes == 1)
            calc_source_pnt_unpol(w, st->srctype, st->dirflux[i - 1] * secmu0, st->ylmsun,
                                  total_planck, (float)alb, w->legent, nr, rad, sourcet);
        else
            calc_source_pnt(w, st->srctype, st->dirflux[i - 1] * secmu0, st->ylmsun,
                            total_planck, (float)alb, w->legent, nr, rad, sourcet);
    } else {
        for (t = 0; t < nsl; t++) sourcet[t] = 0.0f;
        for (ipa = 1; ipa <= npart; ipa++) {
            const int *iph = &st->iphase[(size_t)nq * ((i - 1) + (size_t)npts * (ipa - 1))];
            const float *pw = &st->phaseinterpwt[(size_t)nq * ((i - 1) + (size_t)npts * (ipa - 1))];
            float wgt, pl = st->planck ? st->planck[(i - 1) + (size_t)npts * (ipa - 1)] : 0.0f;
            float a = st->albedo[(i - 1) + (size_t)npts * (ipa - 1)];
            const float *lgu;
            if (ext == 0.0f) wgt = 1.0f;
            else wgt = st->extinct[(i - 1) + (size_t)npts * (ipa - 1)] / ext;
            if (!st->interp_new) {
                lgu = &st->legen[(size_t)nlt * (iph[0] - 1)];
            } else {
                if (pw[0] >= st->phasemax) {
                    const float *lg = &st->legen[(size_t)nlt * (iph[0] - 1)];
                    for (t = 0; t < nlt; t++) w->legent[t] = lg[t];
                } else {
                    for (t = 0; t < nlt; t++) w->legent[t] = 0.0f;
                    for (q = 0; q < nq; q++) {
                        const float *lg;
                        if (pw[q] <= 1e-5f) continue;
                        lg = &st->legen[(size_t)nlt * (iph[q] - 1)];
                        for (t = 0; t < nlt; t++) w->legent[t] = w->legent[t] + lg[t] * pw[q];
                    }
                }
                if (st->deltam) {
                    f = LG(w->legent, 1, ml + 1);
                    for (l = 0; l <= ml; l++)
                        for (k = 1; k <= w->nstleg; k++)
                            LG(w->legent, k, l) = LG(w->legent, k, l) / (1 - f);
                }
                lgu = w->legent;
            }
            if (w->nstokes == 1)
                calc_source_pnt_unpol(w, st->srctype, st->dirflux[i - 1] * secmu0, st->ylmsun,
                                      pl, a, lgu, nr, rad, w->sourcet1);
            else
                calc_source_pnt(w, st->srctype, st->dirflux[i - 1] * secmu0, st->ylmsun,
                                pl, a, lgu, nr, rad, w->sourcet1);
            for (t = 0; t < nsl; t++) sourcet[t] = sourcet[t] + wgt * w->sourcet1[t];
        }
    }
}

static void point_source(const oracle_state *st, const src_work *w, int i, int newmethod,
                         float secmu0, float *sourcet)
{
    point_source_ld(st, w, i, newmethod, secmu0, sourcet, st->npts);
}

static void src_work_init(src_work *w, const oracle_state *st)
{
    int j = 0, l, m;
    w->nstokes = st->nstokes; w->nstleg = st->nstleg; w->nleg = st->nleg; w->nlm = st->nlm;
    w->ml = st->ml; w->mm = st->mm;
    w->lofj = (int *)malloc(sizeof(int) * st->nlm);
    w->legent = (float *)malloc(sizeof(float) * st->nstleg * (st->nleg + 2));
    w->legent1 = (float *)malloc(sizeof(float) * st->nstleg * (st->nleg + 2));
    w->sourcet = (float *)malloc(sizeof(float) * st->nstokes * st->nlm);
    w->sourcet1 = (float *)malloc(sizeof(float) * st->nstokes * st->nlm);
    for (l = 0; l <= st->ml; l++) {
        int me = l < st->mm ? l : st->mm;
        for (m = -me; m <= me; m++) { w->lofj[j] = l; j++; }
    }
}

/* The source function of one point with the "new method" species mixing, as INTERPOLATE_POINT evaluates it for a
 * new grid point (shdomsub1.f:5133-5211; the same arithmetic as COMPUTE_SOURCE's per-point block).  The point arrays
 * of st have leading dimension ld (MAXIG inside SOLUTION_ITERATIONS).  sourcet[nstokes,nlm]. */
void oracle_point_source(const oracle_state *st, int ld, int i, float *sourcet)
{
    src_work wk;
    float secmu0 = 1.0f / fabsf(st->solarmu);
    src_work_init(&wk, st);
    point_source_ld(st, &wk, i, 1, secmu0, sourcet, ld);
    free(wk.lofj); free(wk.legent); free(wk.legent1); free(wk.sourcet); free(wk.sourcet1);
}

/* The four norms of the last oracle_compute_source call accumulated in double (same per-term REAL products, f64
 * running sums): the rounding-free value of what the reference sums sequentially in REAL. */
static double last_sums64[4];
void oracle_compute_source_sums64(double *out) { int k; for (k = 0; k < 4; k++) out[k] = last_sums64[k]; }

/* COMPUTE_SOURCE  shdomsub1.f:967-1611.  st->shptr/st->source are ignored; the in/out arrays
 * are the explicit arguments (SHPTR, SOURCE, OSHPTR, DELSOURCE are intent(in,out)). */
int oracle_compute_source(const oracle_state *st, int fixsh, float shacc, int maxiv,
                          int first, int accelflag, int newmethod,
                          int *shptr, float *source, int *oshptr, float *delsource,
                          float *deljdot_o, float *deljold_o, float *deljnew_o, float *jnorm_o,
                          char *errmsg)
{
    src_work wk, *w = &wk;
    const int nstokes = st->nstokes, nlm = st->nlm, npts = st->npts, ml = st->ml, mm = st->mm;
    float srcmin = shacc;
    float secmu0 = 1.0f / fabsf(st->solarmu);
    float deljdot = 0.0f, deljold = 0.0f, deljnew = 0.0f, jnorm = 0.0f;
    double s64[4] = {0.0, 0.0, 0.0, 0.0};
    int i, j, k, l, m, is, iso, ns = 0, ierr = 0;
#define SRC(k, j) source[((k) - 1) + (size_t)nstokes * ((j) - 1)]
#define DSRC(k, j) delsource[((k) - 1) + (size_t)nstokes * ((j) - 1)]
    w->nstokes = nstokes; w->nstleg = st->nstleg; w->nleg = st->nleg; w->nlm = nlm;
    w->ml = ml; w->mm = mm;
    w->lofj = (int *)malloc(sizeof(int) * nlm);
    w->legent = (float *)malloc(sizeof(float) * st->nstleg * (st->nleg + 2));
    w->legent1 = (float *)malloc(sizeof(float) * st->nstleg * (st->nleg + 2));
    w->sourcet = (float *)malloc(sizeof(float) * nstokes * nlm);
    w->sourcet1 = (float *)malloc(sizeof(float) * nstokes * nlm);
    j = 0;
    for (l = 0; l <= ml; l++) {
        int me = l < mm ? l : mm;
        for (m = -me; m <= me; m++) { w->lofj[j] = l; j++; }
    }
    if (!first) {
        for (i = 1; i <= npts; i++) {
            point_source(st, w, i, newmethod, secmu0, w->sourcet);
            if (accelflag) {
                int nso;
                is = shptr[i - 1];
                iso = oshptr[i - 1];
                ns = shptr[i] - is;
                nso = oshptr[i] - iso;
                if (nso < ns) ns = nso;
                for (k = 1; k <= nstokes; k++)
                    for (j = 1; j <= ns; j++) {
                        float d = ST(w->sourcet, k, j) - SRC(k, is + j);
                        deljdot = deljdot + d * DSRC(k, iso + j);
                        deljold = deljold + DSRC(k, iso + j) * DSRC(k, iso + j);
                        deljnew = deljnew + d * d;
                        jnorm = jnorm + SRC(k, is + j) * SRC(k, is + j);
                        s64[0] += (double)(d * DSRC(k, iso + j));
                        s64[1] += (double)(DSRC(k, iso + j) * DSRC(k, iso + j));
                        s64[2] += (double)(d * d);
                        s64[3] += (double)(SRC(k, is + j) * SRC(k, is + j));
                    }
            } else {
                is = shptr[i - 1];
                ns = shptr[i] - is;
                for (k = 1; k <= nstokes; k++)
                    for (j = 1; j <= ns; j++) {
                        float d = ST(w->sourcet, k, j) - SRC(k, is + j);
                        deljnew = deljnew + d * d;
                        jnorm = jnorm + SRC(k, is + j) * SRC(k, is + j);
                        s64[2] += (double)(d * d);
                        s64[3] += (double)(SRC(k, is + j) * SRC(k, is + j));
                    }
            }
        }
    }
    for (k = 0; k < 4; k++) last_sums64[k] = s64[k];
    if (!first && accelflag) {
        for (i = 1; i <= npts; i++) {
            point_source(st, w, i, newmethod, secmu0, w->sourcet);
            is = shptr[i - 1];
            ns = shptr[i] - is;
            oshptr[i - 1] = is;
            for (j = 1; j <= ns; j++)
                for (k = 1; k <= nstokes; k++)
                    DSRC(k, is + j) = ST(w->sourcet, k, j) - SRC(k, is + j);
        }
        oshptr[npts] = shptr[npts];
    }
    is = 0;
    for (i = 1; i <= npts; i++) {
        int nr = st->rshptr[i] - st->rshptr[i - 1];
        if (nr > nlm) {
            if (errmsg) snprintf(errmsg, 600, "COMPUTE_SOURCE: NR>NLM 3 %d", i);
            ierr = 1;
            break;
        }
        point_source(st, w, i, newmethod, secmu0, w->sourcet);
        if (fixsh) {
            ns = shptr[i] - is;
        } else {
            int js = (st->srctype == 'S') ? 0 : 1;
            for (j = 1; j <= nlm; j++)
                for (k = 1; k <= nstokes; k++)
                    if (fabsf(ST(w->sourcet, k, j)) > srcmin) js = j;
            if (js == 0) {
                ns = 0;
            } else {
                int ls = w->lofj[js - 1];
                if (ls <= mm) ns = (ls * (ls + 1)) + ls + 1;
                else ns = (2 * mm + 1) * ls - (mm * (1 + (mm - 1))) + mm + 1;
            }
            shptr[i - 1] = is;
        }
        if (is + ns > maxiv) {
            if (errmsg) snprintf(errmsg, 600, "COMPUTE_SOURCE: MAXIV exceeded %d Out of memory for "
                                 "more spherical harmonic terms.", maxiv);
            ierr = 2;
            break;
        }
        for (j = 1; j <= ns; j++)
            for (k = 1; k <= nstokes; k++) SRC(k, is + j) = ST(w->sourcet, k, j);
        is = is + ns;
    }
    if (!ierr) shptr[npts] = is;
    *deljdot_o = deljdot; *deljold_o = deljold; *deljnew_o = deljnew; *jnorm_o = jnorm;
    free(w->lofj); free(w->legent); free(w->legent1); free(w->sourcet); free(w->sourcet1);
#undef SRC
#undef DSRC
    return ierr;
}
