/*
 * oracle_solver.c -- TEST INFRASTRUCTURE ONLY (see shdom_oracle.h).
 *
 * Fixed-grid SHDOM solution iterations, restated so that the oracle can be pinned against the
 * reference's own SHDOM verification outputs (tests/data/brdf_*.out), which need a solved state.
 * Follows (paths relative to /root/reference):
 *   src/polarized/shdomsub1.f:445-822    SOLUTION_ITERATIONS (SPLITACC=0 branch: no cell splitting)
 *   src/polarized/shdomsub1.f:1615-1805  RADIANCE_TRUNCATION
 *   src/polarized/shdomsub1.f:1807-1832  ACCELERATE_SOLUTION
 *   src/polarized/shdomsub1.f:1836-2167  PATH_INTEGRATION
 *   src/polarized/shdomsub1.f:2789-3260  SH_TO_DO[_UNPOL], DO_TO_SH[_UNPOL]
 *   src/polarized/shdomsub1.f:3261-3353  SWEEPING_ORDER
 *   src/polarized/shdomsub1.f:3354-4036  BACK_INT_GRID3D[_UNPOL] (3-D grids, IPFLAG 0 or 1)
 *   src/polarized/shdomsub1.f:4295-4468  BACK_INT_GRID1D (IPFLAG=3: independent columns)
 *   src/polarized/shdomsub1.f:4529-4700  SWEEP_BASE_CELL, SWEEP_NEXT_CELL
 *   src/polarized/shdomsub2.f:1146-1220  MAKE_SH_DO_COEF
 *   src/shdom_nompi.f:317-349            CALC_ACCEL_SOLCRIT
 * Deliberate differences, both immaterial for the converged state:
 *   - the first guess is a zero radiance field with 4 SH terms per point instead of the Eddington
 *     two-stream field of INIT_RADIANCE (shdomsub2.f:614-996); the fixed point of the iteration does
 *     not depend on it;
 *   - the azimuthal FFTs (FFTPACK RFFTB/RFFTF) are evaluated as the direct real DFT sums they equal.
 *   src/polarized/shdomsub1.f:4039-4293  BACK_INT_GRID2D (IPFLAG=2: independent pixels in Y)
 * Sweeps covered: IPFLAG=3 (BACK_INT_GRID1D), IPFLAG=2 (BACK_INT_GRID2D), IPFLAG 0/1 (BACK_INT_GRID3D); the
 * multi-processor boundary flags return 3.
 * Pinning: BACK_INT_GRID1D and everything around it against SHDOM's brdf_*.out (tests/test_shdom_verification.py);
 * BACK_INT_GRID3D has no SHDOM output in the checkout that does not also need SPLIT_GRID -- it is pinned by
 * consistency with the pinned column solve (horizontally uniform slab, second-order convergence), same file.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

void oracle_plmall(int transpose, float mu, int ml, int mm, float *prc);

typedef struct {
    int nstleg, ml, mm, nlm, nmu, nphi0max;
    float *cmu1;    /* [nstleg,nlm,nmu] */
    float *cmu2;    /* [nstleg,nmu,nlm] */
    int *fftflag;   /* [nmu] */
    int *mofj;      /* [nlm] */
} shdo_coef;

#define CMU1(c, q, j, i) (c)->cmu1[((q) - 1) + (size_t)(c)->nstleg * (((j) - 1) + (size_t)(c)->nlm * ((i) - 1))]
#define CMU2(c, q, i, j) (c)->cmu2[((q) - 1) + (size_t)(c)->nstleg * (((i) - 1) + (size_t)(c)->nmu * ((j) - 1))]

/* MAKE_SH_DO_COEF  shdomsub2.f:1146-1220 (the azimuthal part is evaluated in sh_to_do / do_to_sh) */
static shdo_coef *make_sh_do_coef(const oracle_state *st, const float *wtmu)
{
    shdo_coef *c = (shdo_coef *)calloc(1, sizeof(shdo_coef));
    float *prc = (float *)malloc(sizeof(float) * 6 * st->nlm);
    int i, j, q, l, m, mmax;
    c->nstleg = st->nstleg; c->ml = st->ml; c->mm = st->mm; c->nlm = st->nlm; c->nmu = st->nmu;
    c->nphi0max = st->nphi0max;
    c->cmu1 = (float *)calloc((size_t)c->nstleg * c->nlm * c->nmu, sizeof(float));
    c->cmu2 = (float *)calloc((size_t)c->nstleg * c->nlm * c->nmu, sizeof(float));
    c->fftflag = (int *)calloc(c->nmu, sizeof(int));
    c->mofj = (int *)calloc(c->nlm, sizeof(int));
    for (i = 1; i <= c->nmu; i++) {
        memset(prc, 0, sizeof(float) * 6 * st->nlm);
        oracle_plmall(0, st->mu[i - 1], c->ml, c->mm, prc);
        for (j = 1; j <= c->nlm; j++)
            for (q = 1; q <= c->nstleg; q++) CMU1(c, q, j, i) = prc[(q - 1) + 6 * (j - 1)];
        memset(prc, 0, sizeof(float) * 6 * st->nlm);
        oracle_plmall(1, st->mu[i - 1], c->ml, c->mm, prc);
        for (j = 1; j <= c->nlm; j++)
            for (q = 1; q <= c->nstleg; q++) CMU2(c, q, i, j) = prc[(q - 1) + 6 * (j - 1)] * wtmu[i - 1];
        /* FFTFLAG of MAKE_ANGLE_SET  shdomsub2.f:1131 */
        mmax = st->nphi0max / 2 - 1; if (mmax < 0) mmax = 0;
        c->fftflag[i - 1] = (st->nphi0[i - 1] > 14) || (mmax > 15);
    }
    j = 0;
    for (l = 0; l <= c->ml; l++) {
        const int me = l < c->mm ? l : c->mm;
        for (m = -me; m <= me; m++) { c->mofj[j] = m; j++; }
    }
    free(prc);
    return c;
}

static void free_sh_do_coef(shdo_coef *c)
{
    free(c->cmu1); free(c->cmu2); free(c->fftflag); free(c->mofj); free(c);
}

/* azimuthal basis value: cos(m phi_k) for m >= 0, sin(|m| phi_k) for m < 0.  With FFTFLAG the angles
 * are the exact 2 pi (k-1)/N of RFFTB/RFFTF; otherwise CPHI1 = COS(M*PHI(I,K)) in REAL. */
static double az_basis(const oracle_state *st, const shdo_coef *c, int imu, int k, int m)
{
    const int n = st->nphi0[imu - 1];
    if (m == 0) return 1.0;
    if (c->fftflag[imu - 1]) {
        const double ang = 2.0 * acos(-1.0) * (double)(abs(m) * (k - 1) % n) / (double)n;
        return m > 0 ? cos(ang) : sin(ang);
    } else {
        const float ph = st->phi[(imu - 1) + st->nmu * (k - 1)];
        return m > 0 ? (double)cosf(m * ph) : (double)sinf(-m * ph);
    }
}

/* SH_TO_DO / SH_TO_DO_UNPOL  shdomsub1.f:2789-3040.  out is OUTDATA(NSTOKES,NPHI0MAX,NPTS). */
static void sh_to_do(const oracle_state *st, const shdo_coef *c, int imu, const int *shptr,
                     const float *indata, float *outdata, const double *aztab)
{
    const int ns = st->nstokes, mm = st->mm, nphi0 = st->nphi0[imu - 1], na = st->nphi0max;
    float *sumuv = (float *)malloc(sizeof(float) * ns * (2 * mm + 1));
    float *sumcs = (float *)malloc(sizeof(float) * ns * (2 * mm + 1));
    int i, j, k, m, n;
#define UV(n, m) sumuv[((n) - 1) + ns * ((m) + mm)]
#define CS(n, m) sumcs[((n) - 1) + ns * ((m) + mm)]
#define IN(n, j) indata[((n) - 1) + (size_t)ns * ((j) - 1)]
#define OUT(n, k, i) outdata[((n) - 1) + (size_t)ns * (((k) - 1) + (size_t)na * ((i) - 1))]
    for (i = 1; i <= st->npts; i++) {
        const int is = shptr[i - 1], nsh = shptr[i] - is;
        int me, ms;
        if (nsh == 0) {
            for (k = 1; k <= nphi0; k++) for (n = 1; n <= ns; n++) OUT(n, k, i) = 0.0f;
            continue;
        }
        me = c->mofj[nsh - 1]; if (me > nphi0 / 2 - 1) me = nphi0 / 2 - 1; if (me < 0) me = 0;
        ms = -me;
        memset(sumuv, 0, sizeof(float) * ns * (2 * mm + 1));
        for (j = 1; j <= nsh; j++) {
            m = c->mofj[j - 1];
            UV(1, m) = UV(1, m) + CMU1(c, 1, j, imu) * IN(1, is + j);
            if (ns > 1) {
                UV(2, m) = UV(2, m) + CMU1(c, 2, j, imu) * IN(2, is + j) + CMU1(c, 5, j, imu) * IN(3, is + j);
                UV(3, m) = UV(3, m) + CMU1(c, 6, j, imu) * IN(2, is + j) + CMU1(c, 3, j, imu) * IN(3, is + j);
            }
        }
        for (n = 1; n <= ns; n++) CS(n, 0) = UV(n, 0);
        for (m = 1; m <= me; m++) {
            CS(1, m) = UV(1, m) + UV(1, -m);
            CS(1, -m) = UV(1, -m) - UV(1, m);
            if (ns > 1) {
                CS(2, m) = UV(2, m) + UV(2, -m);
                CS(2, -m) = UV(2, -m) - UV(2, m);
                CS(3, -m) = UV(3, m) - UV(3, -m);
                CS(3, m) = UV(3, m) + UV(3, -m);
            }
        }
        for (k = 1; k <= nphi0; k++) {
            for (n = 1; n <= ns; n++) {
                if (c->fftflag[imu - 1]) {
                    double s = 0.0;
                    for (m = ms; m <= me; m++) s += aztab[(m + mm) + (2 * mm + 1) * (k - 1)] * (double)CS(n, m);
                    OUT(n, k, i) = (float)s;
                } else {
                    float s = 0.0f;
                    for (m = ms; m <= me; m++) s = s + (float)aztab[(m + mm) + (2 * mm + 1) * (k - 1)] * CS(n, m);
                    OUT(n, k, i) = s;
                }
            }
        }
    }
#undef IN
    free(sumuv); free(sumcs);
}

/* DO_TO_SH / DO_TO_SH_UNPOL  shdomsub1.f:3041-3260; accumulates into OUTDATA(NSTOKES,*) */
static void do_to_sh(const oracle_state *st, const shdo_coef *c, int imu, const int *rshptr,
                     const float *indata, float *outdata, const double *aztab)
{
    const int ns = st->nstokes, mm = st->mm, nphi0 = st->nphi0[imu - 1], na = st->nphi0max;
    float *sumuv = (float *)malloc(sizeof(float) * ns * (2 * mm + 1));
    float *sumcs = (float *)malloc(sizeof(float) * ns * (2 * mm + 1));
    const float delphi = 2.0f * acosf(-1.0f) / nphi0;
    int i, j, k, m, n, is = 0;
#define INP(n, k, i) indata[((n) - 1) + (size_t)ns * (((k) - 1) + (size_t)na * ((i) - 1))]
#define OUTD(n, j) outdata[((n) - 1) + (size_t)ns * ((j) - 1)]
    for (i = 1; i <= st->npts; i++) {
        const int nsh = rshptr[i] - rshptr[i - 1];
        int me = (nsh > 0) ? c->mofj[nsh - 1] : 0, ms;
        if (me > nphi0 / 2 - 1) me = nphi0 / 2 - 1;
        if (me < 0) me = 0;
        ms = -me;
        for (m = ms; m <= me; m++) {
            for (n = 1; n <= ns; n++) {
                if (c->fftflag[imu - 1]) {
                    double s = 0.0;
                    for (k = 1; k <= nphi0; k++) s += aztab[(m + mm) + (2 * mm + 1) * (k - 1)] * (double)INP(n, k, i);
                    CS(n, m) = (float)s * delphi;
                } else {
                    /* CPHI2(K,M,I) = CPHI1(M,K,I)*WTDO(I,K)/WTMU(I) */
                    float s = 0.0f;
                    for (k = 1; k <= nphi0; k++)
                        s = s + ((float)aztab[(m + mm) + (2 * mm + 1) * (k - 1)] * delphi) * INP(n, k, i);
                    CS(n, m) = s;
                }
            }
        }
        memset(sumuv, 0, sizeof(float) * ns * (2 * mm + 1));
        for (n = 1; n <= ns; n++) UV(n, 0) = CS(n, 0);
        for (m = 1; m <= me; m++) {
            UV(1, m) = CS(1, m) - CS(1, -m);
            UV(1, -m) = CS(1, m) + CS(1, -m);
            if (ns > 1) {
                UV(2, m) = CS(2, m) - CS(2, -m);
                UV(2, -m) = CS(2, m) + CS(2, -m);
                UV(3, m) = CS(3, m) + CS(3, -m);
                UV(3, -m) = CS(3, m) - CS(3, -m);
            }
        }
        for (j = 1; j <= nsh; j++) {
            m = c->mofj[j - 1];
            OUTD(1, is + j) = OUTD(1, is + j) + CMU2(c, 1, imu, j) * UV(1, m);
            if (ns > 1) {
                OUTD(2, is + j) = OUTD(2, is + j) + CMU2(c, 2, imu, j) * UV(2, m) + CMU2(c, 5, imu, j) * UV(3, m);
                OUTD(3, is + j) = OUTD(3, is + j) + CMU2(c, 6, imu, j) * UV(2, m) + CMU2(c, 3, imu, j) * UV(3, m);
            }
        }
        is = is + nsh;
    }
#undef INP
#undef OUTD
#undef UV
#undef CS
#undef OUT
    free(sumuv); free(sumcs);
}

/* ---- sweeping order ---- */
typedef struct { int sp, stack[50], ix, iy, iz, six, siy, siz, eix, eiy, eiz, dix, diy, diz; } sweep_state;

/* SWEEP_BASE_CELL  shdomsub1.f:4529-4620 */
static int sweep_base_cell(int bcflag, int nxc, int nyc, int nz, int ioct, int *icell, sweep_state *s)
{
    if (*icell == 0) {
        if (BTEST(ioct - 1, 0)) {
            s->dix = +1;
            if (BTEST(bcflag, 0)) { s->six = nxc; s->eix = nxc - 1 > 1 ? nxc - 1 : 1; }
            else { s->six = 1; s->eix = nxc; }
        } else {
            s->dix = -1;
            if (BTEST(bcflag, 0)) { s->six = 1; s->eix = 2 < nxc ? 2 : nxc; }
            else { s->six = nxc; s->eix = 1; }
        }
        if (BTEST(ioct - 1, 1)) {
            s->diy = +1;
            if (BTEST(bcflag, 1)) { s->siy = nyc; s->eiy = nyc - 1 > 1 ? nyc - 1 : 1; }
            else { s->siy = 1; s->eiy = nyc; }
        } else {
            s->diy = -1;
            if (BTEST(bcflag, 1)) { s->siy = 1; s->eiy = 2 < nyc ? 2 : nyc; }
            else { s->siy = nyc; s->eiy = 1; }
        }
        if (BTEST(ioct - 1, 2)) { s->diz = +1; s->siz = 1; s->eiz = nz - 1; }
        else { s->diz = -1; s->siz = nz - 1; s->eiz = 1; }
        s->ix = s->six; s->iy = s->siy; s->iz = s->siz;
    } else {
        if (s->ix == s->eix) {
            if (s->iy == s->eiy) {
                if (s->iz == s->eiz) return 0;
                s->iz = s->iz + s->diz;
                s->iy = s->siy;
            } else {
                s->iy = (s->iy + s->diy + nyc - 1) % nyc + 1;
            }
            s->ix = s->six;
        } else {
            s->ix = (s->ix + s->dix + nxc - 1) % nxc + 1;
        }
    }
    *icell = s->iz + (nz - 1) * (s->iy - 1) + (nz - 1) * nyc * (s->ix - 1);
    return 1;
}

/* SWEEP_NEXT_CELL  shdomsub1.f:4624-4700 */
static int sweep_next_cell(const oracle_state *st, int nxc, int nyc, int ioct, int *icell, sweep_state *s)
{
    int ic = *icell, done = 0;
    if (*icell == 0) {
        ic = 0;
        sweep_base_cell(st->bcflag, nxc, nyc, st->nz, ioct, &ic, s);
        s->sp = 0;
    }
    while (!done) {
        if (TREEPTR(st, 2, ic) == 0) {
            if (ic != *icell) done = 1;
            else {
                if (s->sp == 0) {
                    if (!sweep_base_cell(st->bcflag, nxc, nyc, st->nz, ioct, &ic, s)) return 0;
                } else {
                    ic = s->stack[s->sp - 1];
                    s->sp = s->sp - 1;
                }
            }
        } else {
            int idir;
            s->sp = s->sp + 1;
            if (s->sp > 50) return 0;
            idir = IBITS2(CELLFLAGS(st, ic)) - 1;
            if (BTEST(ioct - 1, idir)) {
                s->stack[s->sp - 1] = TREEPTR(st, 2, ic) + 1;
                ic = TREEPTR(st, 2, ic);
            } else {
                s->stack[s->sp - 1] = TREEPTR(st, 2, ic);
                ic = TREEPTR(st, 2, ic) + 1;
            }
        }
    }
    *icell = ic;
    return 1;
}

/* SWEEPING_ORDER  shdomsub1.f:3261-3353.  sweepord is SWEEPORD(NPTS,NOCT). */
static int sweeping_order(const oracle_state *st, int *sweepord)
{
    static const int corndog[8][8] = {{8,7,6,5,4,3,2,1},{7,8,5,6,3,4,1,2},{6,5,8,7,2,1,4,3},{5,6,7,8,1,2,3,4},
                                      {4,3,2,1,8,7,6,5},{3,4,1,2,7,8,5,6},{2,1,4,3,6,5,8,7},{1,2,3,4,5,6,7,8}};
    static const int ioctorder[8] = {1, 5, 2, 6, 3, 7, 4, 8};
    const int npts = st->npts;
    float *visited = (float *)malloc(sizeof(float) * npts);
    int noct, joct, nxc, nyc;
    if (BTEST(st->ipflag, 1) && BTEST(st->ipflag, 0)) noct = 2;
    else if (BTEST(st->ipflag, 1)) noct = 4;
    else noct = 8;
    nxc = st->nx;
    if (BTEST(st->bcflag, 0)) nxc = st->nx + 1;
    if (BTEST(st->bcflag, 2) && !BTEST(st->ipflag, 0)) nxc = st->nx - 1;
    nyc = st->ny;
    if (BTEST(st->bcflag, 1)) nyc = st->ny + 1;
    if (BTEST(st->bcflag, 3) && !BTEST(st->ipflag, 1)) nyc = st->ny - 1;
    for (joct = 1; joct <= noct; joct++) {
        const int ioct = ioctorder[joct - 1];
        sweep_state s;
        int ipt, iorder = 1, ipcell = 0, indexcorn = 8, icorner;
        memset(&s, 0, sizeof(s));
        for (ipt = 0; ipt < npts; ipt++) visited[ipt] = -1.0f;
        for (;;) {
            if (indexcorn == 8) {
                indexcorn = 1;
                if (!sweep_next_cell(st, nxc, nyc, ioct, &ipcell, &s)) break;
            } else {
                indexcorn = indexcorn + 1;
            }
            icorner = corndog[ioct - 1][indexcorn - 1];
            ipt = GRIDPTR(st, icorner, ipcell);
            if (visited[ipt - 1] >= 0.0f) continue;
            visited[ipt - 1] = 1.0f;
            sweepord[(iorder - 1) + (size_t)npts * (joct - 1)] = (ipcell << 3) | (icorner - 1);
            iorder = iorder + 1;
        }
        if (iorder - 1 != npts) { free(visited); return 1; }
    }
    free(visited);
    return 0;
}

/* BACK_INT_GRID1D  shdomsub1.f:4295-4468.  source is SOURCE(NSTOKES,NA,NPTS) (discrete-ordinate source of
 * this zenith angle), gridrad is GRIDRAD(NSTOKES,NPTS). */
static int back_int_grid1d(const oracle_state *st, const int *sweepord, float mu, int kang,
                           const float *extinct, const float *source, float *gridrad, char *errmsg)
{
    static const int gridface[6] = {0, 0, 0, 0, 1, 5};
    static const int joctorder[8] = {1, 1, 1, 1, 2, 2, 2, 2};
    const int ns = st->nstokes, na = st->nphi0max, npts = st->npts;
    double eps, cz, czinv, xe, ye, ze, so, ext, ext0, ext1, ext0p, tau, transcell, abscell, transmit;
    double src[4], srcext0[4], srcext1[4], srcext0p[4], rad[4], rad0[4];
    int bitz, ioct, joct, iorder, k;
#define SRC(n, k, i) source[((n) - 1) + (size_t)ns * (((k) - 1) + (size_t)na * ((i) - 1))]
#define GR(n, i) gridrad[((n) - 1) + (size_t)ns * ((i) - 1)]
    eps = 1.0E-3f * (GRIDPOS(st, 3, GRIDPTR(st, 8, 1)) - GRIDPOS(st, 3, GRIDPTR(st, 1, 1)));
    cz = -mu;
    czinv = 1.0 / cz;
    if (cz < -1.0E-3f) bitz = 1;
    else if (cz > 1.0E-3f) bitz = 0;
    else { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID3D: Bad MU"); return 1; }
    ioct = 1 + 4 * bitz;
    joct = joctorder[ioct - 1];
    (void)xe; (void)ye;
    for (iorder = 1; iorder <= npts; iorder++) {
        const int so_entry = sweepord[(iorder - 1) + (size_t)npts * (joct - 1)];
        const int ipcell = so_entry >> 3, icorner = (so_entry & 7) + 1;
        const int ipt = GRIDPTR(st, icorner, ipcell);
        int icell, validrad, inextcell = 0;
        if (GR(1, ipt) >= 0.0f) continue;
        icell = ipcell;
        transmit = 1.0;
        ext1 = extinct[ipt - 1];
        for (k = 0; k < ns; k++) { rad[k] = 0.0; srcext1[k] = ext1 * SRC(k + 1, kang, ipt); }
        ze = GRIDPOS(st, 3, ipt);
        validrad = 0;
        while (!validrad) {
            int iopp, iface, i1;
            if (icell <= 0) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID: ICELL=0"); return 1; }
            iopp = GRIDPTR(st, 9 - ioct, icell);
            so = (GRIDPOS(st, 3, iopp) - ze) * czinv;
            if (so < -eps) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID1D: SO<0"); return 1; }
            ze = ze + so * cz;
            iface = 6 - bitz;
            inextcell = NEIGHPTR(st, iface, icell);
            i1 = GRIDPTR(st, gridface[iface - 1], icell);
            if (inextcell > 0) ze = GRIDPOS(st, 3, GRIDPTR(st, ioct, inextcell));
            ext0 = extinct[i1 - 1];
            for (k = 0; k < ns; k++) srcext0[k] = extinct[i1 - 1] * SRC(k + 1, kang, i1);
            ext = 0.5 * (ext0 + ext1);
            tau = ext * so;
            if (tau >= 0.5) {
                transcell = exp(-tau);
                abscell = 1.0 - transcell;
            } else {
                abscell = tau * (1.0 - 0.5 * tau * (1.0 - 0.33333333333 * tau * (1 - 0.25 * tau)));
                transcell = 1.0 - abscell;
            }
            if (tau <= 2.0) {
                if (ext == 0.0) { for (k = 0; k < ns; k++) src[k] = 0.0; }
                else {
                    for (k = 0; k < ns; k++)
                        src[k] = (0.5 * (srcext0[k] + srcext1[k])
                                  + 0.08333333333 * (ext0 * srcext1[k] - ext1 * srcext0[k]) * so) / ext;
                }
            } else {
                ext0p = ext0;
                for (k = 0; k < ns; k++) srcext0p[k] = srcext0[k];
                if (tau > 4.0) {
                    ext0p = ext1 + (ext0 - ext1) * 4.0 / tau;
                    if (ext0 > 0.0) for (k = 0; k < ns; k++) srcext0p[k] = srcext0[k] * ext0p / ext0;
                }
                for (k = 0; k < ns; k++)
                    src[k] = 1.0 / (ext0p + ext1) * (srcext0p[k] + srcext1[k]
                             + (ext0p * srcext1[k] - ext1 * srcext0p[k]) * 2.0 / (ext0p + ext1)
                               * (1 - 2 / tau + 2 * transcell / abscell));
            }
            src[0] = fmax(src[0], 0.0);
            if (GR(1, i1) >= -0.1f) {
                validrad = 1;
                for (k = 0; k < ns; k++) {
                    rad0[k] = GR(k + 1, i1);
                    rad[k] = rad[k] + transmit * (rad0[k] * transcell + src[k] * abscell);
                }
            } else {
                for (k = 0; k < ns; k++) {
                    rad[k] = rad[k] + transmit * src[k] * abscell;
                    srcext1[k] = srcext0[k];
                }
                transmit = transmit * transcell;
                ext1 = ext0;
                icell = inextcell;
            }
        }
        for (k = 0; k < ns; k++) GR(k + 1, ipt) = (float)rad[k];
    }
    return 0;
}

static float oracle_transmin = 1.00f;
void oracle_set_transmin(float t) { oracle_transmin = t; }

/* BACK_INT_GRID3D / BACK_INT_GRID3D_UNPOL  shdomsub1.f:3354-3690 / 3696-4036 (identical apart from NSTOKES).
   source is SOURCE(NSTOKES,NA,NPTS) (the discrete-ordinate source of one zenith angle), gridrad GRIDRAD(NSTOKES,NPTS)
   with GRIDRAD(1,.) < 0 for points without a value yet.  The multi-processor branch (BCFLAG bits 2,3) is not restated. */
static int back_int_grid3d(const oracle_state *st, const int *sweepord, float mu, float phi, float transmin, int kang,
                           const float *extinct, const float *source, float *gridrad, char *errmsg)
{
    static const int gridface[6][4] = {{1, 3, 5, 7}, {2, 4, 6, 8}, {1, 2, 5, 6}, {3, 4, 7, 8}, {1, 2, 3, 4}, {5, 6, 7, 8}};
    static const int oppface[6] = {2, 1, 4, 3, 6, 5};
    static const int joctorder3[8] = {1, 3, 5, 7, 2, 4, 6, 8}, joctorder2[8] = {1, 3, 1, 3, 2, 4, 2, 4};
    const int ns = st->nstokes, na = st->nphi0max, npts = st->npts;
    double eps, pi, cx, cy, cz, cxinv, cyinv, czinv, xe, ye, ze, so, sox, soy, soz, u, v, f1, f2, f3, f4;
    double ext, ext0, ext1, ext0p, tau, transcell, abscell, transmit;
    double src[4], srcext0[4], srcext1[4], srcext0p[4], rad[4], rad0[4];
    int bitx, bity, bitz, ioct, joct, iorder, k;
    eps = 1.0E-3f * (GRIDPOS(st, 3, GRIDPTR(st, 8, 1)) - GRIDPOS(st, 3, GRIDPTR(st, 1, 1)));
    pi = acosf(-1.0f);
    cx = sqrtf(1.0f - mu * mu) * cos(phi + pi);
    cy = sqrtf(1.0f - mu * mu) * sin(phi + pi);
    cz = -mu;
    if (fabs(cx) > 1.0E-5f) cxinv = 1.0 / cx; else { cx = 0.0; cxinv = 1.0E6f; }
    if (fabs(cy) > 1.0E-5f) cyinv = 1.0 / cy; else { cy = 0.0; cyinv = 1.0E6f; }
    czinv = 1.0 / cz;
    bitx = cx < 0.0 ? 1 : 0;
    bity = cy < 0.0 ? 1 : 0;
    if (cz < -1.0E-3f) bitz = 1;
    else if (cz > 1.0E-3f) bitz = 0;
    else { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID: Bad MU"); return 1; }
    ioct = 1 + bitx + 2 * bity + 4 * bitz;
    joct = BTEST(st->ipflag, 1) ? joctorder2[ioct - 1] : joctorder3[ioct - 1];
    for (iorder = 1; iorder <= npts; iorder++) {
        const int so_entry = sweepord[(iorder - 1) + (size_t)npts * (joct - 1)];
        const int ipcell = so_entry >> 3, icorner = (so_entry & 7) + 1;
        const int ipt = GRIDPTR(st, icorner, ipcell);
        int icell, validrad, inextcell = 0;
        if (GR(1, ipt) >= 0.0f) continue;
        icell = ipcell;
        transmit = 1.0;
        ext1 = extinct[ipt - 1];
        for (k = 0; k < ns; k++) { rad[k] = 0.0; srcext1[k] = ext1 * SRC(k + 1, kang, ipt); }
        xe = GRIDPOS(st, 1, ipt);
        ye = GRIDPOS(st, 2, ipt);
        ze = GRIDPOS(st, 3, ipt);
        validrad = 0;
        while (!validrad) {
            int ipinx, ipiny, iopp, iface, jface, kface, ic, i1, i2, i3, i4, validface;
            if (icell <= 0) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID: ICELL=0"); return 1; }
            ipinx = BTEST(CELLFLAGS(st, icell), 0);
            ipiny = BTEST(CELLFLAGS(st, icell), 1);
            iopp = GRIDPTR(st, 9 - ioct, icell);
            sox = ipinx ? 1.0E20f : (GRIDPOS(st, 1, iopp) - xe) * cxinv;
            soy = ipiny ? 1.0E20f : (GRIDPOS(st, 2, iopp) - ye) * cyinv;
            soz = (GRIDPOS(st, 3, iopp) - ze) * czinv;
            so = fmin(fmin(sox, soy), soz);
            if (so < -eps) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID: SO<0"); return 1; }
            xe = xe + so * cx;
            ye = ye + so * cy;
            ze = ze + so * cz;
            if (sox <= soz && sox <= soy) { iface = 2 - bitx; jface = 1; }
            else if (soy <= soz) { iface = 4 - bity; jface = 2; }
            else { iface = 6 - bitz; jface = 3; }
            inextcell = oracle_next_cell(st, xe, ye, ze, iface, jface, icell);
            if (NEIGHPTR(st, iface, icell) >= 0) { kface = iface; ic = icell; }
            else { kface = oppface[iface - 1]; ic = inextcell; }
            i1 = GRIDPTR(st, gridface[kface - 1][0], ic);
            i2 = GRIDPTR(st, gridface[kface - 1][1], ic);
            i3 = GRIDPTR(st, gridface[kface - 1][2], ic);
            i4 = GRIDPTR(st, gridface[kface - 1][3], ic);
            if (jface == 1) {
                u = (ze - GRIDPOS(st, 3, i1)) / (GRIDPOS(st, 3, i3) - GRIDPOS(st, 3, i1));
                v = ipiny ? 0.5 : (ye - GRIDPOS(st, 2, i1)) / (GRIDPOS(st, 2, i2) - GRIDPOS(st, 2, i1));
            } else if (jface == 2) {
                u = (ze - GRIDPOS(st, 3, i1)) / (GRIDPOS(st, 3, i3) - GRIDPOS(st, 3, i1));
                v = ipinx ? 0.5 : (xe - GRIDPOS(st, 1, i1)) / (GRIDPOS(st, 1, i2) - GRIDPOS(st, 1, i1));
            } else {
                u = ipiny ? 0.5 : (ye - GRIDPOS(st, 2, i1)) / (GRIDPOS(st, 2, i3) - GRIDPOS(st, 2, i1));
                v = ipinx ? 0.5 : (xe - GRIDPOS(st, 1, i1)) / (GRIDPOS(st, 1, i2) - GRIDPOS(st, 1, i1));
            }
            if (inextcell > 0) {
                if (jface == 1) xe = GRIDPOS(st, 1, GRIDPTR(st, ioct, inextcell));
                else if (jface == 2) ye = GRIDPOS(st, 2, GRIDPTR(st, ioct, inextcell));
                else ze = GRIDPOS(st, 3, GRIDPTR(st, ioct, inextcell));
            }
            f1 = (1 - u) * (1 - v);
            f2 = (1 - u) * v;
            f3 = u * (1 - v);
            f4 = u * v;
            ext0 = f1 * extinct[i1 - 1] + f2 * extinct[i2 - 1] + f3 * extinct[i3 - 1] + f4 * extinct[i4 - 1];
            for (k = 0; k < ns; k++)
                srcext0[k] = f1 * SRC(k + 1, kang, i1) * extinct[i1 - 1] + f2 * SRC(k + 1, kang, i2) * extinct[i2 - 1]
                           + f3 * SRC(k + 1, kang, i3) * extinct[i3 - 1] + f4 * SRC(k + 1, kang, i4) * extinct[i4 - 1];
            ext = 0.5 * (ext0 + ext1);
            tau = ext * so;
            if (tau >= 0.5) {
                transcell = exp(-tau);
                abscell = 1.0 - transcell;
            } else {
                abscell = tau * (1.0 - 0.5 * tau * (1.0 - 0.33333333333 * tau * (1 - 0.25 * tau)));
                transcell = 1.0 - abscell;
            }
            if (tau <= 2.0) {
                if (ext == 0.0) { for (k = 0; k < ns; k++) src[k] = 0.0; }
                else {
                    for (k = 0; k < ns; k++)
                        src[k] = (0.5 * (srcext0[k] + srcext1[k])
                                  + 0.08333333333 * (ext0 * srcext1[k] - ext1 * srcext0[k]) * so) / ext;
                }
            } else {
                ext0p = ext0;
                for (k = 0; k < ns; k++) srcext0p[k] = srcext0[k];
                if (tau > 4.0) {
                    ext0p = ext1 + (ext0 - ext1) * 4.0 / tau;
                    if (ext0 > 0.0) for (k = 0; k < ns; k++) srcext0p[k] = srcext0[k] * ext0p / ext0;
                }
                for (k = 0; k < ns; k++)
                    src[k] = 1.0 / (ext0p + ext1) * (srcext0p[k] + srcext1[k]
                             + (ext0p * srcext1[k] - ext1 * srcext0p[k]) * 2.0 / (ext0p + ext1)
                               * (1 - 2 / tau + 2 * transcell / abscell));
            }
            src[0] = fmax(src[0], 0.0);
            for (k = 0; k < ns; k++) rad[k] = rad[k] + transmit * src[k] * abscell;
            transmit = transmit * transcell;
            if (rad[0] < -1.0E-5f) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID3D: RAD<0"); return 1; }
            validface = GR(1, i1) >= -0.1f && GR(1, i2) >= -0.1f && GR(1, i3) >= -0.1f && GR(1, i4) >= -0.1f;
            if (inextcell <= 0 || (transmit <= transmin && validface)) {
                if (validface) {
                    validrad = 1;
                    for (k = 0; k < ns; k++) {
                        rad0[k] = f1 * GR(k + 1, i1) + f2 * GR(k + 1, i2) + f3 * GR(k + 1, i3) + f4 * GR(k + 1, i4);
                        rad[k] = rad[k] + transmit * rad0[k];
                    }
                } else {
                    if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID3D: INEXTCELL=0 without a valid face");
                    return 1;
                }
            } else {
                ext1 = ext0;
                for (k = 0; k < ns; k++) srcext1[k] = srcext0[k];
                icell = inextcell;
            }
        }
        for (k = 0; k < ns; k++) GR(k + 1, ipt) = (float)rad[k];
    }
    return 0;
}

/* BACK_INT_GRID2D  shdomsub1.f:4039-4293: independent pixels in Y (IPFLAG bit 1), rays in the X-Z plane, faces of two
   grid points. */
static int back_int_grid2d(const oracle_state *st, const int *sweepord, float mu, float phi, float transmin, int kang,
                           const float *extinct, const float *source, float *gridrad, char *errmsg)
{
    static const int gridface[6][2] = {{1, 5}, {2, 6}, {0, 0}, {0, 0}, {1, 2}, {5, 6}};
    static const int oppface[6] = {2, 1, 4, 3, 6, 5};
    static const int joctorder[8] = {1, 3, 1, 3, 2, 4, 2, 4};
    const int ns = st->nstokes, na = st->nphi0max, npts = st->npts;
    double eps, pi, cx, cz, cxinv, czinv, xe, ye, ze, so, sox, soz, u, f1, f2;
    double ext, ext0, ext1, ext0p, tau, transcell, abscell, transmit;
    double src[4], srcext0[4], srcext1[4], srcext0p[4], rad[4], rad0[4];
    int bitx, bitz, ioct, joct, iorder, k;
    eps = 1.0E-3f * (GRIDPOS(st, 3, GRIDPTR(st, 8, 1)) - GRIDPOS(st, 3, GRIDPTR(st, 1, 1)));
    pi = acosf(-1.0f);
    cx = sqrtf(1.0f - mu * mu) * cos(phi + pi);
    cz = -mu;
    if (fabs(cx) > 1.0E-5f) cxinv = 1.0 / cx; else { cx = 0.0; cxinv = 1.0E6f; }
    czinv = 1.0 / cz;
    bitx = cx < 0.0 ? 1 : 0;
    if (cz < -1.0E-3f) bitz = 1;
    else if (cz > 1.0E-3f) bitz = 0;
    else { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID: Bad MU"); return 1; }
    ioct = 1 + bitx + 4 * bitz;
    joct = joctorder[ioct - 1];
    for (iorder = 1; iorder <= npts; iorder++) {
        const int so_entry = sweepord[(iorder - 1) + (size_t)npts * (joct - 1)];
        const int ipcell = so_entry >> 3, icorner = (so_entry & 7) + 1;
        const int ipt = GRIDPTR(st, icorner, ipcell);
        int icell, validrad, inextcell = 0;
        if (GR(1, ipt) >= 0.0f) continue;
        icell = ipcell;
        transmit = 1.0;
        ext1 = extinct[ipt - 1];
        for (k = 0; k < ns; k++) { rad[k] = 0.0; srcext1[k] = ext1 * SRC(k + 1, kang, ipt); }
        xe = GRIDPOS(st, 1, ipt);
        ye = GRIDPOS(st, 2, ipt);
        ze = GRIDPOS(st, 3, ipt);
        validrad = 0;
        while (!validrad) {
            int ipinx, iopp, iface, jface, kface, ic, i1, i2, validface;
            if (icell <= 0) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID: ICELL=0"); return 1; }
            ipinx = BTEST(CELLFLAGS(st, icell), 0);
            iopp = GRIDPTR(st, 9 - ioct, icell);
            sox = ipinx ? 1.0E20f : (GRIDPOS(st, 1, iopp) - xe) * cxinv;
            soz = (GRIDPOS(st, 3, iopp) - ze) * czinv;
            so = fmin(sox, soz);
            if (so < -eps) { if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID2D: SO<0"); return 1; }
            xe = xe + so * cx;
            ze = ze + so * cz;
            if (sox <= soz) { iface = 2 - bitx; jface = 1; }
            else { iface = 6 - bitz; jface = 3; }
            inextcell = oracle_next_cell(st, xe, ye, ze, iface, jface, icell);
            if (NEIGHPTR(st, iface, icell) >= 0) { kface = iface; ic = icell; }
            else { kface = oppface[iface - 1]; ic = inextcell; }
            i1 = GRIDPTR(st, gridface[kface - 1][0], ic);
            i2 = GRIDPTR(st, gridface[kface - 1][1], ic);
            if (jface == 1) u = (ze - GRIDPOS(st, 3, i1)) / (GRIDPOS(st, 3, i2) - GRIDPOS(st, 3, i1));
            else u = ipinx ? 0.5 : (xe - GRIDPOS(st, 1, i1)) / (GRIDPOS(st, 1, i2) - GRIDPOS(st, 1, i1));
            if (inextcell > 0) {
                if (jface == 1) xe = GRIDPOS(st, 1, GRIDPTR(st, ioct, inextcell));
                else ze = GRIDPOS(st, 3, GRIDPTR(st, ioct, inextcell));
            }
            f1 = 1 - u;
            f2 = u;
            ext0 = f1 * extinct[i1 - 1] + f2 * extinct[i2 - 1];
            for (k = 0; k < ns; k++)
                srcext0[k] = f1 * SRC(k + 1, kang, i1) * extinct[i1 - 1] + f2 * SRC(k + 1, kang, i2) * extinct[i2 - 1];
            ext = 0.5 * (ext0 + ext1);
            tau = ext * so;
            if (tau >= 0.5) {
                transcell = exp(-tau);
                abscell = 1.0 - transcell;
            } else {
                abscell = tau * (1.0 - 0.5 * tau * (1.0 - 0.33333333333 * tau * (1 - 0.25 * tau)));
                transcell = 1.0 - abscell;
            }
            if (tau <= 2.0) {
                if (ext == 0.0) { for (k = 0; k < ns; k++) src[k] = 0.0; }
                else {
                    for (k = 0; k < ns; k++)
                        src[k] = (0.5 * (srcext0[k] + srcext1[k])
                                  + 0.08333333333 * (ext0 * srcext1[k] - ext1 * srcext0[k]) * so) / ext;
                }
            } else {
                ext0p = ext0;
                for (k = 0; k < ns; k++) srcext0p[k] = srcext0[k];
                if (tau > 4.0) {
                    ext0p = ext1 + (ext0 - ext1) * 4.0 / tau;
                    if (ext0 > 0.0) for (k = 0; k < ns; k++) srcext0p[k] = srcext0[k] * ext0p / ext0;
                }
                for (k = 0; k < ns; k++)
                    src[k] = 1.0 / (ext0p + ext1) * (srcext0p[k] + srcext1[k]
                             + (ext0p * srcext1[k] - ext1 * srcext0p[k]) * 2.0 / (ext0p + ext1)
                               * (1 - 2 / tau + 2 * transcell / abscell));
            }
            src[0] = fmax(src[0], 0.0);
            for (k = 0; k < ns; k++) rad[k] = rad[k] + transmit * src[k] * abscell;
            transmit = transmit * transcell;
            validface = GR(1, i1) >= -0.1f && GR(1, i2) >= -0.1f;
            if (inextcell <= 0 || (transmit <= transmin && validface)) {
                if (validface) {
                    validrad = 1;
                    for (k = 0; k < ns; k++) {
                        rad0[k] = f1 * GR(k + 1, i1) + f2 * GR(k + 1, i2);
                        rad[k] = rad[k] + transmit * rad0[k];
                    }
                } else {
                    if (errmsg) snprintf(errmsg, 600, "BACK_INT_GRID2D: INEXTCELL=0 without a valid face");
                    return 1;
                }
            } else {
                ext1 = ext0;
                for (k = 0; k < ns; k++) srcext1[k] = srcext0[k];
                icell = inextcell;
            }
        }
        for (k = 0; k < ns; k++) GR(k + 1, ipt) = (float)rad[k];
    }
    return 0;
}

/* RADIANCE_TRUNCATION  shdomsub1.f:1615-1805 */
static int radiance_truncation(const oracle_state *st, int highorderrad, const int *shptr, const float *radiance,
                               int maxir, int fixsh, float shacc, int *rshptr, const int *lofj)
{
    const int npts = st->npts, ml = st->ml, mm = st->mm, ns = st->nstokes, nq = 8 * st->maxnmicro;
    const int nlt = st->nstleg * (st->nleg + 1);
    int i, ir, iro, notend = 1;
#define LEG1(l, iph) st->legen[(size_t)nlt * ((iph) - 1) + st->nstleg * (l)]
#define IPH(q, i, ipa) st->iphase[((q) - 1) + (size_t)nq * (((i) - 1) + (size_t)npts * ((ipa) - 1))]
#define PWT(q, i, ipa) st->phaseinterpwt[((q) - 1) + (size_t)nq * (((i) - 1) + (size_t)npts * ((ipa) - 1))]
    if (!fixsh) {
        const float radmin = shacc;
        float f[64];
        int outofmem = 0;
        iro = 0; ir = 0;
        rshptr[0] = 0;
        for (i = 1; i <= npts; i++) {
            const int nro = rshptr[i] - iro;
            const float ext = st->total_ext[i - 1];
            int lr, l, ipa, q, nsx, ls, nr;
            if (nro == 0) notend = 0;
            if (notend) {
                if (st->interp_new && st->deltam) {
                    for (ipa = 1; ipa <= st->npart; ipa++) {
                        f[ipa - 1] = 0.0f;
                        if (PWT(1, i, ipa) >= st->phasemax) f[ipa - 1] = LEG1(ml + 1, IPH(1, i, ipa));
                        else
                            for (q = 1; q <= nq; q++) {
                                if (PWT(q, i, ipa) <= 1e-5f) continue;
                                /* the reference indexes LEGEN(Q,ML+1,.) here (shdomsub1.f:1672), i.e. the Q-th
                                 * Stokes element (linear storage order beyond NSTLEG); restated as is */
                                size_t off = (size_t)nlt * (IPH(q, i, ipa) - 1) + st->nstleg * (ml + 1) + (q - 1);
                                if (off >= (size_t)nlt * st->numphase) off = (size_t)nlt * st->numphase - 1;
                                f[ipa - 1] = f[ipa - 1] + st->legen[off] * PWT(q, i, ipa);
                            }
                        f[ipa - 1] = 1.0f / (1 - f[ipa - 1]);
                    }
                }
                lr = 1;
                for (l = 1; l <= ml; l++) {
                    float rad = 0.0f;
                    for (ipa = 1; ipa <= st->npart; ipa++) {
                        float w, legent;
                        if (ext == 0.0f) w = 1.0f; else w = st->extinct[(i - 1) + (size_t)npts * (ipa - 1)] / ext;
                        if (w == 0.0f) continue;
                        if (!st->interp_new) {
                            rad = rad + w * st->albedo[(i - 1) + (size_t)npts * (ipa - 1)] * LEG1(l, IPH(1, i, ipa))
                                        * radiance[0 + (size_t)ns * iro];
                        } else {
                            if (PWT(1, i, ipa) >= st->phasemax) legent = LEG1(l, IPH(1, i, ipa));
                            else {
                                legent = 0.0f;
                                for (q = 1; q <= nq; q++) {
                                    if (PWT(q, i, ipa) <= 1e-5f) continue;
                                    legent = legent + LEG1(l, IPH(q, i, ipa)) * PWT(q, i, ipa);
                                }
                            }
                            if (st->deltam) legent = legent * f[ipa - 1];
                            rad = rad + w * st->albedo[(i - 1) + (size_t)npts * (ipa - 1)] * legent
                                        * radiance[0 + (size_t)ns * iro];
                        }
                    }
                    if (rad > radmin) lr = l;
                }
            } else {
                lr = ml;
            }
            iro = iro + nro;
            nsx = shptr[i] - shptr[i - 1]; if (nsx < 1) nsx = 1;
            ls = lofj[nsx - 1];
            if (lr > ls + ml / 8 + 2) lr = ls + ml / 8 + 2;
            if (highorderrad) lr = ml;
            if (lr <= mm) nr = (lr * (lr + 1)) + lr + 1;
            else nr = (2 * mm + 1) * lr - (mm * (1 + (mm - 1))) + mm + 1;
            ir = ir + nr;
            rshptr[i] = ir;
            if (ir > maxir) { outofmem = 1; break; }
        }
        if (!outofmem) { rshptr[npts + 1] = ir; return 0; }
    }
    rshptr[0] = 0;
    ir = 0;
    for (i = 1; i <= npts; i++) {
        int nr = shptr[i] - shptr[i - 1]; if (nr < 4) nr = 4;
        if (highorderrad) {
            if (ml <= mm) nr = (ml * (ml + 1)) + ml + 1;
            else nr = (2 * mm + 1) * ml - (mm * (1 + (mm - 1))) + mm + 1;
        }
        ir = ir + nr;
        if (ir > maxir) return 2;
        rshptr[i] = ir;
    }
    rshptr[npts + 1] = ir;
    return 0;
#undef LEG1
#undef IPH
#undef PWT
}

/* PATH_INTEGRATION  shdomsub1.f:1836-2167 */
static int path_integration(oracle_state *st, const shdo_coef *c, const int *sweepord, const int *shptr,
                            const float *source, const int *rshptr, float *radiance, float *fluxes, float *bcrad,
                            float *work, float *gridrad, char *errmsg)
{
    const int ns = st->nstokes, npts = st->npts, na = st->nphi0max, mm = st->mm;
    const int lambertian = st->sfctype1 == 'L';
    double *aztab = (double *)malloc(sizeof(double) * (2 * mm + 1) * na);
    int i, k, imu, iphi, ibc, iang, iupdown, nr, m, ierr = 0;
#define WORK(n, k, i) work[((n) - 1) + (size_t)ns * (((k) - 1) + (size_t)na * ((i) - 1))]
    for (i = 0; i < npts; i++) { fluxes[2 * i] = 0.0f; fluxes[2 * i + 1] = 0.0f; gridrad[(size_t)ns * i] = -1.0f; }
    nr = rshptr[npts];
    memset(radiance, 0, sizeof(float) * (size_t)ns * nr);
    iang = 1;
    for (imu = 1; imu <= st->nmu && !ierr; imu++) {
        const int nphi0 = st->nphi0[imu - 1];
        if (imu == st->nmu / 2 + 1 && lambertian) {
            st->fluxes = fluxes;
            oracle_lambertian_boundary(st, bcrad);
        }
        for (k = 1; k <= nphi0; k++)
            for (m = -mm; m <= mm; m++) aztab[(m + mm) + (2 * mm + 1) * (k - 1)] = az_basis(st, c, imu, k, m);
        sh_to_do(st, c, imu, shptr, source, work, aztab);
        for (i = 1; i <= npts; i++)
            for (iphi = 1; iphi <= nphi0; iphi++) WORK(1, iphi, i) = fmaxf(0.0f, WORK(1, iphi, i));
        for (iphi = 1; iphi <= nphi0 && !ierr; iphi++) {
            if (st->mu[imu - 1] < 0.0f) {
                float top[4];
                oracle_compute_top_radiances(st, st->skyrad, imu, iphi, -2.0f, -2.0f, -1, top);
                iupdown = 1;
                for (ibc = 1; ibc <= st->ntoppts; ibc++) {
                    i = st->bcptr[ibc - 1];
                    for (k = 0; k < ns; k++) {
                        bcrad[k + ns * (ibc - 1)] = top[k];
                        gridrad[k + (size_t)ns * (i - 1)] = top[k];
                    }
                }
            } else {
                iupdown = 2;
                if (!lambertian) {
                    if (oracle_variable_brdf_surface(st, 1, st->nbotpts, st->mu[imu - 1],
                                                     st->phi[(imu - 1) + st->nmu * (iphi - 1)],
                                                     bcrad + (size_t)ns * st->ntoppts)) {
                        if (errmsg) snprintf(errmsg, 600, "SURFACE_BRDF: unsupported BRDF call");
                        ierr = 1; break;
                    }
                }
                for (ibc = 1; ibc <= st->nbotpts; ibc++) {
                    const float sg = st->sfcgridrad
                        ? st->sfcgridrad[(iang - st->nang / 2) + (size_t)(st->nang / 2 + 1) * (ibc - 1)] : 0.0f;
                    i = st->bcptr[st->maxnbc + ibc - 1];
                    for (k = 0; k < ns; k++)
                        gridrad[k + (size_t)ns * (i - 1)] = bcrad[k + ns * (ibc + st->ntoppts - 1)] + sg;
                }
            }
            if (BTEST(st->ipflag, 1) && BTEST(st->ipflag, 0)) {
                ierr = back_int_grid1d(st, sweepord, st->mu[imu - 1], iphi, st->total_ext, work, gridrad, errmsg);
            } else if (BTEST(st->ipflag, 1) && !BTEST(st->bcflag, 2) && !BTEST(st->bcflag, 3)) {
                ierr = back_int_grid2d(st, sweepord, st->mu[imu - 1], st->phi[(imu - 1) + st->nmu * (iphi - 1)], oracle_transmin,
                                       iphi, st->total_ext, work, gridrad, errmsg);
            } else if (!BTEST(st->ipflag, 1) && !BTEST(st->bcflag, 2) && !BTEST(st->bcflag, 3)) {
                /* TRANSMIN: numerical parameter `transmin` of at3d (configuration.py:28, default 1.0) */
                ierr = back_int_grid3d(st, sweepord, st->mu[imu - 1], st->phi[(imu - 1) + st->nmu * (iphi - 1)], oracle_transmin,
                                       iphi, st->total_ext, work, gridrad, errmsg);
            } else {
                if (errmsg) snprintf(errmsg, 600, "oracle solver: the multi-processor sweeps are not restated");
                ierr = 3;
            }
            if (ierr) break;
            for (i = 1; i <= npts; i++) {
                if (gridrad[(size_t)ns * (i - 1)] < -0.0001f) {
                    gridrad[(size_t)ns * (i - 1)] = 0.0f;
                } else {
                    for (k = 0; k < ns; k++) WORK(k + 1, iphi, i) = gridrad[k + (size_t)ns * (i - 1)];
                    gridrad[(size_t)ns * (i - 1)] = -1.0f;
                }
            }
            {
                const float a = fabsf(st->mu[imu - 1]) * st->wtdo[(imu - 1) + st->nmu * (iphi - 1)];
                for (i = 1; i <= npts; i++)
                    fluxes[(iupdown - 1) + 2 * (size_t)(i - 1)] = fluxes[(iupdown - 1) + 2 * (size_t)(i - 1)]
                                                                  + a * WORK(1, iphi, i);
            }
            if (!lambertian && imu <= st->nmu / 2) {
                const int i1 = st->ntoppts + st->nbotpts * iang;
                for (ibc = 1; ibc <= st->nbotpts; ibc++) {
                    i = st->bcptr[st->maxnbc + ibc - 1];
                    for (k = 0; k < ns; k++) bcrad[k + (size_t)ns * (ibc + i1 - 1)] = WORK(k + 1, iphi, i);
                }
            }
            iang = iang + 1;
        }
        if (!ierr) do_to_sh(st, c, imu, rshptr, work, radiance, aztab);
    }
#undef WORK
    free(aztab);
    return ierr;
}

/* SOLUTION_ITERATIONS (fixed grid)  shdomsub1.f:445-822 + CALC_ACCEL_SOLCRIT shdom_nompi.f:317-349 */
int oracle_solve_fixed_grid(const oracle_state *st_in, const float *wtmu, int maxiter, float solacc, float shacc,
                            int accelflag, int highorderrad, int iterfixsh, int maxiv,
                            int *shptr, float *source, int *rshptr, float *radiance, float *fluxes, float *bcrad,
                            int *iters_out, float *solcrit_out, char *errmsg)
{
    return oracle_solve_fixed_grid_from(st_in, wtmu, maxiter, solacc, shacc, accelflag, highorderrad, iterfixsh, maxiv, 0,
                                        shptr, source, rshptr, radiance, fluxes, bcrad, iters_out, solcrit_out, errmsg);
}

/* restore != 0: the solution iterations continue from the SHPTR / SOURCE / RSHPTR / RADIANCE passed in (a solution of a
 * nearby medium on the same grid) instead of the first guess -- INIT_SOLUTION with INRADFLAG=.FALSE. after
 * RTE.load_solution (at3d/solver.py:2654-2666, shdomsub1.f:356-391): no INIT_RADIANCE, no first COMPUTE_SOURCE,
 * OSHPTR = SHPTR and DELSOURCE = 0, then SOLUTION_ITERATIONS from ITER = 0. */
int oracle_solve_fixed_grid_from(const oracle_state *st_in, const float *wtmu, int maxiter, float solacc, float shacc,
                                 int accelflag, int highorderrad, int iterfixsh, int maxiv, int restore,
                                 int *shptr, float *source, int *rshptr, float *radiance, float *fluxes, float *bcrad,
                                 int *iters_out, float *solcrit_out, char *errmsg)
{
    oracle_state st = *st_in;
    const int npts = st.npts, ns = st.nstokes, maxir = maxiv + npts;
    shdo_coef *c = make_sh_do_coef(&st, wtmu);
    int *sweepord = (int *)malloc(sizeof(int) * (size_t)npts * 8);
    int *oshptr = (int *)calloc(npts + 2, sizeof(int));
    int *lofj = (int *)malloc(sizeof(int) * st.nlm);
    float *delsource = (float *)calloc((size_t)ns * maxiv, sizeof(float));
    float *work = (float *)calloc((size_t)ns * st.nphi0max * npts, sizeof(float));
    float *gridrad = (float *)calloc((size_t)ns * npts, sizeof(float));
    float deljdot = 0, deljold = 0, deljnew = 0, jnorm = 0, solcrit = 1.0f, a = 0.0f, accelpar, albmax = 0.0f;
    int iter = 0, ierr = 0, fixsh = 0, i, j, l, m, k;
    j = 0;
    for (l = 0; l <= st.ml; l++) {
        const int me = l < st.mm ? l : st.mm;
        for (m = -me; m <= me; m++) lofj[j++] = l;
    }
    for (i = 0; i < npts * st.npart; i++) if (st.albedo[i] > albmax) albmax = st.albedo[i];
    ierr = sweeping_order(&st, sweepord);
    if (ierr) { if (errmsg) snprintf(errmsg, 600, "SWEEPING_ORDER: not every grid point was reached"); goto done; }
    st.rshptr = rshptr; st.radiance = radiance; st.shptr = shptr; st.source = source;
    st.fluxes = fluxes; st.bcrad = bcrad;
    if (!restore) {
        /* first guess (see header): zero radiance, 4 terms per point; source from COMPUTE_SOURCE(FIRST=.TRUE.) */
        for (i = 0; i <= npts; i++) rshptr[i] = 4 * i;
        rshptr[npts + 1] = rshptr[npts];
        memset(radiance, 0, sizeof(float) * (size_t)ns * rshptr[npts]);
        ierr = oracle_compute_source(&st, 0, shacc, maxiv, 1, accelflag, 1, shptr, source, oshptr, delsource,
                                     &deljdot, &deljold, &deljnew, &jnorm, errmsg);
        if (ierr) goto done;
    } else {
        rshptr[npts + 1] = rshptr[npts];
        if (shptr[npts] > maxiv || rshptr[npts] > maxir) { ierr = 2; if (errmsg) snprintf(errmsg, 600, "restored solution larger than MAXIV"); goto done; }
    }
    if (accelflag) {
        memcpy(oshptr, shptr, sizeof(int) * (npts + 1));
        memset(delsource, 0, sizeof(float) * (size_t)ns * oshptr[npts]);
    }
    while (iter < maxiter && solcrit > solacc) {
        iter = iter + 1;
        ierr = radiance_truncation(&st, highorderrad, shptr, radiance, maxir, fixsh, shacc, rshptr, lofj);
        if (ierr) { if (errmsg) snprintf(errmsg, 600, "RADIANCE_TRUNCATION: out of memory"); goto done; }
        ierr = path_integration(&st, c, sweepord, shptr, source, rshptr, radiance, fluxes, bcrad, work, gridrad, errmsg);
        if (ierr) goto done;
        if (solcrit < 0.001f || iter > iterfixsh) fixsh = 1;
        ierr = oracle_compute_source(&st, fixsh, shacc, maxiv, 0, accelflag, 1, shptr, source, oshptr, delsource,
                                     &deljdot, &deljold, &deljnew, &jnorm, errmsg);
        if (ierr) goto done;
        /* CALC_ACCEL_SOLCRIT */
        if (accelflag && a == 0.0f && deljnew < deljold) {
            const float r = sqrtf(deljnew / deljold);
            const float theta = acosf(deljdot / sqrtf(deljold * deljnew));
            a = (1 - r * cosf(theta) + powf(r, 1 + 0.5f * 3.14159f / theta)) / (1 + r * r - 2 * r * cosf(theta)) - 1.0f;
            a = fminf(10.0f, fmaxf(0.0f, a));
        } else {
            a = 0.0f;
        }
        accelpar = a;
        if (jnorm > 0.0f) solcrit = sqrtf(deljnew / jnorm);
        else if (deljnew == 0.0f) solcrit = 0.0f;
        /* ACCELERATE_SOLUTION */
        if (accelpar > 0.0f) {
            for (i = 1; i <= npts; i++) {
                const int is = shptr[i - 1], nsx = shptr[i] - is, isd = oshptr[i - 1], nsd = oshptr[i] - isd;
                const int nsc = nsx < nsd ? nsx : nsd;
                for (j = 1; j <= nsc; j++)
                    for (k = 0; k < ns; k++)
                        source[k + (size_t)ns * (is + j - 1)] = source[k + (size_t)ns * (is + j - 1)]
                                                                + accelpar * delsource[k + (size_t)ns * (isd + j - 1)];
            }
        }
        if (albmax < solacc) solcrit = solacc;
    }
done:
    if (iters_out) *iters_out = iter;
    if (solcrit_out) *solcrit_out = solcrit;
    free_sh_do_coef(c); free(sweepord); free(oshptr); free(lofj); free(delsource); free(work); free(gridrad);
    return ierr;
}

/* ---- the two transforms alone, for all zenith angles (parity checks of the GPU transforms) ----
 * dofield is DOFIELD(NPTS, NSTOKES, NANG): ordinate IANG = (IMU, IPHI) in PATH_INTEGRATION order. */
int oracle_sh_to_do_all(const oracle_state *st, const float *wtmu, const int *shptr, const float *indata,
                        float *dofield)
{
    shdo_coef *c = make_sh_do_coef(st, wtmu);
    const int ns = st->nstokes, npts = st->npts, na = st->nphi0max, mm = st->mm;
    float *work = (float *)calloc((size_t)ns * na * npts, sizeof(float));
    double *aztab = (double *)malloc(sizeof(double) * (2 * mm + 1) * na);
    int imu, k, m, i, n, iang = 0;
    for (imu = 1; imu <= st->nmu; imu++) {
        const int nphi0 = st->nphi0[imu - 1];
        for (k = 1; k <= nphi0; k++)
            for (m = -mm; m <= mm; m++) aztab[(m + mm) + (2 * mm + 1) * (k - 1)] = az_basis(st, c, imu, k, m);
        sh_to_do(st, c, imu, shptr, indata, work, aztab);
        for (k = 1; k <= nphi0; k++, iang++)
            for (i = 0; i < npts; i++)
                for (n = 0; n < ns; n++)
                    dofield[i + (size_t)npts * (n + (size_t)ns * iang)] = work[n + (size_t)ns * ((k - 1) + (size_t)na * i)];
    }
    free(work); free(aztab); free_sh_do_coef(c);
    return 0;
}

int oracle_do_to_sh_all(const oracle_state *st, const float *wtmu, const int *rshptr, const float *dofield,
                        float *outdata)
{
    shdo_coef *c = make_sh_do_coef(st, wtmu);
    const int ns = st->nstokes, npts = st->npts, na = st->nphi0max, mm = st->mm;
    float *work = (float *)calloc((size_t)ns * na * npts, sizeof(float));
    double *aztab = (double *)malloc(sizeof(double) * (2 * mm + 1) * na);
    int imu, k, m, i, n, iang = 0;
    memset(outdata, 0, sizeof(float) * (size_t)ns * rshptr[npts]);
    for (imu = 1; imu <= st->nmu; imu++) {
        const int nphi0 = st->nphi0[imu - 1];
        for (k = 1; k <= nphi0; k++)
            for (m = -mm; m <= mm; m++) aztab[(m + mm) + (2 * mm + 1) * (k - 1)] = az_basis(st, c, imu, k, m);
        for (k = 1; k <= nphi0; k++, iang++)
            for (i = 0; i < npts; i++)
                for (n = 0; n < ns; n++)
                    work[n + (size_t)ns * ((k - 1) + (size_t)na * i)] = dofield[i + (size_t)npts * (n + (size_t)ns * iang)];
        do_to_sh(st, c, imu, rshptr, work, outdata, aztab);
    }
    free(work); free(aztab); free_sh_do_coef(c);
    return 0;
}

/* RADIANCE_TRUNCATION alone (parity check of the host-side restatement in at3d_b200/solver.py) */
int oracle_radiance_truncation(const oracle_state *st, int highorderrad, const int *shptr, const float *radiance,
                               int maxir, int fixsh, float shacc, int *rshptr)
{
    int *lofj = (int *)malloc(sizeof(int) * st->nlm);
    int j = 0, l, m, rc;
    for (l = 0; l <= st->ml; l++) {
        const int me = l < st->mm ? l : st->mm;
        for (m = -me; m <= me; m++) lofj[j++] = l;
    }
    rc = radiance_truncation(st, highorderrad, shptr, radiance, maxir, fixsh, shacc, rshptr, lofj);
    free(lofj);
    return rc;
}

/* One PATH_INTEGRATION alone (parity check of the GPU sweeps): radiance[nstokes, rshptr[npts]], fluxes[2,npts],
   bcrad as in oracle_solve_fixed_grid. */
int oracle_path_integration_once(const oracle_state *st_in, const float *wtmu, const int *shptr, const float *source,
                                 const int *rshptr, float *radiance, float *fluxes, float *bcrad, char *errmsg)
{
    oracle_state st = *st_in;
    const int npts = st.npts, ns = st.nstokes;
    shdo_coef *c = make_sh_do_coef(&st, wtmu);
    int *sweepord = (int *)malloc(sizeof(int) * (size_t)npts * 8);
    float *work = (float *)calloc((size_t)ns * st.nphi0max * npts, sizeof(float));
    float *gridrad = (float *)calloc((size_t)ns * npts, sizeof(float));
    int ierr = sweeping_order(&st, sweepord);
    if (ierr) { if (errmsg) snprintf(errmsg, 600, "SWEEPING_ORDER: not every grid point was reached"); }
    else {
        st.fluxes = fluxes; st.bcrad = bcrad;
        ierr = path_integration(&st, c, sweepord, shptr, source, rshptr, radiance, fluxes, bcrad, work, gridrad, errmsg);
    }
    free_sh_do_coef(c); free(sweepord); free(work); free(gridrad);
    return ierr;
}

/* SWEEPING_ORDER alone: sweepord is SWEEPORD(NPTS,8) (parity check of the product's host-side order) */
int oracle_sweeping_order(const oracle_state *st, int *sweepord)
{
    return sweeping_order(st, sweepord);
}

/* ---- INIT_SOLUTION + SOLUTION_ITERATIONS with adaptive cell splitting  shdomsub1.f:113-822 ----
 * The point arrays of `st` (extinct, albedo, planck, iphase, phaseinterpwt) have leading dimension maxig; the other
 * routines of this oracle expect leading dimension npts, so for NPART > 1 a compact copy is refreshed whenever the
 * number of points changes (for NPART = 1 the two layouts coincide). */
typedef struct {
    float *extinct, *albedo, *planck, *pwt;
    int *iphase;
} compact_view;

static void refresh_view(oracle_state *v, const oracle_adapt *a, compact_view *cv, int npart, int nq)
{
    const int npts = a->npts, ld = a->maxig;
    int ipa;
    v->npts = npts; v->ncells = a->ncells;
    if (npart == 1) {
        v->extinct = a->extinct; v->albedo = a->albedo; v->planck = a->planck;
        v->iphase = a->iphase; v->phaseinterpwt = a->phaseinterpwt;
        return;
    }
    for (ipa = 0; ipa < npart; ipa++) {
        memcpy(cv->extinct + (size_t)npts * ipa, a->extinct + (size_t)ld * ipa, sizeof(float) * npts);
        memcpy(cv->albedo + (size_t)npts * ipa, a->albedo + (size_t)ld * ipa, sizeof(float) * npts);
        if (a->planck) memcpy(cv->planck + (size_t)npts * ipa, a->planck + (size_t)ld * ipa, sizeof(float) * npts);
        memcpy(cv->iphase + (size_t)nq * npts * ipa, a->iphase + (size_t)nq * ld * ipa, sizeof(int) * (size_t)nq * npts);
        memcpy(cv->pwt + (size_t)nq * npts * ipa, a->phaseinterpwt + (size_t)nq * ld * ipa,
               sizeof(float) * (size_t)nq * npts);
    }
    v->extinct = cv->extinct; v->albedo = cv->albedo; v->planck = a->planck ? cv->planck : NULL;
    v->iphase = cv->iphase; v->phaseinterpwt = cv->pwt;
}

int oracle_solve_adaptive(oracle_state *st, oracle_prop *pg, const float *wtmu, float *temp,
                          int maxig, int maxic, int maxiv, int maxido, int maxbcrad, int nbpts, int nbcells,
                          int maxiter, float solacc, float splitacc, float shacc, int accelflag, int highorderrad,
                          int iterfixsh, int inradflag, float *extdirp, int *iters_out, float *solcrit_out,
                          float *splitcrit_out, char *errmsg)
{
    const int ns = st->nstokes, npart = st->npart, nq = 8 * st->maxnmicro, maxir = maxiv + maxig;
    const int lambertian = st->sfctype1 == 'L';
    oracle_state v = *st;
    oracle_adapt a;
    compact_view cv = {0};
    shdo_coef *c = make_sh_do_coef(st, wtmu);
    int *sweepord = (int *)malloc(sizeof(int) * (size_t)maxig * 8);
    int *oshptr = (int *)calloc(maxig + 2, sizeof(int));
    int *lofj = (int *)malloc(sizeof(int) * st->nlm);
    float *delsource = (float *)calloc((size_t)ns * maxiv, sizeof(float));
    float *work = (float *)calloc((size_t)ns * st->nphi0max * maxig, sizeof(float));
    float *gridrad = (float *)calloc((size_t)ns * maxig, sizeof(float));
    float deljdot = 0, deljold = 0, deljnew = 0, jnorm = 0, solcrit = 1.0f, acc = 0.0f, accelpar, albmax = 0.0f;
    float splitcrit = 0.0f, skyradalb;
    const float endadaptsol = 0.001f, startadaptsol = 0.1f;
    float adaptrange, cursplitacc, startsplitacc, avgsolcrit, beta;
    int iter = 0, ierr = 0, fixsh = 0, splittesting = 1, outofmem = 0, oldnpts = 0, i, j, l, m, k, imu, iphi;
    if (st->sfctype0 == 'V' && splitacc > 0.0f) {
        if (errmsg) snprintf(errmsg, 600, "oracle adaptive solve: SURFACE_PARM_INTERP for new bottom points is not restated");
        ierr = 3; goto done;
    }
    if (npart > 1) {
        cv.extinct = (float *)malloc(sizeof(float) * (size_t)maxig * npart);
        cv.albedo = (float *)malloc(sizeof(float) * (size_t)maxig * npart);
        cv.planck = (float *)malloc(sizeof(float) * (size_t)maxig * npart);
        cv.iphase = (int *)malloc(sizeof(int) * (size_t)nq * maxig * npart);
        cv.pwt = (float *)malloc(sizeof(float) * (size_t)nq * maxig * npart);
    }
    j = 0;
    for (l = 0; l <= st->ml; l++) {
        const int me = l < st->mm ? l : st->mm;
        for (m = -me; m <= me; m++) lofj[j++] = l;
    }
    memset(&a, 0, sizeof(a));
    a.maxig = maxig; a.maxic = maxic; a.maxiv = maxiv; a.maxido = maxido;
    a.npts = st->npts; a.ncells = st->ncells; a.accelflag = accelflag;
    a.gridptr = (int *)st->gridptr; a.neighptr = (int *)st->neighptr; a.treeptr = (int *)st->treeptr;
    a.cellflags = (short *)st->cellflags; a.gridpos = (float *)st->gridpos;
    a.temp = temp; a.planck = (float *)st->planck; a.extinct = (float *)st->extinct; a.albedo = (float *)st->albedo;
    a.total_ext = (float *)st->total_ext; a.dirflux = (float *)st->dirflux;
    a.iphase = (int *)st->iphase; a.phaseinterpwt = (float *)st->phaseinterpwt;
    a.shptr = (int *)st->shptr; a.rshptr = (int *)st->rshptr; a.oshptr = oshptr;
    a.source = (float *)st->source; a.radiance = (float *)st->radiance;
    a.pg = pg; a.extdirp = extdirp;
    oracle_prop_extmin(pg, &a.extmin, &a.scatmin);
    /* ALBMAX of TRANSFER_PA_TO_GRID */
    for (k = 0; k < npart; k++)
        for (i = 0; i < a.npts; i++) if (a.albedo[i + (size_t)maxig * k] > albmax) albmax = a.albedo[i + (size_t)maxig * k];
    /* ---- INIT_SOLUTION ---- */
    if (st->srctype != 'T') {
        ierr = oracle_make_direct(a.npts, st->bcflag, st->ipflag, st->deltam, st->ml, st->nstleg, pg->nlegp,
                                  st->solarflux, st->solarmu, st->solaraz, a.gridpos, pg->npx, pg->npy, pg->npz,
                                  pg->delx, pg->dely, pg->xstart, pg->ystart, pg->zlevels, pg->extinctp, pg->albedop,
                                  pg->legenp, pg->iphasep, pg->phasewtp, pg->maxnmicro, npart, pg->nzckd, pg->zckd,
                                  pg->gasabs, extdirp, a.dirflux, a.beam_d, a.beam_i, errmsg);
        if (ierr) goto done;
    }
    refresh_view(&v, &a, &cv, npart, nq);
    v.rshptr = a.rshptr; v.radiance = a.radiance; v.shptr = a.shptr; v.source = a.source;
    if (inradflag) {
        const int nx1ny1 = nbpts / st->nz;
        skyradalb = 0.0f;
        for (imu = 1; imu <= st->nmu / 2; imu++)
            for (iphi = 1; iphi <= st->nphi0[imu - 1]; iphi++)
                skyradalb = skyradalb + fabsf(st->mu[imu - 1]) * st->wtdo[(imu - 1) + st->nmu * (iphi - 1)]
                            * st->skyrad[ns * ((imu - 1) + (size_t)(st->nmu / 2) * (iphi - 1))];
        ierr = oracle_init_radiance(st, maxig, nx1ny1, st->nz, a.extinct, a.albedo, a.total_ext, temp, a.iphase,
                                    a.phaseinterpwt, skyradalb, 0.0f, a.rshptr, a.radiance);
        if (ierr) { if (errmsg) snprintf(errmsg, 600, "INIT_RADIANCE/EDDRTF failed (%d)", ierr); goto done; }
        ierr = oracle_interp_radiance(ns, nbpts, nbcells, a.ncells, a.treeptr, a.gridptr, a.gridpos, a.rshptr, a.radiance);
        if (ierr) { if (errmsg) snprintf(errmsg, 600, "INTERP_RADIANCE: point not on an edge"); goto done; }
    } else {
        for (i = 0; i <= a.npts; i++) a.rshptr[i] = 4 * i;
        memset(a.radiance, 0, sizeof(float) * (size_t)ns * a.rshptr[a.npts]);
    }
    a.rshptr[a.npts + 1] = a.rshptr[a.npts];
    ierr = oracle_compute_source(&v, 0, shacc, maxiv, 1, accelflag, 1, a.shptr, a.source, oshptr, delsource,
                                 &deljdot, &deljold, &deljnew, &jnorm, errmsg);
    if (ierr) goto done;
    if (accelflag) {
        memcpy(oshptr, a.shptr, sizeof(int) * (a.npts + 1));
        memset(delsource, 0, sizeof(float) * (size_t)ns * oshptr[a.npts]);
    }
    /* ---- SOLUTION_ITERATIONS ---- */
    adaptrange = startadaptsol / (3.0f * endadaptsol);
    cursplitacc = splitacc * adaptrange;
    startsplitacc = cursplitacc;
    avgsolcrit = solcrit;
    while (iter < maxiter && (solcrit > solacc || (splitcrit > splitacc && cursplitacc > splitacc && !outofmem))) {
        iter = iter + 1;
        if (splitacc > 0.0f) {
            int dosplit;
            avgsolcrit = sqrtf(avgsolcrit * solcrit);
            dosplit = solcrit <= startadaptsol && (solcrit > endadaptsol || cursplitacc > splitacc) && !outofmem;
            beta = logf(startsplitacc / splitacc) / logf(adaptrange);
            cursplitacc = fminf(cursplitacc, fmaxf(splitacc, splitacc * powf(avgsolcrit / (3.0f * endadaptsol), beta)));
            if (solcrit <= endadaptsol) cursplitacc = splitacc;
            if (splittesting) {
                ierr = oracle_split_grid(&a, st, dosplit, &outofmem, cursplitacc, &splitcrit, errmsg);
                if (ierr) goto done;
                if (solcrit > startadaptsol) startsplitacc = splitcrit;
            }
            if (solcrit <= endadaptsol) splittesting = 0;
        }
        refresh_view(&v, &a, &cv, npart, nq);
        ierr = radiance_truncation(&v, highorderrad, a.shptr, a.radiance, maxir, fixsh, shacc, a.rshptr, lofj);
        if (ierr) { if (errmsg) snprintf(errmsg, 600, "RADIANCE_TRUNCATION: out of memory"); goto done; }
        if (a.npts != oldnpts) {
            ierr = sweeping_order(&v, sweepord);
            if (ierr) { if (errmsg) snprintf(errmsg, 600, "SWEEPING_ORDER: not every grid point was reached"); goto done; }
            if (oracle_boundary_pnts(a.npts, st->nang, lambertian, st->maxnbc, maxbcrad, st->zgrid[0],
                                     st->zgrid[st->nz - 1], a.gridpos, &v.ntoppts, &v.nbotpts, (int *)st->bcptr)) {
                if (errmsg) snprintf(errmsg, 600, "BOUNDARY_PNTS: MAXNBC exceeded");
                ierr = 1; goto done;
            }
        }
        ierr = path_integration(&v, c, sweepord, a.shptr, a.source, a.rshptr, a.radiance, (float *)st->fluxes,
                                st->bcrad, work, gridrad, errmsg);
        if (ierr) goto done;
        oldnpts = a.npts;
        if (solcrit < endadaptsol || iter > iterfixsh) fixsh = 1;
        ierr = oracle_compute_source(&v, fixsh, shacc, maxiv, 0, accelflag, 1, a.shptr, a.source, oshptr, delsource,
                                     &deljdot, &deljold, &deljnew, &jnorm, errmsg);
        if (ierr) goto done;
        if (accelflag && acc == 0.0f && deljnew < deljold) {
            const float r = sqrtf(deljnew / deljold);
            const float theta = acosf(deljdot / sqrtf(deljold * deljnew));
            acc = (1 - r * cosf(theta) + powf(r, 1 + 0.5f * 3.14159f / theta)) / (1 + r * r - 2 * r * cosf(theta)) - 1.0f;
            acc = fminf(10.0f, fmaxf(0.0f, acc));
        } else {
            acc = 0.0f;
        }
        accelpar = acc;
        if (jnorm > 0.0f) solcrit = sqrtf(deljnew / jnorm);
        else if (deljnew == 0.0f) solcrit = 0.0f;
        if (accelpar > 0.0f) {
            for (i = 1; i <= a.npts; i++) {
                const int is = a.shptr[i - 1], nsx = a.shptr[i] - is, isd = oshptr[i - 1], nsd = oshptr[i] - isd;
                const int nsc = nsx < nsd ? nsx : nsd;
                for (j = 1; j <= nsc; j++)
                    for (k = 0; k < ns; k++)
                        a.source[k + (size_t)ns * (is + j - 1)] = a.source[k + (size_t)ns * (is + j - 1)]
                                                                  + accelpar * delsource[k + (size_t)ns * (isd + j - 1)];
            }
        }
        if (albmax < solacc) solcrit = solacc;
        if (getenv("ORACLE_VERBOSE"))
            fprintf(stderr, "  %4d %8.3f %10.3E %8d %8.2f %6.3f\n", iter, log10f(fmaxf(solcrit, 1.0e-20f)), splitcrit,
                    a.npts, (float)a.shptr[a.npts] / a.npts, (float)a.shptr[a.npts] / (a.npts * (float)st->nlm));
    }
    st->npts = a.npts; st->ncells = a.ncells; st->ntoppts = v.ntoppts; st->nbotpts = v.nbotpts;
done:
    if (iters_out) *iters_out = iter;
    if (solcrit_out) *solcrit_out = solcrit;
    if (splitcrit_out) *splitcrit_out = splitcrit;
    free_sh_do_coef(c); free(sweepord); free(oshptr); free(lofj); free(delsource); free(work); free(gridrad);
    free(cv.extinct); free(cv.albedo); free(cv.planck); free(cv.iphase); free(cv.pwt);
    return ierr;
}
