/*
 * oracle_ray.c -- TEST INFRASTRUCTURE (see shdom_oracle.h).
 * Radiance ray integration through the adaptive grid:
 *   LOCATE_GRID_CELL   /root/reference/src/polarized/shdomsub2.f:4043-4206
 *   NEXT_CELL          /root/reference/src/polarized/shdomsub1.f:4470-4522
 *   COMPUTE_SOURCE_1CELL[_UNPOL]  shdomsub2.f:2868-3192
 *   ROTATE_POL_PLANE   shdomsub2.f:3277-3314
 *   INTEGRATE_1RAY     shdomsub2.f:2311-2743
 *   FIND_BOUNDARY_RADIANCE  shdomsub2.f:2748-2863 (Lambertian surfaces; BRDF surfaces -> IERR)
 *   COMPUTE_TOP_RADIANCES, FIXED/VARIABLE_LAMBERTIAN_BOUNDARY  shdomsub1.f:2336-2529
 *   RENDER             shdomsub4.f:93-286
 * Solar source (SRCTYPE='S') only; thermal sources return IERR=3 (out of scope).
 * float/double usage mirrors the Fortran declarations expression by expression.
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

/* ------------------------------------------------------------------ */
/* LOCATE_GRID_CELL  shdomsub2.f:4043-4206                             */
/* ------------------------------------------------------------------ */
int oracle_locate_grid_cell(const oracle_state *st, double *x0p, double *y0p, double *z0p)
{
    const int nx = st->nx, ny = st->ny, nz = st->nz;
    const int bcflag = st->bcflag, ipflag = st->ipflag;
    const float *xgrid = st->xgrid, *ygrid = st->ygrid, *zgrid = st->zgrid;
    double x0 = *x0p, y0 = *y0p, z0 = *z0p;
    int il, iu, im, ix, iy, iz, nxc, nyc, ic, iptr, dir, icell;
#define XG(i) xgrid[(i) - 1]
#define YG(i) ygrid[(i) - 1]
#define ZG(i) zgrid[(i) - 1]
    if (!(BTEST(bcflag, 0) || BTEST(bcflag, 2))) {
        double xdomain = XG(nx + 1) - XG(1);   /* REAL subtraction, widened */
        if (x0 < XG(1))
            x0 = x0 - xdomain * ((int)((x0 - XG(1)) / xdomain) - 1);
        else if (x0 > XG(nx + 1))
            x0 = x0 - xdomain * (int)((x0 - XG(1)) / xdomain);
    }
    if (!(BTEST(bcflag, 1) || BTEST(bcflag, 3))) {
        double ydomain = YG(ny + 1) - YG(1);
        if (y0 < YG(1))
            y0 = y0 - ydomain * ((int)((y0 - YG(1)) / ydomain) - 1);
        else if (y0 > YG(ny + 1))
            y0 = y0 - ydomain * (int)((y0 - YG(1)) / ydomain);
    }
    il = 0;
    if (BTEST(ipflag, 0)) {
        iu = nx;
        while (iu - il > 1) {
            im = (iu + il) / 2;
            if (x0 >= 0.5f * (XG(im) + XG(im + 1))) il = im; else iu = im;
        }
        il = il + 1;
    } else {
        iu = nx + 1;
        while (iu - il > 1) {
            im = (iu + il) / 2;
            if (x0 >= XG(im)) il = im; else iu = im;
        }
    }
    ix = il > 1 ? il : 1;
    il = 0;
    if (BTEST(ipflag, 1)) {
        iu = ny;
        while (iu - il > 1) {
            im = (iu + il) / 2;
            if (y0 >= 0.5f * (YG(im) + YG(im + 1))) il = im; else iu = im;
        }
        il = il + 1;
    } else {
        iu = ny + 1;
        while (iu - il > 1) {
            im = (iu + il) / 2;
            if (y0 >= YG(im)) il = im; else iu = im;
        }
    }
    iy = il > 1 ? il : 1;
    il = 0;
    iu = nz;
    while (iu - il > 1) {
        im = (iu + il) / 2;
        if (z0 >= ZG(im)) il = im; else iu = im;
    }
    iz = il > 1 ? il : 1;

    nxc = nx;
    if (BTEST(bcflag, 0)) {
        nxc = nx + 1;
        if (x0 < XG(1)) ix = 1;
        else if (x0 > XG(nx)) ix = nx + 1;
        else ix = ix + 1;
    }
    if (BTEST(bcflag, 2) && !BTEST(ipflag, 0)) {
        nxc = nx - 1;
        ix = ix < nxc ? ix : nxc;
    }
    nyc = ny;
    if (BTEST(bcflag, 1)) {
        nyc = ny + 1;
        if (y0 < YG(1)) iy = 1;
        else if (y0 > YG(ny)) iy = ny + 1;
        else iy = iy + 1;
    }
    if (BTEST(bcflag, 3) && !BTEST(ipflag, 1)) {
        nyc = ny - 1;
        iy = iy < nyc ? iy : nyc;
    }
    (void)nxc;
    icell = iz + (nz - 1) * (iy - 1) + (nz - 1) * nyc * (ix - 1);

    while (TREEPTR(st, 2, icell) > 0) {
        dir = IBITS2(CELLFLAGS(st, icell));
        ic = TREEPTR(st, 2, icell) + 1;
        iptr = GRIDPTR(st, 1, ic);
        if (dir == 1) { if (x0 < GRIDPOS(st, 1, iptr)) ic = ic - 1; }
        else if (dir == 2) { if (y0 < GRIDPOS(st, 2, iptr)) ic = ic - 1; }
        else if (dir == 3) { if (z0 < GRIDPOS(st, 3, iptr)) ic = ic - 1; }
        icell = ic;
    }
#undef XG
#undef YG
#undef ZG
    *x0p = x0; *y0p = y0; *z0p = z0;
    return icell;
}

/* NEXT_CELL  shdomsub1.f:4470-4522 */
int oracle_next_cell(const oracle_state *st, double xe, double ye, double ze,
                     int iface, int jface, int icell)
{
    int inext = NEIGHPTR(st, iface, icell);
    if (inext < 0) {
        int ic = -inext;
        while (TREEPTR(st, 2, ic) > 0) {
            int dir = IBITS2(CELLFLAGS(st, ic));
            int ic1 = TREEPTR(st, 2, ic);
            if (dir == jface) {
                ic = ic1 + 1 - ((iface - 1) % 2);
            } else {
                ic = ic1;
                if (dir == 1) { if (xe > GRIDPOS(st, 1, GRIDPTR(st, 8, ic1))) ic = ic + 1; }
                else if (dir == 2) { if (ye > GRIDPOS(st, 2, GRIDPTR(st, 8, ic1))) ic = ic + 1; }
                else { if (ze > GRIDPOS(st, 3, GRIDPTR(st, 8, ic1))) ic = ic + 1; }
            }
        }
        inext = ic;
    }
    return inext;
}

/* ROTATE_POL_PLANE  shdomsub2.f:3277-3314 */
void oracle_rotate_pol_plane(int nstokes, double cosscat, float solarmu, float mu,
                             float delphi, float *scatvect)
{
    if (nstokes > 1) {
        float b1 = scatvect[1];
        double sin_scat = sqrt(fmax(0.0, 1.0 - cosscat * cosscat));
        double sin_theta1 = sqrt(1.0 - (double)(solarmu * solarmu));
        double sin_theta2 = sqrt(1.0 - (double)(mu * mu));
        double sinphi = sin((double)delphi);
        double cosphi = cos((double)delphi);
        double sin2, cos2, sin22, cos22;
        if (sin_scat == 0.0) {
            sin2 = 0.0;
            cos2 = -1.0;
        } else {
            sin2 = sin_theta1 * sinphi / sin_scat;
            cos2 = (sin_theta2 * solarmu - sin_theta1 * mu * cosphi) / sin_scat;
        }
        sin22 = 2.0 * sin2 * cos2;
        cos22 = 1.0 - 2.0 * (sin2 * sin2);
        scatvect[1] = (float)(b1 * cos22);
        scatvect[2] = (float)(b1 * sin22);
    }
    if (nstokes == 4) scatvect[3] = 0.0f;
}

/* per-ray direction setup shared by INTEGRATE_1RAY and ADJOINT_INTEGRATE_1RAY
 * (shdomsub2.f:2407-2481 == shdomsub4.f:3409-3490) */
void oracle_ray_setup(const oracle_state *st, double mu2, double phi2, ray_dir *rd,
                      float *ylmdir, float *singscat,
                      const float *dphasetab, int dnumphase, float *dsingscat)
{
    const int nstokes = st->nstokes, nstphase = st->nstphase, numphase = st->numphase;
    const int nscatangle = st->nscatangle;
    double pi = acos(-1.0);
    int i, k;
    oracle_ylmall(0, (float)mu2, (float)phi2, st->ml, st->mm, st->nstleg, ylmdir);
    rd->cosscat = 0.0;
    if (st->srctype != 'T' && st->deltam) {
        double cosscat = st->solarmu * mu2
            + sqrt((1.0f - st->solarmu * st->solarmu) * (1.0 - mu2 * mu2))
              * cos(st->solaraz - phi2);
        float f;
        int j;
        cosscat = fmax(fmin(1.0, cosscat), -1.0);
        rd->cosscat = cosscat;
        f = (float)((nscatangle - 1) * (acos(cosscat) / pi) + 1);
        j = (int)f;
        if (j > nscatangle - 1) j = nscatangle - 1;
        f = f - (float)j;
        for (i = 1; i <= numphase; i++) {
            for (k = 1; k <= nstphase; k++)
                singscat[(k - 1) + nstokes * (i - 1)] =
                    (1 - f) * st->phasetab[(k - 1) + nstphase * ((i - 1) + (size_t)numphase * (j - 1))]
                    + f * st->phasetab[(k - 1) + nstphase * ((i - 1) + (size_t)numphase * j)];
            if (nstokes > 1)
                oracle_rotate_pol_plane(nstokes, cosscat, st->solarmu, (float)mu2,
                                        st->solaraz - (float)phi2, &singscat[nstokes * (i - 1)]);
        }
        if (dsingscat) {
            for (i = 1; i <= dnumphase; i++) {
                for (k = 1; k <= nstphase; k++)
                    dsingscat[(k - 1) + nstokes * (i - 1)] =
                        (1 - f) * dphasetab[(k - 1) + nstphase * ((i - 1) + (size_t)dnumphase * (j - 1))]
                        + f * dphasetab[(k - 1) + nstphase * ((i - 1) + (size_t)dnumphase * j)];
                if (nstokes > 1)
                    oracle_rotate_pol_plane(nstokes, cosscat, st->solarmu, (float)mu2,
                                            st->solaraz - (float)phi2, &dsingscat[nstokes * (i - 1)]);
            }
        }
    }
    rd->cx = sqrt(1.0 - mu2 * mu2) * cos(phi2 - pi);
    rd->cy = sqrt(1.0 - mu2 * mu2) * sin(phi2 - pi);
    rd->cz = -mu2;
    if (fabs(rd->cx) > 1.0e-6f) rd->cxinv = 1.0 / rd->cx; else { rd->cx = 0.0; rd->cxinv = 1.0e6f; }
    if (fabs(rd->cy) > 1.0e-6f) rd->cyinv = 1.0 / rd->cy; else { rd->cy = 0.0; rd->cyinv = 1.0e6f; }
    if (fabs(rd->cz) > 1.0e-6f) rd->czinv = 1.0 / rd->cz; else { rd->cz = 0.0; rd->czinv = 1.0e6f; }
    rd->bitx = rd->cx < 0.0 ? 1 : 0;
    rd->bity = rd->cy < 0.0 ? 1 : 0;
    rd->bitz = rd->cz < 0.0 ? 1 : 0;
    rd->ioct = 1 + rd->bitx + 2 * rd->bity + 4 * rd->bitz;
    rd->xm = 0.5f * (st->xgrid[0] + st->xgrid[st->nx - 1]);
    rd->ym = 0.5f * (st->ygrid[0] + st->ygrid[st->ny - 1]);
}

/* ------------------------------------------------------------------ */
/* COMPUTE_SOURCE_1CELL[_UNPOL]  shdomsub2.f:2868-3192                 */
/* ------------------------------------------------------------------ */
static void compute_source_1cell(const oracle_state *st, int icell,
                                 const float *ylmdir, const float *singscat,
                                 const int *donethis, int *oldipts,
                                 const float *oextinct8, const float *osrcext8,
                                 float *extinct8, float *srcext8, int singlescatter,
                                 float *legent /* scratch [nstleg*(nleg+1)] */)
{
    const int nstokes = st->nstokes, nstleg = st->nstleg, ml = st->ml, mm = st->mm;
    const int nleg = st->nleg, npart = st->npart, npts = st->npts;
    const int nq = 8 * st->maxnmicro;
    const int nlt = nstleg * (nleg + 1);
    float secmu0 = (float)(1.0 / fabs((double)st->solarmu));
    int n, k, j, l, m, q, ipa, t;
#define SRC8(kk, nn) srcext8[((kk) - 1) + nstokes * ((nn) - 1)]
#define OSRC8(kk, nn) osrcext8[((kk) - 1) + nstokes * ((nn) - 1)]
#define YD(i, jj) ylmdir[((i) - 1) + nstleg * ((jj) - 1)]
#define YS(i, jj) st->ylmsun[((i) - 1) + nstleg * ((jj) - 1)]
#define LT(i, ll) legent[((i) - 1) + nstleg * (ll)]
    for (n = 1; n <= 8; n++) {
        int ip = GRIDPTR(st, n, icell);
        int i = donethis[n - 1];
        if (i > 0 && ip == oldipts[n - 1]) {
            extinct8[n - 1] = oextinct8[i - 1];
            for (k = 1; k <= nstokes; k++) SRC8(k, n) = OSRC8(k, i);
        } else if (i < 0) {
            extinct8[n - 1] = extinct8[-i - 1];
            for (k = 1; k <= nstokes; k++) SRC8(k, n) = SRC8(k, -i);
        } else {
            float ext = st->total_ext[ip - 1];
            int is, ns;
            oldipts[n - 1] = ip;
            is = st->shptr[ip - 1];
            ns = st->shptr[ip] - is;
            for (k = 1; k <= nstokes; k++) SRC8(k, n) = 0.0f;
            if (!singlescatter) {
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
                if (nstokes == 4) {
                    for (j = 1; j <= ns; j++)
                        SRC8(4, n) = SRC8(4, n) + SOURCE(st, 4, is + j) * YD(4, j);
                }
            }
            if (st->srctype != 'T' && st->deltam) {
                for (ipa = 1; ipa <= npart; ipa++) {
                    float w, f = 0.0f, da;
                    const int *iph = &st->iphase[(size_t)nq * ((ip - 1) + (size_t)npts * (ipa - 1))];
                    const float *pw = &st->phaseinterpwt[(size_t)nq * ((ip - 1) + (size_t)npts * (ipa - 1))];
                    if (ext == 0.0f) w = 1.0f;
                    else w = st->extinct[(ip - 1) + (size_t)npts * (ipa - 1)] / ext;
                    if (w == 0.0f) continue;
                    if (!st->interp_new) {
                        const float *lg = &st->legen[(size_t)nlt * (iph[0] - 1)];
                        for (t = 0; t < nlt; t++) legent[t] = lg[t];
                        f = LT(1, ml + 1);
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
                        f = LT(1, ml + 1);
                        for (t = 0; t < nlt; t++) legent[t] = legent[t] / (1 - f);
                    }
                    da = st->albedo[(ip - 1) + (size_t)npts * (ipa - 1)] * st->dirflux[ip - 1] * secmu0 * w;
                    j = 1;
                    if (!singlescatter) {
                        for (l = 0; l <= ml; l++) {
                            int me = l < mm ? l : mm;
                            int ms = -me;
                            float a1 = da * LT(1, l);
                            float b1 = nstleg > 1 ? da * LT(5, l) : 0.0f;
                            if (j <= ns) {
                                int jt = j;
                                for (m = ms; m <= me; m++) {
                                    SRC8(1, n) = SRC8(1, n) - a1 * YD(1, j) * YS(1, j);
                                    j = j + 1;
                                }
                                if (nstokes > 1) {
                                    j = jt;
                                    for (m = ms; m <= me; m++) {
                                        SRC8(2, n) = SRC8(2, n) - b1 * YD(2, j) * YS(1, j);
                                        SRC8(3, n) = SRC8(3, n) - b1 * YD(6, j) * YS(1, j);
                                        j = j + 1;
                                    }
                                }
                            }
                        }
                    }
                    /* NUMPHASE > 0 always (NUMPHASE=0 STOPs in the reference) */
                    if (pw[0] >= st->phasemax) {
                        for (k = 1; k <= nstokes; k++)
                            SRC8(k, n) = SRC8(k, n)
                                + da * singscat[(k - 1) + nstokes * (iph[0] - 1)] / (1 - f);
                    } else {
                        for (q = 0; q < nq; q++) {
                            if (pw[q] <= 1e-5f) continue;
                            for (k = 1; k <= nstokes; k++)
                                SRC8(k, n) = SRC8(k, n)
                                    + da * singscat[(k - 1) + nstokes * (iph[q] - 1)] * pw[q] / (1 - f);
                        }
                    }
                }
            }
            for (k = 1; k <= nstokes; k++) SRC8(k, n) = SRC8(k, n) * ext;
            extinct8[n - 1] = ext;
        }
    }
#undef SRC8
#undef OSRC8
#undef YD
#undef YS
#undef LT
}

/* ------------------------------------------------------------------ */
/* Boundary radiance (Lambertian)  shdomsub2.f:2748-2863               */
/* ------------------------------------------------------------------ */
int oracle_bc_search(const int *bcptr_col, int n, int ip)
{   /* binary search of shdomsub2.f:2791-2804; returns 1-based IBC or 0 (STOP in reference) */
    int il = 1, iu = n, im, ibc;
    while (iu - il > 1) {
        im = (iu + il) / 2;
        if (ip >= bcptr_col[im - 1]) il = im; else iu = im;
    }
    ibc = il;
    if (bcptr_col[ibc - 1] != ip) ibc = iu;
    if (bcptr_col[ibc - 1] != ip) return 0;
    return ibc;
}

static const int GRIDFACE[6][4] = {{1,3,5,7},{2,4,6,8},{1,2,5,6},{3,4,7,8},{1,2,3,4},{5,6,7,8}};

/* COMPUTE_TOP_RADIANCES (shdomsub1.f:2336-2433) for one boundary point.  flag 1: inverse-distance-cubed
 * interpolation over the downward ordinates, flag 2: over the upward ordinates (the "surface emission hack"
 * of FIND_BOUNDARY_RADIANCE), else the ordinate (imu,iphi).  skyrad is SKYRAD(NSTOKES,NMU/2,NPHI0MAX).
 * Only the first Stokes component is interpolated (SKYRAD3(1:1) = WEIGHTEDSUM/WEIGHTSUM). */
void oracle_compute_top_radiances(const oracle_state *st, const float *skyrad, int imu, int iphi,
                                  float mu, float phi, int flag, float *out)
{
    const int ns = st->nstokes, nh = st->nmu / 2;
    float skyrad3[4] = {0, 0, 0, 0};
    int i, j, k;
    if (flag == 1 || flag == 2) {
        const double power = 3.0;
        double weightedsum = 0.0, weightsum = 0.0, weight, distance;
        const int i0 = (flag == 1) ? 1 : nh + 1, i1 = (flag == 1) ? nh : st->nmu;
        for (i = i0; i <= i1; i++) {
            for (j = 1; j <= st->nphi0[i - 1]; j++) {
                const float mus = st->mu[i - 1];
                const float phis = st->phi[(i - 1) + st->nmu * (j - 1)];
                const int is = (flag == 1) ? i : i - nh;
                distance = (double)acosf(mu * mus + sqrtf((1.0f - mu * mu) * (1.0f - mus * mus)) * cosf(phi - phis));
                if (fabs(distance) < 1e-6f) weight = 1.0e8;
                else weight = 1.0 / pow(distance, power);
                weightedsum = weightedsum + skyrad[0 + ns * ((is - 1) + nh * (j - 1))] * weight;
                weightsum = weightsum + weight;
            }
        }
        skyrad3[0] = (float)(weightedsum / weightsum);
    } else {
        for (k = 0; k < ns; k++) skyrad3[k] = skyrad[k + ns * ((imu - 1) + nh * (iphi - 1))];
    }
    if (st->srctype == 'T') {
        const float wn[2] = {st->waveno0, st->waveno1};
        for (k = 0; k < ns; k++) out[k] = 0.0f;
        out[0] = oracle_planck_function(skyrad3[0], st->units, wn, st->wavelen);
    } else {
        for (k = 0; k < ns; k++) out[k] = skyrad3[k];
    }
}

/* FIND_BOUNDARY_RADIANCE  shdomsub2.f:2748-2863.  bcrad is the caller's private, mutable BCRAD. */
static int find_boundary_radiance(const oracle_state *st, float *bcrad, double xb, double yb,
                                  float mu2, float phi2, int icell, int kface, float *radbnd,
                                  char *errmsg)
{
    const int nstokes = st->nstokes;
    const int lambertian = st->sfctype1 == 'L';
    float x[4], y[4], rad[4][4], u, v;
    int j, k;
    for (j = 0; j < 4; j++) {
        int ip = GRIDPTR(st, GRIDFACE[kface - 1][j], icell);
        int ibc;
        x[j] = GRIDPOS(st, 1, ip);
        y[j] = GRIDPOS(st, 2, ip);
        if (mu2 < 0.0f) {
            ibc = oracle_bc_search(st->bcptr, st->ntoppts, ip);
            if (!ibc) { if (errmsg) snprintf(errmsg, 600, "FIND_BOUNDARY_RADIANCE: Not at boundary"); return 1; }
            for (k = 0; k < nstokes; k++) rad[j][k] = bcrad[k + nstokes * (ibc - 1)];
        } else {
            ibc = oracle_bc_search(st->bcptr + st->maxnbc, st->nbotpts, ip);
            if (!ibc) { if (errmsg) snprintf(errmsg, 600, "FIND_BOUNDARY_RADIANCE: Not at boundary"); return 1; }
            if (!lambertian) {
                if (oracle_variable_brdf_surface(st, ibc, ibc, mu2, phi2, bcrad + (size_t)nstokes * st->ntoppts)) {
                    if (errmsg) snprintf(errmsg, 600, "SURFACE_BRDF: Unknown BRDF type / polarized call of a scalar BRDF");
                    return 1;
                }
            }
            /* surface emission "hack" (shdomsub2.f:2832-2846): COMPUTE_TOP_RADIANCES, flag 2, on
             * SFCGRIDRAD(2:,IBC); identically zero while SFCGRIDRAD is (solar problems). */
            for (k = 0; k < nstokes; k++) rad[j][k] = 0.0f;
            if (st->sfcgridrad) {
                float tmp[4 * 64 * 128];
                const int nh = st->nmu / 2;
                int i, ja, iang = 1;
                if ((size_t)nstokes * nh * st->nphi0max > sizeof(tmp) / sizeof(float)) {
                    if (errmsg) snprintf(errmsg, 600, "oracle: ordinate set too large for the emission scratch");
                    return 3;
                }
                memset(tmp, 0, sizeof(float) * (size_t)nstokes * nh * st->nphi0max);
                for (i = 1; i <= nh; i++)
                    for (ja = 1; ja <= st->nphi0[i - 1]; ja++) {
                        tmp[0 + nstokes * ((i - 1) + nh * (ja - 1))] =
                            st->sfcgridrad[iang + (size_t)(st->nang / 2 + 1) * (ibc - 1)];
                        iang = iang + 1;
                    }
                oracle_compute_top_radiances(st, tmp, -1, -1, mu2, phi2, 2, rad[j]);
            } else if (st->srctype == 'T') {
                const float wn[2] = {st->waveno0, st->waveno1};
                rad[j][0] = oracle_planck_function(0.0f, st->units, wn, st->wavelen);
            }
            for (k = 0; k < nstokes; k++)
                rad[j][k] = rad[j][k] + bcrad[k + nstokes * (st->ntoppts + ibc - 1)];
        }
    }
    if (x[1] - x[0] > 0.0f) u = (float)((xb - x[0]) / (x[1] - x[0])); else u = 0.0f;
    if (y[2] - y[0] > 0.0f) v = (float)((yb - y[0]) / (y[2] - y[0])); else v = 0.0f;
    for (k = 0; k < nstokes; k++)
        radbnd[k] = (1 - u) * (1 - v) * rad[0][k] + u * (1 - v) * rad[1][k]
                    + (1 - u) * v * rad[2][k] + u * v * rad[3][k];
    return 0;
}

/* COMPUTE_TOP_RADIANCES with INTERPOLATE_FLAG=1 (shdomsub1.f:2375-2395): the (I only) sky radiance RENDER
 * writes into BCRAD(:,1:NTOPPTS) for an upward-looking ray */
float oracle_sky_radiance(const oracle_state *st, float mu, float phi)
{
    float out[4];
    oracle_compute_top_radiances(st, st->skyrad, 1, 1, mu, phi, 1, out);
    return out[0];
}

/* FIXED / VARIABLE_LAMBERTIAN_BOUNDARY  shdomsub1.f:2438-2529 */
void oracle_lambertian_boundary(const oracle_state *st, float *bcrad)
{
    const int nstokes = st->nstokes;
    const float wn[2] = {st->waveno0, st->waveno1};
    int ibc, k;
    if (st->sfctype0 == 'F' && st->sfctype1 == 'L') {
        float alb = st->gndalbedo / acosf(-1.0f), gndrad = 0.0f;
        if (st->srctype == 'T' || st->srctype == 'B') {
            gndrad = oracle_planck_function(st->gndtemp, st->units, wn, st->wavelen);
            gndrad = gndrad * (1.0f - st->gndalbedo);
        }
        for (ibc = 1; ibc <= st->nbotpts; ibc++) {
            int i = st->bcptr[st->maxnbc + ibc - 1];
            float *b = &bcrad[nstokes * (st->ntoppts + ibc - 1)];
            if (st->srctype == 'S') b[0] = alb * (st->dirflux[i - 1] + st->fluxes[0 + 2 * (i - 1)]);
            else if (st->srctype == 'T') b[0] = gndrad + alb * st->fluxes[0 + 2 * (i - 1)];
            else if (st->srctype == 'B') b[0] = alb * (st->dirflux[i - 1] + st->fluxes[0 + 2 * (i - 1)]) + gndrad;
            for (k = 1; k < nstokes; k++) b[k] = 0.0f;
        }
    } else if (st->sfctype0 == 'V' && st->sfctype1 == 'L') {
        float opi = 1.0f / acosf(-1.0f);
        for (ibc = 1; ibc <= st->nbotpts; ibc++) {
            int i = st->bcptr[st->maxnbc + ibc - 1];
            float alb = st->sfcgridparms[1 + st->nsfcpar * (ibc - 1)];
            float *b = &bcrad[nstokes * (st->ntoppts + ibc - 1)];
            if (st->srctype == 'S') {
                b[0] = opi * alb * (st->dirflux[i - 1] + st->fluxes[0 + 2 * (i - 1)]);
            } else {
                float gndrad = st->sfcgridparms[0 + st->nsfcpar * (ibc - 1)] * (1 - alb);
                if (st->srctype == 'T') b[0] = gndrad + opi * alb * st->fluxes[0 + 2 * (i - 1)];
                else if (st->srctype == 'B')
                    b[0] = opi * alb * (st->dirflux[i - 1] + st->fluxes[0 + 2 * (i - 1)]) + gndrad;
            }
            for (k = 1; k < nstokes; k++) b[k] = 0.0f;
        }
    }
}

/* ------------------------------------------------------------------ */
/* INTEGRATE_1RAY  shdomsub2.f:2311-2743                               */
/* ------------------------------------------------------------------ */
static const int OPPFACE[6] = {2, 1, 4, 3, 6, 5};
static const int ONEY[8] = {0, 0, -1, -2, 0, 0, -5, -6};
static const int ONEX[8] = {0, -1, 0, -3, 0, -5, 0, -7};
/* DONEFACE(8,7) in Fortran column order */
static const int DONEFACE[7][8] = {
    {0,0,0,0,0,0,0,0}, {0,1,0,3,0,5,0,7}, {2,0,4,0,6,0,8,0},
    {0,0,1,2,0,0,5,6}, {3,4,0,0,5,6,0,0},
    {0,0,0,0,1,2,3,4}, {5,6,7,8,0,0,0,0}};

void oracle_donethis(const oracle_state *st, int iface, int *donethis)
{
    int i;
    for (i = 0; i < 8; i++) {
        donethis[i] = DONEFACE[iface][i];
        if (st->nx == 1 && ONEX[i] < 0) donethis[i] = ONEX[i];
        if (st->ny == 1 && ONEY[i] < 0) donethis[i] = ONEY[i];
    }
}

#define TRILERP(A, u, v, w) \
    ((1 - (w)) * ((1 - (v)) * ((1 - (u)) * A(1) + (u) * A(2)) + (v) * ((1 - (u)) * A(3) + (u) * A(4))) \
     + (w) * ((1 - (v)) * ((1 - (u)) * A(5) + (u) * A(6)) + (v) * ((1 - (u)) * A(7) + (u) * A(8))))

int oracle_integrate_1ray(const oracle_state *st, float *bcrad, float skyrad_top,
                          double mu2, double phi2, double x0, double y0, double z0,
                          double *transmit_io, double *radiance,
                          int correctinterpolate, int singlescatter, int nosurface,
                          ray_scratch *sc, int *trace_cells, int trace_cap, int *trace_n,
                          int *nsub_out, char *errmsg)
{
    const int nstokes = st->nstokes;
    ray_dir rd;
    float *ylmdir = sc->ylmdir, *singscat = sc->singscat;
    int oldipts[8] = {0,0,0,0,0,0,0,0}, donethis[8];
    float oextinct8[8], osrcext8[4 * 8], extinct8[8], srcext8[4 * 8];
    float ext0, ext1 = 0.0f, extn, srcext0[4], srcext1[4] = {0, 0, 0, 0}, radbnd[4];
    double xe, ye, ze, xn, yn, zn, xi, yi, zi, so, sox, soy, soz, eps;
    double taugrid, s, dels, ext, tau, transcell, abscell, src[4];
    double u, v, w, delx, dely, delz, invdelx, invdely, invdelz;
    double transmit = *transmit_io;
    int icell, inextcell, iface, jface, kface, ic, iopp, ntau, it, i, k, ngrid, maxcellscross;
    int ipt1, ipt2, ipinx, ipiny, openbcface, validrad, nsub = 0, ntrace = 0;
    const double tautol = st->tautol, transcut = st->transcut;
    (void)skyrad_top;

    memset(extinct8, 0, sizeof(extinct8));
    memset(srcext8, 0, sizeof(srcext8));
    for (k = 0; k < nstokes; k++) radiance[k] = 0.0;
    eps = 1.0e-5f * (GRIDPOS(st, 3, GRIDPTR(st, 8, 1)) - GRIDPOS(st, 3, GRIDPTR(st, 1, 1)));
    maxcellscross = 500 * IMAX3(st->nx, st->ny, st->nz);
    oracle_ray_setup(st, mu2, phi2, &rd, ylmdir, singscat, NULL, 0, NULL);

    xe = x0; ye = y0; ze = z0;
    icell = oracle_locate_grid_cell(st, &xe, &ye, &ze);
    iface = 0;
    ngrid = 0;
    validrad = 0;
    while (!validrad && icell > 0) {
        ngrid = ngrid + 1;
        if (trace_cells && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        oracle_donethis(st, iface, donethis);
        for (i = 0; i < 8; i++) {
            oextinct8[i] = extinct8[i];
            for (k = 0; k < nstokes; k++) osrcext8[k + nstokes * i] = srcext8[k + nstokes * i];
        }
        compute_source_1cell(st, icell, ylmdir, singscat, donethis, oldipts,
                             oextinct8, osrcext8, extinct8, srcext8, singlescatter, sc->legent);
        ipt1 = GRIDPTR(st, 1, icell);
        ipt2 = GRIDPTR(st, 8, icell);
        delx = GRIDPOS(st, 1, ipt2) - GRIDPOS(st, 1, ipt1);
        if (delx <= 0.0) invdelx = 1.0; else invdelx = 1.0 / delx;
        dely = GRIDPOS(st, 2, ipt2) - GRIDPOS(st, 2, ipt1);
        if (dely <= 0.0) invdely = 1.0; else invdely = 1.0 / dely;
        delz = GRIDPOS(st, 3, ipt2) - GRIDPOS(st, 3, ipt1);
        invdelz = 1.0 / delz;
        u = (xe - GRIDPOS(st, 1, ipt1)) * invdelx;
        v = (ye - GRIDPOS(st, 2, ipt1)) * invdely;
        w = (ze - GRIDPOS(st, 3, ipt1)) * invdelz;
#define E8(n) extinct8[(n) - 1]
        if (correctinterpolate || ngrid == 1) {
            for (k = 0; k < nstokes; k++) {
#define S8(n) srcext8[k + nstokes * ((n) - 1)]
                srcext1[k] = (float)TRILERP(S8, u, v, w);
#undef S8
            }
            srcext1[0] = fmaxf(0.0f, srcext1[0]);
            ext1 = (float)TRILERP(E8, u, v, w);
        }
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
            if (errmsg) snprintf(errmsg, 600, "INTEGRATE_1RAY: SO<0 %g %g %g %g %g %g %d",
                                 mu2, phi2, xe, ye, ze, so, icell);
            return 1;
        }
        xn = xe + so * rd.cx;
        yn = ye + so * rd.cy;
        zn = ze + so * rd.cz;
        u = (xn - GRIDPOS(st, 1, ipt1)) * invdelx;
        v = (yn - GRIDPOS(st, 2, ipt1)) * invdely;
        w = (zn - GRIDPOS(st, 3, ipt1)) * invdelz;
        extn = (float)TRILERP(E8, u, v, w);
        taugrid = so * 0.5f * (ext1 + extn);
        ntau = 1 + (int)(taugrid / tautol);
        if (ntau < 1) ntau = 1;
        dels = so / ntau;
        for (it = 1; it <= ntau; it++) {
            s = it * dels;
            xi = xe + s * rd.cx;
            yi = ye + s * rd.cy;
            zi = ze + s * rd.cz;
            u = (xi - GRIDPOS(st, 1, ipt1)) * invdelx;
            v = (yi - GRIDPOS(st, 2, ipt1)) * invdely;
            w = (zi - GRIDPOS(st, 3, ipt1)) * invdelz;
            ext0 = (float)TRILERP(E8, u, v, w);
            for (k = 0; k < nstokes; k++) {
#define S8(n) srcext8[k + nstokes * ((n) - 1)]
                srcext0[k] = (float)TRILERP(S8, u, v, w);
#undef S8
            }
            srcext0[0] = fmaxf(0.0f, srcext0[0]);
            ext = 0.5f * (ext0 + ext1);
            if (ext != 0.0) {
                tau = ext * dels;
                abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                transcell = 1.0f - abscell;
                for (k = 0; k < nstokes; k++)
                    src[k] = (0.5f * (srcext0[k] + srcext1[k])
                              + 0.08333333333f * (ext0 * srcext1[k] - ext1 * srcext0[k]) * dels
                                * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext;
            } else {
                abscell = 0.0;
                transcell = 1.0;
                for (k = 0; k < nstokes; k++) src[k] = 0.0;
            }
            for (k = 0; k < nstokes; k++) radiance[k] = radiance[k] + transmit * src[k] * abscell;
            transmit = transmit * transcell;
            ext1 = ext0;
            for (k = 0; k < nstokes; k++) srcext1[k] = srcext0[k];
            nsub++;
        }
#undef E8
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
        if (transmit < transcut || ngrid > maxcellscross) {
            validrad = 1;
        } else if (inextcell == 0 && iface >= 5) {
            int ierr;
            validrad = 1;
            ierr = find_boundary_radiance(st, bcrad, xn, yn, (float)mu2, (float)phi2, ic, kface,
                                          radbnd, errmsg);
            if (ierr) return ierr;
            if (!nosurface)
                for (k = 0; k < nstokes; k++) radiance[k] = radiance[k] + transmit * radbnd[k];
        } else {
            icell = inextcell;
        }
        xe = xn; ye = yn; ze = zn;
    }
    *transmit_io = transmit;
    if (trace_n) *trace_n = ntrace;
    if (nsub_out) *nsub_out = nsub;
    return 0;
}

ray_scratch *oracle_scratch_new(const oracle_state *st, int dnumphase)
{
    ray_scratch *sc = (ray_scratch *)calloc(1, sizeof(ray_scratch));
    sc->ylmdir = (float *)calloc((size_t)st->nstleg * st->nlm + 8, sizeof(float));
    sc->singscat = (float *)calloc((size_t)st->nstokes * (st->numphase > 0 ? st->numphase : 1) + 8, sizeof(float));
    sc->dsingscat = (float *)calloc((size_t)st->nstokes * (dnumphase > 0 ? dnumphase : 1) + 8, sizeof(float));
    sc->legent = (float *)calloc((size_t)st->nstleg * (st->nleg + 2) * 8, sizeof(float));
    return sc;
}

void oracle_scratch_free(ray_scratch *sc)
{
    if (!sc) return;
    free(sc->ylmdir); free(sc->singscat); free(sc->dsingscat); free(sc->legent);
    free(sc);
}

/* start-point handling of RENDER (shdomsub4.f:214-236); returns 1 if the ray sees nothing */
int oracle_ray_start(const oracle_state *st, double mu2, double phi2,
                     double *x0, double *y0, double *z0, int *ierr)
{
    double pi = acos(-1.0);
    double muray = -mu2, phiray = phi2 - pi, r;
    float ztop = st->zgrid[st->nz - 1];
    *ierr = 0;
    if (*z0 > ztop) {
        if (muray >= 0.0) return 1;
        r = (ztop - *z0) / muray;
        *x0 = *x0 + r * sqrt(1 - muray * muray) * cos(phiray);
        *y0 = *y0 + r * sqrt(1 - muray * muray) * sin(phiray);
        *z0 = ztop;
    } else if (*z0 < st->zgrid[0]) {
        *ierr = 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* RENDER  shdomsub4.f:93-286                                          */
/* ------------------------------------------------------------------ */
int oracle_render(const oracle_state *st, const oracle_rays *rays, float *stokes,
                  int correctinterpolate, int singlescatter, int nosurface,
                  oracle_trace *trace, int nthreads, char *errmsg)
{
    const int nstokes = st->nstokes;
    int ierr_all = 0;
    /* bottom boundary radiances (shdomsub4.f:201-209); BCRAD is mutated like the reference */
    oracle_lambertian_boundary(st, st->bcrad);
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        ray_scratch *sc = oracle_scratch_new(st, 0);
        /* private BCRAD: the reference rewrites BCRAD(:,1:NTOPPTS) per ray (shdomsub4.f:238-248),
         * which is why at3d deep-copies it per thread (solver.py:747) */
        size_t nbc = (size_t)nstokes * (st->ntoppts + (size_t)st->nbotpts *
                                        (st->sfctype1 == 'L' ? 1 : 1 + st->nang / 2));
        float *bcrad = (float *)malloc(sizeof(float) * (nbc + 1));
        char lmsg[600];
        int ivis, k, itop;
        memcpy(bcrad, st->bcrad, sizeof(float) * nbc);
        lmsg[0] = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (ivis = 0; ivis < rays->nrays; ivis++) {
            double x0 = rays->camx[ivis], y0 = rays->camy[ivis], z0 = rays->camz[ivis];
            double mu2 = rays->cammu[ivis], phi2 = rays->camphi[ivis];
            double muray = -mu2, transmit = 1.0, visrad[4] = {0, 0, 0, 0};
            int ierr = 0, dark, ntr = 0, nsub = 0;
            if (ierr_all) continue;
            dark = oracle_ray_start(st, mu2, phi2, &x0, &y0, &z0, &ierr);
            if (ierr) {
                snprintf(lmsg, 600, "RENDER: Level below domain");
            } else if (!dark) {
                if (muray > 0.0) {
                    float sky = oracle_sky_radiance(st, (float)mu2, (float)phi2);
                    for (itop = 0; itop < st->ntoppts; itop++) {
                        bcrad[nstokes * itop] = sky;
                        for (k = 1; k < nstokes; k++) bcrad[k + nstokes * itop] = 0.0f;
                    }
                } else {
                    for (itop = 0; itop < st->ntoppts * nstokes; itop++) bcrad[itop] = 0.0f;
                }
                ierr = oracle_integrate_1ray(st, bcrad, 0.0f, mu2, phi2, x0, y0, z0, &transmit, visrad,
                                             correctinterpolate, singlescatter, nosurface, sc,
                                             trace ? trace->cells + (size_t)trace->max_per_ray * ivis : NULL,
                                             trace ? trace->max_per_ray : 0, &ntr, &nsub, lmsg);
            }
            if (trace) { trace->ncells[ivis] = ntr; trace->nsub[ivis] = nsub; }
            for (k = 0; k < nstokes; k++) stokes[k + nstokes * ivis] = (float)visrad[k];
            if (ierr) {
#ifdef _OPENMP
#pragma omp critical
#endif
                { if (!ierr_all) { ierr_all = ierr; if (errmsg) { strncpy(errmsg, lmsg, 599); errmsg[599] = 0; } } }
            }
        }
        free(bcrad);
        oracle_scratch_free(sc);
    }
    return ierr_all;
}
