/*
 * oracle_direct.c -- TEST INFRASTRUCTURE (see shdom_oracle.h).
 * Direct solar beam on the property grid and its extinction derivative paths:
 *   MAKE_DIRECT / DIRECT_BEAM_PROP   /root/reference/src/polarized/shdomsub2.f:393-478,
 *                                    /root/reference/src/polarized/shdom90.f90:352-867
 *   MAKE_DIRECT_DERIVATIVE / DIRECT_BEAM_AND_PATHS_PROP  /root/reference/src/shdomsub5.f:1553-2004
 * Pinned by /root/reference/tests/data/dirflux_gradient_*.npy (tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "shdom_oracle.h"

typedef struct {
    int bcflag, npx, npy, npz, ipdirect, di, dj, dk;
    double cx, cy, cz, cxinv, cyinv, czinv, epss, epsz, xdomain, ydomain, delxd, delyd;
    float xstart, ystart;
    const float *zlevels;
} beam_geom;

#define BT(x, b) ((((int)(x)) >> (b)) & 1)

/* One walk from (xi,yi,zi) toward the sun through the property grid.
 * extdirp != NULL : accumulate the optical path (DIRECT_BEAM_PROP INIT=0, shdom90.f90:563-867)
 * dpath   != NULL : record DPATH/DPTR (DIRECT_BEAM_AND_PATHS_PROP, shdomsub5.f:1646-2003)   */
static int beam_walk(const beam_geom *g, float xi, float yi, float zi,
                     const float *extdirp, double *path_io, int *path_pts,
                     float *dpath, int *dptr, int longest_path_pts, char *errmsg)
{
    const int npx = g->npx, npy = g->npy, npz = g->npz, bcflag = g->bcflag;
    const double cx = g->cx, cy = g->cy, cz = g->cz, delxd = g->delxd, delyd = g->delyd;
    const float *zl = g->zlevels;
    double x, y, z, xe, ye, ze, xp, yp, zp, x0, x1, y0, y1, z0, z1, so, sox, soy, soz;
    double xoffs, yoffs, ax, ay, az, u0, v0, w0, u1, v1, w1, u0m, v0m, w0m, du, dv, dw;
    double uv, umv, uvm, umvm, uw, umw, uwm, umwm, vw, vmw, vwm, vmwm;
    double b1, b2, b3, b4, b5, b6, b7, b8, c1, c2, c3, c4, c5, c6, c7, c8;
    double vwu, vwum, uwv, uwvm, uvw, uvwm;
    double path = path_io ? *path_io : 0.0;
    int il, iu, im, i, j, k, ip, jp, i1, i2, i3, i4, constx, consty, hitboundary, idp = 0, npp = 0;

    z = zi;
    x = xi - g->xstart;
    y = yi - g->ystart;
    il = 0; iu = npz;
    while (iu - il > 1) {
        im = (iu + il) / 2;
        if (z >= zl[im - 1]) il = im; else iu = im;
    }
    k = il > 1 ? il : 1;
    i = (int)(x / delxd) + 1;
    if (i > npx && fabs(x - g->xdomain) < 0.001f * delxd) i = npx;
    if (i < 1 || i > npx) {
        if (errmsg) snprintf(errmsg, 600, "DIRECT_BEAM_PROP: Beyond X domain %d %g %g %g", i, xi, yi, zi);
        return 1;
    }
    j = (int)(y / delyd) + 1;
    if (j > npy && fabs(y - g->ydomain) < 0.001f * delyd) j = npy;
    if (j < 1 || j > npy) {
        if (errmsg) snprintf(errmsg, 600, "DIRECT_BEAM_PROP: Beyond Y domain %d %g %g %g", j, xi, yi, zi);
        return 1;
    }
    xe = x; ye = y; ze = z;
    xp = xe; yp = ye; zp = ze;
    constx = BT(g->ipdirect, 0);
    consty = BT(g->ipdirect, 1);
    if (cx == 0.0) constx = 1;
    if (cy == 0.0) consty = 1;
    if (BT(bcflag, 0) && (fabs(x) < 0.01f * delxd || fabs(x - (npx - 1) * delxd) < 0.01f * delxd)) constx = 1;
    if (BT(bcflag, 1) && (fabs(y) < 0.01f * delyd || fabs(y - (npy - 1) * delyd) < 0.01f * delyd)) consty = 1;
    hitboundary = 0;
    if (BT(bcflag, 2)) {
        if (cx > 0.0 && fabs(x - g->xdomain) < 0.001f * delxd) hitboundary = 1;
        else if (cx < 0.0 && fabs(x) < 0.001f * delxd) hitboundary = 1;
    }
    if (BT(bcflag, 3)) {
        if (cy > 0.0 && fabs(y - g->ydomain) < 0.001f * delyd) hitboundary = 1;
        if (cy < 0.0 && fabs(y) < 0.001f * delyd) hitboundary = 1;
    }
    while (!hitboundary && fabs(ze - zl[npz - 1]) > g->epsz) {
        ip = i + 1;
        if (i == npx) ip = (BT(bcflag, 0) || BT(bcflag, 2)) ? npx : 1;
        jp = j + 1;
        if (j == npy) jp = (BT(bcflag, 1) || BT(bcflag, 3)) ? npy : 1;
        x0 = delxd * (i - 1);
        x1 = x0 + delxd;
        y0 = delyd * (j - 1);
        y1 = y0 + delyd;
        if (i < 1 || i > npx || j < 1 || j > npy || k < 1 || k >= npz) {
            if (errmsg) snprintf(errmsg, 600, "DIRECT_BEAM_PROP: beyond grid! %d %d %d", i, j, k);
            return 1;
        }
        z0 = zl[k - 1];
        z1 = zl[k];
        i1 = k + npz * (j - 1) + npz * npy * (i - 1);
        i2 = k + npz * (j - 1) + npz * npy * (ip - 1);
        i3 = k + npz * (jp - 1) + npz * npy * (i - 1);
        i4 = k + npz * (jp - 1) + npz * npy * (ip - 1);
        if (constx) sox = 1.0e30f;
        else if (cx > 0.0) { sox = (x1 - xe) * g->cxinv; xp = x1; }
        else { sox = (x0 - xe) * g->cxinv; xp = x0; }
        if (consty) soy = 1.0e30f;
        else if (cy > 0.0) { soy = (y1 - ye) * g->cyinv; yp = y1; }
        else { soy = (y0 - ye) * g->cyinv; yp = y0; }
        if (cz > 0.0) { soz = (z1 - ze) * g->czinv; zp = z1; }
        else if (cz < 0.0) { soz = (z0 - ze) * g->czinv; zp = z0; }
        else soz = 1.0e30f;
        xoffs = 0.0;
        yoffs = 0.0;
        if (soz <= sox && soz <= soy) {
            so = soz;
            if (!constx) xp = xe + so * cx;
            if (!consty) yp = ye + so * cy;
            k = k + g->dk;
        } else if (sox <= soy) {
            so = sox;
            if (!consty) yp = ye + so * cy;
            zp = ze + so * cz;
            i = i + g->di;
            if (i == 0) {
                if (BT(bcflag, 0)) { i = 1; constx = 1; }
                else if (BT(bcflag, 2)) hitboundary = 1;
                else { i = npx; xoffs = g->xdomain; }
            } else if (i >= npx && BT(bcflag, 2)) {
                hitboundary = 1;
            } else if (i == npx + 1) {
                if (BT(bcflag, 0)) { i = npx; constx = 1; }
                else { i = 1; xoffs = -g->xdomain; }
            }
        } else {
            so = soy;
            if (!constx) xp = xe + so * cx;
            zp = ze + so * cz;
            j = j + g->dj;
            if (j == 0) {
                if (BT(bcflag, 1)) { j = 1; consty = 1; }
                else if (BT(bcflag, 3)) hitboundary = 1;
                else { j = npy; yoffs = g->ydomain; }
            } else if (j >= npy && BT(bcflag, 3)) {
                hitboundary = 1;
            } else if (j == npy + 1) {
                if (BT(bcflag, 1)) { j = npy; consty = 1; }
                else { j = 1; yoffs = -g->ydomain; }
            }
        }
        if (so < -g->epss) {
            if (errmsg) snprintf(errmsg, 600, "DIRECT_BEAM_PROP: SO<0 %g %g %g", x, y, z);
            return 1;
        }
        so = fmax(so, 0.0);
        ax = 1.0 / (x1 - x0);
        ay = 1.0 / (y1 - y0);
        az = 1.0 / (z1 - z0);
        u0 = (xe - x0) * ax; v0 = (ye - y0) * ay; w0 = (ze - z0) * az;
        u1 = (xp - x0) * ax; v1 = (yp - y0) * ay; w1 = (zp - z0) * az;
        u0m = 1.0f - u0; v0m = 1.0f - v0; w0m = 1.0f - w0;
        du = u1 - u0; dv = v1 - v0; dw = w1 - w0;
        uv = u0 * v0; umv = u0m * v0; uvm = u0 * v0m; umvm = u0m * v0m;
        uw = u0 * w0; umw = u0m * w0; uwm = u0 * w0m; umwm = u0m * w0m;
        vw = v0 * w0; vmw = v0m * w0; vwm = v0 * w0m; vmwm = v0m * w0m;
        b1 = -du * vmwm - dv * umwm - dw * umvm;
        b2 = du * vmwm - dv * uwm - dw * uvm;
        b3 = -du * vwm + dv * umwm - dw * umv;
        b4 = du * vwm + dv * uwm - dw * uv;
        b5 = -du * vmw - dv * umw + dw * umvm;
        b6 = du * vmw - dv * uw + dw * uvm;
        b7 = -du * vw + dv * umw + dw * umv;
        b8 = du * vw + dv * uw + dw * uv;
        if (extdirp) {
            double e1 = extdirp[i1 - 1], e2 = extdirp[i2 - 1], e3 = extdirp[i3 - 1], e4 = extdirp[i4 - 1];
            double e5 = extdirp[i1], e6 = extdirp[i2], e7 = extdirp[i3], e8 = extdirp[i4];
            double a, b, c, d, vw2, uw2, uv2;
            a = (e1 * u0m + e2 * u0) * vmwm + (e3 * u0m + e4 * u0) * vwm
              + (e5 * u0m + e6 * u0) * vmw + (e7 * u0m + e8 * u0) * vw;
            b = b1 * e1 + b2 * e2 + b3 * e3 + b4 * e4 + b5 * e5 + b6 * e6 + b7 * e7 + b8 * e8;
            vw2 = dv * dw; vwu = vw2 * u0; vwum = vw2 * u0m;
            uw2 = du * dw; uwv = uw2 * v0; uwvm = uw2 * v0m;
            uv2 = du * dv; uvw = uv2 * w0; uvwm = uv2 * w0m;
            c1 = +vwum + uwvm + uvwm; c2 = +vwu - uwvm - uvwm;
            c3 = -vwum + uwv - uvwm;  c4 = -vwu - uwv + uvwm;
            c5 = -vwum - uwvm + uvw;  c6 = -vwu + uwvm - uvw;
            c7 = +vwum - uwv - uvw;   c8 = +vwu + uwv + uvw;
            c = c1 * e1 + c2 * e2 + c3 * e3 + c4 * e4 + c5 * e5 + c6 * e6 + c7 * e7 + c8 * e8;
            d = du * dv * dw * (e2 + e3 + e5 + e8 - e1 - e4 - e6 - e7);
            path = path + so * (a + 0.5 * b + 0.3333333333333333 * c + 0.25 * d);
            npp += 8;
        }
        if (dpath) {
            double a1 = u0m * vmwm, a2 = u0 * vmwm, a3 = u0m * vwm, a4 = u0 * vwm;
            double a5 = u0m * vmw, a6 = u0 * vmw, a7 = u0m * vw, a8 = u0 * vw;
            double vw2, uw2, uv2;
            vw2 = dv * dw; vwu = vw2 * u0; vwum = vw2 * u0m;
            uw2 = du * dw; uwv = uw2 * v0; uwvm = uw2 * v0m;
            uv2 = du * dv; uvw = uv2 * w0; uvwm = uv2 * w0m;
            c1 = +vwum + uwvm + uvwm; c2 = +vwu - uwvm - uvwm;
            c3 = -vwum + uwv - uvwm;  c4 = -vwu - uwv + uvwm;
            c5 = -vwum - uwvm + uvw;  c6 = -vwu + uwvm - uvw;
            c7 = +vwum - uwv - uvw;   c8 = +vwu + uwv + uvw;
            if (idp + 8 > longest_path_pts) {
                if (errmsg) snprintf(errmsg, 600, "DIRECT_BEAM_AND_PATHS_PROP: Max number of property "
                                     "points to pass exceeded: IDP=%d", longest_path_pts);
                return 1;
            }
            dpath[idp + 0] = (float)(so * (a1 + 0.5 * b1 + 0.3333333333333333 * c1 - 0.25 * du * dv * dw));
            dpath[idp + 1] = (float)(so * (a2 + 0.5 * b2 + 0.3333333333333333 * c2 + 0.25 * du * dv * dw));
            dpath[idp + 2] = (float)(so * (a3 + 0.5 * b3 + 0.3333333333333333 * c3 + 0.25 * du * dv * dw));
            dpath[idp + 3] = (float)(so * (a4 + 0.5 * b4 + 0.3333333333333333 * c4 - 0.25 * du * dv * dw));
            dpath[idp + 4] = (float)(so * (a5 + 0.5 * b5 + 0.3333333333333333 * c5 + 0.25 * du * dv * dw));
            dpath[idp + 5] = (float)(so * (a6 + 0.5 * b6 + 0.3333333333333333 * c6 - 0.25 * du * dv * dw));
            dpath[idp + 6] = (float)(so * (a7 + 0.5 * b7 + 0.3333333333333333 * c7 - 0.25 * du * dv * dw));
            dpath[idp + 7] = (float)(so * (a8 + 0.5 * b8 + 0.3333333333333333 * c8 + 0.25 * du * dv * dw));
            dptr[idp + 0] = i1; dptr[idp + 1] = i2; dptr[idp + 2] = i3; dptr[idp + 3] = i4;
            dptr[idp + 4] = i1 + 1; dptr[idp + 5] = i2 + 1; dptr[idp + 6] = i3 + 1; dptr[idp + 7] = i4 + 1;
            idp += 8;
        }
        xe = xp + xoffs;
        ye = yp + yoffs;
        ze = zp;
    }
    if (path_io) *path_io = path;
    if (path_pts) *path_pts = npp;
    return 0;
}

/* MAKE_DIRECT: DIRECT_BEAM_PROP(INIT=1) then one walk per grid point.
 * out_d[13] = CX,CY,CZ,CXINV,CYINV,CZINV,EPSS,EPSZ,XDOMAIN,YDOMAIN,UNIFORMZLEV,DELXD,DELYD
 * out_i[5]  = IPDIRECT,DI,DJ,DK,LONGEST_PATH_PTS */
int oracle_make_direct(int npts, int bcflag, int ipflag, int deltam, int ml, int nstleg, int nlegp,
                       float solarflux, float solarmu, float solaraz, const float *gridpos,
                       int npx, int npy, int npz, float delx, float dely, float xstart, float ystart,
                       const float *zlevels, const float *extinctp, const float *albedop,
                       const float *legenp, const int *iphasep, const float *phasewtp,
                       int maxnmicro, int npart, int nzckd, const float *zckd, const float *gasabs,
                       float *extdirp, float *dirflux, double *out_d, int *out_i, char *errmsg)
{
    const int maxpg = npx * npy * npz;
    beam_geom g;
    float *gasext = (float *)calloc(npz, sizeof(float));
    float *extmin = (float *)calloc(npz, sizeof(float)), *extmax = (float *)calloc(npz, sizeof(float));
    double sunmu, sunaz, epss, epsz, uniformzlev;
    int ix, iy, iz, ip, ipa, q, jz, longest = 0, ierr = 0;
    for (iz = 1; iz <= npz; iz++) {
        if (nzckd > 0) {
            int il = 1, iu = nzckd, im, i;
            double w0;
            while (iu - il > 1) {
                im = (iu + il) / 2;
                if (zlevels[iz - 1] <= zckd[im - 1]) il = im; else iu = im;
            }
            i = il > 1 ? il : 1;
            if (i > nzckd - 1) i = nzckd - 1;
            w0 = (zlevels[iz - 1] - zckd[i - 1]) / (zckd[i] - zckd[i - 1]);
            w0 = fmin(fmax(w0, 0.0), 1.0);
            gasext[iz - 1] = (float)((1.0f - w0) * gasabs[i - 1] + w0 * gasabs[i]);
        } else gasext[iz - 1] = 0.0f;
    }
    ip = 0;
    for (ix = 1; ix <= npx; ix++)
        for (iy = 1; iy <= npy; iy++)
            for (iz = 1; iz <= npz; iz++) {
                ip++;
                extdirp[ip - 1] = 0.0f;
                for (ipa = 1; ipa <= npart; ipa++) {
                    double extinct = extinctp[(ip - 1) + (size_t)maxpg * (ipa - 1)];
                    double albedo = albedop[(ip - 1) + (size_t)maxpg * (ipa - 1)];
                    if (gasext[iz - 1] > 0.0f) {
                        albedo = albedo * extinct / (extinct + gasext[iz - 1]);
                        extinct = extinct + gasext[iz - 1];
                    }
                    if (deltam) {
                        int l = ml + 1;
                        double f = 0.0;
                        for (q = 1; q <= maxnmicro; q++) {
                            int iph = iphasep[(q - 1) + maxnmicro * ((ip - 1) + (size_t)maxpg * (ipa - 1))];
                            float pw = phasewtp[(q - 1) + maxnmicro * ((ip - 1) + (size_t)maxpg * (ipa - 1))];
                            f = f + pw * legenp[nstleg * (l + (size_t)(nlegp + 1) * (iph - 1))] / (2 * l + 1);
                        }
                        extinct = (1.0f - albedo * f) * extinct;
                    }
                    extdirp[ip - 1] = (float)(extdirp[ip - 1] + extinct);
                }
            }
    g.bcflag = bcflag; g.npx = npx; g.npy = npy; g.npz = npz;
    g.xstart = xstart; g.ystart = ystart; g.zlevels = zlevels;
    g.ipdirect = ipflag;
    if (BT(ipflag, 2)) g.ipdirect = 0;
    sunmu = -solarmu;
    sunaz = solaraz + acosf(-1.0f);
    g.cx = sqrt(1.0f - sunmu * sunmu) * cos(sunaz);
    g.cy = sqrt(1.0f - sunmu * sunmu) * sin(sunaz);
    g.cz = fabs(sunmu);
    if (fabs(g.cx) > 1.0e-6f) g.cxinv = 1.0 / g.cx; else { g.cx = 0.0; g.cxinv = 1.0e20f; }
    if (fabs(g.cy) > 1.0e-6f) g.cyinv = 1.0 / g.cy; else { g.cy = 0.0; g.cyinv = 1.0e20f; }
    if (fabs(g.cz) > 1.0e-6f) g.czinv = 1.0 / g.cz; else { g.cz = 0.0; g.czinv = 1.0e20f; }
    g.di = signbit(g.cx) ? -1 : 1;
    g.dj = signbit(g.cy) ? -1 : 1;
    g.dk = signbit(g.cz) ? -1 : 1;
    epsz = 1.0e-6f * (zlevels[npz - 1] - zlevels[0]);
    epss = 1.0e-3f * (zlevels[npz - 1] - zlevels[0]) / npz;
    if (!BT(g.ipdirect, 0)) epss = fmax(epss, 1.0e-4 * delx);
    if (!BT(g.ipdirect, 1)) epss = fmax(epss, 1.0e-4 * dely);
    g.delxd = (double)delx;
    g.delyd = (double)dely;
    epss = fmax(fmax(0.001f * g.delxd, 0.001f * g.delyd), epss);
    g.epss = epss; g.epsz = epsz;
    g.xdomain = g.delxd * npx;
    if (BT(bcflag, 2)) g.xdomain = g.delxd * (npx - 1);
    g.ydomain = g.delyd * npy;
    if (BT(bcflag, 3)) g.ydomain = g.delyd * (npy - 1);
    for (iz = 0; iz < npz; iz++) { extmin[iz] = 1.0e20f; extmax[iz] = 0.0f; }
    ip = 0;
    for (ix = 1; ix <= npx; ix++)
        for (iy = 1; iy <= npy; iy++)
            for (iz = 1; iz <= npz; iz++) {
                float s = 0.0f;
                ip++;
                for (ipa = 1; ipa <= npart; ipa++) s = s + extinctp[(ip - 1) + (size_t)maxpg * (ipa - 1)];
                extmin[iz - 1] = fminf(s, extmin[iz - 1]);
                extmax[iz - 1] = fmaxf(s, extmax[iz - 1]);
            }
    jz = 0;
    for (iz = 1; iz <= npz; iz++)
        if (extmax[iz - 1] - extmin[iz - 1] > 1.0e-4f) jz = iz;
    jz = jz + 1 < npz ? jz + 1 : npz;
    uniformzlev = zlevels[jz - 1];
    for (ip = 1; ip <= npts && !ierr; ip++) {
        double path = 0.0;
        int npp = 0;
        ierr = beam_walk(&g, gridpos[3 * (size_t)(ip - 1)], gridpos[1 + 3 * (size_t)(ip - 1)],
                         gridpos[2 + 3 * (size_t)(ip - 1)], extdirp, &path, &npp, NULL, NULL, 0, errmsg);
        dirflux[ip - 1] = (float)(solarflux * exp(-path));
        if (npp > longest) longest = npp;
    }
    out_d[0] = g.cx; out_d[1] = g.cy; out_d[2] = g.cz; out_d[3] = g.cxinv; out_d[4] = g.cyinv;
    out_d[5] = g.czinv; out_d[6] = g.epss; out_d[7] = g.epsz; out_d[8] = g.xdomain; out_d[9] = g.ydomain;
    out_d[10] = uniformzlev; out_d[11] = g.delxd; out_d[12] = g.delyd;
    out_i[0] = g.ipdirect; out_i[1] = g.di; out_i[2] = g.dj; out_i[3] = g.dk; out_i[4] = longest;
    free(gasext); free(extmin); free(extmax);
    return ierr;
}

/* DIRECT_BEAM_PROP(INIT=0) for one point (the call of INTERPOLATE_POINT, shdomsub1.f:5093-5107), with the beam
 * constants oracle_make_direct returned in out_d / out_i. */
int oracle_direct_beam_point(const double *out_d, const int *out_i, int bcflag, int npx, int npy, int npz,
                             float xstart, float ystart, const float *zlevels, const float *extdirp,
                             float solarflux, float x, float y, float z, float *dirflux, char *errmsg)
{
    beam_geom g;
    double path = 0.0;
    int npp = 0, ierr;
    g.bcflag = bcflag; g.npx = npx; g.npy = npy; g.npz = npz;
    g.xstart = xstart; g.ystart = ystart; g.zlevels = zlevels;
    g.cx = out_d[0]; g.cy = out_d[1]; g.cz = out_d[2]; g.cxinv = out_d[3]; g.cyinv = out_d[4]; g.czinv = out_d[5];
    g.epss = out_d[6]; g.epsz = out_d[7]; g.xdomain = out_d[8]; g.ydomain = out_d[9];
    g.delxd = out_d[11]; g.delyd = out_d[12];
    g.ipdirect = out_i[0]; g.di = out_i[1]; g.dj = out_i[2]; g.dk = out_i[3];
    ierr = beam_walk(&g, x, y, z, extdirp, &path, &npp, NULL, NULL, 0, errmsg);
    *dirflux = (float)(solarflux * exp(-path));
    return ierr;
}

/* MAKE_DIRECT_DERIVATIVE  shdomsub5.f:1553-1605 */
int oracle_make_direct_derivative(int npts, int bcflag, int npx, int npy, int npz,
                                  float delx, float dely, float xstart, float ystart,
                                  const float *gridpos, const float *zlevels,
                                  int ipdirect, int di, int dj, int dk,
                                  double cx, double cy, double cz,
                                  double cxinv, double cyinv, double czinv,
                                  double epss, double epsz, double xdomain, double ydomain,
                                  double uniformzlev, double delxd, double delyd,
                                  float *dpath, int *dptr, int longest_path_pts, char *errmsg)
{
    beam_geom g;
    int ip, ierr = 0;
    (void)delx; (void)dely; (void)uniformzlev;
    g.bcflag = bcflag; g.npx = npx; g.npy = npy; g.npz = npz; g.ipdirect = ipdirect;
    g.di = di; g.dj = dj; g.dk = dk; g.cx = cx; g.cy = cy; g.cz = cz;
    g.cxinv = cxinv; g.cyinv = cyinv; g.czinv = czinv; g.epss = epss; g.epsz = epsz;
    g.xdomain = xdomain; g.ydomain = ydomain; g.delxd = delxd; g.delyd = delyd;
    g.xstart = xstart; g.ystart = ystart; g.zlevels = zlevels;
    memset(dptr, 0, sizeof(int) * (size_t)longest_path_pts * npts);
    memset(dpath, 0, sizeof(float) * (size_t)longest_path_pts * npts);
    for (ip = 1; ip <= npts && !ierr; ip++)
        ierr = beam_walk(&g, gridpos[3 * (size_t)(ip - 1)], gridpos[1 + 3 * (size_t)(ip - 1)],
                         gridpos[2 + 3 * (size_t)(ip - 1)], NULL, NULL, NULL,
                         &dpath[(size_t)longest_path_pts * (ip - 1)],
                         &dptr[(size_t)longest_path_pts * (ip - 1)], longest_path_pts, errmsg);
    return ierr;
}
