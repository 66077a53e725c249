// Direct-beam walks through the property grid, shared by MAKE_DIRECT / MAKE_DIRECT_DERIVATIVE (at3d_prep.cu) and the
// streaming direct-beam derivative of the gradient (at3d_grad.cu).
#pragma once

struct BeamGeom {
    int bcflag, npx, npy, npz, ipdirect, di, dj, dk;
    double cx, cy, cz, cxinv, cyinv, czinv, epss, epsz, xdomain, ydomain, delxd, delyd;
    float xstart, ystart;
};

#define BT(x, b) ((((int)(x)) >> (b)) & 1)

// One walk from a grid point toward the sun through the property grid (shdom90.f90:563-867 with the
// closed-form trilinear path integrals; shdomsub5.f:1646-2003 for the recorded DPATH/DPTR variant).
// Returns 0 or an error code (1 beyond X, 2 beyond Y, 3 beyond grid, 4 SO<0, 5 list overflow).
// PATHS: every crossed property cell hands its 8 (property point, path integral) entries to `sink.put(ptr, val)`
// (false from the sink: list overflow, code 5).
template <bool PATHS, class Sink>
__device__ int beam_walk(const BeamGeom &g, const float *zl, float xi, float yi, float zi,
                         const float *extdirp, double &path, int &npp, Sink &sink)
{
    const int npx = g.npx, npy = g.npy, npz = g.npz, bcflag = g.bcflag;
    const double cx = g.cx, cy = g.cy, cz = g.cz, delxd = g.delxd, delyd = g.delyd;
    double x, y, z, xe, ye, ze, xp, yp, zp, x0, x1, y0, y1, z0, z1, so, sox, soy, soz;
    int il, iu, im, i, j, k, ip, jp;
    path = 0.0; npp = 0;
    z = zi; x = xi - g.xstart; y = yi - g.ystart;
    il = 0; iu = npz;
    while (iu - il > 1) { im = (iu + il) / 2; if (z >= zl[im - 1]) il = im; else iu = im; }
    k = il > 1 ? il : 1;
    i = (int)(x / delxd) + 1;
    if (i > npx && fabs(x - g.xdomain) < 0.001f * delxd) i = npx;
    if (i < 1 || i > npx) return 1;
    j = (int)(y / delyd) + 1;
    if (j > npy && fabs(y - g.ydomain) < 0.001f * delyd) j = npy;
    if (j < 1 || j > npy) return 2;
    xe = x; ye = y; ze = z;
    xp = xe; yp = ye; zp = ze;
    bool constx = BT(g.ipdirect, 0), consty = BT(g.ipdirect, 1);
    if (cx == 0.0) constx = true;
    if (cy == 0.0) consty = true;
    if (BT(bcflag, 0) && (fabs(x) < 0.01f * delxd || fabs(x - (npx - 1) * delxd) < 0.01f * delxd)) constx = true;
    if (BT(bcflag, 1) && (fabs(y) < 0.01f * delyd || fabs(y - (npy - 1) * delyd) < 0.01f * delyd)) consty = true;
    bool hitboundary = false;
    if (BT(bcflag, 2)) {
        if (cx > 0.0 && fabs(x - g.xdomain) < 0.001f * delxd) hitboundary = true;
        else if (cx < 0.0 && fabs(x) < 0.001f * delxd) hitboundary = true;
    }
    if (BT(bcflag, 3)) {
        if (cy > 0.0 && fabs(y - g.ydomain) < 0.001f * delyd) hitboundary = true;
        if (cy < 0.0 && fabs(y) < 0.001f * delyd) hitboundary = true;
    }
    while (!hitboundary && fabs(ze - zl[npz - 1]) > g.epsz) {
        ip = i + 1;
        if (i == npx) ip = (BT(bcflag, 0) || BT(bcflag, 2)) ? npx : 1;
        jp = j + 1;
        if (j == npy) jp = (BT(bcflag, 1) || BT(bcflag, 3)) ? npy : 1;
        x0 = delxd * (i - 1); x1 = x0 + delxd;
        y0 = delyd * (j - 1); y1 = y0 + delyd;
        if (i < 1 || i > npx || j < 1 || j > npy || k < 1 || k >= npz) return 3;
        z0 = zl[k - 1]; z1 = zl[k];
        const int i1 = k + npz * (j - 1) + npz * npy * (i - 1);
        const int i2 = k + npz * (j - 1) + npz * npy * (ip - 1);
        const int i3 = k + npz * (jp - 1) + npz * npy * (i - 1);
        const int i4 = k + npz * (jp - 1) + npz * npy * (ip - 1);
        if (constx) sox = 1.0e30f;
        else if (cx > 0.0) { sox = (x1 - xe) * g.cxinv; xp = x1; }
        else { sox = (x0 - xe) * g.cxinv; xp = x0; }
        if (consty) soy = 1.0e30f;
        else if (cy > 0.0) { soy = (y1 - ye) * g.cyinv; yp = y1; }
        else { soy = (y0 - ye) * g.cyinv; yp = y0; }
        if (cz > 0.0) { soz = (z1 - ze) * g.czinv; zp = z1; }
        else if (cz < 0.0) { soz = (z0 - ze) * g.czinv; zp = z0; }
        else soz = 1.0e30f;
        double xoffs = 0.0, yoffs = 0.0;
        if (soz <= sox && soz <= soy) {
            so = soz;
            if (!constx) xp = xe + so * cx;
            if (!consty) yp = ye + so * cy;
            k = k + g.dk;
        } else if (sox <= soy) {
            so = sox;
            if (!consty) yp = ye + so * cy;
            zp = ze + so * cz;
            i = i + g.di;
            if (i == 0) {
                if (BT(bcflag, 0)) { i = 1; constx = true; }
                else if (BT(bcflag, 2)) hitboundary = true;
                else { i = npx; xoffs = g.xdomain; }
            } else if (i >= npx && BT(bcflag, 2)) {
                hitboundary = true;
            } else if (i == npx + 1) {
                if (BT(bcflag, 0)) { i = npx; constx = true; }
                else { i = 1; xoffs = -g.xdomain; }
            }
        } else {
            so = soy;
            if (!constx) xp = xe + so * cx;
            zp = ze + so * cz;
            j = j + g.dj;
            if (j == 0) {
                if (BT(bcflag, 1)) { j = 1; consty = true; }
                else if (BT(bcflag, 3)) hitboundary = true;
                else { j = npy; yoffs = g.ydomain; }
            } else if (j >= npy && BT(bcflag, 3)) {
                hitboundary = true;
            } else if (j == npy + 1) {
                if (BT(bcflag, 1)) { j = npy; consty = true; }
                else { j = 1; yoffs = -g.ydomain; }
            }
        }
        if (so < -g.epss) return 4;
        so = fmax(so, 0.0);
        const double ax = 1.0 / (x1 - x0), ay = 1.0 / (y1 - y0), az = 1.0 / (z1 - z0);
        const double u0 = (xe - x0) * ax, v0 = (ye - y0) * ay, w0 = (ze - z0) * az;
        const double u1 = (xp - x0) * ax, v1 = (yp - y0) * ay, w1 = (zp - z0) * az;
        const double u0m = 1.0f - u0, v0m = 1.0f - v0, w0m = 1.0f - w0;
        const double du = u1 - u0, dv = v1 - v0, dw = w1 - w0;
        const double uv = u0 * v0, umv = u0m * v0, uvm = u0 * v0m, umvm = u0m * v0m;
        const double uw = u0 * w0, umw = u0m * w0, uwm = u0 * w0m, umwm = u0m * w0m;
        const double vw = v0 * w0, vmw = v0m * w0, vwm = v0 * w0m, vmwm = v0m * w0m;
        const double b1 = -du * vmwm - dv * umwm - dw * umvm;
        const double b2 = du * vmwm - dv * uwm - dw * uvm;
        const double b3 = -du * vwm + dv * umwm - dw * umv;
        const double b4 = du * vwm + dv * uwm - dw * uv;
        const double b5 = -du * vmw - dv * umw + dw * umvm;
        const double b6 = du * vmw - dv * uw + dw * uvm;
        const double b7 = -du * vw + dv * umw + dw * umv;
        const double b8 = du * vw + dv * uw + dw * uv;
        const double vw2 = dv * dw, vwu = vw2 * u0, vwum = vw2 * u0m;
        const double uw2 = du * dw, uwv = uw2 * v0, uwvm = uw2 * v0m;
        const double uv2 = du * dv, uvw = uv2 * w0, uvwm = uv2 * w0m;
        const double c1 = +vwum + uwvm + uvwm, c2 = +vwu - uwvm - uvwm;
        const double c3 = -vwum + uwv - uvwm, c4 = -vwu - uwv + uvwm;
        const double c5 = -vwum - uwvm + uvw, c6 = -vwu + uwvm - uvw;
        const double c7 = +vwum - uwv - uvw, c8 = +vwu + uwv + uvw;
        if (!PATHS) {
            const double e1 = extdirp[i1 - 1], e2 = extdirp[i2 - 1], e3 = extdirp[i3 - 1], e4 = extdirp[i4 - 1];
            const double e5 = extdirp[i1], e6 = extdirp[i2], e7 = extdirp[i3], e8 = extdirp[i4];
            const double a = (e1 * u0m + e2 * u0) * vmwm + (e3 * u0m + e4 * u0) * vwm
                           + (e5 * u0m + e6 * u0) * vmw + (e7 * u0m + e8 * u0) * vw;
            const double b = b1 * e1 + b2 * e2 + b3 * e3 + b4 * e4 + b5 * e5 + b6 * e6 + b7 * e7 + b8 * e8;
            const double c = c1 * e1 + c2 * e2 + c3 * e3 + c4 * e4 + c5 * e5 + c6 * e6 + c7 * e7 + c8 * e8;
            const double dd = du * dv * dw * (e2 + e3 + e5 + e8 - e1 - e4 - e6 - e7);
            path = path + so * (a + 0.5 * b + 0.3333333333333333 * c + 0.25 * dd);
            npp += 8;
        } else {
            const double a1 = u0m * vmwm, a2 = u0 * vmwm, a3 = u0m * vwm, a4 = u0 * vwm;
            const double a5 = u0m * vmw, a6 = u0 * vmw, a7 = u0m * vw, a8 = u0 * vw;
            const double q = 0.25 * du * dv * dw;
            float dv8[8];
            dv8[0] = (float)(so * (a1 + 0.5 * b1 + 0.3333333333333333 * c1 - q));
            dv8[1] = (float)(so * (a2 + 0.5 * b2 + 0.3333333333333333 * c2 + q));
            dv8[2] = (float)(so * (a3 + 0.5 * b3 + 0.3333333333333333 * c3 + q));
            dv8[3] = (float)(so * (a4 + 0.5 * b4 + 0.3333333333333333 * c4 - q));
            dv8[4] = (float)(so * (a5 + 0.5 * b5 + 0.3333333333333333 * c5 + q));
            dv8[5] = (float)(so * (a6 + 0.5 * b6 + 0.3333333333333333 * c6 - q));
            dv8[6] = (float)(so * (a7 + 0.5 * b7 + 0.3333333333333333 * c7 - q));
            dv8[7] = (float)(so * (a8 + 0.5 * b8 + 0.3333333333333333 * c8 + q));
            const int dp8[8] = {i1, i2, i3, i4, i1 + 1, i2 + 1, i3 + 1, i4 + 1};
            if (!sink.put(dp8, dv8)) return 5;
        }
        xe = xp + xoffs; ye = yp + yoffs; ze = zp;
    }
    return 0;
}

// Sinks of beam_walk<true>: the dense zero-terminated DPATH/DPTR lists of MAKE_DIRECT_DERIVATIVE, and a counter.
struct BeamNoSink { __device__ bool put(const int *, const float *) { return true; } };
struct BeamDenseSink {
    float *dpath; int *dptr; int cap, idp;
    __device__ bool put(const int *p, const float *v)
    {
        if (idp + 8 > cap) return false;
#pragma unroll
        for (int c = 0; c < 8; c++) { dpath[idp + c] = v[c]; dptr[idp + c] = p[c]; }
        idp += 8;
        return true;
    }
};
struct BeamCountSink { int n; __device__ bool put(const int *, const float *) { n += 8; return true; } };
