// at3d_ray.cuh -- warp-per-ray march shared by the RENDER and gradient kernels.
//
// One warp integrates one ray.  The per-ray direction tables (YLMDIR) live in shared memory,
// lanes split the spherical-harmonic sums (float4 coalesced loads of the planar, 16-byte aligned
// source blocks), and the cell-walk geometry is evaluated redundantly (warp-uniform) in FP64
// with the exact operation order of INTEGRATE_1RAY (shdomsub2.f:2311-2743) /
// ADJOINT_INTEGRATE_1RAY (shdomsub4.f:3223-3967) so that the visited-cell sequence is bit-exact.
#pragma once
#include "at3d_device.cuh"

// Per-warp shared scratch for the 8 corner values of the current and previous cell.
template <int NST>
struct CornerCache {
    int pt[8];
    float ext[8];
    float src[NST][8];
    float ss[NST][8];
};

struct RayErr {
    int code;       // 0 ok, 1 SO<0, 2 below domain, 3 not at boundary, 4 too many sub-intervals
    int ray;
};

__device__ __forceinline__ void set_err(RayErr *err, int code, int ray)
{
    if (atomicCAS(&err->code, 0, code) == 0) err->ray = ray;
}

// Evaluate source*extinction (and the exact single-scatter part separately) of one grid point in
// the ray direction: COMPUTE_SOURCE_1CELL (shdomsub2.f:2911-3038) for one corner, with the
// TMS-corrected SH source prepared by prep_source_kernel.  Returns reduced values on every lane.
template <int NST>
__device__ __forceinline__ void eval_point(const DevState &S, int ip, const float *Ysh, const RayDir &rd,
                                           bool singlescatter, float &ext, float (&src)[NST], float (&ss)[NST])
{
    const int lane = lane_id();
    const float4 pr = __ldg(&S.ptrec[ip - 1]);
    const int2 sr = __ldg(&S.srcrec[ip - 1]);
    ext = pr.w;
    const int nsp = (sr.y + 3) & ~3;
    float acc[NST], sacc[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { acc[k] = 0.0f; sacc[k] = 0.0f; }
    if (!singlescatter) {
        const float *base = S.shsrc + sr.x;
        for (int j4 = lane * 4; j4 < nsp; j4 += 128) {
            const float4 s = __ldg((const float4 *)(base + j4));
            const float4 y = *(const float4 *)(Ysh + j4);
            acc[0] = fmaf(s.x, y.x, acc[0]); acc[0] = fmaf(s.y, y.y, acc[0]);
            acc[0] = fmaf(s.z, y.z, acc[0]); acc[0] = fmaf(s.w, y.w, acc[0]);
            if (NST > 1) {
                const float4 q = __ldg((const float4 *)(base + nsp + j4));
                const float4 u = __ldg((const float4 *)(base + 2 * nsp + j4));
                const float4 y2 = *(const float4 *)(Ysh + 1 * S.nlmp + j4);
                const float4 y5 = *(const float4 *)(Ysh + 2 * S.nlmp + j4);
                const float4 y6 = *(const float4 *)(Ysh + 3 * S.nlmp + j4);
                const float4 y3 = *(const float4 *)(Ysh + 4 * S.nlmp + j4);
                acc[1] = fmaf(q.x, y2.x, acc[1]); acc[1] = fmaf(u.x, y5.x, acc[1]);
                acc[1] = fmaf(q.y, y2.y, acc[1]); acc[1] = fmaf(u.y, y5.y, acc[1]);
                acc[1] = fmaf(q.z, y2.z, acc[1]); acc[1] = fmaf(u.z, y5.z, acc[1]);
                acc[1] = fmaf(q.w, y2.w, acc[1]); acc[1] = fmaf(u.w, y5.w, acc[1]);
                acc[NST - 1] = fmaf(q.x, y6.x, acc[NST - 1]); acc[NST - 1] = fmaf(u.x, y3.x, acc[NST - 1]);
                acc[NST - 1] = fmaf(q.y, y6.y, acc[NST - 1]); acc[NST - 1] = fmaf(u.y, y3.y, acc[NST - 1]);
                acc[NST - 1] = fmaf(q.z, y6.z, acc[NST - 1]); acc[NST - 1] = fmaf(u.z, y3.z, acc[NST - 1]);
                acc[NST - 1] = fmaf(q.w, y6.w, acc[NST - 1]); acc[NST - 1] = fmaf(u.w, y3.w, acc[NST - 1]);
            }
        }
    }
    const int cnt = __ldg(&S.sscount[ip - 1]);
    for (int k = lane; k < cnt; k += 32) {
        const int2 e = __ldg(&S.ssent[(size_t)(ip - 1) * S.kmax + k]);
        const float coef = __int_as_float(e.y);
        const float *p0 = S.phasetab + (size_t)S.nstphase * ((e.x - 1) + (size_t)S.numphase * (rd.j - 1));
        const float *p1 = p0 + (size_t)S.nstphase * S.numphase;
        const float a = (1 - rd.f) * __ldg(p0) + rd.f * __ldg(p1);
        sacc[0] = fmaf(coef, a, sacc[0]);
        if (NST > 1) {
            const float b1 = (1 - rd.f) * __ldg(p0 + 1) + rd.f * __ldg(p1 + 1);
            const float q = (float)(b1 * rd.cos22);
            const float u = (float)(b1 * rd.sin22);
            sacc[1] = fmaf(coef, q, sacc[1]);
            sacc[NST - 1] = fmaf(coef, u, sacc[NST - 1]);
        }
    }
#pragma unroll
    for (int k = 0; k < NST; k++) {
        const float a = warp_sum(acc[k]);
        const float b = warp_sum(sacc[k]);
        ss[k] = b * ext;
        src[k] = (a + b) * ext;
    }
}

// nested-lerp trilinear interpolation of INTEGRATE_1RAY (shdomsub2.f:2563-2571), double weights
__device__ __forceinline__ double trilerp(const float *a, double u, double v, double w)
{
    return (1 - w) * ((1 - v) * ((1 - u) * a[0] + u * a[1]) + v * ((1 - u) * a[2] + u * a[3]))
         + w * ((1 - v) * ((1 - u) * a[4] + u * a[5]) + v * ((1 - u) * a[6] + u * a[7]));
}
// GET_INTERP_KERNEL (shdomsub5.f:1526-1533)
__device__ __forceinline__ void interp_kernel(double u, double v, double w, double *f)
{
    f[0] = (1 - w) * (1 - v) * (1 - u);
    f[1] = (1 - w) * (1 - v) * u;
    f[2] = (1 - w) * v * (1 - u);
    f[3] = (1 - w) * v * u;
    f[4] = w * (1 - v) * (1 - u);
    f[5] = w * (1 - v) * u;
    f[6] = w * v * (1 - u);
    f[7] = w * v * u;
}
__device__ __forceinline__ double fcsum(const double *f, const float *a)
{
    return f[0] * a[0] + f[1] * a[1] + f[2] * a[2] + f[3] * a[3]
         + f[4] * a[4] + f[5] * a[5] + f[6] * a[6] + f[7] * a[7];
}

// Boundary radiance at the ray exit: FIND_BOUNDARY_RADIANCE (shdomsub2.f:2748-2863, REAL u,v) when
// GRADMODE is false, FIND_BOUNDARY_RADIANCE_GRAD (shdomsub4.f:2151-2347, DOUBLE u,v) otherwise.
// Lambertian surfaces; solar source (the surface-emission term is identically zero).
template <int NST, bool GRADMODE>
__device__ int boundary_radiance(const DevState &S, double xb, double yb, float mu2, float sky,
                                 int icell, int kface, float (&radbnd)[NST],
                                 int *boundpts, double *boundinterp, double *dirrad1)
{
    const int gf[6][4] = {{1,3,5,7},{2,4,6,8},{1,2,5,6},{3,4,7,8},{1,2,3,4},{5,6,7,8}};
    float x[4], y[4], rad[4][NST];
    const float opi = 1.0f / acosf(-1.0f);
    for (int j = 0; j < 4; j++) {
        const int ip = cell_gp(S, icell, gf[kface - 1][j]);
        if (boundpts) { boundpts[j] = ip; dirrad1[j] = 0.0; }
        x[j] = pt_coord(S, ip, 1);
        y[j] = pt_coord(S, ip, 2);
        if (mu2 < 0.0f) {
            const int ibc = dev_bc_search(S.bcptr, S.ntoppts, ip);
            if (!ibc) return 3;
            // RENDER fills BCRAD(:,1:NTOPPTS) with the (I only) sky radiance per ray (shdomsub4.f:238-248)
            rad[j][0] = sky;
#pragma unroll
            for (int k = 1; k < NST; k++) rad[j][k] = 0.0f;
        } else {
            const int ibc = dev_bc_search(S.bcptr + S.maxnbc, S.nbotpts, ip);
            if (!ibc) return 3;
#pragma unroll
            for (int k = 0; k < NST; k++)
                rad[j][k] = 0.0f + __ldg(&S.bcrad[k + NST * (S.ntoppts + ibc - 1)]);
            if (GRADMODE && boundpts) {
                if (S.sfctype0 == 'V')
                    dirrad1[j] = (double)(opi * __ldg(&S.sfcgridparms[1 + S.nsfcpar * (ibc - 1)]) * __ldg(&S.dirflux[ip - 1]));
                else if (S.sfctype0 == 'F')
                    dirrad1[j] = (double)(opi * S.gndalbedo * __ldg(&S.dirflux[ip - 1]));
            }
        }
    }
    if (!GRADMODE) {
        float u, v;
        if (x[1] - x[0] > 0.0f) u = (float)((xb - x[0]) / (x[1] - x[0])); else u = 0.0f;
        if (y[2] - y[0] > 0.0f) v = (float)((yb - y[0]) / (y[2] - y[0])); else v = 0.0f;
#pragma unroll
        for (int k = 0; k < NST; k++)
            radbnd[k] = (1 - u) * (1 - v) * rad[0][k] + u * (1 - v) * rad[1][k]
                        + (1 - u) * v * rad[2][k] + u * v * rad[3][k];
    } else {
        double u, v;
        if (x[1] - x[0] > 0.0f) u = (xb - x[0]) / (x[1] - x[0]); else u = 0.0;
        if (y[2] - y[0] > 0.0f) v = (yb - y[0]) / (y[2] - y[0]); else v = 0.0;
        if (boundinterp) {
            boundinterp[0] = (1 - u) * (1 - v);
            boundinterp[1] = u * (1 - v);
            boundinterp[2] = (1 - u) * v;
            boundinterp[3] = u * v;
        }
#pragma unroll
        for (int k = 0; k < NST; k++)
            radbnd[k] = (float)((1 - u) * (1 - v) * rad[0][k] + u * (1 - v) * rad[1][k]
                                + (1 - u) * v * rad[2][k] + u * v * rad[3][k]);
    }
    return 0;
}

// Refresh the 8 corner values of cell `c` in the per-warp cache: values of points shared with the
// previous cell are reused (they are bit-identical to a recomputation, so this is the reference's
// OLDIPTS/DONEFACE shortcut, shdomsub2.f:2509-2515,2914-2919, without its slot restriction).
template <int NST>
__device__ __forceinline__ void refresh_corners(const DevState &S, const CellRec &c, CornerCache<NST> *cc,
                                                const float *Ysh, const RayDir &rd, bool singlescatter,
                                                bool first, int &npt_eval, int &nsh_eval)
{
    const int lane = lane_id();
    int myp = 0, hit = -1;
    float oext = 0.0f, osrc[NST], oss[NST];
    if (lane < 8) {
        myp = c.gp[0];
#pragma unroll
        for (int n = 1; n < 8; n++) if (lane == n) myp = c.gp[n];
        if (!first) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (cc->pt[k] == myp) hit = k;
        }
        if (hit >= 0) {
            oext = cc->ext[hit];
#pragma unroll
            for (int k = 0; k < NST; k++) { osrc[k] = cc->src[k][hit]; oss[k] = cc->ss[k][hit]; }
        }
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
    unsigned need = __ballot_sync(FULLMASK, lane < 8 && hit < 0);
    // duplicate corners inside one cell (IP-mode / open-boundary end cells): evaluate once
    while (need) {
        const int n = __ffs(need) - 1;
        need &= need - 1;
        const int ip = __shfl_sync(FULLMASK, myp, n);
        float ext, src[NST], ss[NST];
        eval_point<NST>(S, ip, Ysh, rd, singlescatter, ext, src, ss);
        npt_eval++; nsh_eval += __ldg(&S.srcrec[ip - 1]).y;
        const unsigned same = __ballot_sync(FULLMASK, lane < 8 && myp == ip);
        if (lane < 8 && myp == ip) {
            cc->ext[lane] = ext;
#pragma unroll
            for (int k = 0; k < NST; k++) { cc->src[k][lane] = src[k]; cc->ss[k][lane] = ss[k]; }
        }
        need &= ~same;
    }
    __syncwarp();
}

// Integrate one ray.  MODE 0: INTEGRATE_1RAY arithmetic; MODE 1: the forward part of
// ADJOINT_INTEGRATE_1RAY (GET_INTERP_KERNEL weights, EXT0=EXTN on the last sub-interval,
// no MAXCELLSCROSS stop).  Returns the radiance in rad[] (all lanes) and an error code.
template <int NST, int MODE>
__device__ int march_ray(const DevState &S, CornerCache<NST> *cc, const float *Ysh, const RayDir &rd,
                         double mu2, double x0, double y0, double z0, float sky,
                         bool correctinterpolate, bool singlescatter, bool nosurface, int maxsub,
                         double (&rad)[NST], int *trace_cells, int trace_cap, int &ntrace, int &nsub)
{
    const int lane = lane_id();
    double xe = x0, ye = y0, ze = z0, transmit = 1.0;
    float ext1 = 0.0f, srcext1[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { rad[k] = 0.0; srcext1[k] = 0.0f; }
    const int p1c = cell_gp(S, 1, 1), p8c = cell_gp(S, 1, 8);
    const double eps = (double)(1.0e-5f * (pt_coord(S, p8c, 3) - pt_coord(S, p1c, 3)));
    const int maxcellscross = 500 * max(S.nx, max(S.ny, S.nz));
    int icell = dev_locate_grid_cell(S, xe, ye, ze);
    int iface = 0, ngrid = 0, npt_eval = 0, nsh_eval = 0;
    bool done = false, first = true;
    ntrace = 0; nsub = 0;
    while (!done && icell > 0) {
        ngrid++;
        if (trace_cells && lane == 0 && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        const CellRec c = load_cell(S, icell);
        refresh_corners<NST>(S, c, cc, Ysh, rd, singlescatter, first, npt_eval, nsh_eval);
        first = false;
        float e8[8], s8[NST][8];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            e8[n] = cc->ext[n];
#pragma unroll
            for (int k = 0; k < NST; k++) s8[k][n] = cc->src[k][n];
        }
        const float4 q1 = __ldg(&S.ptrec[c.gp[0] - 1]);
        const float4 q8 = __ldg(&S.ptrec[c.gp[7] - 1]);
        double delx = (double)(q8.x - q1.x), dely = (double)(q8.y - q1.y), delz = (double)(q8.z - q1.z);
        double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        double invdelz = 1.0 / delz;
        double u = (xe - q1.x) * invdelx, v = (ye - q1.y) * invdely, w = (ze - q1.z) * invdelz;
        double fc[8];
        if (MODE == 1) {
            interp_kernel(u, v, w, fc);
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1[k] = (float)fcsum(fc, s8[k]);
            srcext1[0] = fmaxf(0.0f, srcext1[0]);
            ext1 = (float)fcsum(fc, e8);
        } else if (correctinterpolate || ngrid == 1) {
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1[k] = (float)trilerp(s8[k], u, v, w);
            srcext1[0] = fmaxf(0.0f, srcext1[0]);
            ext1 = (float)trilerp(e8, u, v, w);
        }
        const bool ipinx = DBTEST(c.flags, 0) &&
            !(DBTEST(S.bcflag, 0) && ((rd.cx > 0 && xe < rd.xm) || (rd.cx < 0 && xe > rd.xm)));
        const bool ipiny = DBTEST(c.flags, 1) &&
            !(DBTEST(S.bcflag, 1) && ((rd.cy > 0 && ye < rd.ym) || (rd.cy < 0 && ye > rd.ym)));
        int iopp = c.gp[0];
#pragma unroll
        for (int n = 1; n < 8; n++) if (8 - rd.ioct == n) iopp = c.gp[n];
        const float4 qo = __ldg(&S.ptrec[iopp - 1]);
        double sox = ipinx ? (double)1.0e20f : (qo.x - xe) * rd.cxinv;
        double soy = ipiny ? (double)1.0e20f : (qo.y - ye) * rd.cyinv;
        double soz = (qo.z - ze) * rd.czinv;
        double so = fmin(fmin(sox, soy), soz);
        if (so < -eps) return 1;
        double xn = xe + so * rd.cx, yn = ye + so * rd.cy, zn = ze + so * rd.cz;
        u = (xn - q1.x) * invdelx; v = (yn - q1.y) * invdely; w = (zn - q1.z) * invdelz;
        float extn;
        if (MODE == 1) { double fcn[8]; interp_kernel(u, v, w, fcn); extn = (float)fcsum(fcn, e8); }
        else extn = (float)trilerp(e8, u, v, w);
        const double taugrid = so * 0.5f * (ext1 + extn);
        int ntau = 1 + (int)(taugrid / S.tautol);
        if (ntau < 1) ntau = 1;
        const double dels = so / ntau;
        for (int it = 1; it <= ntau; it++) {
            const double s = it * dels;
            const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
            u = (xi - q1.x) * invdelx; v = (yi - q1.y) * invdely; w = (zi - q1.z) * invdelz;
            float ext0, srcext0[NST];
            if (MODE == 1) {
                interp_kernel(u, v, w, fc);
#pragma unroll
                for (int k = 0; k < NST; k++) srcext0[k] = (float)fcsum(fc, s8[k]);
                ext0 = (it != ntau) ? (float)fcsum(fc, e8) : extn;
            } else {
                ext0 = (float)trilerp(e8, u, v, w);
#pragma unroll
                for (int k = 0; k < NST; k++) srcext0[k] = (float)trilerp(s8[k], u, v, w);
            }
            srcext0[0] = fmaxf(0.0f, srcext0[0]);
            const double ext = (double)(0.5f * (ext0 + ext1));
            if (ext != 0.0) {
                const double tau = ext * dels;
                const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                const double transcell = 1.0f - abscell;
#pragma unroll
                for (int k = 0; k < NST; k++) {
                    const double src = (0.5f * (srcext0[k] + srcext1[k])
                        + 0.08333333333f * (ext0 * srcext1[k] - ext1 * srcext0[k]) * dels
                          * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext;
                    rad[k] = rad[k] + transmit * src * abscell;
                }
                transmit = transmit * transcell;
                if (MODE == 1) { nsub++; if (nsub + 1 > maxsub) return 4; }
            }
            if (MODE == 0) nsub++;
            ext1 = ext0;
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1[k] = srcext0[k];
        }
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
        if (transmit < S.transcut || (MODE == 0 && ngrid > maxcellscross)) {
            done = true;
        } else if (inextcell == 0 && iface >= 5) {
            done = true;
            float radbnd[NST];
            const int e = boundary_radiance<NST, MODE == 1>(S, xn, yn, (float)mu2, sky, ic, kface, radbnd,
                                                            nullptr, nullptr, nullptr);
            if (e) return e;
            if (!nosurface) {
#pragma unroll
                for (int k = 0; k < NST; k++) rad[k] = rad[k] + transmit * radbnd[k];
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
        atomicAdd(&S.counts[4], (unsigned long long)nsub);
        atomicAdd(&S.counts[5], 1ull);
    }
    return 0;
}
