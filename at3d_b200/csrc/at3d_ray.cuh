// at3d_ray.cuh -- octet-per-ray march shared by the RENDER and gradient kernels (sm_100a).
//
// Eight lanes (an "octet") integrate one ray; a warp carries four neighbouring rays.  The cell walk
// (geometry, extinction interpolation, NTAU, transmission) is evaluated by every lane of the octet in
// FP64 with the exact operation order of INTEGRATE_1RAY (shdomsub2.f:2311-2743) /
// ADJOINT_INTEGRATE_1RAY (shdomsub4.f:3223-3967), so the visited-cell sequence and the
// sub-interval counts are bit-exact.  The spherical-harmonic sums of COMPUTE_SOURCE_1CELL are
// split over the 8 lanes (float4 loads of 128-byte aligned planar blocks, 3-step shuffle
// reduction); lane n owns corner n of the current cell (values in registers, exchanged by
// shuffles).  The per-ray direction table YLMDIR lives in shared memory.
#pragma once
#include "at3d_device.cuh"

struct RayErr {
    int code;       // 0 ok, 1 SO<0, 2 below domain, 3 not at boundary, 4 too many sub-intervals
    int ray;
};

__device__ __forceinline__ void set_err(RayErr *err, int code, int ray)
{
    if (atomicCAS(&err->code, 0, code) == 0) err->ray = ray;
}

// lanes of the calling thread's octet
struct Oct {
    unsigned m;     // lane mask of the octet inside its warp
    int ol;         // lane index inside the octet (0..7) = the cell corner this lane owns
    int ob;         // first lane of the octet in the warp
};
__device__ __forceinline__ Oct oct_id()
{
    Oct o;
    const int lane = threadIdx.x & 31;
    o.ol = lane & 7; o.ob = lane & ~7; o.m = 0xFFu << o.ob;
    return o;
}
__device__ __forceinline__ unsigned oct_ballot(const Oct &o, bool p)
{ return (__ballot_sync(o.m, p) >> o.ob) & 0xFFu; }
__device__ __forceinline__ unsigned oct_or(const Oct &o, unsigned v)
{
    v |= __shfl_xor_sync(o.m, v, 4); v |= __shfl_xor_sync(o.m, v, 2); v |= __shfl_xor_sync(o.m, v, 1);
    return v;
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

// Spherical-harmonic part of COMPUTE_SOURCE_1CELL (shdomsub2.f:2930-2947) for one planar block of
// `nsp` (multiple of 32) coefficients per Stokes component at `base`: the 8 lanes of the octet split
// j (float4 per lane and 128-byte line); returns the lane-partial sums (reduce with oct_sum).
template <int NST>
__device__ __forceinline__ void sh_dot_partial(const float *base, int nsp, const float *Ysh, int nlmp, const Oct &o,
                                               float (&a)[NST])
{
    base += o.ol * 4;
    const float *yb = Ysh + o.ol * 4;
#pragma unroll 4
    for (int j = 0; j < nsp; j += 32) {
        const float4 s = __ldg((const float4 *)(base + j));
        const float4 y = *(const float4 *)(yb + j);
        a[0] = fmaf(s.x, y.x, a[0]); a[0] = fmaf(s.y, y.y, a[0]);
        a[0] = fmaf(s.z, y.z, a[0]); a[0] = fmaf(s.w, y.w, a[0]);
        if (NST > 1) {
            const float4 q = __ldg((const float4 *)(base + nsp + j));
            const float4 u = __ldg((const float4 *)(base + 2 * nsp + j));
            const float4 y2 = *(const float4 *)(yb + 1 * nlmp + j);
            const float4 y5 = *(const float4 *)(yb + 2 * nlmp + j);
            const float4 y6 = *(const float4 *)(yb + 3 * nlmp + j);
            const float4 y3 = *(const float4 *)(yb + 4 * nlmp + j);
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
}

// The values lane n keeps for corner n of the current cell.
template <int NST>
struct Corner {
    int pt;                 // grid point (1-based), 0 = none yet
    float x, y, z, ext;     // GRIDPOS, TOTAL_EXT
    float src[NST];         // SRCEXT8(:,n)
};

// Records of a corner's grid point that is new in this cell, loaded by the lane that owns the corner
// (all new corners of a cell in parallel): coordinates/extinction, SH block and the exact
// single-scatter sum (shdomsub2.f:3008-3019) from the per-point list.
template <int NST>
__device__ __forceinline__ void load_corner(const DevState &S, int ip, const RayDir &rd, float &x, float &y, float &z,
                                            float &ext, int &soff, int &sns, float (&b)[NST])
{
    const float4 pr = __ldg(&S.ptrec[ip - 1]);
    const int4 ps = __ldg(&S.ptsrc[ip - 1]);
    x = pr.x; y = pr.y; z = pr.z; ext = pr.w;
    soff = ps.x; sns = ps.y & 0xFFFF;
    int cnt = (ps.y >> 16) & 0x7FFF;
    if (ps.y < 0) { sns = 0; cnt = 0; }       // dark point (build_ptsrc_kernel): SRCEXT8 is exactly 0
#pragma unroll
    for (int k = 0; k < NST; k++) b[k] = 0.0f;
    if (cnt > 0) {
        float sv[NST];
        ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, ps.z, rd, sv);
        const float coef = __int_as_float(ps.w);
#pragma unroll
        for (int k = 0; k < NST; k++) b[k] = fmaf(coef, sv[k], b[k]);
        for (int e = 1; e < cnt; e++) {
            const int2 en = __ldg(&S.ssent[(size_t)(ip - 1) * S.kmax + e]);
            ray_singscat<NST>(S.phasetab, S.nstphase, S.numphase, en.x, rd, sv);
            const float c2 = __int_as_float(en.y);
#pragma unroll
            for (int k = 0; k < NST; k++) b[k] = fmaf(c2, sv[k], b[k]);
        }
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
// element `corner` of GET_INTERP_KERNEL (same products in the same order)
__device__ __forceinline__ double interp_kernel_own(double u, double v, double w, int corner)
{
    const double fw = (corner & 4) ? w : (1 - w);
    const double fv = (corner & 2) ? v : (1 - v);
    const double fu = (corner & 1) ? u : (1 - u);
    return fw * fv * fu;
}
__device__ __forceinline__ double fcsum(const double *f, const float *a)
{
    return f[0] * a[0] + f[1] * a[1] + f[2] * a[2] + f[3] * a[3]
         + f[4] * a[4] + f[5] * a[5] + f[6] * a[6] + f[7] * a[7];
}

// Boundary radiance at the ray exit: FIND_BOUNDARY_RADIANCE (shdomsub2.f:2748-2863, REAL u,v) when
// GRADMODE is false, FIND_BOUNDARY_RADIANCE_GRAD (shdomsub4.f:2151-2347, DOUBLE u,v) otherwise.
// Lambertian surfaces (other BRDFs: surface_kernel, at3d_surface.cu).
template <int NST, bool GRADMODE>
__device__ int boundary_radiance(const DevState &S, double xb, double yb, float mu2, float phi2, float sky,
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
            if (S.sfcgridrad || S.srctype == 'T')
                rad[j][0] = dev_surface_emission(S, ibc, mu2, phi2) + __ldg(&S.bcrad[NST * (S.ntoppts + ibc - 1)]);
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

// GRIDPTR entry of the corner this lane owns
__device__ __forceinline__ int own_corner(const CellRec &c, int ol)
{
    int p = c.gp[0];
#pragma unroll
    for (int n = 1; n < 8; n++) if (ol == n) p = c.gp[n];
    return p;
}

// Refresh the corner values of the lanes for cell `c`.  Points shared with the previous cell are found
// with the reference's DONEFACE rule (shdomsub2.f:2395-2397, 2509-2515): after crossing a face normal to
// axis `jf`, corner n of the new cell can only coincide with corner n^bit of the old one; the ids decide.
// Reused values are bit-identical to a recomputation.  New corners: every owner lane loads its point's
// records at once, then the octet evaluates the SH sums point by point.  The thread-per-ray kernels
// (at3d_tray.cuh) use the same rule, so all marches evaluate the same (ray, point) pairs.
template <int NST>
__device__ __forceinline__ void refresh_corners(const DevState &S, const CellRec &c, const float *Ysh,
                                                const RayDir &rd, bool singlescatter, int jf /*0: first cell*/,
                                                const Oct &o, Corner<NST> &K, int &npt_eval, int &nsh_eval)
{
    const int myp = own_corner(c, o.ol);
    const int from = o.ol ^ (jf == 1 ? 1 : jf == 2 ? 2 : 4);
    const int cand = __shfl_sync(o.m, K.pt, from, 8);
    const bool hit = (jf != 0) && (cand == myp);
    const int src_lane = hit ? from : o.ol;
    K.x = __shfl_sync(o.m, K.x, src_lane, 8); K.y = __shfl_sync(o.m, K.y, src_lane, 8);
    K.z = __shfl_sync(o.m, K.z, src_lane, 8); K.ext = __shfl_sync(o.m, K.ext, src_lane, 8);
#pragma unroll
    for (int k = 0; k < NST; k++) K.src[k] = __shfl_sync(o.m, K.src[k], src_lane, 8);
    K.pt = myp;
    int soff = 0, sns = 0;
    float b[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) b[k] = 0.0f;
    if (S.viewsrc) {
        // all rays of this launch share their direction: SRCEXT of the point was evaluated once (view_source_kernel_oct)
        if (!hit) {
            const float4 pr = __ldg(&S.ptrec[myp - 1]);
            K.x = pr.x; K.y = pr.y; K.z = pr.z; K.ext = pr.w;
#pragma unroll
            for (int k = 0; k < NST; k++) K.src[k] = __ldg(&S.viewsrc[k + NST * (size_t)(myp - 1)]);
        }
        const unsigned nw = oct_ballot(o, !hit);
        npt_eval += __popc(nw);
        int nsm = hit ? 0 : (__ldg(&S.ptsrc[myp - 1]).y & 0xFFFF);         // work counters as in the generic path
        nsm += __shfl_xor_sync(o.m, nsm, 4, 8); nsm += __shfl_xor_sync(o.m, nsm, 2, 8); nsm += __shfl_xor_sync(o.m, nsm, 1, 8);
        nsh_eval += nsm;
        return;
    }
    if (!hit) load_corner<NST>(S, myp, rd, K.x, K.y, K.z, K.ext, soff, sns, b);
    unsigned need = oct_ballot(o, !hit);
    npt_eval += __popc(need);
    while (need) {
        const int n = __ffs(need) - 1;
        const int ipn = __shfl_sync(o.m, myp, n, 8);
        const int off = __shfl_sync(o.m, soff, n, 8), ns = __shfl_sync(o.m, sns, n, 8);
        float a[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) a[k] = 0.0f;
        if (!singlescatter) {
            sh_dot_partial<NST>(S.shsrc + off, AT3D_SHPAD(ns), Ysh, S.nlmp, o, a);
#pragma unroll
            for (int k = 0; k < NST; k++) a[k] = oct_sum(o.m, a[k]);
        }
        // new corners that are the same grid point (zero-width open-boundary cells) share the evaluation
        const bool mine = !hit && (myp == ipn);
        const unsigned same = oct_ballot(o, mine);
        nsh_eval += ns * __popc(same);
        if (mine) {
#pragma unroll
            for (int k = 0; k < NST; k++) K.src[k] = (a[k] + b[k]) * K.ext;
        }
        need &= ~same;
    }
}

// Forward integration of one ray by one octet.
// MODES bit 0: INTEGRATE_1RAY arithmetic (result radA, nsubA counts every sub-interval);
// MODES bit 1: the forward part of ADJOINT_INTEGRATE_1RAY (GET_INTERP_KERNEL weights, EXT0=EXTN on
//              the last sub-interval, no MAXCELLSCROSS stop; result radB, nsubB counts EXT!=0).
// Both share the walk (it depends on the geometry only) and the corner evaluations.  The record of
// the next cell is requested as soon as the exit face is known, before the sub-interval loop.
template <int NST, int MODES>
__device__ int march_forward(const DevState &S, const float *Ysh, const RayDir &rd, double mu2,
                             double x0, double y0, double z0, float sky, bool correctinterpolate,
                             bool singlescatter, bool nosurface, int maxsub, const Oct &o,
                             double (&radA)[NST], double (&radB)[NST],
                             int *trace_cells, int trace_cap, int &ntrace, int &nsubA, int &nsubB, int &nptB)
{
    double xe = x0, ye = y0, ze = z0, trA = 1.0, trB = 1.0;
    float ext1A = 0.0f, srcext1A[NST], ext1B = 0.0f, srcext1B[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) { radA[k] = 0.0; radB[k] = 0.0; srcext1A[k] = 0.0f; srcext1B[k] = 0.0f; }
    const int p1c = cell_gp(S, 1, 1), p8c = cell_gp(S, 1, 8);
    const double eps = (double)(1.0e-5f * (pt_coord(S, p8c, 3) - pt_coord(S, p1c, 3)));
    const int maxcellscross = 500 * max(S.nx, max(S.ny, S.nz));
    int icell = dev_locate_grid_cell(S, xe, ye, ze);
    int iface = 0, ngrid = 0, npt_eval = 0, nsh_eval = 0, jf = 0;
    bool doneA = !(MODES & 1), doneB = !(MODES & 2);
    nptB = 0;
    Corner<NST> K;
    K.pt = 0; K.x = K.y = K.z = K.ext = 0.0f;
#pragma unroll
    for (int k = 0; k < NST; k++) K.src[k] = 0.0f;
    ntrace = 0; nsubA = 0; nsubB = 0;
    CellRec c;
    if (icell > 0) c = load_cell(S, icell);
    while (!(doneA && doneB) && icell > 0) {
        ngrid++;
        if (trace_cells && o.ol == 0 && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        const int ne0 = npt_eval;
        refresh_corners<NST>(S, c, Ysh, rd, singlescatter, jf, o, K, npt_eval, nsh_eval);
        if ((MODES & 2) && !doneB) nptB += npt_eval - ne0;
        float e8[8], s8[NST][8];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            e8[n] = __shfl_sync(o.m, K.ext, n, 8);
#pragma unroll
            for (int k = 0; k < NST; k++) s8[k][n] = __shfl_sync(o.m, K.src[k], n, 8);
        }
        const float q1x = __shfl_sync(o.m, K.x, 0, 8), q1y = __shfl_sync(o.m, K.y, 0, 8), q1z = __shfl_sync(o.m, K.z, 0, 8);
        const float q8x = __shfl_sync(o.m, K.x, 7, 8), q8y = __shfl_sync(o.m, K.y, 7, 8), q8z = __shfl_sync(o.m, K.z, 7, 8);
        const float qox = __shfl_sync(o.m, K.x, 8 - rd.ioct, 8), qoy = __shfl_sync(o.m, K.y, 8 - rd.ioct, 8),
                    qoz = __shfl_sync(o.m, K.z, 8 - rd.ioct, 8);
        const double delx = (double)(q8x - q1x), dely = (double)(q8y - q1y), delz = (double)(q8z - q1z);
        const double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        const double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        const double invdelz = 1.0 / delz;
        double u = (xe - q1x) * invdelx, v = (ye - q1y) * invdely, w = (ze - q1z) * invdelz;
        double fc[8];
        if ((MODES & 2) && !doneB) {
            interp_kernel(u, v, w, fc);
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1B[k] = (float)fcsum(fc, s8[k]);
            srcext1B[0] = fmaxf(0.0f, srcext1B[0]);
            ext1B = (float)fcsum(fc, e8);
        }
        if ((MODES & 1) && !doneA && (correctinterpolate || ngrid == 1)) {
#pragma unroll
            for (int k = 0; k < NST; k++) srcext1A[k] = (float)trilerp(s8[k], u, v, w);
            srcext1A[0] = fmaxf(0.0f, srcext1A[0]);
            ext1A = (float)trilerp(e8, u, v, w);
        }
        const bool ipinx = DBTEST(c.flags, 0) &&
            !(DBTEST(S.bcflag, 0) && ((rd.cx > 0 && xe < rd.xm) || (rd.cx < 0 && xe > rd.xm)));
        const bool ipiny = DBTEST(c.flags, 1) &&
            !(DBTEST(S.bcflag, 1) && ((rd.cy > 0 && ye < rd.ym) || (rd.cy < 0 && ye > rd.ym)));
        const double sox = ipinx ? (double)1.0e20f : (qox - xe) * rd.cxinv;
        const double soy = ipiny ? (double)1.0e20f : (qoy - ye) * rd.cyinv;
        const double soz = (qoz - ze) * rd.czinv;
        const double so = fmin(fmin(sox, soy), soz);
        if (so < -eps) return 1;
        double xn = xe + so * rd.cx, yn = ye + so * rd.cy, zn = ze + so * rd.cz;
        // ---- exit face and next cell (shdomsub2.f:2668-2716); its record is requested now ----
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
        CellRec cn = c;
        float snap = 0.0f;
        if (inextcell > 0) {
            cn = load_cell(S, inextcell);
            int pn = cn.gp[0];
#pragma unroll
            for (int n = 1; n < 8; n++) if (rd.ioct - 1 == n) pn = cn.gp[n];
            snap = pt_coord(S, pn, jface);
        }
        u = (xn - q1x) * invdelx; v = (yn - q1y) * invdely; w = (zn - q1z) * invdelz;
        if ((MODES & 1) && !doneA) {
            const float extn = (float)trilerp(e8, u, v, w);
            const double taugrid = so * 0.5f * (ext1A + extn);
            int ntau = 1 + (int)(taugrid / S.tautol);
            if (ntau < 1) ntau = 1;
            const double dels = so / ntau;
            for (int it = 1; it <= ntau; it++) {
                const double s = it * dels;
                const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
                const double ui = (xi - q1x) * invdelx, vi = (yi - q1y) * invdely, wi = (zi - q1z) * invdelz;
                const float ext0 = (float)trilerp(e8, ui, vi, wi);
                float srcext0[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) srcext0[k] = (float)trilerp(s8[k], ui, vi, wi);
                srcext0[0] = fmaxf(0.0f, srcext0[0]);
                const double ext = (double)(0.5f * (ext0 + ext1A));
                if (ext != 0.0) {
                    const double tau = ext * dels;
                    const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                    const double transcell = 1.0f - abscell;
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        const double src = (0.5f * (srcext0[k] + srcext1A[k])
                            + 0.08333333333f * (ext0 * srcext1A[k] - ext1A * srcext0[k]) * dels
                              * (1.0f - 0.05f * (ext1A - ext0) * dels)) / ext;
                        radA[k] = radA[k] + trA * src * abscell;
                    }
                    trA = trA * transcell;
                }
                nsubA++;
                ext1A = ext0;
#pragma unroll
                for (int k = 0; k < NST; k++) srcext1A[k] = srcext0[k];
            }
        }
        if ((MODES & 2) && !doneB) {
            float extn;
            { double fcn[8]; interp_kernel(u, v, w, fcn); extn = (float)fcsum(fcn, e8); }
            const double taugrid = so * 0.5f * (ext1B + extn);
            int ntau = 1 + (int)(taugrid / S.tautol);
            if (ntau < 1) ntau = 1;
            const double dels = so / ntau;
            for (int it = 1; it <= ntau; it++) {
                const double s = it * dels;
                const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
                const double ui = (xi - q1x) * invdelx, vi = (yi - q1y) * invdely, wi = (zi - q1z) * invdelz;
                interp_kernel(ui, vi, wi, fc);
                float srcext0[NST];
#pragma unroll
                for (int k = 0; k < NST; k++) srcext0[k] = (float)fcsum(fc, s8[k]);
                const float ext0 = (it != ntau) ? (float)fcsum(fc, e8) : extn;
                srcext0[0] = fmaxf(0.0f, srcext0[0]);
                const double ext = (double)(0.5f * (ext0 + ext1B));
                if (ext != 0.0) {
                    const double tau = ext * dels;
                    const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                    const double transcell = 1.0f - abscell;
#pragma unroll
                    for (int k = 0; k < NST; k++) {
                        const double src = (0.5f * (srcext0[k] + srcext1B[k])
                            + 0.08333333333f * (ext0 * srcext1B[k] - ext1B * srcext0[k]) * dels
                              * (1.0f - 0.05f * (ext1B - ext0) * dels)) / ext;
                        radB[k] = radB[k] + trB * src * abscell;
                    }
                    trB = trB * transcell;
                    nsubB++;
                    if (nsubB + 1 > maxsub) return 4;
                }
                ext1B = ext0;
#pragma unroll
                for (int k = 0; k < NST; k++) srcext1B[k] = srcext0[k];
            }
        }
        if (inextcell > 0) {
            if (jface == 1) xn = (double)snap;
            else if (jface == 2) yn = (double)snap;
            else zn = (double)snap;
        }
        const bool atbnd = (inextcell == 0 && iface >= 5);
        if ((MODES & 1) && !doneA) {
            if (trA < S.transcut || ngrid > maxcellscross) doneA = true;
            else if (atbnd) {
                doneA = true;
                if (rd.hit && !((float)mu2 < 0.0f)) {
                    // general BRDF: the reflected radiance is added by surface_kernel (at3d_surface.cu)
                    if (!nosurface && o.ol == 0) {
                        SurfHit h;
                        h.xb = xn; h.yb = yn; h.transmit = trA; h.icell = ic; h.kface = kface;
#pragma unroll
                        for (int k = 0; k < 3; k++) h.rad[k] = k < NST ? radA[k < NST ? k : 0] : 0.0;
                        *rd.hit = h;
                    }
                } else {
                float radbnd[NST];
                const int e = boundary_radiance<NST, false>(S, xn, yn, (float)mu2, rd.phi2, sky, ic, kface, radbnd,
                                                            nullptr, nullptr, nullptr);
                if (e) return e;
                if (!nosurface) {
#pragma unroll
                    for (int k = 0; k < NST; k++) radA[k] = radA[k] + trA * radbnd[k];
                }
                }
            }
        }
        if ((MODES & 2) && !doneB) {
            if (trB < S.transcut) doneB = true;
            else if (atbnd) {
                doneB = true;
                float radbnd[NST];
                const int e = boundary_radiance<NST, true>(S, xn, yn, (float)mu2, rd.phi2, sky, ic, kface, radbnd,
                                                           nullptr, nullptr, nullptr);
                if (e) return e;
                if (!nosurface) {
#pragma unroll
                    for (int k = 0; k < NST; k++) radB[k] = radB[k] + trB * radbnd[k];
                }
            }
        }
        if (!atbnd) { icell = inextcell; c = cn; jf = jface; }
        xe = xn; ye = yn; ze = zn;
    }
    if (S.counts && o.ol == 0) {
        atomicAdd(&S.counts[0], (unsigned long long)ntrace);
        atomicAdd(&S.counts[1], (unsigned long long)npt_eval);
        atomicAdd(&S.counts[2], (unsigned long long)nsh_eval);
        atomicAdd(&S.counts[4], (unsigned long long)((MODES & 1) ? nsubA : nsubB));
        atomicAdd(&S.counts[5], 1ull);
    }
    return 0;
}
