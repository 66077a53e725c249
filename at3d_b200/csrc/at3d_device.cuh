// at3d_device.cuh -- device-side state layout and shared device functions (sm_100a).
//
// Data layout in HBM (DESIGN.md "Data layout"): the reference's pointer-chasing arrays are
// re-packed once per solved state into records sized for single 32/64-byte sector reads:
//   cellrec : 64 B per cell  = GRIDPTR(8) | NEIGHPTR(6) | TREEPTR(2,.) | CELLFLAGS
//   ptrec   : 16 B per point = GRIDPOS(3) | TOTAL_EXT
//   srcrec  :  8 B per point = offset/ns into shsrc (16-byte aligned planar SH blocks)
//   ssent   :  8 B per entry = (phase-table index, DA*w/(1-F)) single-scatter list per point
// Index CONTENTS stay 1-based as in the reference so cell/point ids compare bit-exactly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AT3D_MAX_NSTOKES 3
// Ray kernels: 8 lanes ("octet") integrate one ray, 4 rays per warp.
#define AT3D_OCT 8
#define AT3D_RAY_THREADS 128
#define AT3D_RAYS_PER_BLOCK (AT3D_RAY_THREADS / AT3D_OCT)
// resident blocks per SM the ray kernels are compiled for (register budget): scalar / polarized
#ifndef AT3D_MINB_FWD1
#define AT3D_MINB_FWD1 3
#endif
#ifndef AT3D_MINB_FWD3
#define AT3D_MINB_FWD3 2
#endif
#ifndef AT3D_MINB_ADJ1
#define AT3D_MINB_ADJ1 3
#endif
#ifndef AT3D_MINB_ADJ3
#define AT3D_MINB_ADJ3 2
#endif
// planar SH blocks are padded to full 128-byte lines (32 floats): an octet reads them with
// 4 float4 per lane and iteration, no tail handling
#define AT3D_SHPAD(ns) (((ns) + 31) & ~31)

struct DevState {
    int nstokes, nstleg, nx, ny, nz, npts, ncells;
    int ml, mm, nlm, nlmp, nleg, numphase, npart, maxnmicro, nq;
    int bcflag, ipflag, nmu, nphi0max, maxnbc, ntoppts, nbotpts, nsfcpar;
    int nscatangle, nstphase, deltam, srctype, sfctype0, sfctype1, interp_new;
    int kmax;                 // single-scatter entries stride per point
    int ny_comp;              // number of YLMDIR components staged per ray (1 or 5)
    float solarmu, solaraz, gndalbedo, phasemax;
    double tautol, transcut;
    const int4 *cellrec;      // [ncells*4]
    const float4 *ptrec;      // [npts]
    const int2 *srcrec;       // [npts] (offset in floats, ns)
    const int4 *ptsrc;        // [npts] (offset, ns | count<<16, first single-scatter entry): one load per new corner
    const float *shsrc;       // TMS-corrected source, planar per point, padded to 4
    const int2 *radrec;       // [npts] radiance SH (gradient)
    const float *shrad;
    const int *sscount;       // [npts]
    const int2 *ssent;        // [npts*kmax] (iph, __float_as_int(coef))
    const float *phasetab;    // [nstphase,numphase,nscatangle]
    const float *xgrid, *ygrid, *zgrid;
    const int *bcptr;         // [maxnbc,2]
    const float *bcrad;       // [nstokes, ntoppts+nbotpts]  (bottom = Lambertian boundary)
    const int *nphi0;
    const float *mu, *phi, *skyrad;
    const int *lofj;          // [nlm]
    // raw optics kept for the gradient kernels
    const float *extinct, *albedo, *dirflux, *legen, *phaseinterpwt, *ylmsun, *sfcgridparms;
    const int *iphase;
    // surfaces other than Lambertian and thermal sources (at3d_surface.cu)
    int nang, units;
    float wavelen, gndtemp;
    const float *ord_mu, *ord_phi, *ord_w;   // [nang/2] downward ordinates: MU, PHI, OPI*|MU|*WTDO
    const float *up_mu, *up_phi;             // [nang/2] upward ordinates (surface-emission interpolation)
    const int *up_src;                       // [nang/2] row of SFCGRIDRAD each upward ordinate reads, or -1
    const float *sfcgridrad;                 // [nang/2+1, nbotpts] or null when identically zero
    const float *temp;                       // [npts] grid-point temperatures (thermal sources with a gradient), or null
    void *surfhits;                          // SurfHit[nrays] of the current RENDER call (non-Lambertian only)
    int ray_base;                            // index of the launch's first ray in the caller's ray arrays (error reports)
    const float *viewsrc;                    // [npts] SRCEXT of every grid point for the direction shared by all rays of the
                                             // current launch (orthographic views, NSTOKES=1), or null
    // optional work counters of the last call: [0] cells visited, [1] grid points evaluated,
    // [2] sum of NS over evaluated points, [3] sum of NR (gradient), [4] sub-intervals, [5] rays marched,
    // [6] rays that ended on a general-BRDF surface
    unsigned long long *counts;
    // source stream of the gradient's forward pass (thread-per-ray kernels): every corner evaluated while the adjoint
    // arithmetic is still integrating leaves (SRCEXT8/EXT, single-scatter part * EXT), in evaluation order, in chunks of
    // AT3D_SRC_CHUNK entries drawn from a pool; the last entry of a chunk links to the ray's next chunk.  nullptr: off.
    // Chunk cursors: same-address atomics serialise in L2 (one cursor for all blocks cost 7.5 ms per 1.2 M chunks), so
    // the pool is split into AT3D_SRC_REGIONS regions of srcpool_region chunks with a cursor each (block b draws from
    // region b mod AT3D_SRC_REGIONS) and a shared tail behind them for blocks that exhaust their region.
    float2 *srcpool;
    unsigned srcpool_chunks;                 // all chunks
    unsigned srcpool_region;                 // chunks per region
    unsigned srcpool_nreg;                   // regions in use (<= AT3D_SRC_REGIONS): the SM count
    unsigned *srcpool_top;                   // [0] shared-tail cursor, [1] overflow flag, [32 * (r + 1)] cursor of region r
    long long *srcstart;                     // [nrays of the launch] first entry of the ray, -1: none
};
#define AT3D_SRC_CHUNK 256
#define AT3D_SRC_REGIONS 256
#define AT3D_SRC_TOP_WORDS (32 * (AT3D_SRC_REGIONS + 1))

// Derivative tables of LEVISAPPROX_GRADIENT resident in HBM (reference layouts, see at3d_grad_desc).
#include "at3d_beam.cuh"

struct DevGrad {
    int maxpg, numder, dnumphase, deriv_maxnmicro, pmaxnmicro, longest_path_pts;
    int exact_single_scatter, singlescatter, maxsub;
    double scatmin;
    const int *partder, *doexact;
    const float *dext, *dalb, *dextm;       // [maxpg,numder]
    const float *dalbm, *dfj;               // [8,npts,numder]
    const float *optinterpwt;               // [8,npts]
    const int *interpptr;                   // [8,npts]
    const float *dleg;                      // [nstleg,0:nleg,dnumphase]
    const float *dphasetab;                 // [nstphase,dnumphase,nscatangle]
    const int *diphasep;                    // [deriv_maxnmicro,maxpg,numder]
    const float *dphasewtp;
    const int *iphasep;                     // [pmaxnmicro,maxpg,npart]
    const float *phasewtp;
    const float *extinctp, *albedop;        // [maxpg,npart]
    const float *dpath;                     // [longest_path_pts,npts]
    const int *dptr;
    const float *dtemp;                     // [maxpg,numder] (thermal sources) or null
    int stream_beam;                        // 1: no dense lists, the beam walks run inside the gradient call
    BeamGeom bg;                            // property-grid beam constants (stream_beam)
    const float *bzl;                       // [bg.npz] property-grid levels (stream_beam)
    // ray-independent tables built once per attach by the grad_prep kernels (at3d_grad.cu).  A grid point
    // owns NNZ*NUMDER "rows" (unknown-major, then its property corners with a non-zero weight).
    int ncomp, ntup;                        // Legendre components that reach I,Q,U (1|4); padded table length
    int prow_stride;                        // int4 per row record
    int sp_stride;                          // int4 per species record
    const int4 *gptrec;                     // [npts]: first row, nrows | nnz<<16, DSH block (units of 32 floats), SHPAD(NR)/32
    const int4 *rowrec;                     // [rows][prow_stride]: c_src, Cj, DEXTM*XI, IB | nb, npl, 0, 0 | (iph, coef) list
    const int4 *sprec;                      // [npts,numder][sp_stride]: alb, F, SCATTERJ, count | (iph, coef) list of SINGSCATJ
    const float *dsh;                       // rows of XI*DLEGT(l_j)*RADIANCE(.,j), planar like the source blocks
    const float *dlegt;                     // [rows][ntup]: DLEGT(comp + ncomp*l)          (no delta-M only, else null)
    const float *legs;                      // [npts,numder][ntup]: table for SOURCET (NPART>1 and no delta-M only)
};

#define FULLMASK 0xffffffffu

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- arithmetic helpers with the reference's rounding (no FMA contraction; file built with
// ---- --fmad=false, explicit fmaf() only where op order is free) ----
__device__ __forceinline__ double ipow_d(double x, int n)
{
    double r = 1.0, b = x;
    int e = n;
    while (e > 0) {
        if (e & 1) r *= b;
        e >>= 1;
        if (e) b *= b;
    }
    return r;
}

// DM_M10_N0  (shdomsub2.f:4580-4600)
__device__ __forceinline__ double dev_dm_m10_n0(double x, int m)
{
    if (m == 0) return 1.0;
    double cm = (m & 1) ? -1.0 : 1.0;
    double prod = 1.0;
    for (int p = 1; p <= m; p++) prod = prod * sqrt((double)(m + p) / (double)p);
    return cm * prod * ipow_d(0.5 * sqrt((1.0 - x) * (1.0 + x)), m);
}

// DMM1_N0 (shdomsub2.f:4452-4486)
__device__ __forceinline__ double dev_dmm1_n0(double x, int m, int m1)
{
    if (m == m1) return ipow_d((1.0 + x) / 2.0, m);
    double cmm1 = (m1 > m) ? 1.0 : (((m - m1) & 1) ? -1.0 : 1.0);
    int maxm = m > m1 ? m : m1, minm = m < m1 ? m : m1;
    double prod = 1.0;
    for (int p = 1; p <= maxm - minm; p++) prod = prod * sqrt((double)(m + m1 + p) / (double)p);
    double fact = sqrt((1.0 - x) / 2.0);
    double r = cmm1 * prod * ipow_d(fact, maxm - minm);
    fact = sqrt((1.0 + x) / 2.0);
    return r * ipow_d(fact, maxm + minm);
}

__device__ __forceinline__ int sh_index(int l, int m, int mm)
{   // J = l(l+1)+m+1 (l<=MM) else (2MM+1)l - MM^2 + m + 1   (shdomsub2.f:4290-4294); returns 0-based
    return (l <= mm) ? (l * (l + 1) + m) : ((2 * mm + 1) * l - mm * mm + m);
}

// YLMALL for one direction, cooperative over a lane group (member gl of gsize lanes, lanes over m).
// Ysh layout [ncomp][nlmp]:
// comp 0 = YR(1,:), and for polarized 1 = YR(2,:), 2 = YR(5,:), 3 = YR(6,:), 4 = YR(3,:).
// Follows YLMALL_UNPOL (shdomsub2.f:4490-4539) / YLMALL (shdomsub2.f:4244-4360), TRANSPOSE=.FALSE.
static __device__ void group_ylmall(const DevState &S, float mu, float phi, float *Ysh,
                                    const int lane, const int gsize, const unsigned gmask)
{
    const int ml = S.ml, mm = S.mm, nlmp = S.nlmp;
    const double x = (double)mu;
    const double pi = 3.14159265358979323846;   // DACOS(-1.D0)
    const double fct = 1.0 / sqrt(2.0 * pi);
    // zero padding entries
    for (int j = S.nlm + lane; j < nlmp; j += gsize)
        for (int c = 0; c < S.ny_comp; c++) Ysh[c * nlmp + j] = 0.0f;
    if (S.nstleg == 1) {
        for (int m = lane; m <= mm; m += gsize) {
            double cosm, sinm;
            if (m > 0) { cosm = (double)cosf((float)m * phi); sinm = (double)sinf((float)m * phi); }
            else { cosm = 1.0; sinm = 0.0; }
            double dprev = 0.0, dcur;
            if (m == 0) dcur = 1.0; else dcur = dev_dm_m10_n0(x, m);
            for (int n = m; n <= ml; n++) {
                // emit l = n
                double t = sqrt(n + 0.5) * dcur;
                t = fct * t;
                Ysh[sh_index(n, m, mm)] = (float)((cosm - sinm) * t);
                Ysh[sh_index(n, -m, mm)] = (float)((cosm + sinm) * t);
                // advance recurrence
                double dnext;
                if (m == 0) {
                    if (n == 0) dnext = x;
                    else dnext = ((2 * n + 1) * x * dcur - n * dprev) / (n + 1);
                } else {
                    dnext = ((2 * n + 1) * x * dcur - sqrt((double)(n * n - m * m)) * dprev)
                            / sqrt((double)((n + 1) * (n + 1) - m * m));
                }
                dprev = dcur;
                dcur = dnext;
            }
        }
    } else {
        for (int m = lane; m <= mm; m += gsize) {
            double cosm = 1.0, sinm = 0.0;
            if (m > 0) { cosm = (double)cosf((float)m * phi); sinm = (double)sinf((float)m * phi); }
            const double xp = x, xm = -x;
            const int n0 = m > 2 ? m : 2;
            // WIGNERFCT02P2M_NORMALIZED recurrences (shdomsub2.f:4363-4448), streamed over n
            double d0p = 0.0, d0c = 0.0;      // DM0(n-1), DM0(n)
            double pp = 0.0, pc = 0.0;        // DM2P(n-1), DM2P(n)
            double qp = 0.0, qc = 0.0;        // DM2M(n-1), DM2M(n)
            for (int n = 0; n <= ml; n++) {
                // --- value of DM0(n) ---
                if (n < m) d0c = 0.0;
                else if (n == m) d0c = (m == 0) ? 1.0 : dev_dmm1_n0(xp, m, 0);
                // (for n > m d0c was advanced at the end of the previous iteration)
                if (n < n0) { pc = 0.0; qc = 0.0; }
                else if (n == n0 && ml >= 2) { pc = dev_dmm1_n0(xp, m, 2); qc = dev_dmm1_n0(xm, m, 2); }
                if (n >= m) {
                    double dm0 = sqrt(n + 0.5) * d0c;
                    double dm2m = (((n + m) & 1) ? -1.0 : 1.0) * qc;
                    double dm2p = sqrt(n + 0.5) * pc;
                    dm2m = sqrt(n + 0.5) * dm2m;
                    double p1 = fct * dm0;
                    double p2 = -0.5 * fct * (dm2p + dm2m);
                    double p3 = -0.5 * fct * (dm2p - dm2m);
                    int jp = sh_index(n, m, mm);
                    if (m == 0) {
                        Ysh[0 * nlmp + jp] = (float)p1;
                        Ysh[1 * nlmp + jp] = (float)p2;   // YR(2)
                        Ysh[2 * nlmp + jp] = (float)p3;   // YR(5)
                        Ysh[3 * nlmp + jp] = (float)p3;   // YR(6)
                        Ysh[4 * nlmp + jp] = (float)p2;   // YR(3)
                    } else {
                        int jn = sh_index(n, -m, mm);
                        Ysh[0 * nlmp + jp] = (float)(p1 * cosm - p1 * sinm);
                        Ysh[1 * nlmp + jp] = (float)(p2 * cosm - p2 * sinm);
                        Ysh[4 * nlmp + jp] = (float)(p2 * cosm + p2 * sinm);
                        Ysh[2 * nlmp + jp] = (float)(p3 * cosm - p3 * sinm);
                        Ysh[3 * nlmp + jp] = (float)(p3 * cosm + p3 * sinm);
                        Ysh[0 * nlmp + jn] = (float)(p1 * sinm + p1 * cosm);
                        Ysh[1 * nlmp + jn] = (float)(p2 * sinm + p2 * cosm);
                        Ysh[4 * nlmp + jn] = (float)(p2 * sinm - p2 * cosm);
                        Ysh[2 * nlmp + jn] = (float)(p3 * sinm + p3 * cosm);
                        Ysh[3 * nlmp + jn] = (float)(p3 * sinm - p3 * cosm);
                    }
                }
                // --- advance DM0 to n+1 ---
                if (n >= m && n < ml) {
                    double dnext;
                    if (m == 0) {
                        if (n == 0) dnext = xp;
                        else {
                            double fact1 = (double)(2 * n + 1) * xp / (double)(n + 1);
                            double fact2 = (double)n / (double)(n + 1);
                            dnext = fact1 * d0c - fact2 * d0p;
                        }
                    } else {
                        double fact1 = (double)(n * (n + 1)) * xp;
                        fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - m * m));
                        fact1 = fact1 / (double)(n + 1);
                        fact1 = fact1 * (double)(2 * n + 1) / (double)n;
                        double fact2 = sqrt((double)(n * n - m * m)) * (double)n;
                        fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m * m));
                        fact2 = fact2 / (double)(n + 1);
                        fact2 = fact2 * (double)(n + 1) / (double)n;
                        dnext = fact1 * d0c - fact2 * d0p;
                    }
                    d0p = d0c;
                    d0c = dnext;
                }
                // --- advance DM2P/DM2M to n+1 ---
                if (n >= n0 && n < ml) {
                    double factp = (double)(n * (n + 1)) * xp - (double)(2 * m);
                    double factm = (double)(n * (n + 1)) * xm - (double)(2 * m);
                    double fact1 = 1.0 / sqrt((double)((n + 1) * (n + 1) - m * m));
                    fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - 4));
                    fact1 = fact1 * (double)(2 * n + 1) / (double)n;
                    double fact2 = sqrt((double)(n * n - m * m)) * sqrt((double)(n * n - 4));
                    fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m * m));
                    fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - 4));
                    fact2 = fact2 * (double)(n + 1) / (double)n;
                    double pn = factp * fact1 * pc - fact2 * pp;
                    double qn = factm * fact1 * qc - fact2 * qp;
                    pp = pc; pc = pn;
                    qp = qc; qc = qn;
                }
            }
        }
    }
    __syncwarp(gmask);
}

// ---------------- grid traversal helpers (uniform per warp; every lane runs them) ----------------
struct CellRec {
    int gp[8];     // GRIDPTR(1:8)
    int nb[6];     // NEIGHPTR(1:6)
    int child;     // TREEPTR(2)
    int flags;     // CELLFLAGS
};

__device__ __forceinline__ CellRec load_cell(const DevState &S, int icell)
{
    const int4 *p = S.cellrec + 4 * (size_t)(icell - 1);
    int4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    CellRec r;
    r.gp[0] = a.x; r.gp[1] = a.y; r.gp[2] = a.z; r.gp[3] = a.w;
    r.gp[4] = b.x; r.gp[5] = b.y; r.gp[6] = b.z; r.gp[7] = b.w;
    r.nb[0] = c.x; r.nb[1] = c.y; r.nb[2] = c.z; r.nb[3] = c.w;
    r.nb[4] = d.x; r.nb[5] = d.y; r.child = d.z; r.flags = d.w;
    return r;
}

__device__ __forceinline__ int cell_child(const DevState &S, int icell)
{ return __ldg(&S.cellrec[4 * (size_t)(icell - 1) + 3]).z; }
__device__ __forceinline__ int cell_flags(const DevState &S, int icell)
{ return __ldg(&S.cellrec[4 * (size_t)(icell - 1) + 3]).w; }
__device__ __forceinline__ int cell_gp(const DevState &S, int icell, int n /*1..8*/)
{
    const int *p = (const int *)(S.cellrec + 4 * (size_t)(icell - 1));
    return __ldg(p + (n - 1));
}
__device__ __forceinline__ float pt_coord(const DevState &S, int ip, int axis /*1..3*/)
{
    const float *p = (const float *)(S.ptrec + (size_t)(ip - 1));
    return __ldg(p + (axis - 1));
}

#define DBTEST(x, b) ((((int)(x)) >> (b)) & 1)

// LOCATE_GRID_CELL (shdomsub2.f:4043-4206); may wrap x0,y0 for periodic boundaries
static __device__ int dev_locate_grid_cell(const DevState &S, double &x0, double &y0, double &z0)
{
    const int nx = S.nx, ny = S.ny, nz = S.nz, bcflag = S.bcflag, ipflag = S.ipflag;
#define XG(i) __ldg(&S.xgrid[(i) - 1])
#define YG(i) __ldg(&S.ygrid[(i) - 1])
#define ZG(i) __ldg(&S.zgrid[(i) - 1])
    if (!(DBTEST(bcflag, 0) || DBTEST(bcflag, 2))) {
        double xdomain = (double)(XG(nx + 1) - XG(1));
        if (x0 < XG(1)) x0 = x0 - xdomain * ((int)((x0 - XG(1)) / xdomain) - 1);
        else if (x0 > XG(nx + 1)) x0 = x0 - xdomain * (int)((x0 - XG(1)) / xdomain);
    }
    if (!(DBTEST(bcflag, 1) || DBTEST(bcflag, 3))) {
        double ydomain = (double)(YG(ny + 1) - YG(1));
        if (y0 < YG(1)) y0 = y0 - ydomain * ((int)((y0 - YG(1)) / ydomain) - 1);
        else if (y0 > YG(ny + 1)) y0 = y0 - ydomain * (int)((y0 - YG(1)) / ydomain);
    }
    int il = 0, iu, im, ix, iy, iz;
    if (DBTEST(ipflag, 0)) {
        iu = nx;
        while (iu - il > 1) { im = (iu + il) / 2; if (x0 >= 0.5f * (XG(im) + XG(im + 1))) il = im; else iu = im; }
        il = il + 1;
    } else {
        iu = nx + 1;
        while (iu - il > 1) { im = (iu + il) / 2; if (x0 >= XG(im)) il = im; else iu = im; }
    }
    ix = il > 1 ? il : 1;
    il = 0;
    if (DBTEST(ipflag, 1)) {
        iu = ny;
        while (iu - il > 1) { im = (iu + il) / 2; if (y0 >= 0.5f * (YG(im) + YG(im + 1))) il = im; else iu = im; }
        il = il + 1;
    } else {
        iu = ny + 1;
        while (iu - il > 1) { im = (iu + il) / 2; if (y0 >= YG(im)) il = im; else iu = im; }
    }
    iy = il > 1 ? il : 1;
    il = 0; iu = nz;
    while (iu - il > 1) { im = (iu + il) / 2; if (z0 >= ZG(im)) il = im; else iu = im; }
    iz = il > 1 ? il : 1;
    int nyc = ny;
    if (DBTEST(bcflag, 0)) {
        if (x0 < XG(1)) ix = 1;
        else if (x0 > XG(nx)) ix = nx + 1;
        else ix = ix + 1;
    }
    if (DBTEST(bcflag, 2) && !DBTEST(ipflag, 0)) { int nxc = nx - 1; ix = ix < nxc ? ix : nxc; }
    if (DBTEST(bcflag, 1)) {
        nyc = ny + 1;
        if (y0 < YG(1)) iy = 1;
        else if (y0 > YG(ny)) iy = ny + 1;
        else iy = iy + 1;
    }
    if (DBTEST(bcflag, 3) && !DBTEST(ipflag, 1)) { nyc = ny - 1; iy = iy < nyc ? iy : nyc; }
#undef XG
#undef YG
#undef ZG
    int icell = iz + (nz - 1) * (iy - 1) + (nz - 1) * nyc * (ix - 1);
    int child;
    while ((child = cell_child(S, icell)) > 0) {
        int dir = (cell_flags(S, icell) >> 2) & 3;
        int ic = child + 1;
        int iptr = cell_gp(S, ic, 1);
        if (dir == 1) { if (x0 < pt_coord(S, iptr, 1)) ic = ic - 1; }
        else if (dir == 2) { if (y0 < pt_coord(S, iptr, 2)) ic = ic - 1; }
        else if (dir == 3) { if (z0 < pt_coord(S, iptr, 3)) ic = ic - 1; }
        icell = ic;
    }
    return icell;
}

// NEXT_CELL (shdomsub1.f:4470-4522) for a negative neighbour pointer
static __device__ int dev_next_cell(const DevState &S, double xe, double ye, double ze,
                             int iface, int jface, int inext_neg)
{
    int ic = -inext_neg, child;
    while ((child = cell_child(S, ic)) > 0) {
        int dir = (cell_flags(S, ic) >> 2) & 3;
        int ic1 = child;
        if (dir == jface) {
            ic = ic1 + 1 - ((iface - 1) % 2);
        } else {
            ic = ic1;
            int p8 = cell_gp(S, ic1, 8);
            if (dir == 1) { if (xe > pt_coord(S, p8, 1)) ic = ic + 1; }
            else if (dir == 2) { if (ye > pt_coord(S, p8, 2)) ic = ic + 1; }
            else { if (ze > pt_coord(S, p8, 3)) ic = ic + 1; }
        }
    }
    return ic;
}

// binary search of FIND_BOUNDARY_RADIANCE (shdomsub2.f:2791-2804); 1-based index or 0
__device__ __forceinline__ int dev_bc_search(const int *col, int n, int ip)
{
    int il = 1, iu = n, im;
    while (iu - il > 1) { im = (iu + il) / 2; if (ip >= __ldg(&col[im - 1])) il = im; else iu = im; }
    int ibc = il;
    if (__ldg(&col[ibc - 1]) != ip) ibc = iu;
    if (__ldg(&col[ibc - 1]) != ip) return 0;
    return ibc;
}

// ---- per-ray setup: start point (RENDER, shdomsub4.f:214-236) and direction quantities
// ---- (INTEGRATE_1RAY, shdomsub2.f:2413-2481; ROTATE_POL_PLANE, :3277-3314).
// The cell walk is bit-exact only if CX,CY,CZ and the scattering-angle index agree to the last
// bit with the reference, whose libm is glibc.  CUDA's double sin/cos/acos are not bit-identical to
// glibc's, so for HOST ray arrays the library evaluates this function on the host (same libm as
// the reference's gfortran build) and ships one 80-byte RayPack per ray; for DEVICE-resident rays
// the same function runs on the device (ulp-level differences possible, documented in DESIGN.md).
struct RayGeom {
    float solarmu, solaraz, ztop, zbot;
    int nscatangle, srctype, deltam, nstokes;
};

struct __align__(16) RayPack {
    double x0, y0, z0;      // start point after the top-of-domain slide
    double cx, cy, cz;      // direction cosines of the backward march (small ones zeroed)
    double cos22, sin22;    // polarization-plane rotation
    float f;                // scattering-angle interpolation weight
    int j;                  // scattering-angle table index (1-based)
    int status;             // 0 ok, 1 looks away from the domain (radiance 0), 2 below the domain (error)
    int pad;
};

__host__ __device__ inline void make_ray_pack(const RayGeom &g, double x0, double y0, double z0,
                                              double mu2, double phi2, RayPack &p)
{
    const double pi = 3.14159265358979323846;      // = ACOS(-1.0D0), correctly rounded
    p.status = 0; p.pad = 0;
    // SQRT(1-MU2**2), COS(PHI2-PI), SIN(PHI2-PI) appear three / two / two times in the reference with the same operands
    // (MURAY = -MU2, PHIRAY = PHI2-PI): evaluated once here, bit for bit the same values
    const double sinth = sqrt(1.0 - mu2 * mu2), cphi = cos(phi2 - pi), sphi = sin(phi2 - pi);
    {
        const double muray = -mu2;
        if (z0 > g.ztop) {
            if (muray >= 0.0) p.status = 1;
            else {
                const double r = (g.ztop - z0) / muray;
                x0 = x0 + r * sinth * cphi;
                y0 = y0 + r * sinth * sphi;
                z0 = g.ztop;
            }
        } else if (z0 < g.zbot) {
            p.status = 2;
        }
    }
    p.x0 = x0; p.y0 = y0; p.z0 = z0;
    p.f = 0.0f; p.j = 1; p.cos22 = 1.0; p.sin22 = 0.0;
    if (g.srctype != 'T' && g.deltam) {
        double cosscat = g.solarmu * mu2
            + sqrt((1.0f - g.solarmu * g.solarmu) * (1.0 - mu2 * mu2)) * cos(g.solaraz - phi2);
        cosscat = fmax(fmin(1.0, cosscat), -1.0);
        float f = (float)((g.nscatangle - 1) * (acos(cosscat) / pi) + 1);
        int j = (int)f;
        if (j > g.nscatangle - 1) j = g.nscatangle - 1;
        p.f = f - (float)j;
        p.j = j;
        if (g.nstokes > 1) {
            // ROTATE_POL_PLANE with MU=SNGL(MU2), DELPHI=SOLARAZ-SNGL(PHI2)
            const float mu = (float)mu2;
            const float delphi = g.solaraz - (float)phi2;
            const double sin_scat = sqrt(fmax(0.0, 1.0 - cosscat * cosscat));
            const double sin_theta1 = sqrt(1.0 - (double)(g.solarmu * g.solarmu));
            const double sin_theta2 = sqrt(1.0 - (double)(mu * mu));
            const double sinphi = sin((double)delphi), cosphi = cos((double)delphi);
            double sin2, cos2;
            if (sin_scat == 0.0) { sin2 = 0.0; cos2 = -1.0; }
            else {
                sin2 = sin_theta1 * sinphi / sin_scat;
                cos2 = (sin_theta2 * g.solarmu - sin_theta1 * mu * cosphi) / sin_scat;
            }
            p.sin22 = 2.0 * sin2 * cos2;
            p.cos22 = 1.0 - 2.0 * (sin2 * sin2);
        }
    }
    p.cx = sinth * cphi;
    p.cy = sinth * sphi;
    p.cz = -mu2;
    if (!(fabs(p.cx) > 1.0e-6f)) p.cx = 0.0;
    if (!(fabs(p.cy) > 1.0e-6f)) p.cy = 0.0;
    if (!(fabs(p.cz) > 1.0e-6f)) p.cz = 0.0;
}

// A ray that ended on a non-Lambertian bottom boundary: the reflected radiance needs NANG/2 BRDF evaluations for
// each of the 4 face points, which a warp per ray does afterwards (surface_kernel) instead of one lane here.
struct __align__(16) SurfHit {
    double xb, yb, transmit;
    double rad[3];          // radiance accumulated along the ray (INTEGRATE_1RAY arithmetic)
    int icell, kface;       // kface = 0: no surface hit
};

struct RayDir {
    double cx, cy, cz, cxinv, cyinv, czinv;
    double cos22, sin22;    // polarization-plane rotation (ROTATE_POL_PLANE)
    float f;                // scattering-angle interpolation weight
    int j;                  // scattering-angle table index (1-based)
    int bitx, bity, bitz, ioct;
    float xm, ym;
    float phi2;             // SNGL(PHI2) (surface emission interpolation)
    SurfHit *hit;           // where to leave a non-Lambertian surface hit (null: Lambertian surface)
};

__device__ __forceinline__ RayGeom dev_ray_geom(const DevState &S)
{
    RayGeom g;
    g.solarmu = S.solarmu; g.solaraz = S.solaraz;
    g.ztop = __ldg(&S.zgrid[S.nz - 1]); g.zbot = __ldg(&S.zgrid[0]);
    g.nscatangle = S.nscatangle; g.srctype = S.srctype; g.deltam = S.deltam; g.nstokes = S.nstokes;
    return g;
}

__device__ __forceinline__ void dev_ray_dir(const DevState &S, const RayPack &p, RayDir &rd)
{
    rd.phi2 = 0.0f; rd.hit = nullptr;
    rd.f = p.f; rd.j = p.j; rd.cos22 = p.cos22; rd.sin22 = p.sin22;
    rd.cx = p.cx; rd.cy = p.cy; rd.cz = p.cz;
    rd.cxinv = (rd.cx != 0.0) ? 1.0 / rd.cx : (double)1.0e6f;
    rd.cyinv = (rd.cy != 0.0) ? 1.0 / rd.cy : (double)1.0e6f;
    rd.czinv = (rd.cz != 0.0) ? 1.0 / rd.cz : (double)1.0e6f;
    rd.bitx = rd.cx < 0.0 ? 1 : 0;
    rd.bity = rd.cy < 0.0 ? 1 : 0;
    rd.bitz = rd.cz < 0.0 ? 1 : 0;
    rd.ioct = 1 + rd.bitx + 2 * rd.bity + 4 * rd.bitz;
    rd.xm = 0.5f * (__ldg(&S.xgrid[0]) + __ldg(&S.xgrid[S.nx - 1]));
    rd.ym = 0.5f * (__ldg(&S.ygrid[0]) + __ldg(&S.ygrid[S.ny - 1]));
}

// the RayPack of ray iray: shipped from the host, or evaluated here for device-resident rays
__device__ __forceinline__ RayPack dev_get_pack(const DevState &S, const RayPack *packs, int iray,
                                                const float *camx, const float *camy, const float *camz,
                                                double mu2, double phi2)
{
    if (packs) {
        const double2 *q = (const double2 *)(packs + iray);
        RayPack p;
        double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
        const int4 e = __ldg((const int4 *)(q + 4));
        p.x0 = a.x; p.y0 = a.y; p.z0 = b.x; p.cx = b.y; p.cy = c.x; p.cz = c.y; p.cos22 = d.x; p.sin22 = d.y;
        p.f = __int_as_float(e.x); p.j = e.y; p.status = e.z; p.pad = 0;
        return p;
    }
    RayPack p;
    make_ray_pack(dev_ray_geom(S), (double)__ldg(&camx[iray]), (double)__ldg(&camy[iray]),
                  (double)__ldg(&camz[iray]), mu2, phi2, p);
    return p;
}

// PLANCK_FUNCTION (shdomsub2.f:4756-4790), UNITS 'T' or radiance units
__device__ __forceinline__ float dev_planck(float temp, int units, float wavelen)
{
    if (units == 'T') return temp;
    if (temp > 0.0f)
        return 1.1911e8f / (wavelen * wavelen * wavelen * wavelen * wavelen) / (expf(1.4388e4f / (wavelen * temp)) - 1);
    return 0.0f;
}

// COMPUTE_TOP_RADIANCES, INTERPOLATE_FLAG=1 (shdomsub1.f:2375-2395, :2421-2427)
static __device__ float dev_sky_radiance(const DevState &S, float mu, float phi)
{
    double weightedsum = 0.0, weightsum = 0.0, weight, distance;
    for (int i = 1; i <= S.nmu / 2; i++) {
        int n = __ldg(&S.nphi0[i - 1]);
        float mus = __ldg(&S.mu[i - 1]);
        for (int j = 1; j <= n; j++) {
            float phis = __ldg(&S.phi[(i - 1) + S.nmu * (j - 1)]);
            distance = (double)acosf(mu * mus + sqrtf((1.0f - mu * mu) * (1.0f - mus * mus)) * cosf(phi - phis));
            if (fabs(distance) < 1e-6f) weight = 1.0e8;
            else weight = 1.0 / pow(distance, 3.0);
            weightedsum = weightedsum + __ldg(&S.skyrad[0 + S.nstokes * ((i - 1) + (S.nmu / 2) * (j - 1))]) * weight;
            weightsum = weightsum + weight;
        }
    }
    const float sky3 = (float)(weightedsum / weightsum);
    if (S.srctype == 'T') return dev_planck(sky3, S.units, S.wavelen);
    return sky3;
}

// surface emission term of FIND_BOUNDARY_RADIANCE (shdomsub2.f:2832-2846): COMPUTE_TOP_RADIANCES with
// INTERPOLATE_FLAG=2 on SFCGRIDRAD(2:,IBC), i.e. inverse-distance-cubed weights over the upward ordinates
static __device__ float dev_surface_emission(const DevState &S, int ibc, float mu, float phi)
{
    if (!S.sfcgridrad) return (S.srctype == 'T') ? dev_planck(0.0f, S.units, S.wavelen) : 0.0f;
    double weightedsum = 0.0, weightsum = 0.0, weight, distance;
    const int nh = S.nang / 2;
    for (int q = 0; q < nh; q++) {
        const float mus = __ldg(&S.up_mu[q]), phis = __ldg(&S.up_phi[q]);
        distance = (double)acosf(mu * mus + sqrtf((1.0f - mu * mu) * (1.0f - mus * mus)) * cosf(phi - phis));
        if (fabs(distance) < 1e-6f) weight = 1.0e8;
        else weight = 1.0 / pow(distance, 3.0);
        const int src = __ldg(&S.up_src[q]);
        const float v = src >= 0 ? __ldg(&S.sfcgridrad[src + (size_t)(nh + 1) * (ibc - 1)]) : 0.0f;
        weightedsum = weightedsum + v * weight;
        weightsum = weightsum + weight;
    }
    const float e = (float)(weightedsum / weightsum);
    if (S.srctype == 'T') return dev_planck(e, S.units, S.wavelen);
    return e;
}

// sums over the 8 lanes of an octet (m = the octet's lane mask)
__device__ __forceinline__ float oct_sum(unsigned m, float v)
{
    v += __shfl_xor_sync(m, v, 4); v += __shfl_xor_sync(m, v, 2); v += __shfl_xor_sync(m, v, 1);
    return v;
}
__device__ __forceinline__ double oct_sum_d(unsigned m, double v)
{
    v += __shfl_xor_sync(m, v, 4); v += __shfl_xor_sync(m, v, 2); v += __shfl_xor_sync(m, v, 1);
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
