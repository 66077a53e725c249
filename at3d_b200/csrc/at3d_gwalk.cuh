// at3d_gwalk.cuh -- thread-per-ray ADJOINT_INTEGRATE_1RAY walk for unpolarized delta-M states (NSTOKES=1), without any
// spherical-harmonic work: the sources of the corners come from the stream the forward pass left.
//
// LEVISAPPROX_GRADIENT (src/polarized/shdomsub4.f:633-804) walks every ray twice: INTEGRATE_1RAY for the pixel value,
// then ADJOINT_INTEGRATE_1RAY (:3223-3967) once the adjoint weight of the pixel is known, and the second walk
// re-evaluates the source function of every corner it meets (COMPUTE_SOURCE_GRAD_1CELL's forward part, :1700-1780).
// Here the forward pass (forward_kernel_t, modes 3) writes, for every corner it evaluates while the adjoint arithmetic
// is still integrating, the pair (SRCEXT8/EXT, single-scatter part * EXT) to a per-ray stream (DevState::srcpool) in
// evaluation order.  The second walk is bit-identical in geometry and extinction, so it meets the same corners in the
// same order and reads its sources back sequentially: no YLMDIR table in shared memory (1 KB per ray at NLM=256), no SH
// code, many more resident warps.  One thread per ray: nothing of the FP64 walk is evaluated redundantly (the octet
// kernel weights_kernel repeats it in 8 lanes), the per-corner weights
//   W = SUM T*A*w(FC0,FC1,E0,E1,ds)   (source term, :3692-3764),   G (radiance term, :4037-4114),
//   B = ADJ * SUM SRCSINGSCAT         (direct-beam weight, :3725-3731, :3815-3826)
// live in shared memory and follow their grid point from cell to cell through a slot permutation held in one register;
// a VisitRec leaves when a point stops being a corner.  Divisions by the sub-interval's extinction are multiplications
// by its reciprocal here (<= 1 ulp of a double; the gradient's parity bar is 1e-4).
#pragma once
#include "at3d_tray.cuh"

// per-thread view of the block's shared memory: element (slot s) of this thread at ptr[s * bt]
#ifndef GW_FC_UNROLL
#define GW_FC_UNROLL 8          // corners per iteration of the sub-interval update (8: fully unrolled, weights in registers)
#endif
constexpr int kGwFcUnroll = GW_FC_UNROLL;
struct GwShared {
    double *W, *Gr, *B;
    float *sf, *ss, *bw;           // SRCEXT8/EXT, single-scatter part times EXT, SRCSINGSCAT of the current cell
    double *fa, *fb;               // interpolation weights of the two ends of a sub-interval (GW_FC_UNROLL < 8)
    int bt;
};

__host__ __device__ inline size_t gw_smem_per_thread() { return 8 * (3 * 8 + 3 * 4) + (GW_FC_UNROLL < 8 ? 2 * 8 * 8 : 0); }

// sequential reader of a ray's source stream
struct SrcReader {
    const float2 *pool;
    const float2 *p;
    int left;
};

__device__ __forceinline__ float2 src_next(SrcReader &r)
{
    if (r.left == 0) {
        r.p = r.pool + (long long)__float_as_int(__ldg(r.p).x) * AT3D_SRC_CHUNK;
        r.left = AT3D_SRC_CHUNK - 1;
    }
    const float2 v = __ldg(r.p);
    r.p++;
    r.left--;
    return v;
}

__device__ __forceinline__ void gw_write(VisitRec *dst, int ip, float srcfull, double W, double Gr, double B)
{
    *(int4 *)dst = make_int4(ip, __float_as_int(srcfull), 0, 0);
    *((double2 *)dst + 1) = make_double2(W, Gr);
    *((double2 *)dst + 2) = make_double2(B, 0.0);
}

#define GW_SLOT(perm, n) (((perm) >> (4 * (n))) & 7u)

// Corner refresh: DONEFACE inheritance as thread_refresh (at3d_tray.cuh); the accumulator slot of an inherited corner
// follows it, a corner that is not inherited leaves its record and frees its slot; new corners read the stream.
__device__ __forceinline__ int gw_refresh(const DevState &S, const DevGrad &G, const CellRec &c, const GwShared &M,
                                          bool singlescatter, int jf, bool first_cell, TCorners &K, unsigned &perm,
                                          SrcReader &rd, VisitRec *rec, int cap, int &nrec, int &npairs, int &npt_eval,
                                          int &nsh_eval)
{
    TCorners N;
    unsigned need = 0, used_old = 0, newperm = 0;
    const int bt = M.bt;
    int err = 0;
    // one copy of the inheritance per entry face (the face fixes which old corner n ^ 1 | 2 | 4 a new corner can inherit
    // from): three times the code, a third of the instructions executed per cell
#define GW_INHERIT(MASK)                                                                            \
    _Pragma("unroll")                                                                               \
    for (int n = 0; n < 8; n++) {                                                                   \
        const int ip = c.gp[n];                                                                     \
        const int k = n ^ (MASK);                                                                   \
        N.x[n] = K.x[k]; N.y[n] = K.y[k]; N.z[n] = K.z[k]; N.ext[n] = K.ext[k]; N.src[n] = K.src[k]; \
        N.pt[n] = ip;                                                                               \
        if (K.pt[k] == ip) { used_old |= 1u << k; newperm |= GW_SLOT(perm, k) << (4 * n); }         \
        else need |= 1u << n;                                                                       \
    }
    if (jf == 1) { GW_INHERIT(1) }
    else if (jf == 2) { GW_INHERIT(2) }
    else if (jf == 3) { GW_INHERIT(4) }
    else {
#pragma unroll
        for (int n = 0; n < 8; n++) { N.x[n] = N.y[n] = N.z[n] = N.ext[n] = N.src[n] = 0.0f; N.pt[n] = c.gp[n]; }
        need = 0xFFu;
    }
#undef GW_INHERIT
    // old corners without an heir leave their record
    unsigned freeslots = 0;
    if (first_cell) freeslots = 0xFFu;
    else {
        // (rolled: one copy of the record write instead of eight keeps the loop body of the march in the instruction cache)
        unsigned leave = ~used_old & 0xFFu;
#ifndef GW_UNROLL_EVICT
#pragma unroll 1
#endif
        while (leave) {
            const int k = __ffs(leave) - 1;
            leave &= leave - 1;
            const unsigned s = GW_SLOT(perm, k);
            freeslots |= 1u << s;
            const double w_ = M.W[s * bt], g_ = M.Gr[s * bt], b_ = M.B[s * bt];
            if (w_ != 0.0 || g_ != 0.0 || b_ != 0.0) {         // a visit in clear air contributes exactly nothing
                const int ptk = SEL8(K.pt, k);
                if (nrec < cap) gw_write(rec + nrec, ptk, M.sf[s * bt], w_, g_, b_);
                else err = 5;
                nrec++;
                npairs += visit_pairs(G, ptk);
            }
        }
    }
    K = N;
    while (need) {
        const int n = __ffs(need) - 1;
        need &= need - 1;
        const unsigned s = __ffs(freeslots) - 1;
        freeslots &= freeslots - 1;
        newperm |= s << (4 * n);
        const int ip = SEL8(K.pt, n);
        const float4 pr = __ldg(&S.ptrec[ip - 1]);
        // dark points (build_ptsrc_kernel) have no stream entry: their SRCEXT8 is exactly 0 and their record stays empty
        const int psy = __ldg(&S.ptsrc[ip - 1].y);
        const float2 sv = psy < 0 ? make_float2(0.0f, 0.0f) : src_next(rd);
        npt_eval++; nsh_eval += psy & 0xFFFF;
        const float src = singlescatter ? sv.y : sv.x * pr.w;
        M.sf[s * bt] = sv.x; M.ss[s * bt] = sv.y;
        M.W[s * bt] = 0.0; M.Gr[s * bt] = 0.0; M.B[s * bt] = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k == n) { K.x[k] = pr.x; K.y[k] = pr.y; K.z[k] = pr.z; K.ext[k] = pr.w; K.src[k] = src; }
    }
    perm = newperm;
    return err;
}

// One ray in ADJOINT_INTEGRATE_1RAY's arithmetic (shdomsub4.f:3223-3967).  Returns the ray error code (0: fine).
__device__ int thread_march_weights(const DevState &S, const DevGrad &G, const GwShared &M, const RayDir &rd, double mu2,
                                    double x0, double y0, double z0, float sky, double adj, double total,
                                    SrcReader &sr, VisitRec *rec, int cap, int *trace_cells, int trace_cap,
                                    int &ntrace, int &nsub, int &npt_eval, int &nsh_eval, int &nrec, int &npairs)
{
    double xe = x0, ye = y0, ze = z0, tr = 1.0, radout = 0.0;
    float ext1 = 0.0f, srcext1 = 0.0f;
    double ext1d = 0.0;
    const int bt = M.bt;
    const int p1c = cell_gp(S, 1, 1), p8c = cell_gp(S, 1, 8);
    const double eps = (double)(1.0e-5f * (pt_coord(S, p8c, 3) - pt_coord(S, p1c, 3)));
    const bool exact_ss = G.exact_single_scatter != 0, singlescatter = G.singlescatter != 0;
    int icell = dev_locate_grid_cell(S, xe, ye, ze);
    int iface = 0, jf = 0, err = 0;
    bool done = false, first_cell = true;
    unsigned perm = 0x76543210u;
    TCorners K;
    npt_eval = 0; nsh_eval = 0; nrec = 0; npairs = 0; ntrace = 0; nsub = 0;
#pragma unroll
    for (int n = 0; n < 8; n++) { K.pt[n] = 0; K.x[n] = K.y[n] = K.z[n] = K.ext[n] = K.src[n] = 0.0f; }
    int boundpts[4] = {0, 0, 0, 0};
    double bval[4] = {0.0, 0.0, 0.0, 0.0};
    CellRec c;
    if (icell > 0) c = load_cell(S, icell);
    while (!done && icell > 0) {
        if (trace_cells && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        err = gw_refresh(S, G, c, M, singlescatter, jf, first_cell, K, perm, sr, rec, cap, nrec, npairs, npt_eval, nsh_eval);
        if (err) return err;
        first_cell = false;
        const float q1x = K.x[0], q1y = K.y[0], q1z = K.z[0];
        const float q8x = K.x[7], q8y = K.y[7], q8z = K.z[7];
        const int io = 8 - rd.ioct;
        const float qox = SEL8(K.x, io), qoy = SEL8(K.y, io), qoz = SEL8(K.z, io);
        const double delx = (double)(q8x - q1x), dely = (double)(q8y - q1y), delz = (double)(q8z - q1z);
        const double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        const double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        const double invdelz = 1.0 / delz;
        double u = (xe - q1x) * invdelx, v = (ye - q1y) * invdely, w = (ze - q1z) * invdelz;
        double f1[8];                                           // FC1: interpolation weights at the previous point
        interp_kernel(u, v, w, f1);
        srcext1 = fmaxf(0.0f, (float)fcsum(f1, K.src));
        ext1d = fcsum(f1, K.ext);
#if GW_FC_UNROLL < 8
        // the weights of the two ends of a sub-interval live in shared memory, so that the corner update below is a short
        // loop instead of eight copies (the loop body of the march has to come out of the instruction cache every cell)
        double *fprev = M.fa, *fcur = M.fb;
#pragma unroll
        for (int n = 0; n < 8; n++) fprev[n * bt] = f1[n];
#endif
        ext1 = (float)ext1d;
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
        float extn;
        { double fcn[8]; interp_kernel(u, v, w, fcn); extn = (float)fcsum(fcn, K.ext); }
        const double taugrid = so * 0.5f * (ext1 + extn);
        int ntau = 1 + (int)(taugrid / S.tautol);
        if (ntau < 1) ntau = 1;
        const double dels = so / ntau;
        if (exact_ss) {
#pragma unroll
            for (int n = 0; n < 8; n++) M.bw[n * bt] = 0.0f;
        }
        for (int it = 1; it <= ntau; it++) {
            const double s = it * dels;
            const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
            const double ui = (xi - q1x) * invdelx, vi = (yi - q1y) * invdely, wi = (zi - q1z) * invdelz;
            double fc[8];
            interp_kernel(ui, vi, wi, fc);
            const float srcext0 = fmaxf(0.0f, (float)fcsum(fc, K.src));
            const double ext0d = fcsum(fc, K.ext);
            const float ext0 = (it != ntau) ? (float)ext0d : extn;
            const double ext = (double)(0.5f * (ext0 + ext1));
            if (ext != 0.0) {
                const double tau = ext * dels;
                const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                const double transcell = 1.0f - abscell;
                const double corr = dels * (1.0f - 0.05f * (ext1 - ext0) * dels);
                const double rcur = adj * (total - radout) / tr;                 // adj . PASSEDRAD(kk)
                const double src = (0.5f * (srcext0 + srcext1)
                    + 0.08333333333f * (ext0 * srcext1 - ext1 * srcext0) * dels
                      * (1.0f - 0.05f * (ext1 - ext0) * dels)) / ext;
                radout = radout + tr * src * abscell;
                const double tnext = tr * transcell;
                const double rnext = adj * (total - radout) / tnext;             // adj . PASSEDRAD(kk+1)
                const double wgt = tr * abscell, rext = 1.0 / ext;
                const double aext = 0.5f * (ext0d + ext1d);
                const double raext = aext != 0.0 ? 1.0 / aext : 0.0;
                const double cg = 0.08333333333f * dels * (1.0f - 0.05f * (ext1d - ext0d) * dels);
#if GW_FC_UNROLL < 8
#pragma unroll
                for (int n = 0; n < 8; n++) fcur[n * bt] = fc[n];
#pragma unroll kGwFcUnroll
#else
#pragma unroll
#endif
                for (int n = 0; n < 8; n++) {
                    const unsigned sl = GW_SLOT(perm, n);
#if GW_FC_UNROLL < 8
                    const double f0n = fcur[n * bt], f1n = fprev[n * bt];
#else
                    const double f0n = fc[n], f1n = f1[n];
#endif
                    M.W[sl * bt] += wgt * ((0.5f * (f0n + f1n) + 0.08333333333f * (ext0 * f1n - ext1 * f0n) * corr) * rext);
                    if (exact_ss) {
                        const float ssn = M.ss[sl * bt];
                        const float ss0 = fmaxf(0.0f, (float)(f0n * ssn)), ss1 = fmaxf(0.0f, (float)(f1n * ssn));
                        M.bw[n * bt] = (float)(M.bw[n * bt] + wgt *
                                       (0.5f * (ss0 + ss1) + 0.08333333333f * (ext0 * ss1 - ext1 * ss0) * corr) * rext);
                    }
                    // radiance term (COMPUTE_RADIANCE_DERIVATIVE_ADJOINT): extinctions re-interpolated in double
                    const double g0 = -rnext * f0n, g1 = -rcur * f1n;
                    M.Gr[sl * bt] += ((0.5f * (g0 + g1) + cg * (ext0d * g1 - ext1d * g0)) * raext) * wgt;
                }
                tr = tnext;
                nsub++;
                if (nsub + 1 > G.maxsub) return 4;
            } else if (exact_ss) {
#pragma unroll
                for (int n = 0; n < 8; n++) M.bw[n * bt] = 0.0f;
            }
            ext1 = ext0; ext1d = ext0d;
            srcext1 = srcext0;
#if GW_FC_UNROLL < 8
            if (ext != 0.0) { double *t_ = fprev; fprev = fcur; fcur = t_; }
            else {
#pragma unroll
                for (int n = 0; n < 8; n++) fprev[n * bt] = fc[n];
            }
#else
#pragma unroll
            for (int n = 0; n < 8; n++) f1[n] = fc[n];
#endif
        }
        if (exact_ss) {
#pragma unroll 1
            for (int n = 0; n < 8; n++) {
                const unsigned sl = GW_SLOT(perm, n);
                M.B[sl * bt] += adj * (double)M.bw[n * bt];
            }
        }
        if (inextcell > 0) {
            if (jface == 1) xn = (double)snap;
            else if (jface == 2) yn = (double)snap;
            else zn = (double)snap;
        }
        if (tr < S.transcut) done = true;
        else if (inextcell == 0 && iface >= 5) {
            done = true;
            float radbnd[1];
            double boundinterp[4], dirrad1[4];
            const int e = boundary_radiance<1, true>(S, xn, yn, (float)mu2, rd.phi2, sky, ic, kface, radbnd,
                                                     boundpts, boundinterp, dirrad1);
            if (e) return e;
            if (exact_ss) {
#pragma unroll
                for (int j = 0; j < 4; j++) bval[j] = adj * tr * boundinterp[j] * dirrad1[j];
            }
        } else { icell = inextcell; c = cn; }
        jf = jface;
        xe = xn; ye = yn; ze = zn;
    }
    if (!first_cell) {
        // the corners of the last cell, then the surface points (FIND_BOUNDARY_RADIANCE_GRAD's beam weights)
#pragma unroll 1
        for (int k = 0; k < 8; k++) {
            const unsigned sl = GW_SLOT(perm, k);
            const double w_ = M.W[sl * bt], g_ = M.Gr[sl * bt], b_ = M.B[sl * bt];
            if (w_ != 0.0 || g_ != 0.0 || b_ != 0.0) {
                const int ptk = SEL8(K.pt, k);
                if (nrec < cap) gw_write(rec + nrec, ptk, M.sf[sl * bt], w_, g_, b_);
                else err = 5;
                nrec++;
                npairs += visit_pairs(G, ptk);
            }
        }
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            const double bv = j == 0 ? bval[0] : j == 1 ? bval[1] : j == 2 ? bval[2] : bval[3];
            const int bp = j == 0 ? boundpts[0] : j == 1 ? boundpts[1] : j == 2 ? boundpts[2] : boundpts[3];
            if (bv != 0.0) {
                if (nrec < cap) gw_write(rec + nrec, bp, 0.0f, 0.0, 0.0, bv);
                else err = 5;
                nrec++;
                npairs += visit_pairs(G, bp);
            }
        }
    }
    return err;
}
