// at3d_tray.cuh -- thread-per-ray march for unpolarized states (NSTOKES=1) on sm_100a.
//
// One thread integrates one ray; a warp carries 32 neighbouring rays (neighbouring pixels cross mostly
// the same cells, so the warp stays largely converged and its loads hit the same lines).  Nothing is
// evaluated redundantly: the FP64 walk of INTEGRATE_1RAY (shdomsub2.f:2311-2743) runs once per ray
// in the reference's operation order (bit-exact cell sequence and sub-interval counts), and each thread
// contracts the spherical-harmonic source of its own new corner points with its own YLMDIR, which lives
// in shared memory as float4 columns [j/4][thread] (conflict-free LDS.128; 4*NLM bytes per thread).
// The octet kernels (at3d_ray.cuh) remain the path for NSTOKES=3 and for NLM too large for this layout.
#pragma once
#include "at3d_ray.cuh"

// YLMALL_UNPOL (shdomsub2.f:4490-4539) for one direction by one thread, written to the thread's column
// of the shared float4 table: element j lives at Y4[(j>>2)*bt + tid], component j&3.
static __device__ void thread_ylmall_unpol(const DevState &S, float mu, float phi, float *Ycol, int bt)
{
    const int ml = S.ml, mm = S.mm;
    const double x = (double)mu;
    const double pi = 3.14159265358979323846;   // DACOS(-1.D0)
    const double fct = 1.0 / sqrt(2.0 * pi);
#define YST(j, v) Ycol[(size_t)((j) >> 2) * bt * 4 + ((j) & 3)] = (v)
    for (int j = S.nlm; j < S.nlmp; j++) YST(j, 0.0f);
    for (int m = 0; m <= mm; m++) {
        double cosm, sinm;
        if (m > 0) { cosm = (double)cosf((float)m * phi); sinm = (double)sinf((float)m * phi); }
        else { cosm = 1.0; sinm = 0.0; }
        double dprev = 0.0, dcur;
        if (m == 0) dcur = 1.0; else dcur = dev_dm_m10_n0(x, m);
        for (int n = m; n <= ml; n++) {
            double t = sqrt(n + 0.5) * dcur;
            t = fct * t;
            YST(sh_index(n, m, mm), (float)((cosm - sinm) * t));
            YST(sh_index(n, -m, mm), (float)((cosm + sinm) * t));
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
#undef YST
}

// The eight corner points of the current cell, kept by the thread (static register indexing only).
struct TCorners {
    int pt[8];
    float x[8], y[8], z[8], ext[8], src[8];
};

#define SEL8(arr, n) ((n) == 0 ? arr[0] : (n) == 1 ? arr[1] : (n) == 2 ? arr[2] : (n) == 3 ? arr[3] : \
                      (n) == 4 ? arr[4] : (n) == 5 ? arr[5] : (n) == 6 ? arr[6] : arr[7])

// exact single-scatter sum of one point from its record / list (shdomsub2.f:3172-3183)
__device__ __forceinline__ float thread_singscat_sum(const DevState &S, int ip, const int4 ps, const RayDir &rd)
{
    const int cnt = (ps.y >> 16) & 0x7FFF;
    float b = 0.0f;
    if (cnt > 0) {
        float sv[1];
        ray_singscat<1>(S.phasetab, S.nstphase, S.numphase, ps.z, rd, sv);
        b = fmaf(__int_as_float(ps.w), sv[0], b);
        for (int e = 1; e < cnt; e++) {
            const int2 en = __ldg(&S.ssent[(size_t)(ip - 1) * S.kmax + e]);
            ray_singscat<1>(S.phasetab, S.nstphase, S.numphase, en.x, rd, sv);
            b = fmaf(__int_as_float(en.y), sv[0], b);
        }
    }
    return b;
}

// SRCEXT of one grid point for the direction whose YLMDIR is in Y4 (column stride bt): COMPUTE_SOURCE_1CELL_UNPOL
// (shdomsub2.f:3046-3192) with the TMS-corrected SH block of the point, times the extinction
// Per-thread writer of the gradient's source stream (DevState::srcpool): the state is ONE register pair, the next entry
// `cur` (-1: nothing written yet, SRC_OVF: the pool overflowed).  The common case is a store and an increment; drawing a
// chunk (once per AT3D_SRC_CHUNK - 1 entries) is kept out of line, so that the march's register allocation is the one
// of the kernel without a stream (inlined, the cursor code cost 58 % more executed instructions in the whole kernel).
typedef long long SrcWriter;
#define SRC_OVF (-1ll - AT3D_SRC_CHUNK)

static __device__ __noinline__ long long src_draw_chunk(float2 *pool, unsigned *top, unsigned nreg, unsigned region, unsigned chunks,
                                                 long long cur, long long *startslot)
{
    if (cur == SRC_OVF) return cur;
    const unsigned reg = blockIdx.x % nreg;
    unsigned c = atomicAdd(top + 32 * (reg + 1), 1u);
    if (c < region) c += reg * region;
    else c = nreg * region + atomicAdd(top, 1u);
    if (c >= chunks) { atomicExch(top + 1, 1u); return SRC_OVF; }
    const long long base = (long long)c * AT3D_SRC_CHUNK;
    if (cur < 0) *startslot = base;
    else __stcs(&pool[cur], make_float2(__int_as_float((int)c), 0.0f));     // link slot of the full chunk
    return base;
}

__device__ __forceinline__ void src_emit(const DevState &S, SrcWriter &cur, long long *startslot, float srcfull, float ss)
{
    if ((cur & (AT3D_SRC_CHUNK - 1)) == AT3D_SRC_CHUNK - 1)
        cur = src_draw_chunk(S.srcpool, S.srcpool_top, S.srcpool_nreg, S.srcpool_region, S.srcpool_chunks, cur, startslot);
    if (cur >= 0) {
        __stcs(&S.srcpool[cur], make_float2(srcfull, ss));   // streaming: the L1 left beside the YLMDIR table holds SH rows
        cur++;
    }
}

__device__ __forceinline__ float thread_point_source(const DevState &S, const float4 *Y4, int bt, const RayDir &rd,
                                                     bool singlescatter, int ip, float ext, int &ns,
                                                     SrcWriter &sw, long long *startslot, bool stream)
{
    const int4 ps = __ldg(&S.ptsrc[ip - 1]);
    ns = ps.y & 0xFFFF;
    if (ps.y < 0) return 0.0f;                // dark point (build_ptsrc_kernel): SRCEXT8 is exactly 0, no stream entry
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    if (!singlescatter || stream) {           // the gradient's source stream wants the SH part in any case
        const float4 *base = (const float4 *)(S.shsrc + ps.x);
        const int n4 = AT3D_SHPAD(ns) >> 2;                 // multiple of 8
#ifndef AT3D_SH_UNROLL4
        // eight 16-byte loads of the row in flight per iteration (n4 is a multiple of 8); same summation order as below
#pragma unroll 1
        for (int j = 0; j < n4; j += 8) {
            const float4 s0 = __ldg(base + j), s1 = __ldg(base + j + 1), s2 = __ldg(base + j + 2), s3 = __ldg(base + j + 3);
            const float4 s4 = __ldg(base + j + 4), s5 = __ldg(base + j + 5), s6 = __ldg(base + j + 6), s7 = __ldg(base + j + 7);
            {
                const float4 y0 = Y4[(size_t)j * bt], y1 = Y4[(size_t)(j + 1) * bt];
                const float4 y2 = Y4[(size_t)(j + 2) * bt], y3 = Y4[(size_t)(j + 3) * bt];
                a0 = fmaf(s0.x, y0.x, a0); a1 = fmaf(s0.y, y0.y, a1); a2 = fmaf(s0.z, y0.z, a2); a3 = fmaf(s0.w, y0.w, a3);
                a0 = fmaf(s1.x, y1.x, a0); a1 = fmaf(s1.y, y1.y, a1); a2 = fmaf(s1.z, y1.z, a2); a3 = fmaf(s1.w, y1.w, a3);
                a0 = fmaf(s2.x, y2.x, a0); a1 = fmaf(s2.y, y2.y, a1); a2 = fmaf(s2.z, y2.z, a2); a3 = fmaf(s2.w, y2.w, a3);
                a0 = fmaf(s3.x, y3.x, a0); a1 = fmaf(s3.y, y3.y, a1); a2 = fmaf(s3.z, y3.z, a2); a3 = fmaf(s3.w, y3.w, a3);
            }
            {
                const float4 y0 = Y4[(size_t)(j + 4) * bt], y1 = Y4[(size_t)(j + 5) * bt];
                const float4 y2 = Y4[(size_t)(j + 6) * bt], y3 = Y4[(size_t)(j + 7) * bt];
                a0 = fmaf(s4.x, y0.x, a0); a1 = fmaf(s4.y, y0.y, a1); a2 = fmaf(s4.z, y0.z, a2); a3 = fmaf(s4.w, y0.w, a3);
                a0 = fmaf(s5.x, y1.x, a0); a1 = fmaf(s5.y, y1.y, a1); a2 = fmaf(s5.z, y1.z, a2); a3 = fmaf(s5.w, y1.w, a3);
                a0 = fmaf(s6.x, y2.x, a0); a1 = fmaf(s6.y, y2.y, a1); a2 = fmaf(s6.z, y2.z, a2); a3 = fmaf(s6.w, y2.w, a3);
                a0 = fmaf(s7.x, y3.x, a0); a1 = fmaf(s7.y, y3.y, a1); a2 = fmaf(s7.z, y3.z, a2); a3 = fmaf(s7.w, y3.w, a3);
            }
        }
#else
#pragma unroll 1
        for (int j = 0; j < n4; j += 4) {
            const float4 s0 = __ldg(base + j), s1 = __ldg(base + j + 1), s2 = __ldg(base + j + 2), s3 = __ldg(base + j + 3);
            const float4 y0 = Y4[(size_t)j * bt], y1 = Y4[(size_t)(j + 1) * bt];
            const float4 y2 = Y4[(size_t)(j + 2) * bt], y3 = Y4[(size_t)(j + 3) * bt];
            a0 = fmaf(s0.x, y0.x, a0); a1 = fmaf(s0.y, y0.y, a1); a2 = fmaf(s0.z, y0.z, a2); a3 = fmaf(s0.w, y0.w, a3);
            a0 = fmaf(s1.x, y1.x, a0); a1 = fmaf(s1.y, y1.y, a1); a2 = fmaf(s1.z, y1.z, a2); a3 = fmaf(s1.w, y1.w, a3);
            a0 = fmaf(s2.x, y2.x, a0); a1 = fmaf(s2.y, y2.y, a1); a2 = fmaf(s2.z, y2.z, a2); a3 = fmaf(s2.w, y2.w, a3);
            a0 = fmaf(s3.x, y3.x, a0); a1 = fmaf(s3.y, y3.y, a1); a2 = fmaf(s3.z, y3.z, a2); a3 = fmaf(s3.w, y3.w, a3);
        }
#endif
    }
    const float b = thread_singscat_sum(S, ip, ps, rd);
    if (stream) {
        src_emit(S, sw, startslot, ((a0 + a1) + (a2 + a3)) + b, b * ext);
        if (singlescatter) return b * ext;
    }
    return (((a0 + a1) + (a2 + a3)) + b) * ext;
}
__device__ __forceinline__ float thread_point_source(const DevState &S, const float4 *Y4, int bt, const RayDir &rd,
                                                     bool singlescatter, int ip, float ext, int &ns)
{
    SrcWriter none = SRC_OVF;
    return thread_point_source(S, Y4, bt, rd, singlescatter, ip, ext, ns, none, nullptr, false);
}

// Corner refresh of one thread.  Points shared with the previous cell are found with the reference's
// DONEFACE rule (shdomsub2.f:2395-2397, 2509-2515): after crossing a face normal to axis `jf`, corner n
// of the new cell can only coincide with corner n^bit of the old one; the ids decide.  Reused values
// are bit-identical to a recomputation.  New corners: COMPUTE_SOURCE_1CELL_UNPOL
// (shdomsub2.f:3046-3192) with the TMS-corrected SH block of the point.
__device__ __forceinline__ void thread_refresh(const DevState &S, const CellRec &c, const float4 *Y4, int bt,
                                               const RayDir &rd, bool singlescatter, int jf /*0: first cell*/,
                                               TCorners &K, int &npt_eval, int &nsh_eval, SrcWriter &sw, long long *startslot,
                                               bool stream)
{
    TCorners N;
    unsigned need = 0;
    // one copy of the inheritance per entry face (candidate: old corner n^1, n^2 or n^4): a third of the instructions per cell
#define T_INHERIT(MASK)                                                                             \
    _Pragma("unroll")                                                                               \
    for (int n = 0; n < 8; n++) {                                                                   \
        const int ip = c.gp[n];                                                                     \
        const int k = n ^ (MASK);                                                                   \
        N.x[n] = K.x[k]; N.y[n] = K.y[k]; N.z[n] = K.z[k]; N.ext[n] = K.ext[k]; N.src[n] = K.src[k]; \
        N.pt[n] = ip;                                                                               \
        if (K.pt[k] != ip) need |= 1u << n;                                                         \
    }
    if (jf == 1) { T_INHERIT(1) }
    else if (jf == 2) { T_INHERIT(2) }
    else if (jf == 3) { T_INHERIT(4) }
    else {
#pragma unroll
        for (int n = 0; n < 8; n++) { N.x[n] = N.y[n] = N.z[n] = N.ext[n] = N.src[n] = 0.0f; N.pt[n] = c.gp[n]; }
        need = 0xFFu;
    }
#undef T_INHERIT
    K = N;
    while (need) {
        const int n = __ffs(need) - 1;
        need &= need - 1;
        const int ip = SEL8(K.pt, n);
        const float4 pr = __ldg(&S.ptrec[ip - 1]);
        float src;
        if (S.viewsrc) {
            // all rays of this launch share their direction: the point's SRCEXT was evaluated once (view_source_kernel)
            src = __ldg(&S.viewsrc[ip - 1]);
            npt_eval++; nsh_eval += __ldg(&S.ptsrc[ip - 1]).y & 0xFFFF;
        } else {
            int ns;
            src = thread_point_source(S, Y4, bt, rd, singlescatter, ip, pr.w, ns, sw, startslot, stream);
            npt_eval++; nsh_eval += ns;
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k == n) { K.x[k] = pr.x; K.y[k] = pr.y; K.z[k] = pr.z; K.ext[k] = pr.w; K.src[k] = src; }
    }
}

// Forward integration of one ray by one thread (NSTOKES=1).  MODES as in march_forward (at3d_ray.cuh).
template <int MODES>
__device__ __forceinline__ int thread_march_forward(const DevState &S, const float4 *Y4, int bt, const RayDir &rd, double mu2,
                                    double x0, double y0, double z0, float sky, bool correctinterpolate,
                                    bool singlescatter, bool nosurface, int maxsub,
                                    double &radA, double &radB,
                                    int *trace_cells, int trace_cap, int &ntrace, int &nsubA, int &nsubB,
                                    int &npt_eval, int &nsh_eval, int &nptB, SrcWriter &sw, long long *startslot,
                                                    bool stream)
{
    double xe = x0, ye = y0, ze = z0, trA = 1.0, trB = 1.0;
    float ext1A = 0.0f, srcext1A = 0.0f, ext1B = 0.0f, srcext1B = 0.0f;
    radA = 0.0; radB = 0.0;
    const int p1c = cell_gp(S, 1, 1), p8c = cell_gp(S, 1, 8);
    const double eps = (double)(1.0e-5f * (pt_coord(S, p8c, 3) - pt_coord(S, p1c, 3)));
    const int maxcellscross = 500 * max(S.nx, max(S.ny, S.nz));
    int icell = dev_locate_grid_cell(S, xe, ye, ze);
    int iface = 0, ngrid = 0, jf = 0;
    bool doneA = !(MODES & 1), doneB = !(MODES & 2);
    TCorners K;
    npt_eval = 0; nsh_eval = 0; nptB = 0;
#pragma unroll
    for (int n = 0; n < 8; n++) { K.pt[n] = 0; K.x[n] = K.y[n] = K.z[n] = K.ext[n] = K.src[n] = 0.0f; }
    ntrace = 0; nsubA = 0; nsubB = 0;
    CellRec c;
    if (icell > 0) c = load_cell(S, icell);
    while (!(doneA && doneB) && icell > 0) {
        ngrid++;
        if (trace_cells && ntrace < trace_cap) trace_cells[ntrace] = icell;
        ntrace++;
        const int ne0 = npt_eval;
        thread_refresh(S, c, Y4, bt, rd, singlescatter, jf, K, npt_eval, nsh_eval, sw, startslot,
                       (MODES & 2) && stream && !doneB);
        if ((MODES & 2) && !doneB) nptB += npt_eval - ne0;
        const float q1x = K.x[0], q1y = K.y[0], q1z = K.z[0];
        const float q8x = K.x[7], q8y = K.y[7], q8z = K.z[7];
        const int io = 8 - rd.ioct;
        const float qox = SEL8(K.x, io), qoy = SEL8(K.y, io), qoz = SEL8(K.z, io);
        const double delx = (double)(q8x - q1x), dely = (double)(q8y - q1y), delz = (double)(q8z - q1z);
        const double invdelx = (delx <= 0.0) ? 1.0 : 1.0 / delx;
        const double invdely = (dely <= 0.0) ? 1.0 : 1.0 / dely;
        const double invdelz = 1.0 / delz;
        double u = (xe - q1x) * invdelx, v = (ye - q1y) * invdely, w = (ze - q1z) * invdelz;
        double fc[8];
        if ((MODES & 2) && !doneB) {
            interp_kernel(u, v, w, fc);
            srcext1B = fmaxf(0.0f, (float)fcsum(fc, K.src));
            ext1B = (float)fcsum(fc, K.ext);
        }
        if ((MODES & 1) && !doneA && (correctinterpolate || ngrid == 1)) {
            srcext1A = fmaxf(0.0f, (float)trilerp(K.src, u, v, w));
            ext1A = (float)trilerp(K.ext, u, v, w);
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
            const float extn = (float)trilerp(K.ext, u, v, w);
            const double taugrid = so * 0.5f * (ext1A + extn);
            int ntau = 1 + (int)(taugrid / S.tautol);
            if (ntau < 1) ntau = 1;
            const double dels = so / ntau;
            for (int it = 1; it <= ntau; it++) {
                const double s = it * dels;
                const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
                const double ui = (xi - q1x) * invdelx, vi = (yi - q1y) * invdely, wi = (zi - q1z) * invdelz;
                const float ext0 = (float)trilerp(K.ext, ui, vi, wi);
                const float srcext0 = fmaxf(0.0f, (float)trilerp(K.src, ui, vi, wi));
                const double ext = (double)(0.5f * (ext0 + ext1A));
                if (ext != 0.0) {
                    const double tau = ext * dels;
                    const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                    const double transcell = 1.0f - abscell;
                    const double src = (0.5f * (srcext0 + srcext1A)
                        + 0.08333333333f * (ext0 * srcext1A - ext1A * srcext0) * dels
                          * (1.0f - 0.05f * (ext1A - ext0) * dels)) / ext;
                    radA = radA + trA * src * abscell;
                    trA = trA * transcell;
                }
                nsubA++;
                ext1A = ext0;
                srcext1A = srcext0;
            }
        }
        if ((MODES & 2) && !doneB) {
            float extn;
            { double fcn[8]; interp_kernel(u, v, w, fcn); extn = (float)fcsum(fcn, K.ext); }
            const double taugrid = so * 0.5f * (ext1B + extn);
            int ntau = 1 + (int)(taugrid / S.tautol);
            if (ntau < 1) ntau = 1;
            const double dels = so / ntau;
            for (int it = 1; it <= ntau; it++) {
                const double s = it * dels;
                const double xi = xe + s * rd.cx, yi = ye + s * rd.cy, zi = ze + s * rd.cz;
                const double ui = (xi - q1x) * invdelx, vi = (yi - q1y) * invdely, wi = (zi - q1z) * invdelz;
                interp_kernel(ui, vi, wi, fc);
                const float srcext0 = fmaxf(0.0f, (float)fcsum(fc, K.src));
                const float ext0 = (it != ntau) ? (float)fcsum(fc, K.ext) : extn;
                const double ext = (double)(0.5f * (ext0 + ext1B));
                if (ext != 0.0) {
                    const double tau = ext * dels;
                    const double abscell = tau * (1.0f - 0.5f * tau * (1.0f - 0.33333333333f * tau));
                    const double transcell = 1.0f - abscell;
                    const double src = (0.5f * (srcext0 + srcext1B)
                        + 0.08333333333f * (ext0 * srcext1B - ext1B * srcext0) * dels
                          * (1.0f - 0.05f * (ext1B - ext0) * dels)) / ext;
                    radB = radB + trB * src * abscell;
                    trB = trB * transcell;
                    nsubB++;
                    if (nsubB + 1 > maxsub) return 4;
                }
                ext1B = ext0;
                srcext1B = srcext0;
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
                    if (!nosurface) {
                        SurfHit h;
                        h.xb = xn; h.yb = yn; h.transmit = trA; h.icell = ic; h.kface = kface;
                        h.rad[0] = radA; h.rad[1] = 0.0; h.rad[2] = 0.0;
                        *rd.hit = h;
                    }
                } else {
                    float radbnd[1];
                    const int e = boundary_radiance<1, false>(S, xn, yn, (float)mu2, rd.phi2, sky, ic, kface, radbnd,
                                                              nullptr, nullptr, nullptr);
                    if (e) return e;
                    if (!nosurface) radA = radA + trA * radbnd[0];
                }
            }
        }
        if ((MODES & 2) && !doneB) {
            if (trB < S.transcut) doneB = true;
            else if (atbnd) {
                doneB = true;
                float radbnd[1];
                const int e = boundary_radiance<1, true>(S, xn, yn, (float)mu2, rd.phi2, sky, ic, kface, radbnd,
                                                         nullptr, nullptr, nullptr);
                if (e) return e;
                if (!nosurface) radB = radB + trB * radbnd[0];
            }
        }
        if (!atbnd) { icell = inextcell; c = cn; jf = jface; }
        xe = xn; ye = yn; ze = zn;
    }
    return 0;
}
