// at3d_prep.cu -- per-evaluation input preparation of the gradient on the device (sm_100a):
//   at3d_prepare_deriv_interps   PREPARE_DERIV_INTERPS / COMPUTE_INTERP_WEIGHTS (src/polarized/shdomsub4.f:2917-3169)
//   at3d_make_direct             MAKE_DIRECT / DIRECT_BEAM_PROP (shdomsub2.f:393-478, shdom90.f90:352-867)
//   at3d_make_direct_derivative  MAKE_DIRECT_DERIVATIVE / DIRECT_BEAM_AND_PATHS_PROP (src/shdomsub5.f:1553-2004)
// These run once per cost-function evaluation (StateGenerator rebuilds the solvers, medium.py:1813-1831).
// All three are embarrassingly parallel over grid points / property points: one thread each.
#include "at3d_mem.h"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <vector>
#include "at3d_host.h"

static void set_msg(char *errmsg, const char *fmt, ...)
{
    if (!errmsg) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, AT3D_ERRMSG_LEN, fmt, ap);
    va_end(ap);
}

namespace {
struct Arena {      // device allocations of one call
    std::vector<void *> p;
    ~Arena() { for (void *q : p) at3d_free(q); }
    template <typename T> T *alloc(size_t n)
    {
        void *q = nullptr;
        if (at3d_malloc(&q, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
        p.push_back(q);
        return (T *)q;
    }
    template <typename T> const T *up(const T *h, size_t n)
    {
        if (!h) return nullptr;
        T *d = alloc<T>(n);
        if (!d) return nullptr;
        if (cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return d;
    }
};
}

// ------------------------------------------------------------------------------------------
// PREPARE_DERIV_INTERPS
// ------------------------------------------------------------------------------------------
struct InterpGeom { int npx, npy, npz; float delx, dely, xstart, ystart; };

// COMPUTE_INTERP_WEIGHTS (shdomsub4.f:3082-3169), one thread per RTE grid point
__global__ void interp_weights_kernel(int npts, InterpGeom g, const float *gridpos, const float *zlevels,
                                      int *interpptr, float *optinterpwt, int *bad)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npts) return;
    const float x = gridpos[3 * (size_t)ip], y = gridpos[3 * (size_t)ip + 1], z = gridpos[3 * (size_t)ip + 2];
    int il = 0, iu = g.npz, im;
    while (iu - il > 1) { im = (iu + il) / 2; if (z >= zlevels[im - 1]) il = im; else iu = im; }
    const int iz = il > 1 ? il : 1;
    double w = (double)(z - zlevels[iz - 1]) / (zlevels[iz] - zlevels[iz - 1]);
    w = fmax(fmin(w, 1.0), 0.0);
    int ix = (int)((x - g.xstart) / g.delx) + 1;
    if (fabsf(x - g.xstart - g.npx * g.delx) < 0.01f * g.delx) ix = g.npx;
    if (ix < 1 || ix > g.npx) { atomicCAS(bad, 0, 2 * (ip + 1)); return; }
    const int ixp = (ix % g.npx) + 1;
    double u = (double)(x - g.xstart - g.delx * (ix - 1)) / g.delx;
    u = fmax(fmin(u, 1.0), 0.0);
    if (u < 1.0e-5) u = 0.0;
    if (u > 1.0 - 1.0e-5) u = 1.0;
    int iy = (int)((y - g.ystart) / g.dely) + 1;
    if (fabsf(y - g.ystart - g.npy * g.dely) < 0.01f * g.dely) iy = g.npy;
    if (iy < 1 || iy > g.npy) { atomicCAS(bad, 0, 2 * (ip + 1) + 1); return; }
    const int iyp = (iy % g.npy) + 1;
    double v = (double)(y - g.ystart - g.dely * (iy - 1)) / g.dely;
    v = fmax(fmin(v, 1.0), 0.0);
    if (v < 1.0e-5) v = 0.0;
    if (v > 1.0 - 1.0e-5) v = 1.0;
    float *wt = optinterpwt + 8 * (size_t)ip;
    int *pt = interpptr + 8 * (size_t)ip;
    wt[0] = (float)((1 - u) * (1 - v) * (1 - w));
    wt[1] = (float)(u * (1 - v) * (1 - w));
    wt[2] = (float)((1 - u) * v * (1 - w));
    wt[3] = (float)(u * v * (1 - w));
    wt[4] = (float)((1 - u) * (1 - v) * w);
    wt[5] = (float)(u * (1 - v) * w);
    wt[6] = (float)((1 - u) * v * w);
    wt[7] = (float)(u * v * w);
    const int i1 = iz + g.npz * (iy - 1) + g.npz * g.npy * (ix - 1);
    const int i2 = iz + g.npz * (iy - 1) + g.npz * g.npy * (ixp - 1);
    const int i3 = iz + g.npz * (iyp - 1) + g.npz * g.npy * (ix - 1);
    const int i4 = iz + g.npz * (iyp - 1) + g.npz * g.npy * (ixp - 1);
    pt[0] = i1; pt[1] = i2; pt[2] = i3; pt[3] = i4;
    pt[4] = i1 + 1; pt[5] = i2 + 1; pt[6] = i3 + 1; pt[7] = i4 + 1;
}

struct PdiArgs {
    int npts, maxpg, numder, ml, nstleg, nleg, pmaxnmicro, dmaxnmicro, nq, deltam, interp_new;
    float phasemax;
    const int *partder, *doexact, *iphasep, *diphasep, *iphase, *interpptr;
    const float *legen, *dleg, *phasewtp, *dphasewtp, *albedop, *extinctp, *dext, *dalb, *albedo, *phaseinterpwt;
    float *fp, *dfp, *dextm, *dalbm, *dfj;
};

// delta-M scaled extinction derivative on the property grid (shdomsub4.f:2994-3025)
__global__ void pdi_property_kernel(PdiArgs a)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.maxpg * a.numder) return;
    const int ib = (int)(t % a.maxpg), idr = (int)(t / a.maxpg);
    const int ipa = a.partder[idr] - 1;
    const int nlt = a.nstleg * (a.nleg + 1);
    float fp = 0.0f, dfp = 0.0f;
    const float albp = a.albedop[ib + (size_t)a.maxpg * ipa], extp = a.extinctp[ib + (size_t)a.maxpg * ipa];
    const float dext = a.dext[t], dalb = a.dalb[t];
    if (a.deltam) {
        for (int q = 0; q < a.dmaxnmicro; q++) {
            const float pwp = a.phasewtp[q + (size_t)a.pmaxnmicro * (ib + (size_t)a.maxpg * ipa)];
            const int iphp = a.iphasep[q + (size_t)a.pmaxnmicro * (ib + (size_t)a.maxpg * ipa)];
            fp = fp + pwp * a.legen[(size_t)nlt * (iphp - 1) + a.nstleg * (a.ml + 1)];
            if (a.doexact[idr] == 1) {
                const int dip = a.diphasep[q + (size_t)a.dmaxnmicro * (ib + (size_t)a.maxpg * idr)];
                dfp = dfp + pwp * a.dleg[(size_t)nlt * (dip - 1) + a.nstleg * (a.ml + 1)];
            } else if (a.doexact[idr] == 0) {
                const float dpw = a.dphasewtp[q + (size_t)a.dmaxnmicro * (ib + (size_t)a.maxpg * idr)];
                dfp = dfp + dpw * a.legen[(size_t)nlt * (iphp - 1) + a.nstleg * (a.ml + 1)];
            }
        }
    }
    a.fp[t] = fp; a.dfp[t] = dfp;
    a.dextm[t] = dext * (1 - fp * albp) - dalb * fp * extp - extp * albp * dfp;
}

// delta-M scaled albedo / truncation-fraction derivatives per RTE point and property corner (:3027-3076)
__global__ void pdi_point_kernel(PdiArgs a)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.npts * a.numder) return;
    const int ip = (int)(t % a.npts), idr = (int)(t / a.npts);
    const int ipa = a.partder[idr] - 1;
    const int nlt = a.nstleg * (a.nleg + 1);
    const int *iph = a.iphase + (size_t)a.nq * (ip + (size_t)a.npts * ipa);
    const float *pw = a.phaseinterpwt + (size_t)a.nq * (ip + (size_t)a.npts * ipa);
    const float alb = a.albedo[ip + (size_t)a.npts * ipa];
    float f = 0.0f, albedoj;
    if (a.deltam) {
        if (!a.interp_new || pw[0] >= a.phasemax) f = a.legen[(size_t)nlt * (iph[0] - 1) + a.nstleg * (a.ml + 1)];
        else
            for (int q = 0; q < a.nq; q++) {
                if (pw[q] < 1e-7f) continue;
                f = f + pw[q] * a.legen[(size_t)nlt * (iph[q] - 1) + a.nstleg * (a.ml + 1)];
            }
        albedoj = alb / (f * (alb - 1) + 1);
    } else albedoj = alb;
    const float divide = 1.0f / (1.0f - albedoj * f);
    for (int nb = 0; nb < 8; nb++) {
        const int ib = a.interpptr[nb + 8 * (size_t)ip] - 1;
        const float albp = a.albedop[ib + (size_t)a.maxpg * ipa], extp = a.extinctp[ib + (size_t)a.maxpg * ipa];
        const float dext = a.dext[ib + (size_t)a.maxpg * idr], dalb = a.dalb[ib + (size_t)a.maxpg * idr];
        const float fp = a.fp[ib + (size_t)a.maxpg * idr], dfp = a.dfp[ib + (size_t)a.maxpg * idr];
        a.dalbm[nb + 8 * ((size_t)ip + (size_t)a.npts * idr)] = divide * (
            dext * ((1 - f) * (albp - albedoj) + (albedoj - 1) * albp * (fp - f))
            + dalb * ((1 - f) * extp + (albedoj - 1) * extp * (fp - f))
            + dfp * (albedoj - 1) * extp * albp);
        a.dfj[nb + 8 * ((size_t)ip + (size_t)a.npts * idr)] =
            (dext * (fp - f) * albp + dalb * (fp - f) * extp + dfp * extp * albp) / (1 - f);
    }
}

extern "C" int at3d_prepare_deriv_interps(const at3d_state_desc *d, int npx, int npy, int npz, int maxpg,
                                          float delx, float dely, float xstart, float ystart,
                                          const float *zlevels, const at3d_grad_desc *g,
                                          float *optinterpwt, int32_t *interpptr,
                                          float *dalbm, float *dextm, float *dfj, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !g || !zlevels || !optinterpwt || !interpptr || !dalbm || !dextm || !dfj) { set_msg(errmsg, "null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    Arena A;
    const size_t npts = d->npts, nd = g->numder, mp = maxpg;
    const size_t nlt = (size_t)d->nstleg * (d->nleg + 1);
    const int nq = 8 * d->maxnmicro;
    PdiArgs a;
    memset(&a, 0, sizeof(a));
    a.npts = d->npts; a.maxpg = maxpg; a.numder = g->numder; a.ml = d->ml; a.nstleg = d->nstleg; a.nleg = d->nleg;
    a.pmaxnmicro = d->maxnmicro; a.dmaxnmicro = g->deriv_maxnmicro; a.nq = nq; a.deltam = d->deltam;
    a.interp_new = d->interp_new; a.phasemax = d->phasemax;
    const float *gridpos = A.up(d->gridpos, 3 * npts);
    const float *zl = A.up(zlevels, (size_t)npz);
    a.partder = A.up(g->partder, nd); a.doexact = A.up(g->doexact, nd);
    a.iphasep = A.up(g->iphasep, (size_t)d->maxnmicro * mp * d->npart);
    a.phasewtp = A.up(g->phasewtp, (size_t)d->maxnmicro * mp * d->npart);
    a.diphasep = A.up(g->diphasep, (size_t)g->deriv_maxnmicro * mp * nd);
    a.dphasewtp = A.up(g->dphasewtp, (size_t)g->deriv_maxnmicro * mp * nd);
    a.iphase = A.up(d->iphase, (size_t)nq * npts * d->npart);
    a.phaseinterpwt = A.up(d->phaseinterpwt, (size_t)nq * npts * d->npart);
    a.legen = A.up(d->legen, nlt * d->numphase);
    a.dleg = A.up(g->dleg, nlt * g->dnumphase);
    a.albedop = A.up(g->albedop, mp * d->npart); a.extinctp = A.up(g->extinctp, mp * d->npart);
    a.dext = A.up(g->dext, mp * nd); a.dalb = A.up(g->dalb, mp * nd);
    a.albedo = A.up(d->albedo, npts * d->npart);
    int *iptr_d = A.alloc<int>(8 * npts); float *wt_d = A.alloc<float>(8 * npts);
    a.fp = A.alloc<float>(mp * nd); a.dfp = A.alloc<float>(mp * nd); a.dextm = A.alloc<float>(mp * nd);
    a.dalbm = A.alloc<float>(8 * npts * nd); a.dfj = A.alloc<float>(8 * npts * nd);
    int *bad = A.alloc<int>(1);
    if (!gridpos || !zl || !a.partder || !a.doexact || !a.iphasep || !a.phasewtp || !a.diphasep || !a.dphasewtp ||
        !a.iphase || !a.phaseinterpwt || !a.legen || !a.dleg || !a.albedop || !a.extinctp || !a.dext || !a.dalb ||
        !a.albedo || !iptr_d || !wt_d || !a.fp || !a.dfp || !a.dextm || !a.dalbm || !a.dfj || !bad) {
        set_msg(errmsg, "at3d_prepare_deriv_interps: NULL input array or device allocation failure");
        return 4;
    }
    a.interpptr = iptr_d;
    if (g->deriv_maxnmicro > d->maxnmicro) { set_msg(errmsg, "DERIV_MAXNMICRO > MAXNMICRO is not supported"); return 3; }
    cudaMemset(bad, 0, sizeof(int));
    InterpGeom ig = {npx, npy, npz, delx, dely, xstart, ystart};
    interp_weights_kernel<<<(unsigned)((npts + 127) / 128), 128>>>((int)npts, ig, gridpos, zl, iptr_d, wt_d, bad);
    int hbad = 0;
    if (cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { set_msg(errmsg, "CUDA error in interp_weights_kernel"); return 4; }
    if (hbad) { set_msg(errmsg, "TRILIN: Beyond %s domain (grid point %d)", (hbad & 1) ? "Y" : "X", hbad / 2); return 1; }
    pdi_property_kernel<<<(unsigned)((mp * nd + 127) / 128), 128>>>(a);
    pdi_point_kernel<<<(unsigned)((npts * nd + 127) / 128), 128>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(optinterpwt, wt_d, 8 * npts * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(interpptr, iptr_d, 8 * npts * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dextm, a.dextm, mp * nd * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dalbm, a.dalbm, 8 * npts * nd * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dfj, a.dfj, 8 * npts * nd * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_prepare_deriv_interps", cudaGetErrorString(e)); return 4; }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Direct beam on the property grid
// ------------------------------------------------------------------------------------------
#include "at3d_beam.cuh"

// delta-M scaled property-grid extinction EXTDIRP (DIRECT_BEAM_PROP INIT=1, shdom90.f90:453-483)
__global__ void extdirp_kernel(int maxpg, int npz, int npart, int pmaxnmicro, int deltam, int ml, int nstleg,
                               int nlegp, const float *extinctp, const float *albedop, const float *legenp,
                               const int *iphasep, const float *phasewtp, const float *gasext, float *extdirp)
{
    const int ib = blockIdx.x * blockDim.x + threadIdx.x;
    if (ib >= maxpg) return;
    const int iz = ib % npz;
    float acc = 0.0f;
    for (int ipa = 0; ipa < npart; ipa++) {
        double extinct = extinctp[ib + (size_t)maxpg * ipa];
        double albedo = albedop[ib + (size_t)maxpg * ipa];
        if (gasext[iz] > 0.0f) {
            albedo = albedo * extinct / (extinct + gasext[iz]);
            extinct = extinct + gasext[iz];
        }
        if (deltam) {
            const int l = ml + 1;
            double f = 0.0;
            for (int q = 0; q < pmaxnmicro; q++) {
                const int iph = iphasep[q + (size_t)pmaxnmicro * (ib + (size_t)maxpg * ipa)];
                const float pw = phasewtp[q + (size_t)pmaxnmicro * (ib + (size_t)maxpg * ipa)];
                f = f + pw * legenp[nstleg * (l + (size_t)(nlegp + 1) * (iph - 1))] / (2 * l + 1);
            }
            extinct = (1.0f - albedo * f) * extinct;
        }
        acc = (float)(acc + extinct);
    }
    extdirp[ib] = acc;
}

__global__ void make_direct_kernel(int npts, BeamGeom g, const float *zl, const float *gridpos,
                                   const float *extdirp, float solarflux, float *dirflux, int *longest, int *bad)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npts) return;
    double path; int npp;
    BeamNoSink none;
    const int e = beam_walk<false>(g, zl, gridpos[3 * (size_t)ip], gridpos[3 * (size_t)ip + 1],
                                   gridpos[3 * (size_t)ip + 2], extdirp, path, npp, none);
    if (e) { atomicCAS(bad, 0, 8 * (ip + 1) + e); return; }
    dirflux[ip] = (float)(solarflux * exp(-path));
    atomicMax(longest, npp);
}

__global__ void make_direct_derivative_kernel(int npts, BeamGeom g, const float *zl, const float *gridpos,
                                              float *dpath, int *dptr, int longest_path_pts, int *bad)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= npts) return;
    double path; int npp;
    BeamDenseSink sink{dpath + (size_t)longest_path_pts * ip, dptr + (size_t)longest_path_pts * ip, longest_path_pts, 0};
    const int e = beam_walk<true>(g, zl, gridpos[3 * (size_t)ip], gridpos[3 * (size_t)ip + 1],
                                  gridpos[3 * (size_t)ip + 2], nullptr, path, npp, sink);
    if (e) atomicCAS(bad, 0, 8 * (ip + 1) + e);
}

static int beam_error(int code, const char *who, char *errmsg)
{
    static const char *txt[] = {"", "Beyond X domain", "Beyond Y domain", "beyond grid!", "SO<0",
                                "Max number of property points to pass exceeded"};
    const int e = code & 7;
    set_msg(errmsg, "%s: %s (grid point %d)", who, txt[e < 6 ? e : 0], code / 8);
    return 1;
}

extern "C" int at3d_make_direct(int npts, int bcflag, int ipflag, int deltam, int ml, int nstleg, int nlegp,
                                float solarflux, float solarmu, float solaraz, const float *gridpos,
                                int npx, int npy, int npz, float delx, float dely, float xstart, float ystart,
                                const float *zlevels, const float *extinctp, const float *albedop,
                                const float *legenp, int numphase, const int32_t *iphasep, const float *phasewtp,
                                int maxnmicro, int npart, int nzckd, const float *zckd, const float *gasabs,
                                float *extdirp, float *dirflux, double *out_d, int32_t *out_i, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!gridpos || !zlevels || !extinctp || !albedop || !legenp || !iphasep || !phasewtp || !extdirp || !dirflux ||
        !out_d || !out_i) { set_msg(errmsg, "null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    const int maxpg = npx * npy * npz;
    // gas extinction per property level (shdom90.f90:425-451), host side: NPZ values
    std::vector<float> gasext(npz, 0.0f);
    for (int iz = 1; iz <= npz && nzckd > 0; iz++) {
        int il = 1, iu = nzckd, im;
        while (iu - il > 1) { im = (iu + il) / 2; if (zlevels[iz - 1] <= zckd[im - 1]) il = im; else iu = im; }
        int i = il > 1 ? il : 1;
        if (i > nzckd - 1) i = nzckd - 1;
        double w0 = (zlevels[iz - 1] - zckd[i - 1]) / (zckd[i] - zckd[i - 1]);
        w0 = fmin(fmax(w0, 0.0), 1.0);
        gasext[iz - 1] = (float)((1.0f - w0) * gasabs[i - 1] + w0 * gasabs[i]);
    }
    // beam geometry constants (shdom90.f90:485-560): a handful of scalars, host libm like the reference
    BeamGeom g;
    g.bcflag = bcflag; g.npx = npx; g.npy = npy; g.npz = npz; g.xstart = xstart; g.ystart = ystart;
    g.ipdirect = ipflag;
    if (BT(ipflag, 2)) g.ipdirect = 0;
    const double sunmu = -solarmu, sunaz = solaraz + acosf(-1.0f);
    g.cx = sqrt(1.0f - sunmu * sunmu) * cos(sunaz);
    g.cy = sqrt(1.0f - sunmu * sunmu) * sin(sunaz);
    g.cz = fabs(sunmu);
    if (fabs(g.cx) > 1.0e-6f) g.cxinv = 1.0 / g.cx; else { g.cx = 0.0; g.cxinv = 1.0e20f; }
    if (fabs(g.cy) > 1.0e-6f) g.cyinv = 1.0 / g.cy; else { g.cy = 0.0; g.cyinv = 1.0e20f; }
    if (fabs(g.cz) > 1.0e-6f) g.czinv = 1.0 / g.cz; else { g.cz = 0.0; g.czinv = 1.0e20f; }
    g.di = std::signbit(g.cx) ? -1 : 1;
    g.dj = std::signbit(g.cy) ? -1 : 1;
    g.dk = std::signbit(g.cz) ? -1 : 1;
    const double epsz = 1.0e-6f * (zlevels[npz - 1] - zlevels[0]);
    double epss = 1.0e-3f * (zlevels[npz - 1] - zlevels[0]) / npz;
    if (!BT(g.ipdirect, 0)) epss = fmax(epss, 1.0e-4 * delx);
    if (!BT(g.ipdirect, 1)) epss = fmax(epss, 1.0e-4 * dely);
    g.delxd = (double)delx; g.delyd = (double)dely;
    epss = fmax(fmax(0.001f * g.delxd, 0.001f * g.delyd), epss);
    g.epss = epss; g.epsz = epsz;
    g.xdomain = g.delxd * npx;
    if (BT(bcflag, 2)) g.xdomain = g.delxd * (npx - 1);
    g.ydomain = g.delyd * npy;
    if (BT(bcflag, 3)) g.ydomain = g.delyd * (npy - 1);
    // UNIFORMZLEV (shdom90.f90:520-548): level above which the medium is horizontally uniform
    double uniformzlev;
    {
        std::vector<float> emin(npz, 1.0e20f), emax(npz, 0.0f);
        for (int ib = 0; ib < maxpg; ib++) {
            float s = 0.0f;
            for (int ipa = 0; ipa < npart; ipa++) s = s + extinctp[ib + (size_t)maxpg * ipa];
            const int iz = ib % npz;
            emin[iz] = fminf(s, emin[iz]); emax[iz] = fmaxf(s, emax[iz]);
        }
        int jz = 0;
        for (int iz = 1; iz <= npz; iz++) if (emax[iz - 1] - emin[iz - 1] > 1.0e-4f) jz = iz;
        jz = jz + 1 < npz ? jz + 1 : npz;
        uniformzlev = zlevels[jz - 1];
    }
    Arena A;
    const float *gp = A.up(gridpos, 3 * (size_t)npts), *zl = A.up(zlevels, (size_t)npz);
    const float *ex = A.up(extinctp, (size_t)maxpg * npart), *al = A.up(albedop, (size_t)maxpg * npart);
    const float *lg = A.up(legenp, (size_t)nstleg * (nlegp + 1) * numphase);
    const int *iq = A.up(iphasep, (size_t)maxnmicro * maxpg * npart);
    const float *pw = A.up(phasewtp, (size_t)maxnmicro * maxpg * npart);
    const float *ge = A.up(gasext.data(), (size_t)npz);
    float *ed = A.alloc<float>(maxpg), *df = A.alloc<float>(npts);
    int *flags = A.alloc<int>(2);
    if (!gp || !zl || !ex || !al || !lg || !iq || !pw || !ge || !ed || !df || !flags) { set_msg(errmsg, "at3d_make_direct: device allocation failure"); return 4; }
    cudaMemset(flags, 0, 2 * sizeof(int));
    extdirp_kernel<<<(maxpg + 127) / 128, 128>>>(maxpg, npz, npart, maxnmicro, deltam, ml, nstleg, nlegp, ex, al, lg, iq, pw, ge, ed);
    make_direct_kernel<<<(npts + 127) / 128, 128>>>(npts, g, zl, gp, ed, solarflux, df, flags, flags + 1);
    int hf[2] = {0, 0};
    cudaError_t e = cudaMemcpy(hf, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(extdirp, ed, (size_t)maxpg * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dirflux, df, (size_t)npts * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_make_direct", cudaGetErrorString(e)); return 4; }
    if (hf[1]) return beam_error(hf[1], "DIRECT_BEAM_PROP", errmsg);
    out_d[0] = g.cx; out_d[1] = g.cy; out_d[2] = g.cz; out_d[3] = g.cxinv; out_d[4] = g.cyinv; out_d[5] = g.czinv;
    out_d[6] = g.epss; out_d[7] = g.epsz; out_d[8] = g.xdomain; out_d[9] = g.ydomain; out_d[10] = uniformzlev;
    out_d[11] = g.delxd; out_d[12] = g.delyd;
    out_i[0] = g.ipdirect; out_i[1] = g.di; out_i[2] = g.dj; out_i[3] = g.dk; out_i[4] = hf[0];
    return 0;
}

// DIRECT_BEAM_PROP for `count` device-resident points (the call of INTERPOLATE_POINT, shdomsub1.f:5093-5107), with the
// beam constants at3d_make_direct returned (out_d / out_i); flags[2] as in make_direct_kernel (flags[1] != 0: error).
cudaError_t launch_direct_points(const double *out_d, const int *out_i, int bcflag, int npx, int npy, int npz,
                                 float xstart, float ystart, const float *zlevels_d, const float *gridpos_d,
                                 const float *extdirp_d, float solarflux, float *dirflux_d, int count, int *flags_d,
                                 cudaStream_t s)
{
    if (count < 1) return cudaSuccess;
    BeamGeom g;
    g.bcflag = bcflag; g.npx = npx; g.npy = npy; g.npz = npz; g.xstart = xstart; g.ystart = ystart;
    g.cx = out_d[0]; g.cy = out_d[1]; g.cz = out_d[2]; g.cxinv = out_d[3]; g.cyinv = out_d[4]; g.czinv = out_d[5];
    g.epss = out_d[6]; g.epsz = out_d[7]; g.xdomain = out_d[8]; g.ydomain = out_d[9]; g.delxd = out_d[11]; g.delyd = out_d[12];
    g.ipdirect = out_i[0]; g.di = out_i[1]; g.dj = out_i[2]; g.dk = out_i[3];
    make_direct_kernel<<<(count + 127) / 128, 128, 0, s>>>(count, g, zlevels_d, gridpos_d, extdirp_d, solarflux, dirflux_d,
                                                           flags_d, flags_d + 1);
    return cudaGetLastError();
}

extern "C" int at3d_make_direct_derivative(int npts, int bcflag, int npx, int npy, int npz,
                                           float delx, float dely, float xstart, float ystart,
                                           const float *gridpos, const float *zlevels,
                                           int ipdirect, int di, int dj, int dk,
                                           double cx, double cy, double cz,
                                           double cxinv, double cyinv, double czinv,
                                           double epss, double epsz, double xdomain, double ydomain,
                                           double uniformzlev, double delxd, double delyd,
                                           float *dpath, int32_t *dptr, int longest_path_pts, char *errmsg)
{
    (void)delx; (void)dely; (void)uniformzlev;
    if (errmsg) errmsg[0] = 0;
    if (!gridpos || !zlevels || !dpath || !dptr || longest_path_pts < 1) { set_msg(errmsg, "bad argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    BeamGeom g;
    g.bcflag = bcflag; g.npx = npx; g.npy = npy; g.npz = npz; g.ipdirect = ipdirect; g.di = di; g.dj = dj; g.dk = dk;
    g.cx = cx; g.cy = cy; g.cz = cz; g.cxinv = cxinv; g.cyinv = cyinv; g.czinv = czinv; g.epss = epss; g.epsz = epsz;
    g.xdomain = xdomain; g.ydomain = ydomain; g.delxd = delxd; g.delyd = delyd; g.xstart = xstart; g.ystart = ystart;
    Arena A;
    const size_t n = (size_t)longest_path_pts * npts;
    const float *gp = A.up(gridpos, 3 * (size_t)npts), *zl = A.up(zlevels, (size_t)npz);
    float *dp = A.alloc<float>(n); int *dq = A.alloc<int>(n); int *bad = A.alloc<int>(1);
    if (!gp || !zl || !dp || !dq || !bad) { set_msg(errmsg, "at3d_make_direct_derivative: device allocation failure"); return 4; }
    cudaMemset(dp, 0, n * sizeof(float)); cudaMemset(dq, 0, n * sizeof(int)); cudaMemset(bad, 0, sizeof(int));
    make_direct_derivative_kernel<<<(npts + 127) / 128, 128>>>(npts, g, zl, gp, dp, dq, longest_path_pts, bad);
    int hbad = 0;
    cudaError_t e = cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dpath, dp, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dptr, dq, n * sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_make_direct_derivative", cudaGetErrorString(e)); return 4; }
    if (hbad) return beam_error(hbad, "DIRECT_BEAM_AND_PATHS_PROP", errmsg);
    return 0;
}

// ------------------------------------------------------------------------------------------
// TRANSFER_PA_TO_GRID: TRILIN_INTERP_PROP (src/polarized/shdom90.f90:17-346) + the delta-M scaling of PREPARE_PROP
// (src/polarized/shdomsub2.f:479-608) in the 'N' interpolation mode, thread = grid point.
// Same operation order as the host restatement at3d_b200/medium.py: transfer_pa_to_grid (double accumulators over the
// 8 property corners in corner order; duplicate phase-table entries merged into their first occurrence; stable sort by
// descending weight; REAL arithmetic in the delta-M step).
// ------------------------------------------------------------------------------------------
// SSORT (src/polarized/shdomsub2.f:4961-5244; Singleton's quicksort, KFLAG=-2: decreasing X carrying Y) for the short
// per-point lists of TRILIN_INTERP_PROP.  The order of equal keys follows the reference's algorithm, so IPHASE comes out
// identical to the Fortran's, entries of weight zero included.  Arrays are 1-based inside, n <= TPA_MAXQ.
__host__ __device__ inline void tpa_ssort_desc(float *x0, int *y0, int n)
{
    float *x = x0 - 1;
    int *y = y0 - 1;
    float r = 0.375f, t, tt;
    int ty, tty, i = 1, j = n, k, l, m = 1, ij;
    int il[24], iu[24];
    if (n < 1) return;
    for (k = 1; k <= n; k++) x[k] = -x[k];
    int state = 110;     // the labels of the Fortran as a small state machine
    for (;;) {
        if (state == 110) {
            if (i == j) { state = 150; continue; }
            if (r <= 0.5898437f) r = r + 3.90625e-2f; else r = r - 0.21875f;
            state = 120;
        }
        if (state == 120) {
            k = i;
            ij = i + (int)((j - i) * r);
            t = x[ij]; ty = y[ij];
            if (x[i] > t) { x[ij] = x[i]; x[i] = t; t = x[ij]; y[ij] = y[i]; y[i] = ty; ty = y[ij]; }
            l = j;
            if (x[j] < t) {
                x[ij] = x[j]; x[j] = t; t = x[ij]; y[ij] = y[j]; y[j] = ty; ty = y[ij];
                if (x[i] > t) { x[ij] = x[i]; x[i] = t; t = x[ij]; y[ij] = y[i]; y[i] = ty; ty = y[ij]; }
            }
            for (;;) {
                do { l = l - 1; } while (x[l] > t);
                do { k = k + 1; } while (x[k] < t);
                if (k <= l) { tt = x[l]; x[l] = x[k]; x[k] = tt; tty = y[l]; y[l] = y[k]; y[k] = tty; }
                else break;
            }
            if (l - i > j - k) { il[m] = i; iu[m] = l; i = k; m = m + 1; }
            else { il[m] = k; iu[m] = j; j = l; m = m + 1; }
            state = 160;
        }
        if (state == 150) {
            m = m - 1;
            if (m == 0) break;
            i = il[m]; j = iu[m];
            state = 160;
        }
        if (state == 160) {
            if (j - i >= 1) { state = 120; continue; }
            if (i == 1) { state = 110; continue; }
            i = i - 1;
            for (;;) {
                i = i + 1;
                if (i == j) break;
                t = x[i + 1]; ty = y[i + 1];
                if (x[i] <= t) continue;
                k = i;
                do { x[k + 1] = x[k]; y[k + 1] = y[k]; k = k - 1; } while (t < x[k]);
                x[k + 1] = t; y[k + 1] = ty;
            }
            state = 150;
        }
    }
    for (k = 1; k <= n; k++) x[k] = -x[k];
}

// TRILIN_INTERP_PROP (src/polarized/shdom90.f90:78-346) + the per-point delta-M scaling of PREPARE_PROP
// (shdomsub2.f:554-571) / INTERPOLATE_POINT (shdomsub1.f:5063-5095), thread = grid point of [first, first+count).
// Every expression keeps the reference's precision (REAL products, DOUBLE PRECISION interpolation factors), so the
// outputs are bit-identical to the Fortran's for both INTERPMETHOD(2:2) modes.  Point arrays have leading dimension ld.
__global__ void tpa_kernel(TpaArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const int i = a.first + t;
    const float x = a.gridpos[3 * (size_t)i], y = a.gridpos[3 * (size_t)i + 1], z = a.gridpos[3 * (size_t)i + 2];
    const int npx = a.npx, npy = a.npy, npz = a.npz, mnm = a.mnm, nq = 8 * a.mnm;
    int il = 0, iu = npz;
    while (iu - il > 1) { const int im = (iu + il) / 2; if (z >= a.zlevels[im - 1]) il = im; else iu = im; }
    const int iz = il > 1 ? il : 1;
    double w = (double)(z - a.zlevels[iz - 1]) / (a.zlevels[iz] - a.zlevels[iz - 1]);
    w = fmax(fmin(w, 1.0), 0.0);
    int ix = (int)((x - a.xstart) / a.delx) + 1;
    if (fabsf(x - a.xstart - npx * a.delx) < 0.01f * a.delx) ix = npx;
    int iy = (int)((y - a.ystart) / a.dely) + 1;
    if (fabsf(y - a.ystart - npy * a.dely) < 0.01f * a.dely) iy = npy;
    if (ix < 1 || ix > npx) { atomicCAS(a.bad, 0, 1); return; }
    if (iy < 1 || iy > npy) { atomicCAS(a.bad, 0, 2); return; }
    const int ixp = ix % npx + 1, iyp = iy % npy + 1;
    double u = (double)(x - a.xstart - a.delx * (ix - 1)) / a.delx;
    u = fmax(fmin(u, 1.0), 0.0); if (u < 1.0e-5) u = 0.0; if (u > 1.0 - 1.0e-5) u = 1.0;
    double v = (double)(y - a.ystart - a.dely * (iy - 1)) / a.dely;
    v = fmax(fmin(v, 1.0), 0.0); if (v < 1.0e-5) v = 0.0; if (v > 1.0 - 1.0e-5) v = 1.0;
    int ptr[8];
    ptr[0] = iz + npz * (iy - 1) + npz * npy * (ix - 1);
    ptr[1] = iz + npz * (iy - 1) + npz * npy * (ixp - 1);
    ptr[2] = iz + npz * (iyp - 1) + npz * npy * (ix - 1);
    ptr[3] = iz + npz * (iyp - 1) + npz * npy * (ixp - 1);
    for (int c = 0; c < 4; c++) ptr[4 + c] = ptr[c] + 1;
    const double f[8] = {(1 - u) * (1 - v) * (1 - w), u * (1 - v) * (1 - w), (1 - u) * v * (1 - w), u * v * (1 - w),
                         (1 - u) * (1 - v) * w, u * (1 - v) * w, (1 - u) * v * w, u * v * w};
    const size_t maxpg = (size_t)npx * npy * npz;
    float tempv = 0.0f;
    if (a.tempp) {
        const float *tp = a.tempp;
        tempv = (float)(f[0] * tp[ptr[0] - 1] + f[1] * tp[ptr[1] - 1] + f[2] * tp[ptr[2] - 1] + f[3] * tp[ptr[3] - 1]
                        + f[4] * tp[ptr[4] - 1] + f[5] * tp[ptr[5] - 1] + f[6] * tp[ptr[6] - 1] + f[7] * tp[ptr[7] - 1]);
    }
    if (a.temp) a.temp[i] = tempv;
    float kg = 0.0f;
    if (a.nzckd > 0) {
        int jl = 1, ju = a.nzckd;
        while (ju - jl > 1) { const int jm = (ju + jl) / 2; if (z <= a.zckd[jm - 1]) jl = jm; else ju = jm; }
        int k = jl > 1 ? jl : 1;
        if (k > a.nzckd - 1) k = a.nzckd - 1;
        double ff = (z - a.zckd[k - 1]) / (a.zckd[k] - a.zckd[k - 1]);
        ff = fmin(fmax(ff, 0.0), 1.0);
        kg = (float)((1.0f - ff) * a.gasabs[k - 1] + ff * a.gasabs[k]);
    }
    float total = 0.0f, total_unscaled = 0.0f, sum_unscaled = 0.0f;
    for (int ipa = 0; ipa < a.npart; ipa++) {
        const float *extp = a.extinctp + maxpg * ipa, *albp = a.albedop + maxpg * ipa;
        const int *iphp = a.iphasep + (size_t)mnm * maxpg * ipa;
        const float *pwp = a.phasewtp + (size_t)mnm * maxpg * ipa;
        double scat[8], scatter;
        float extf = (float)(f[0] * extp[ptr[0] - 1] + f[1] * extp[ptr[1] - 1] + f[2] * extp[ptr[2] - 1] + f[3] * extp[ptr[3] - 1]
                             + f[4] * extp[ptr[4] - 1] + f[5] * extp[ptr[5] - 1] + f[6] * extp[ptr[6] - 1] + f[7] * extp[ptr[7] - 1]);
        for (int c = 0; c < 8; c++) scat[c] = f[c] * extp[ptr[c] - 1] * albp[ptr[c] - 1];
        scatter = scat[0] + scat[1] + scat[2] + scat[3] + scat[4] + scat[5] + scat[6] + scat[7];
        float albf = extf > a.extmin ? (float)(scatter / extf) : (float)(scatter / a.extmin);
        int ip[TPA_MAXQ];
        float pw[TPA_MAXQ];
        for (int c = 0; c < 8; c++)
            for (int m = 0; m < mnm; m++) ip[c * mnm + m] = iphp[m + (size_t)mnm * (ptr[c] - 1)];
        if (a.interp_new) {
            const double den = scatter >= a.scatmin ? scatter : a.scatmin;
            for (int c = 0; c < 8; c++)
                for (int m = 0; m < mnm; m++) pw[c * mnm + m] = (float)(pwp[m + (size_t)mnm * (ptr[c] - 1)] * scat[c] / den);
            for (int q = 0; q < nq; q++)
                for (int q2 = q + 1; q2 < nq; q2++)
                    if (ip[q] == ip[q2]) { pw[q] = pw[q] + pw[q2]; pw[q2] = 0.0f; }
            tpa_ssort_desc(pw, ip, nq);
        } else {
            double maxscat = -1.0f;
            for (int q = 0; q < nq; q++) pw[q] = 0.0f;
            pw[0] = 1.0f;
            for (int c = 0; c < 8; c++)
                if (scat[c] > maxscat || fabs(f[c] - 1) < 0.001f) {
                    // SSORT(PHASEWTP(:,I), IPHASEP(:,I), MAXNMICRO, -2): the dominant entry of the property point
                    int best = 0;
                    if (mnm > 1) {
                        float xs[TPA_MAXQ / 8]; int ys[TPA_MAXQ / 8];
                        for (int m = 0; m < mnm; m++) { xs[m] = pwp[m + (size_t)mnm * (ptr[c] - 1)]; ys[m] = iphp[m + (size_t)mnm * (ptr[c] - 1)]; }
                        tpa_ssort_desc(xs, ys, mnm);
                        best = ys[0];
                    } else best = iphp[(size_t)mnm * (ptr[c] - 1)];
                    maxscat = scat[c];
                    ip[0] = best;
                }
        }
        const size_t po = (size_t)i + (size_t)a.ld * ipa;
        if (ipa == 0) total_unscaled = total_unscaled + kg;
        total_unscaled = total_unscaled + extf;
        sum_unscaled = sum_unscaled + extf;
        for (int q = 0; q < nq; q++) {
            a.iphase[q + (size_t)nq * po] = ip[q];
            a.phaseinterpwt[q + (size_t)nq * po] = pw[q];
        }
        if (a.deltam) {
            float fd;
            if (!a.interp_new || pw[0] >= a.phasemax) fd = a.ftab[ip[0] - 1];
            else { fd = 0.0f; for (int q = 0; q < nq; q++) fd = fd + a.ftab[ip[q] - 1] * pw[q]; }
            const float a0 = albf;
            extf = (1.0f - a0 * fd) * extf;
            albf = (1.0f - fd) * a0 / (1.0f - a0 * fd);
        }
        a.extinct[po] = extf;
        a.albedo[po] = albf;
        if (ipa == 0) total = total + kg;
        total = total + extf;
        if (a.planck && a.srctype != 'S') a.planck[po] = (1.0f - albf) * dev_planck(tempv, a.units, a.wavelen);
    }
    if (a.prepare_prop && a.deltam) {
        // PREPARE_PROP (shdomsub2.f:545-571): TOTAL_EXT - SUM(EXTINCT) of the unscaled values (clipped at 0; this is the gas
        // absorption up to rounding), then the scaled extinctions are added
        float t = total_unscaled - sum_unscaled;
        if (t < 0.0f) t = 0.0f;
        for (int ipa = 0; ipa < a.npart; ipa++) t = t + a.extinct[(size_t)i + (size_t)a.ld * ipa];
        total = t;
    }
    a.total_ext[i] = total;
}

cudaError_t launch_tpa(const TpaArgs &a, cudaStream_t s)
{
    if (a.count < 1) return cudaSuccess;
    tpa_kernel<<<(a.count + 127) / 128, 128, 0, s>>>(a);
    return cudaGetLastError();
}

// EXTMIN, SCATMIN of TRILIN_INTERP_PROP(INIT=.TRUE.) (shdom90.f90:73-75): REAL arithmetic, then DOUBLE PRECISION
void tpa_extmin(const float *zlevels, int npz, double *extmin, double *scatmin)
{
    const float e = 1.0e-5f / ((zlevels[npz - 1] - zlevels[0]) / npz);
    *extmin = (double)e;
    *scatmin = 0.1f * *extmin;
}

extern "C" int at3d_transfer_pa_to_grid(int npts, const float *gridpos, int npx, int npy, int npz, float delx, float dely,
                                        float xstart, float ystart, const float *zlevels, int npart, int maxnmicro,
                                        const float *extinctp, const float *albedop, const int32_t *iphasep,
                                        const float *phasewtp, int numphase, const float *ftab, int ml, int deltam,
                                        float phasemax, float *extinct, float *albedo, float *total_ext, int32_t *iphase,
                                        float *phaseinterpwt, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!gridpos || !zlevels || !extinctp || !albedop || !iphasep || !phasewtp || !extinct || !albedo || !total_ext || !iphase ||
        !phaseinterpwt || (deltam && !ftab)) { set_msg(errmsg, "null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if (8 * maxnmicro > TPA_MAXQ) { set_msg(errmsg, "at3d_transfer_pa_to_grid: MAXNMICRO > %d", TPA_MAXQ / 8); return 3; }
    if (npz < 2 || npts < 1) { set_msg(errmsg, "at3d_transfer_pa_to_grid: bad sizes"); return 1; }
    const size_t maxpg = (size_t)npx * npy * npz, nq = 8 * (size_t)maxnmicro;
    Arena A;
    TpaArgs a;
    memset(&a, 0, sizeof(a));
    a.first = 0; a.count = npts; a.ld = npts; a.interp_new = 1; a.srctype = 'S'; a.prepare_prop = 1;
    a.npart = npart; a.mnm = maxnmicro; a.npx = npx; a.npy = npy; a.npz = npz; a.ml = ml; a.deltam = deltam;
    a.delx = delx; a.dely = dely; a.xstart = xstart; a.ystart = ystart; a.phasemax = phasemax;
    tpa_extmin(zlevels, npz, &a.extmin, &a.scatmin);
    a.gridpos = A.up(gridpos, (size_t)3 * npts); a.zlevels = A.up(zlevels, npz);
    a.extinctp = A.up(extinctp, maxpg * npart); a.albedop = A.up(albedop, maxpg * npart);
    a.iphasep = A.up(iphasep, (size_t)maxnmicro * maxpg * npart);
    a.phasewtp = A.up(phasewtp, (size_t)maxnmicro * maxpg * npart);
    a.ftab = deltam ? A.up(ftab, numphase) : nullptr;
    a.extinct = A.alloc<float>((size_t)npts * npart); a.albedo = A.alloc<float>((size_t)npts * npart);
    a.total_ext = A.alloc<float>(npts);
    a.iphase = A.alloc<int>(nq * npts * npart); a.phaseinterpwt = A.alloc<float>(nq * npts * npart);
    a.bad = A.alloc<int>(1);
    if (!a.gridpos || !a.zlevels || !a.extinctp || !a.albedop || !a.iphasep || !a.phasewtp || (deltam && !a.ftab) || !a.extinct ||
        !a.albedo || !a.total_ext || !a.iphase || !a.phaseinterpwt || !a.bad) { set_msg(errmsg, "device allocation failure"); return 4; }
    cudaMemset(a.bad, 0, sizeof(int));
    launch_tpa(a, 0);
    int bad = 0;
    cudaError_t e = cudaMemcpy(&bad, a.bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(extinct, a.extinct, (size_t)npts * npart * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(albedo, a.albedo, (size_t)npts * npart * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(total_ext, a.total_ext, (size_t)npts * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(iphase, a.iphase, nq * npts * npart * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(phaseinterpwt, a.phaseinterpwt, nq * npts * npart * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_transfer_pa_to_grid", cudaGetErrorString(e)); return 4; }
    if (bad) { set_msg(errmsg, bad == 1 ? "TRILIN: Beyond X domain" : "TRILIN: Beyond Y domain"); return 1; }
    return 0;
}
