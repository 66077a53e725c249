// at3d_solver.cu -- PATH_INTEGRATION for independent-pixel grids (IPFLAG=3) on the device (sm_100a).
// Replaces, for the plane-parallel-column case, PATH_INTEGRATION (src/polarized/shdomsub1.f:1836-2167 of the AT3D
// reference) with BACK_INT_GRID1D (:4295-4468), the top/bottom boundary conditions it applies
// (COMPUTE_TOP_RADIANCES :2336, FIXED/VARIABLE_LAMBERTIAN_BOUNDARY :2438-2529, VARIABLE_BRDF_SURFACE :2597-2669) and
// the SH <-> discrete-ordinate transforms (at3d_transform.cu).  SURVEY.md 8f rank 1, first step: the 3-D sweep
// (BACK_INT_GRID3D) needs the wavefront ordering of SWEEPING_ORDER and is not here yet.
//
// Layout: DOFIELD(NPTS, NSTOKES, NANG) holds the discrete-ordinate source function after SH_TO_DO and is overwritten
// in place by the radiance; a thread integrates one (column, ordinate) from its boundary through the NZ levels with
// the arithmetic of BACK_INT_GRID1D (DOUBLE PRECISION), carrying the previous level's extinction and source in
// registers.  The reference's loop over ordinates only carries a dependence through the surface, so all downward
// ordinates run in one launch, then the surface, then all upward ordinates.
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <vector>
#include "at3d_host.h"
#include "at3d_surface.cuh"

static void set_msg(char *errmsg, const char *fmt, ...)
{
    if (!errmsg) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, AT3D_ERRMSG_LEN, fmt, ap);
    va_end(ap);
}

struct PiArgs {
    int npts, nst, nz, ncol, nang, nmu, nphi0max, ntop, nbot, nsfcpar;
    int srctype, units, sfctype0, sfctype1;
    float wavelen, solarmu, solaraz, gndalbedo, gndtemp;
    float *dofield;            // [npts, nst, nang]
    const float *total_ext, *zlev, *dirflux;
    const float *ang_mu, *ang_phi, *ang_w;     // [nang]: MU, PHI, ABS(MU)*WTDO per ordinate
    const int *ang_imu, *ang_iphi;             // [nang] 0-based
    const float *skyrad;       // [nst, nmu/2, nphi0max]
    const float *sfcgridparms; // [nsfcpar, nbot]
    const float *sfcgridrad;   // [nang/2+1, nbot] or null
    float *bcrad;              // [nst, ntop + nbot*(1 or 1+nang/2)]
    float *botrad;             // [nst, nbot, nang/2] upwelling boundary radiance per upward ordinate (general BRDF)
    float *fluxes;             // [2, npts]
};

// one backward step of BACK_INT_GRID1D (shdomsub1.f:4404-4447): radiance at a point from the known radiance rad0 at the
// neighbouring level (extinction ext0, source*extinction srcext0) over the path length so
template <int NST>
__device__ __forceinline__ void pi_step(double so, double ext0, const double (&srcext0)[NST], double ext1,
                                        const double (&srcext1)[NST], const double (&rad0)[NST], double (&rad)[NST])
{
    const double ext = 0.5 * (ext0 + ext1);
    const double tau = ext * so;
    double transcell, abscell, src[NST];
    if (tau >= 0.5) { transcell = exp(-tau); abscell = 1.0 - transcell; }
    else { abscell = tau * (1.0 - 0.5 * tau * (1.0 - 0.33333333333 * tau * (1 - 0.25 * tau))); transcell = 1.0 - abscell; }
    if (tau <= 2.0) {
        if (ext == 0.0) {
#pragma unroll
            for (int k = 0; k < NST; k++) src[k] = 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < NST; k++)
                src[k] = (0.5 * (srcext0[k] + srcext1[k]) + 0.08333333333 * (ext0 * srcext1[k] - ext1 * srcext0[k]) * so) / ext;
        }
    } else {
        double ext0p = ext0, srcext0p[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) srcext0p[k] = srcext0[k];
        if (tau > 4.0) {
            ext0p = ext1 + (ext0 - ext1) * 4.0 / tau;
            if (ext0 > 0.0) {
#pragma unroll
                for (int k = 0; k < NST; k++) srcext0p[k] = srcext0[k] * ext0p / ext0;
            }
        }
#pragma unroll
        for (int k = 0; k < NST; k++)
            src[k] = 1.0 / (ext0p + ext1) * (srcext0p[k] + srcext1[k]
                     + (ext0p * srcext1[k] - ext1 * srcext0p[k]) * 2.0 / (ext0p + ext1) * (1 - 2 / tau + 2 * transcell / abscell));
    }
    src[0] = fmax(src[0], 0.0);
#pragma unroll
    for (int k = 0; k < NST; k++) rad[k] = 0.0 + 1.0 * (rad0[k] * transcell + src[k] * abscell);
}

// all ordinates of one hemisphere: thread = (column, ordinate); up = 0: downward ordinates [0, nang/2) from the top
// boundary, up = 1: upward ordinates [nang/2, nang) from the bottom boundary.  Points of a column are contiguous in z.
template <int NST>
__global__ void pi_sweep_kernel(PiArgs a, int up)
{
    const int nh = a.nang / 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.ncol * nh) return;
    const int col = t % a.ncol, ia = (up ? nh : 0) + t / a.ncol;
    const float mu = a.ang_mu[ia];
    const double cz = -mu, czinv = 1.0 / cz;
    const int nz = a.nz;
    float *f = a.dofield + (size_t)a.npts * NST * ia;      // plane k of this ordinate at f[p + npts*k]
    double radn[NST], srcextn[NST], extn;
    // boundary point
    const int pb = (up ? 0 : nz - 1) + nz * col;
    if (!up) {
        // COMPUTE_TOP_RADIANCES, INTERPOLATE_FLAG=-1: SKYRAD(:,IMU,IPHI), Planck function of it for SRCTYPE='T'
        const int imu = a.ang_imu[ia], iphi = a.ang_iphi[ia];
#pragma unroll
        for (int k = 0; k < NST; k++) radn[k] = (double)a.skyrad[k + NST * (imu + (a.nmu / 2) * iphi)];
        if (a.srctype == 'T') {
            radn[0] = (double)dev_planck((float)radn[0], a.units, a.wavelen);
#pragma unroll
            for (int k = 1; k < NST; k++) radn[k] = 0.0;
        }
        if (ia == 0) {
#pragma unroll
            for (int k = 0; k < NST; k++) a.bcrad[k + NST * col] = (float)radn[k];
        }
    } else {
        const int iu = ia - nh;
        const float sg = a.sfcgridrad ? a.sfcgridrad[(iu + 1) + (size_t)(nh + 1) * col] : 0.0f;
#pragma unroll
        for (int k = 0; k < NST; k++) {
            const float b = (a.sfctype1 == 'L') ? a.bcrad[k + NST * (a.ntop + col)]
                                                : a.botrad[k + NST * (col + (size_t)a.nbot * iu)];
            radn[k] = (double)(b + sg);
        }
    }
    float extf = a.total_ext[pb];
    extn = (double)extf;
#pragma unroll
    for (int k = 0; k < NST; k++) {
        float s = f[pb + (size_t)a.npts * k];
        if (k == 0) s = fmaxf(0.0f, s);          // PATH_INTEGRATION clamps the I source function (shdomsub1.f:2001-2005)
        srcextn[k] = (double)(extf * s);         // SRCEXT0 = EXTINCT(I1)*SOURCE(:,KANG,I1) is a REAL product (:4402)
        f[pb + (size_t)a.npts * k] = (float)radn[k];
    }
    double zn = (double)a.zlev[up ? 0 : nz - 1];
    for (int step = 1; step < nz; step++) {
        const int iz = up ? step : nz - 1 - step;
        const int p = iz + nz * col;
        extf = a.total_ext[p];
        const double ext1 = (double)extf;
        double srcext1[NST], srcnext[NST], rad[NST];
#pragma unroll
        for (int k = 0; k < NST; k++) {
            float s = f[p + (size_t)a.npts * k];
            if (k == 0) s = fmaxf(0.0f, s);
            srcext1[k] = ext1 * (double)s;       // SRCEXT1 = EXT1*SOURCE(:,KANG,IPT) with EXT1 DOUBLE (:4378)
            srcnext[k] = (double)(extf * s);     // what the next point sees as SRCEXT0
        }
        const double ze = (double)a.zlev[iz];
        const double so = (zn - ze) * czinv;
        pi_step<NST>(so, extn, srcextn, ext1, srcext1, radn, rad);
#pragma unroll
        for (int k = 0; k < NST; k++) {
            const float rf = (float)rad[k];          // GRIDRAD is REAL
            f[p + (size_t)a.npts * k] = rf;
            radn[k] = (double)rf; srcextn[k] = srcnext[k];
        }
        extn = ext1; zn = ze;
    }
    // store the downwelling radiance at the bottom point for the surface reflection (shdomsub1.f:2139-2145)
    if (!up && a.sfctype1 != 'L') {
#pragma unroll
        for (int k = 0; k < NST; k++)
            a.bcrad[k + NST * (a.ntop + col + (size_t)a.nbot * (ia + 1))] = (float)radn[k];
    }
}

// hemispheric fluxes, summed over the ordinates in the reference's order (shdomsub1.f:2131-2135): thread = point
template <int NST>
__global__ void pi_flux_kernel(PiArgs a, int up)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npts) return;
    const int nh = a.nang / 2;
    float s = 0.0f;
    for (int ia = up ? nh : 0; ia < (up ? a.nang : nh); ia++)
        s = s + a.ang_w[ia] * a.dofield[p + (size_t)a.npts * NST * ia];
    a.fluxes[up + 2 * (size_t)p] = s;
}

// FIXED / VARIABLE_LAMBERTIAN_BOUNDARY (shdomsub1.f:2438-2529)
__global__ void pi_lambertian_kernel(PiArgs a)
{
    const int ibc = blockIdx.x * blockDim.x + threadIdx.x;
    if (ibc >= a.nbot) return;
    const int i = a.nz * ibc;                         // bottom point of column ibc (0-based)
    const float down = a.fluxes[2 * (size_t)i];
    float v = 0.0f;
    if (a.sfctype0 == 'F') {
        const float alb = a.gndalbedo / acosf(-1.0f);
        float gndrad = 0.0f;
        if (a.srctype == 'T' || a.srctype == 'B') { gndrad = dev_planck(a.gndtemp, a.units, a.wavelen); gndrad = gndrad * (1.0f - a.gndalbedo); }
        if (a.srctype == 'S') v = alb * (a.dirflux[i] + down);
        else if (a.srctype == 'T') v = gndrad + alb * down;
        else v = alb * (a.dirflux[i] + down) + gndrad;
    } else {
        const float opi = 1.0f / acosf(-1.0f);
        const float alb = a.sfcgridparms[1 + a.nsfcpar * ibc];
        if (a.srctype == 'S') v = opi * alb * (a.dirflux[i] + down);
        else {
            const float gndrad = a.sfcgridparms[0 + a.nsfcpar * ibc] * (1 - alb);
            if (a.srctype == 'T') v = gndrad + opi * alb * down;
            else v = opi * alb * (a.dirflux[i] + down) + gndrad;
        }
    }
    a.bcrad[a.nst * (size_t)(a.ntop + ibc)] = v;
    for (int k = 1; k < a.nst; k++) a.bcrad[k + a.nst * (size_t)(a.ntop + ibc)] = 0.0f;
}

// VARIABLE_BRDF_SURFACE for every (bottom point, upward ordinate) (shdomsub1.f:2597-2669): thread = (ibc, ordinate),
// the incident ordinates are summed sequentially in the reference's order
template <int NST>
__global__ void pi_brdf_kernel(PiArgs a)
{
    const int nh = a.nang / 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nbot * nh) return;
    const int ibc = t % a.nbot, iu = t / a.nbot, ia = nh + iu;
    const float mu2 = a.ang_mu[ia], phi2 = a.ang_phi[ia];
    const float opi = 1.0f / acosf(-1.0f);
    const float *parms = a.sfcgridparms + (size_t)a.nsfcpar * ibc;
    float reflect[16], out[NST];
#pragma unroll
    for (int k = 0; k < NST; k++) out[k] = 0.0f;
    if (a.srctype != 'T') {
        dev_surface_brdf(a.sfctype1, parms + 1, a.wavelen, mu2, phi2, a.solarmu, a.solaraz, NST, reflect);
        const float df = a.dirflux[a.nz * ibc];
#pragma unroll
        for (int k = 0; k < NST; k++) out[k] = out[k] + opi * reflect[k] * df;
    }
    for (int ja = 0; ja < nh; ja++) {
        dev_surface_brdf(a.sfctype1, parms + 1, a.wavelen, mu2, phi2, a.ang_mu[ja], a.ang_phi[ja], NST, reflect);
        const float w = opi * a.ang_w[ja];               // OPI*ABS(MU)*WTDO
        const float *down = a.bcrad + (size_t)NST * (a.ntop + ibc + (size_t)a.nbot * (ja + 1));
#pragma unroll
        for (int k1 = 0; k1 < NST; k1++) {
#pragma unroll
            for (int k = 0; k < NST; k++) out[k] = out[k] + w * reflect[k + 4 * k1] * down[k1];
        }
        out[0] = out[0] + w * (1 - reflect[0]) * parms[0];
#pragma unroll
        for (int k = 1; k < NST; k++) out[k] = out[k] - w * reflect[k] * parms[0];
    }
#pragma unroll
    for (int k = 0; k < NST; k++) {
        a.botrad[k + NST * (ibc + (size_t)a.nbot * iu)] = out[k];
        if (iu == nh - 1) a.bcrad[k + NST * (a.ntop + ibc)] = out[k];      // what BCRAD(:,IBC,1) holds afterwards
    }
}

namespace {
struct Arena {
    std::vector<void *> ptrs;
    ~Arena() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> T *alloc(size_t n)
    {
        void *p = nullptr;
        if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T *)p;
    }
    template <typename T> T *up(const T *h, size_t n)
    {
        T *d = alloc<T>(n);
        if (d && n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return d;
    }
};
}

extern "C" int at3d_path_integration_ip(const at3d_state_desc *d, const float *wtmu, const int32_t *shptr,
                                        const float *source, const int32_t *rshptr, float *radiance, float *fluxes,
                                        float *bcrad, double *kernel_ms, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !wtmu || !shptr || !source || !rshptr || !radiance || !fluxes || !bcrad) { set_msg(errmsg, "null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if ((d->ipflag & 3) != 3) { set_msg(errmsg, "at3d_path_integration_ip: only IPFLAG=3 (independent columns, BACK_INT_GRID1D) is implemented"); return 3; }
    const int nz = d->nz, npts = d->npts, nst = d->nstokes;
    if (npts % nz != 0) { set_msg(errmsg, "at3d_path_integration_ip: the grid must be the unsplit base grid"); return 3; }
    const int ncol = npts / nz;
    // base-grid layout checks: columns contiguous in z, boundary lists in column order
    for (int c = 0; c < ncol; c++) {
        for (int iz = 0; iz < nz; iz++)
            if (d->gridpos[2 + 3 * (size_t)(iz + nz * c)] != d->zgrid[iz]) { set_msg(errmsg, "at3d_path_integration_ip: the grid must be the unsplit base grid"); return 3; }
        if (d->bcptr[c] != nz * c + nz || d->bcptr[d->maxnbc + c] != nz * c + 1) { set_msg(errmsg, "at3d_path_integration_ip: unexpected boundary point lists"); return 3; }
    }
    if (d->ntoppts != ncol || d->nbotpts != ncol) { set_msg(errmsg, "at3d_path_integration_ip: unexpected boundary point counts"); return 3; }
    if (d->srctype != 'S' && d->units == 'B') { set_msg(errmsg, "UNITS='B' is not implemented"); return 3; }
    TrPlan *P = nullptr;
    int rc = tr_plan_create(nst, d->nstleg, d->ml, d->mm, d->nlm, d->nmu, d->nphi0max, d->nphi0, d->mu, d->phi, wtmu, &P, errmsg);
    if (rc) return rc;
    const int nang = tr_plan_nang(P), nh = nang / 2;
    const bool lamb = d->sfctype1 == 'L';
    std::vector<float> amu(nang), aphi(nang), aw(nang);
    std::vector<int> aimu(nang), aiphi(nang);
    for (int i = 0, ia = 0; i < d->nmu; i++)
        for (int k = 0; k < d->nphi0[i]; k++, ia++) {
            amu[ia] = d->mu[i]; aphi[ia] = d->phi[i + (size_t)d->nmu * k];
            aw[ia] = fabsf(d->mu[i]) * d->wtdo[i + (size_t)d->nmu * k];
            aimu[ia] = i; aiphi[ia] = k;
        }
    Arena A;
    PiArgs a;
    memset(&a, 0, sizeof(a));
    a.npts = npts; a.nst = nst; a.nz = nz; a.ncol = ncol; a.nang = nang; a.nmu = d->nmu; a.nphi0max = d->nphi0max;
    a.ntop = ncol; a.nbot = ncol; a.nsfcpar = d->nsfcpar; a.srctype = d->srctype; a.units = d->units;
    a.sfctype0 = d->sfctype0; a.sfctype1 = d->sfctype1; a.wavelen = d->wavelen; a.solarmu = d->solarmu;
    a.solaraz = d->solaraz; a.gndalbedo = d->gndalbedo; a.gndtemp = d->gndtemp;
    const size_t nbc = (size_t)nst * (ncol + (size_t)ncol * (lamb ? 1 : 1 + nh));
    const size_t nsh = (size_t)nst * shptr[npts], nrad = (size_t)nst * rshptr[npts];
    const int *shptr_d = A.up(shptr, (size_t)npts + 1), *rshptr_d = A.up(rshptr, (size_t)npts + 1);
    const float *src_d = A.up(source, nsh);
    float *rad_d = A.alloc<float>(nrad);
    a.dofield = A.alloc<float>((size_t)npts * nst * nang);
    a.total_ext = A.up(d->total_ext, npts); a.zlev = A.up(d->zgrid, nz); a.dirflux = A.up(d->dirflux, npts);
    a.ang_mu = A.up(amu.data(), nang); a.ang_phi = A.up(aphi.data(), nang); a.ang_w = A.up(aw.data(), nang);
    a.ang_imu = A.up(aimu.data(), nang); a.ang_iphi = A.up(aiphi.data(), nang);
    a.skyrad = A.up(d->skyrad, (size_t)nst * (d->nmu / 2) * d->nphi0max);
    a.sfcgridparms = A.up(d->sfcgridparms, (size_t)d->nsfcpar * ncol);
    if (d->sfcgridrad) {
        bool nonzero = false;
        for (size_t i = 0; i < (size_t)(nh + 1) * ncol && !nonzero; i++) nonzero = d->sfcgridrad[i] != 0.0f;
        if (nonzero) a.sfcgridrad = A.up(d->sfcgridrad, (size_t)(nh + 1) * ncol);
    }
    a.bcrad = A.up(bcrad, nbc);
    a.botrad = A.alloc<float>((size_t)nst * ncol * nh);
    a.fluxes = A.alloc<float>((size_t)2 * npts);
    if (!shptr_d || !rshptr_d || !src_d || !rad_d || !a.dofield || !a.total_ext || !a.zlev || !a.dirflux || !a.ang_mu ||
        !a.ang_phi || !a.ang_w || !a.ang_imu || !a.ang_iphi || !a.skyrad || !a.sfcgridparms || !a.bcrad || !a.botrad || !a.fluxes) {
        tr_plan_destroy(P); set_msg(errmsg, "device allocation failure"); return 4;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    cudaError_t e = tr_sh_to_do(P, npts, shptr_d, src_d, a.dofield, 0);
    const int nsw = (ncol * nh + 127) / 128, npb = (npts + 255) / 256;
    if (nst == 1) {
        pi_sweep_kernel<1><<<nsw, 128>>>(a, 0);
        pi_flux_kernel<1><<<npb, 256>>>(a, 0);
        if (lamb) pi_lambertian_kernel<<<(ncol + 127) / 128, 128>>>(a);
        else pi_brdf_kernel<1><<<nsw, 128>>>(a);
        pi_sweep_kernel<1><<<nsw, 128>>>(a, 1);
        pi_flux_kernel<1><<<npb, 256>>>(a, 1);
    } else {
        pi_sweep_kernel<3><<<nsw, 128>>>(a, 0);
        pi_flux_kernel<3><<<npb, 256>>>(a, 0);
        if (lamb) pi_lambertian_kernel<<<(ncol + 127) / 128, 128>>>(a);
        else pi_brdf_kernel<3><<<nsw, 128>>>(a);
        pi_sweep_kernel<3><<<nsw, 128>>>(a, 1);
        pi_flux_kernel<3><<<npb, 256>>>(a, 1);
    }
    if (e == cudaSuccess) e = tr_do_to_sh(P, npts, rshptr_d, a.dofield, rad_d, 0);
    cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.0f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e == cudaSuccess) e = cudaMemcpy(radiance, rad_d, nrad * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(fluxes, a.fluxes, (size_t)2 * npts * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(bcrad, a.bcrad, nbc * sizeof(float), cudaMemcpyDeviceToHost);
    tr_plan_destroy(P);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_path_integration_ip", cudaGetErrorString(e)); return 4; }
    if (kernel_ms) *kernel_ms = ms;
    return 0;
}
