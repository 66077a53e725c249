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
#include "at3d_mem.h"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <ctime>
#include <vector>
#include <cub/cub.cuh>
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
    const int *botpt;          // [nbot] 0-based grid point of bottom boundary point ibc (null: column layout nz*ibc)
    const float *fluxsrc;      // field the hemispheric fluxes are summed from (null: dofield)
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
    const float *f = a.fluxsrc ? a.fluxsrc : a.dofield;
    float s = 0.0f;
    for (int ia = up ? nh : 0; ia < (up ? a.nang : nh); ia++)
        s = s + a.ang_w[ia] * f[p + (size_t)a.npts * NST * ia];
    a.fluxes[up + 2 * (size_t)p] = s;
}

// FIXED / VARIABLE_LAMBERTIAN_BOUNDARY (shdomsub1.f:2438-2529)
__global__ void pi_lambertian_kernel(PiArgs a)
{
    const int ibc = blockIdx.x * blockDim.x + threadIdx.x;
    if (ibc >= a.nbot) return;
    const int i = a.botpt ? a.botpt[ibc] : a.nz * ibc;   // bottom boundary point ibc (0-based)
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
        const float df = a.dirflux[a.botpt ? a.botpt[ibc] : a.nz * ibc];
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
    ~Arena() { for (void *p : ptrs) at3d_free(p); }
    template <typename T> T *alloc(size_t n)
    {
        void *p = nullptr;
        if (at3d_malloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
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

// =====================================================================================================================
// 3-D grids: PATH_INTEGRATION with BACK_INT_GRID3D[_UNPOL] (src/polarized/shdomsub1.f:3354-4036) on a fixed grid.
//
// The reference visits the grid points of an ordinate serially in the order SWEEPING_ORDER built for the ordinate's
// octant (:3261-3352); a point traces a ray backwards until it reaches a cell face whose four points already have a
// radiance, i.e. were visited EARLIER in that order (or are boundary points of the hemisphere), and interpolates there.
// Which face that is depends on the geometry and on the order only, never on the values, so the serial loop becomes a
// data-flow kernel with the same results: thread = (ordinate, position r in the sweep order); a face point is "valid"
// iff it is a preset boundary point or its position is < r (static test), and the thread then waits for the four
// radiances it needs (GRIDRAD(1,.) = -1 marks "not yet", as in the reference).  Blocks take their (ordinate, chunk of
// positions) from a ticket counter in chunk-major order, so everything a thread waits for belongs to a block that is
// already running or done: the waits terminate (and are bounded anyway: a timeout raises an error instead of hanging).
// All ordinates of a hemisphere run in ONE launch; the reference's loop over ordinates carries a dependence only
// through the surface.
// =====================================================================================================================
struct OrdDir {
    double cx, cy, cz, cxinv, cyinv, czinv;
    int bitx, bity, bitz, ioct, joct, pad;
};

struct SwArgs {
    DevState S;                 // cellrec / ptrec only
    int npts, nang, nchunks, group;
    const OrdDir *dir;          // [nang]
    const int *sweepord;        // [noct][npts]  cell<<3 | corner-1 (SWEEPORD)
    const int *rank;            // [noct][npts]  position of a point in the octant's order; -1 for the boundary points that
                                // are preset in the octant's hemisphere (top for downward, bottom for upward ordinates)
    const unsigned char *bflag; // [npts] bit 0: top boundary point, bit 1: bottom boundary point
    const float *srcdo;         // [npts, nst, nang] discrete-ordinate source function
    float *radf;                // [npts, nst, nang] radiance (GRIDRAD of every ordinate)
    double eps, transmin;
    int *ticket, *err;
    const int2 *plan;           // [nang][npts] processing order of an ordinate, sorted by dependency level (the preset boundary
                                // points, level 0, come first and are skipped): x = SWEEPORD entry, y = point-1 | cells to walk
                                // << 24 (255: decide with the rank test); null: the reference order itself
    int two_d;                  // BACK_INT_GRID2D (IPFLAG bit 1): rays stay in the X-Z plane, faces of two points
    int npre[2];                // preset boundary points per hemisphere (down: top points, up: bottom points)
    int *level;                 // [nang][npts] dependency level of (ordinate, point); -1 = not yet (level pass only)
    unsigned char *nwalk;       // [nang][npts] cells walked by (ordinate, point), capped at 255 (level pass only)
};

__device__ __forceinline__ float ld_vol(const float *p) { return *(const volatile float *)p; }
// GRIDFACE(n+1,kface) of BACK_INT_GRID3D: the four corners (1..8) of cell face kface (1..6: -x,+x,-y,+y,-z,+z)
__device__ __forceinline__ int face_corner(int kface, int n)
{
    const int axis = (kface - 1) >> 1, s = (kface - 1) & 1;
    if (axis == 0) return 1 + (s | ((n & 1) << 1) | ((n >> 1) << 2));
    if (axis == 1) return 1 + ((n & 1) | (s << 1) | ((n >> 1) << 2));
    return 1 + (n | (s << 2));
}

// LEVEL = true: the same walk and waits, but what travels is the dependency level 1 + max(level of the four face
// points) instead of the radiance (run once per solver object; the levels order the threads of the real sweep so that
// a thread's face points were finished a whole wavefront earlier and neighbouring lanes never wait for each other)
#ifndef AT3D_SWEEP_SLEEP
#define AT3D_SWEEP_SLEEP 64          // ns between two polls of a waiting thread
#endif
#ifndef AT3D_SWEEP_MINB
#define AT3D_SWEEP_MINB 3
#endif
// A waiting thread gives up after AT3D_SWEEP_WAIT_NS of WALL CLOCK (%globaltimer), not after a number of polls: under
// time slicing, MPS, compute-sanitizer or a co-tenant kernel a legitimate wait can be long.  The waits rely on the
// independent thread scheduling of sm_70+ (a lane may wait for another lane of its own warp at a level boundary).
#ifndef AT3D_SWEEP_WAIT_NS
#define AT3D_SWEEP_WAIT_NS 5000000000ull
#endif
__device__ __forceinline__ unsigned long long sweep_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// true when the wait has to be abandoned: another thread reported an error, or the wall-clock budget is spent
__device__ __forceinline__ bool sweep_give_up(int &spins, unsigned long long &t0, const int *err)
{
    if ((++spins & 255) != 0) return false;
    if (*(volatile const int *)err) return true;
    const unsigned long long now = sweep_now();
    if (t0 == 0) { t0 = now; return false; }
    return now - t0 > AT3D_SWEEP_WAIT_NS;
}
template <int NST, bool LEVEL>
__global__ void __launch_bounds__(256, (NST == 1 ? AT3D_SWEEP_MINB : 2)) sweep3d_kernel(SwArgs a, int up)
{
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
    __syncthreads();
    const bool planned = !LEVEL && a.plan != nullptr;
    const int npre = planned ? a.npre[up] : 0, nact = a.npts - npre, nchunks = (nact + 255) / 256;
    const int nh = a.nang / 2, G = a.group, per = nchunks * G;
    int t = s_ticket, grp = t / per, io, chunk;
    const int ngrp = (nh + G - 1) / G;
    if (grp >= ngrp - 1) {
        grp = ngrp - 1;
        const int gl = nh - grp * G;
        t -= grp * per;
        chunk = t / gl; io = grp * G + t % gl;
    } else {
        t -= grp * per;
        chunk = t / G; io = grp * G + t % G;
    }
    const int kpos = chunk * 256 + threadIdx.x;
    if (kpos >= nact) return;
    const int ia = (up ? nh : 0) + io;
    const OrdDir D = a.dir[ia];
    const int *rank = a.rank + (size_t)a.npts * (D.joct - 1);
    const DevState &S = a.S;
    int r, ipcell, ipt, nwalk = 255;
    if (planned) {
        const int2 pl = __ldg(&a.plan[(size_t)a.npts * ia + npre + kpos]);
        ipcell = pl.x >> 3; ipt = (pl.y & 0xFFFFFF) + 1; nwalk = (int)((unsigned)pl.y >> 24);
        r = nwalk == 255 ? __ldg(&rank[ipt - 1]) : 0;
    } else {
        r = kpos;
        const int entry = a.sweepord[(size_t)a.npts * (D.joct - 1) + r];
        ipcell = entry >> 3;
        ipt = cell_gp(S, ipcell, (entry & 7) + 1);
        if (a.bflag[ipt - 1] & (up ? 2 : 1)) return;          // GRIDRAD(1,IPT) >= 0: preset boundary point
    }
    const size_t fo = (size_t)a.npts * NST * ia;
    const float *src = a.srcdo + fo;
    float *R = a.radf + fo;
    const int ioct = D.ioct;
    int icell = ipcell, fail = 0, step = 0;
    double transmit = 1.0, rad[NST], srcext1[NST], srcext0[NST];
    float4 pp = __ldg(&S.ptrec[ipt - 1]);
    double ext1 = (double)pp.w, ext0 = 0.0, f1 = 0, f2 = 0, f3 = 0, f4 = 0;
    double xe = (double)pp.x, ye = (double)pp.y, ze = (double)pp.z;
    int i1 = 0, i2 = 0, i3 = 0, i4 = 0;
#pragma unroll
    for (int k = 0; k < NST; k++) {
        float s = __ldg(&src[(ipt - 1) + (size_t)a.npts * k]);
        if (k == 0) s = fmaxf(0.0f, s);                       // PATH_INTEGRATION clamps the I source (shdomsub1.f:2001-2005)
        rad[k] = 0.0; srcext1[k] = ext1 * (double)s;
    }
    for (;;) {
        if (icell <= 0) { fail = 1; break; }
        const CellRec c = load_cell(S, icell);
        const bool ipinx = c.flags & 1, ipiny = a.two_d || (c.flags & 2);
        const float4 po = __ldg(&S.ptrec[c.gp[8 - ioct] - 1]);     // GRIDPTR(9-IOCT,ICELL)
        const double sox = ipinx ? (double)1.0E20f : ((double)po.x - xe) * D.cxinv;
        const double soy = ipiny ? (double)1.0E20f : ((double)po.y - ye) * D.cyinv;
        const double soz = ((double)po.z - ze) * D.czinv;
        const double so = fmin(fmin(sox, soy), soz);
        if (so < -a.eps) { fail = 2; break; }
        xe = xe + so * D.cx;
        ye = ye + so * D.cy;
        ze = ze + so * D.cz;
        int iface, jface;
        if (sox <= soz && sox <= soy) { iface = 2 - D.bitx; jface = 1; }
        else if (soy <= soz) { iface = 4 - D.bity; jface = 2; }
        else { iface = 6 - D.bitz; jface = 3; }
        const int nb = c.nb[iface - 1];
        int inextcell = nb;
        if (nb < 0) inextcell = dev_next_cell(S, xe, ye, ze, iface, jface, nb);
        int kface;
        if (nb >= 0) {
            kface = iface;
            i1 = c.gp[face_corner(kface, 0) - 1]; i2 = c.gp[face_corner(kface, 1) - 1];
            i3 = c.gp[face_corner(kface, 2) - 1]; i4 = c.gp[face_corner(kface, 3) - 1];
        } else {
            kface = ((iface - 1) ^ 1) + 1;                     // OPPFACE
            i1 = cell_gp(S, inextcell, face_corner(kface, 0)); i2 = cell_gp(S, inextcell, face_corner(kface, 1));
            i3 = cell_gp(S, inextcell, face_corner(kface, 2)); i4 = cell_gp(S, inextcell, face_corner(kface, 3));
        }
        if (a.two_d) {
            // GRIDFACE(2,6) of BACK_INT_GRID2D: X faces {1,5},{2,6} = (I1,I3) here, Z faces {1,2},{5,6} = (I1,I2); the
            // other two weights are exact zeros below, so the four-point formulas reduce to the two-point ones
            if (jface == 1) { i2 = i1; i4 = i3; } else { i3 = i1; i4 = i2; }
        }
        const float4 p1 = __ldg(&S.ptrec[i1 - 1]), p2 = __ldg(&S.ptrec[i2 - 1]);
        const float4 p3 = __ldg(&S.ptrec[i3 - 1]), p4 = __ldg(&S.ptrec[i4 - 1]);
        double u, v;
        if (jface == 1) {
            u = (ze - (double)p1.z) / (double)(p3.z - p1.z);
            v = ipiny ? 0.5 : (ye - (double)p1.y) / (double)(p2.y - p1.y);
        } else if (jface == 2) {
            u = (ze - (double)p1.z) / (double)(p3.z - p1.z);
            v = ipinx ? 0.5 : (xe - (double)p1.x) / (double)(p2.x - p1.x);
        } else {
            u = ipiny ? 0.5 : (ye - (double)p1.y) / (double)(p3.y - p1.y);
            v = ipinx ? 0.5 : (xe - (double)p1.x) / (double)(p2.x - p1.x);
        }
        if (a.two_d) { if (jface == 1) v = 0.0; else u = 0.0; }     // F = (1-U, U) on the face's two points
        // (the exit point is only needed to go on into the next cell: skipped on the last cell of a recorded walk)
        if (inextcell > 0 && !(nwalk != 255 && step + 1 == nwalk)) {
            const int pn = cell_gp(S, inextcell, ioct);
            if (jface == 1) xe = (double)pt_coord(S, pn, 1);
            else if (jface == 2) ye = (double)pt_coord(S, pn, 2);
            else ze = (double)pt_coord(S, pn, 3);
        }
        f1 = (1 - u) * (1 - v);
        f2 = (1 - u) * v;
        f3 = u * (1 - v);
        f4 = u * v;
        ext0 = f1 * (double)p1.w + f2 * (double)p2.w + f3 * (double)p3.w + f4 * (double)p4.w;
#pragma unroll
        for (int k = 0; k < NST; k++) {
            const size_t ko = (size_t)a.npts * k;
            float s1 = __ldg(&src[(i1 - 1) + ko]), s2 = __ldg(&src[(i2 - 1) + ko]);
            float s3 = __ldg(&src[(i3 - 1) + ko]), s4 = __ldg(&src[(i4 - 1) + ko]);
            if (k == 0) { s1 = fmaxf(0.0f, s1); s2 = fmaxf(0.0f, s2); s3 = fmaxf(0.0f, s3); s4 = fmaxf(0.0f, s4); }
            srcext0[k] = f1 * (double)s1 * (double)p1.w + f2 * (double)s2 * (double)p2.w
                       + f3 * (double)s3 * (double)p3.w + f4 * (double)s4 * (double)p4.w;
        }
        const double ext = 0.5 * (ext0 + ext1);
        const double tau = ext * so;
        double transcell, abscell, sc[NST];
        if (tau >= 0.5) { transcell = exp(-tau); abscell = 1.0 - transcell; }
        else { abscell = tau * (1.0 - 0.5 * tau * (1.0 - 0.33333333333 * tau * (1 - 0.25 * tau))); transcell = 1.0 - abscell; }
        if (tau <= 2.0) {
            if (ext == 0.0) {
#pragma unroll
                for (int k = 0; k < NST; k++) sc[k] = 0.0;
            } else {
#pragma unroll
                for (int k = 0; k < NST; k++)
                    sc[k] = (0.5 * (srcext0[k] + srcext1[k]) + 0.08333333333 * (ext0 * srcext1[k] - ext1 * srcext0[k]) * so) / ext;
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
                sc[k] = 1.0 / (ext0p + ext1) * (srcext0p[k] + srcext1[k]
                        + (ext0p * srcext1[k] - ext1 * srcext0p[k]) * 2.0 / (ext0p + ext1) * (1 - 2 / tau + 2 * transcell / abscell));
        }
        sc[0] = fmax(sc[0], 0.0);
#pragma unroll
        for (int k = 0; k < NST; k++) rad[k] = rad[k] + transmit * sc[k] * abscell;
        transmit = transmit * transcell;
        step++;
        if (nwalk != 255) {
            // the level pass walked this (ordinate, point) with the test below and recorded where it stopped
            if (step == nwalk) break;
        } else {
            // VALIDFACE, statically: all four face points are preset or earlier in the sweep order
            // (preset boundary points of the hemisphere carry rank -1)
            const bool validface = __ldg(&rank[i1 - 1]) < r && __ldg(&rank[i2 - 1]) < r && __ldg(&rank[i3 - 1]) < r && __ldg(&rank[i4 - 1]) < r;
            if (inextcell <= 0 || (transmit <= a.transmin && validface)) {
                if (!validface) fail = 3;
                break;
            }
        }
        ext1 = ext0;
#pragma unroll
        for (int k = 0; k < NST; k++) srcext1[k] = srcext0[k];
        icell = inextcell;
    }
    if (LEVEL) {
        int *L = a.level + (size_t)a.npts * ia;
        int lv = 0;
        if (!fail) {
            int spins = 0;
            unsigned long long t0 = 0;
            for (;;) {
                const int l1 = *(volatile int *)&L[i1 - 1], l2 = *(volatile int *)&L[i2 - 1];
                const int l3 = *(volatile int *)&L[i3 - 1], l4 = *(volatile int *)&L[i4 - 1];
                if (l1 >= 0 && l2 >= 0 && l3 >= 0 && l4 >= 0) { lv = 1 + max(max(l1, l2), max(l3, l4)); break; }
                if (sweep_give_up(spins, t0, a.err)) { fail = 4; break; }
                __nanosleep(AT3D_SWEEP_SLEEP);
            }
        }
        if (fail) atomicCAS(a.err, 0, fail);
        a.nwalk[(size_t)a.npts * ia + (ipt - 1)] = (unsigned char)(step < 255 ? step : 255);
        *(volatile int *)&L[ipt - 1] = lv;
        return;
    }
    float out[NST];
    if (!fail) {
        // wait for the four face radiances (bounded; polls go to L2)
        float g1, g2, g3, g4;
        int spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
            g1 = ld_vol(&R[i1 - 1]); g2 = ld_vol(&R[i2 - 1]); g3 = ld_vol(&R[i3 - 1]); g4 = ld_vol(&R[i4 - 1]);
            if (g1 >= -0.1f && g2 >= -0.1f && g3 >= -0.1f && g4 >= -0.1f) break;
            if (sweep_give_up(spins, t0, a.err)) { fail = 4; break; }
            __nanosleep(AT3D_SWEEP_SLEEP);
        }
        if (!fail) {
            __threadfence();
            rad[0] = rad[0] + transmit * (f1 * (double)g1 + f2 * (double)g2 + f3 * (double)g3 + f4 * (double)g4);
#pragma unroll
            for (int k = 1; k < NST; k++) {
                const size_t ko = (size_t)a.npts * k;
                const double rad0 = f1 * (double)__ldcg(&R[(i1 - 1) + ko]) + f2 * (double)__ldcg(&R[(i2 - 1) + ko])
                                  + f3 * (double)__ldcg(&R[(i3 - 1) + ko]) + f4 * (double)__ldcg(&R[(i4 - 1) + ko]);
                rad[k] = rad[k] + transmit * rad0;
            }
        }
    }
    if (fail) {
        atomicCAS(a.err, 0, fail);
#pragma unroll
        for (int k = 0; k < NST; k++) out[k] = 0.0f;
    } else {
#pragma unroll
        for (int k = 0; k < NST; k++) out[k] = (float)rad[k];
        if (!(out[0] >= 0.0f)) { atomicCAS(a.err, 0, 5); out[0] = 0.0f; }     // RAD<0 / NaN: the reference aborts
    }
#pragma unroll
    for (int k = 1; k < NST; k++) __stcg(&R[(ipt - 1) + (size_t)a.npts * k], out[k]);
    if (NST > 1) __threadfence();
    *(volatile float *)&R[ipt - 1] = out[0];
}

// GRIDRAD(1,.) = -1 everywhere, then the boundary values of every ordinate (shdomsub1.f:1968-1972, 2024-2071):
// thread = (point, ordinate)
__global__ void sweep3d_level_init_kernel(int npts, int nang, const unsigned char *bflag, int *level)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)npts * nang) return;
    const int p = (int)(t % npts), ia = (int)(t / npts);
    level[t] = (bflag[p] & (ia >= nang / 2 ? 2 : 1)) ? 0 : -1;
}

// sort key of (ordinate, position r): ordinate | level | r
__global__ void sweep3d_key_kernel(SwArgs a, unsigned long long *keys)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.npts * a.nang) return;
    const int r = (int)(t % a.npts), ia = (int)(t / a.npts);
    const int entry = a.sweepord[(size_t)a.npts * (a.dir[ia].joct - 1) + r];
    const int ipt = cell_gp(a.S, entry >> 3, (entry & 7) + 1);
    const int lv = a.level[(size_t)a.npts * ia + (ipt - 1)];
    keys[t] = ((unsigned long long)ia << 48) | ((unsigned long long)(lv < 0 ? 0 : lv) << 24) | (unsigned long long)r;
}

__global__ void sweep3d_plan_kernel(SwArgs a, const unsigned long long *keys, int2 *plan)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.npts * a.nang) return;
    const int ia = (int)(t / a.npts), r = (int)(keys[t] & 0xFFFFFFull);
    const int entry = a.sweepord[(size_t)a.npts * (a.dir[ia].joct - 1) + r];
    const int ipt = cell_gp(a.S, entry >> 3, (entry & 7) + 1);
    const int nw = a.nwalk[(size_t)a.npts * ia + (ipt - 1)];
    plan[t] = make_int2(entry, (ipt - 1) | (nw << 24));
}

template <int NST>
__global__ void sweep3d_init_kernel(PiArgs a, float *radf, const int *toppt, int up)
{
    const int nh = a.nang / 2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.npts * nh) return;
    const int p = (int)(t % a.npts), ia = (up ? nh : 0) + (int)(t / a.npts);
    radf[p + (size_t)a.npts * NST * ia] = -1.0f;
}

template <int NST>
__global__ void sweep3d_boundary_kernel(PiArgs a, float *radf, const int *toppt, int up)
{
    const int nh = a.nang / 2, nb = up ? a.nbot : a.ntop;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * nh) return;
    const int ibc = t % nb, io = t / nb, ia = (up ? nh : 0) + io;
    float v[NST];
    int p;
    if (!up) {
        const int imu = a.ang_imu[ia], iphi = a.ang_iphi[ia];
#pragma unroll
        for (int k = 0; k < NST; k++) v[k] = a.skyrad[k + NST * (imu + (a.nmu / 2) * iphi)];
        if (a.srctype == 'T') {
            v[0] = dev_planck(v[0], a.units, a.wavelen);
#pragma unroll
            for (int k = 1; k < NST; k++) v[k] = 0.0f;
        }
        p = toppt[ibc];
        if (io == nh - 1) {
#pragma unroll
            for (int k = 0; k < NST; k++) a.bcrad[k + NST * (size_t)ibc] = v[k];
        }
    } else {
        const float sg = a.sfcgridrad ? a.sfcgridrad[(io + 1) + (size_t)(nh + 1) * ibc] : 0.0f;
#pragma unroll
        for (int k = 0; k < NST; k++) {
            const float b = (a.sfctype1 == 'L') ? a.bcrad[k + NST * (size_t)(a.ntop + ibc)]
                                                : a.botrad[k + NST * (ibc + (size_t)a.nbot * io)];
            v[k] = b + sg;
        }
        p = a.botpt[ibc];
    }
#pragma unroll
    for (int k = 0; k < NST; k++) radf[p + (size_t)a.npts * (k + (size_t)NST * ia)] = v[k];
}

// BCRAD(:,IBC+NTOPPTS+NBOTPTS*IANG) = downwelling radiance at the bottom points, per downward ordinate (:2139-2145)
template <int NST>
__global__ void sweep3d_store_down_kernel(PiArgs a, const float *radf)
{
    const int nh = a.nang / 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nbot * nh) return;
    const int ibc = t % a.nbot, ia = t / a.nbot, p = a.botpt[ibc];
#pragma unroll
    for (int k = 0; k < NST; k++)
        a.bcrad[k + NST * (a.ntop + ibc + (size_t)a.nbot * (ia + 1))] = radf[p + (size_t)a.npts * (k + (size_t)NST * ia)];
}

namespace {
// SWEEPING_ORDER (shdomsub1.f:3261-3352) with SWEEP_BASE_CELL / SWEEP_NEXT_CELL (:4529-4700): base cells in the
// octant's x-fastest order (open boundaries: the boundary column first), leaves of a base cell depth-first with the
// upstream child first, the 8 corners of a leaf upstream first; a point is listed at its first visit.
bool host_sweeping_order(const at3d_state_desc *d, int noct, std::vector<int> &sweepord, std::vector<int> &rank)
{
    static const int ioctorder[8] = {1, 5, 2, 6, 3, 7, 4, 8};
    const int npts = d->npts, nz = d->nz;
    int nxc = d->nx, nyc = d->ny;
    if (d->bcflag & 1) nxc = d->nx + 1;
    if (d->bcflag & 2) nyc = d->ny + 1;
    sweepord.assign((size_t)noct * npts, 0);
    rank.assign((size_t)noct * npts, 0);
    auto axis_seq = [](int n, bool positive, bool open) {
        int s, e, dstep;
        if (positive) { dstep = +1; if (open) { s = n; e = n - 1 > 1 ? n - 1 : 1; } else { s = 1; e = n; } }
        else { dstep = -1; if (open) { s = 1; e = 2 < n ? 2 : n; } else { s = n; e = 1; } }
        std::vector<int> v;
        for (int i = s;; i = (i + dstep + n - 1) % n + 1) { v.push_back(i); if (i == e) break; }
        return v;
    };
    int bad = 0;
    // the octants are independent: one host thread each
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (int joct = 1; joct <= noct; joct++) {
        const int ioct = ioctorder[joct - 1], ob = ioct - 1;
        std::vector<char> seen(npts, 0);
        std::vector<int> stack;
        const std::vector<int> xs = axis_seq(nxc, ob & 1, d->bcflag & 1), ys = axis_seq(nyc, ob & 2, d->bcflag & 2);
        int iorder = 0;
        bool ok = true;
        int *so = sweepord.data() + (size_t)npts * (joct - 1), *rk = rank.data() + (size_t)npts * (joct - 1);
        const int z0 = (ob & 4) ? 1 : nz - 1, z1 = (ob & 4) ? nz - 1 : 1, dz = (ob & 4) ? 1 : -1;
        for (int iz = z0; ok; iz += dz) {
            for (int iy : ys)
                for (int ix : xs) {
                    stack.clear();
                    stack.push_back(iz + (nz - 1) * (iy - 1) + (nz - 1) * nyc * (ix - 1));
                    while (!stack.empty()) {
                        const int ic = stack.back();
                        stack.pop_back();
                        const int child = d->treeptr[1 + 2 * (size_t)(ic - 1)];
                        if (child > 0) {
                            const int idir = ((d->cellflags[ic - 1] >> 2) & 3) - 1;
                            if ((ob >> idir) & 1) { stack.push_back(child + 1); stack.push_back(child); }
                            else { stack.push_back(child); stack.push_back(child + 1); }
                            continue;
                        }
                        for (int index = 0; index < 8; index++) {
                            const int icorner = ((~(index ^ ob)) & 7) + 1;
                            const int ipt = d->gridptr[(icorner - 1) + 8 * (size_t)(ic - 1)];
                            if (seen[ipt - 1]) continue;
                            seen[ipt - 1] = 1;
                            if (iorder >= npts) { ok = false; break; }
                            so[iorder] = (ic << 3) | (icorner - 1);
                            rk[ipt - 1] = iorder;
                            iorder++;
                        }
                    }
                }
            if (iz == z1) break;
        }
        if (!ok || iorder != npts) bad += 1;
    }
    if (bad) return false;
    return true;
}
}

// SWEEPING_ORDER alone (host only, no device needed): SWEEPORD(NPTS,NOCT), NOCT = 8 (4 for IPFLAG=2), entries cell<<3 | corner-1
extern "C" int at3d_sweeping_order(const at3d_state_desc *d, int32_t *sweepord, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !sweepord) { set_msg(errmsg, "null argument"); return 1; }
    if ((d->ipflag & 3) == 3 || (d->bcflag & 12)) { set_msg(errmsg, "at3d_sweeping_order: 3-D (8 octants) and 2-D (IPFLAG=2, 4 octants) grids without multi-processor flags only"); return 3; }
    std::vector<int> so, rank;
    if (!host_sweeping_order(d, (d->ipflag & 2) ? 4 : 8, so, rank)) { set_msg(errmsg, "SWEEPING_ORDER: not every grid point was reached"); return 1; }
    memcpy(sweepord, so.data(), so.size() * sizeof(int));
    return 0;
}

struct at3d_solver {
    Arena A;
    TrPlan *P = nullptr;
    PiArgs a;
    SwArgs w;
    int nst = 0, npts = 0, nang = 0, ntop = 0, nbot = 0;
    size_t nbc = 0;
    bool lamb = true;
    int *toppt = nullptr;
    int *shptr_d = nullptr, *rshptr_d = nullptr;
    int blocks_resident = 0;
    bool ip = false;            // IPFLAG=3: independent columns (BACK_INT_GRID1D), no sweep order, DOFIELD in place
    const float *gpos_d = nullptr;      // GRIDPOS and the point records of the sweep (3-D / 2-D grids), rebuilt by update_medium
    float4 *ptrec_d = nullptr;
    double transmin = 1.0;
};

extern "C" int at3d_solver_destroy(at3d_solver *sv)
{
    if (!sv) return 0;
    if (sv->P) tr_plan_destroy(sv->P);
    delete sv;
    return 0;
}

// the solver object for independent-pixel grids (IPFLAG=3): the state of at3d_path_integration_ip kept resident
static int solver_create_ip(const at3d_state_desc *d, const float *wtmu, at3d_solver **out, char *errmsg)
{
    const int nz = d->nz, npts = d->npts, nst = d->nstokes;
    if (nz < 1 || npts % nz != 0) { set_msg(errmsg, "at3d_solver_create: the independent-pixel grid must be the unsplit base grid"); return 3; }
    const int ncol = npts / nz;
    for (int c = 0; c < ncol; c++) {
        for (int iz = 0; iz < nz; iz++)
            if (d->gridpos[2 + 3 * (size_t)(iz + nz * c)] != d->zgrid[iz]) { set_msg(errmsg, "at3d_solver_create: the independent-pixel grid must be the unsplit base grid"); return 3; }
        if (d->bcptr[c] != nz * c + nz || d->bcptr[d->maxnbc + c] != nz * c + 1) { set_msg(errmsg, "at3d_solver_create: unexpected boundary point lists"); return 3; }
    }
    if (d->ntoppts != ncol || d->nbotpts != ncol) { set_msg(errmsg, "at3d_solver_create: unexpected boundary point counts"); return 3; }
    if (d->srctype != 'S' && d->units == 'B') { set_msg(errmsg, "UNITS='B' is not implemented"); return 3; }
    at3d_solver *sv = new at3d_solver();
    sv->ip = true;
    int rc = tr_plan_create(nst, d->nstleg, d->ml, d->mm, d->nlm, d->nmu, d->nphi0max, d->nphi0, d->mu, d->phi, wtmu, &sv->P, errmsg);
    if (rc) { delete sv; return rc; }
    const int nang = tr_plan_nang(sv->P), nh = nang / 2;
    sv->nst = nst; sv->npts = npts; sv->nang = nang; sv->ntop = ncol; sv->nbot = ncol;
    sv->lamb = d->sfctype1 == 'L';
    std::vector<float> amu(nang), aphi(nang), aw(nang);
    std::vector<int> aimu(nang), aiphi(nang);
    for (int i = 0, ia = 0; i < d->nmu; i++)
        for (int k = 0; k < d->nphi0[i]; k++, ia++) {
            amu[ia] = d->mu[i]; aphi[ia] = d->phi[i + (size_t)d->nmu * k];
            aw[ia] = fabsf(d->mu[i]) * d->wtdo[i + (size_t)d->nmu * k];
            aimu[ia] = i; aiphi[ia] = k;
        }
    Arena &A = sv->A;
    PiArgs &a = sv->a;
    memset(&a, 0, sizeof(a));
    memset(&sv->w, 0, sizeof(sv->w));
    a.npts = npts; a.nst = nst; a.nz = nz; a.ncol = ncol; a.nang = nang; a.nmu = d->nmu; a.nphi0max = d->nphi0max;
    a.ntop = ncol; a.nbot = ncol; a.nsfcpar = d->nsfcpar; a.srctype = d->srctype; a.units = d->units;
    a.sfctype0 = d->sfctype0; a.sfctype1 = d->sfctype1; a.wavelen = d->wavelen; a.solarmu = d->solarmu;
    a.solaraz = d->solaraz; a.gndalbedo = d->gndalbedo; a.gndtemp = d->gndtemp;
    sv->nbc = (size_t)nst * (ncol + (size_t)ncol * (sv->lamb ? 1 : 1 + nh));
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
    a.bcrad = A.alloc<float>(sv->nbc);
    a.botrad = A.alloc<float>((size_t)nst * ncol * nh);
    a.fluxes = A.alloc<float>((size_t)2 * npts);
    sv->shptr_d = A.alloc<int>((size_t)npts + 1);
    sv->rshptr_d = A.alloc<int>((size_t)npts + 1);
    if (!a.dofield || !a.total_ext || !a.zlev || !a.dirflux || !a.ang_mu || !a.ang_phi || !a.ang_w || !a.ang_imu || !a.ang_iphi ||
        !a.skyrad || !a.sfcgridparms || !a.bcrad || !a.botrad || !a.fluxes || !sv->shptr_d || !sv->rshptr_d) {
        at3d_solver_destroy(sv); set_msg(errmsg, "device allocation failure"); return 4;
    }
    *out = sv;
    return 0;
}

static double wall_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}

extern "C" int at3d_solver_create(const at3d_state_desc *d, const float *wtmu, float transmin, at3d_solver **out, char *errmsg)
{
    const bool timing = getenv("AT3D_SOLVER_TIMING") != nullptr;     // developer aid: set-up phases on stderr
    double t_prev = wall_ms();
    auto lap = [&](const char *what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        const double t = wall_ms();
        fprintf(stderr, "at3d_solver_create: %-28s %9.2f ms\n", what, t - t_prev);
        t_prev = t;
    };
    if (errmsg) errmsg[0] = 0;
    if (!d || !wtmu || !out) { set_msg(errmsg, "null argument"); return 1; }
    *out = nullptr;
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if ((d->ipflag & 3) == 3) return solver_create_ip(d, wtmu, out, errmsg);
    const bool two_d = (d->ipflag & 2) != 0;                   // IPFLAG=2: BACK_INT_GRID2D
    if (d->bcflag & 12) { set_msg(errmsg, "at3d_solver_create: multi-processor boundary flags are not supported"); return 3; }
    if (d->srctype != 'S' && d->units == 'B') { set_msg(errmsg, "UNITS='B' is not implemented"); return 3; }
    if (!(transmin >= 0.0f && transmin <= 1.0f)) { set_msg(errmsg, "TRANSMIN must be in [0,1]"); return 1; }
    const int npts = d->npts, nst = d->nstokes, noct = two_d ? 4 : 8;
    std::vector<int> sweepord, rank;
    if (!host_sweeping_order(d, noct, sweepord, rank)) { set_msg(errmsg, "SWEEPING_ORDER: not every grid point was reached"); return 1; }
    lap("host SWEEPING_ORDER");
    at3d_solver *sv = new at3d_solver();
    int rc = tr_plan_create(nst, d->nstleg, d->ml, d->mm, d->nlm, d->nmu, d->nphi0max, d->nphi0, d->mu, d->phi, wtmu, &sv->P, errmsg);
    if (rc) { delete sv; return rc; }
    const int nang = tr_plan_nang(sv->P), nh = nang / 2;
    sv->nst = nst; sv->npts = npts; sv->nang = nang; sv->ntop = d->ntoppts; sv->nbot = d->nbotpts;
    sv->lamb = d->sfctype1 == 'L';
    std::vector<float> amu(nang), aphi(nang), aw(nang);
    std::vector<int> aimu(nang), aiphi(nang);
    std::vector<OrdDir> dirs(nang);
    static const int joctorder3[8] = {1, 3, 5, 7, 2, 4, 6, 8}, joctorder2[8] = {1, 3, 1, 3, 2, 4, 2, 4};
    for (int i = 0, ia = 0; i < d->nmu; i++)
        for (int k = 0; k < d->nphi0[i]; k++, ia++) {
            const float mu = d->mu[i], phi = d->phi[i + (size_t)d->nmu * k];
            amu[ia] = mu; aphi[ia] = phi;
            aw[ia] = fabsf(mu) * d->wtdo[i + (size_t)d->nmu * k];
            aimu[ia] = i; aiphi[ia] = k;
            // the ray direction of BACK_INT_GRID3D (shdomsub1.f:3401-3441), host libm like the reference
            OrdDir &D = dirs[ia];
            const double pi = (double)acosf(-1.0f);
            D.cx = (double)sqrtf(1.0f - mu * mu) * cos((double)phi + pi);
            D.cy = (double)sqrtf(1.0f - mu * mu) * sin((double)phi + pi);
            D.cz = -(double)mu;
            if (fabs(D.cx) > (double)1.0E-5f) D.cxinv = 1.0 / D.cx; else { D.cx = 0.0; D.cxinv = (double)1.0E6f; }
            if (fabs(D.cy) > (double)1.0E-5f) D.cyinv = 1.0 / D.cy; else { D.cy = 0.0; D.cyinv = (double)1.0E6f; }
            if (two_d) { D.cy = 0.0; D.cyinv = (double)1.0E6f; }        // BACK_INT_GRID2D has no Y component (:4073-4095)
            D.czinv = 1.0 / D.cz;
            D.bitx = D.cx < 0.0 ? 1 : 0;
            D.bity = D.cy < 0.0 ? 1 : 0;
            if (D.cz < -(double)1.0E-3f) D.bitz = 1;
            else if (D.cz > (double)1.0E-3f) D.bitz = 0;
            else { at3d_solver_destroy(sv); set_msg(errmsg, "BACK_INT_GRID: Bad MU"); return 1; }
            if ((D.bitz == 1) != (ia >= nh)) { at3d_solver_destroy(sv); set_msg(errmsg, "unexpected ordinate order"); return 1; }
            D.ioct = 1 + D.bitx + 2 * D.bity + 4 * D.bitz;
            D.joct = two_d ? joctorder2[D.ioct - 1] : joctorder3[D.ioct - 1];
            D.pad = 0;
        }
    std::vector<unsigned char> bflag(npts, 0);
    std::vector<int> toppt(d->ntoppts), botpt(d->nbotpts);
    for (int i = 0; i < d->ntoppts; i++) { toppt[i] = d->bcptr[i] - 1; bflag[toppt[i]] |= 1; }
    for (int i = 0; i < d->nbotpts; i++) { botpt[i] = d->bcptr[d->maxnbc + i] - 1; bflag[botpt[i]] |= 2; }
    Arena &A = sv->A;
    PiArgs &a = sv->a;
    memset(&a, 0, sizeof(a));
    a.npts = npts; a.nst = nst; a.nz = d->nz; a.ncol = 0; a.nang = nang; a.nmu = d->nmu; a.nphi0max = d->nphi0max;
    a.ntop = d->ntoppts; a.nbot = d->nbotpts; a.nsfcpar = d->nsfcpar; a.srctype = d->srctype; a.units = d->units;
    a.sfctype0 = d->sfctype0; a.sfctype1 = d->sfctype1; a.wavelen = d->wavelen; a.solarmu = d->solarmu;
    a.solaraz = d->solaraz; a.gndalbedo = d->gndalbedo; a.gndtemp = d->gndtemp;
    sv->nbc = (size_t)nst * (a.ntop + (size_t)a.nbot * (sv->lamb ? 1 : 1 + nh));
    a.dofield = A.alloc<float>((size_t)npts * nst * nang);
    float *radf = A.alloc<float>((size_t)npts * nst * nang);
    a.fluxsrc = radf;
    a.total_ext = A.up(d->total_ext, npts); a.zlev = A.up(d->zgrid, d->nz); a.dirflux = A.up(d->dirflux, npts);
    a.ang_mu = A.up(amu.data(), nang); a.ang_phi = A.up(aphi.data(), nang); a.ang_w = A.up(aw.data(), nang);
    a.ang_imu = A.up(aimu.data(), nang); a.ang_iphi = A.up(aiphi.data(), nang);
    a.skyrad = A.up(d->skyrad, (size_t)nst * (d->nmu / 2) * d->nphi0max);
    a.sfcgridparms = A.up(d->sfcgridparms, (size_t)d->nsfcpar * a.nbot);
    if (d->sfcgridrad) {
        bool nonzero = false;
        for (size_t i = 0; i < (size_t)(nh + 1) * a.nbot && !nonzero; i++) nonzero = d->sfcgridrad[i] != 0.0f;
        if (nonzero) a.sfcgridrad = A.up(d->sfcgridrad, (size_t)(nh + 1) * a.nbot);
    }
    a.bcrad = A.alloc<float>(sv->nbc);
    a.botrad = A.alloc<float>((size_t)nst * a.nbot * nh);
    a.fluxes = A.alloc<float>((size_t)2 * npts);
    a.botpt = A.up(botpt.data(), botpt.size());
    sv->toppt = A.up(toppt.data(), toppt.size());
    sv->shptr_d = A.alloc<int>((size_t)npts + 1);
    sv->rshptr_d = A.alloc<int>((size_t)npts + 1);
    // topology records of the ray kernels
    const int *gp = A.up(d->gridptr, (size_t)8 * d->ncells), *np = A.up(d->neighptr, (size_t)6 * d->ncells);
    const int *tp = A.up(d->treeptr, (size_t)2 * d->ncells);
    const short *cf = A.up((const short *)d->cellflags, (size_t)d->ncells);
    const float *gpos = A.up(d->gridpos, (size_t)3 * npts);
    int4 *cellrec = A.alloc<int4>((size_t)4 * d->ncells);
    float4 *ptrec = A.alloc<float4>((size_t)npts);
    SwArgs &w = sv->w;
    memset(&w, 0, sizeof(w));
    w.npts = npts; w.nang = nang; w.nchunks = (npts + 255) / 256;
    w.npre[0] = d->ntoppts; w.npre[1] = d->nbotpts;
    w.two_d = two_d ? 1 : 0;
    w.dir = A.up(dirs.data(), dirs.size());
    w.sweepord = A.up(sweepord.data(), sweepord.size());
    for (int joct = 1; joct <= noct; joct++) {
        static const int ioctorder[8] = {1, 5, 2, 6, 3, 7, 4, 8};
        const bool upward = ((ioctorder[joct - 1] - 1) >> 2) & 1;          // BITZ = 1: CZ < 0, an upward ordinate
        int *rk = rank.data() + (size_t)npts * (joct - 1);
        for (int p : (upward ? botpt : toppt)) rk[p] = -1;
    }
    w.rank = A.up(rank.data(), rank.size());
    w.bflag = A.up(bflag.data(), bflag.size());
    w.srcdo = a.dofield; w.radf = radf;
    w.eps = (double)(1.0E-3f * (d->gridpos[2 + 3 * (size_t)(d->gridptr[7] - 1)] - d->gridpos[2 + 3 * (size_t)(d->gridptr[0] - 1)]));
    w.transmin = (double)transmin;
    w.ticket = A.alloc<int>(2); w.err = w.ticket ? w.ticket + 1 : nullptr;
    bool ok = a.dofield && radf && a.total_ext && a.zlev && a.dirflux && a.ang_mu && a.ang_phi && a.ang_w && a.ang_imu && a.ang_iphi &&
              a.skyrad && a.sfcgridparms && a.bcrad && a.botrad && a.fluxes && a.botpt && sv->toppt && sv->shptr_d && sv->rshptr_d &&
              gp && np && tp && cf && gpos && cellrec && ptrec && w.dir && w.sweepord && w.rank && w.bflag && w.ticket;
    if (ok) ok = launch_build_cellrec(d->ncells, gp, np, tp, cf, cellrec, 0) == cudaSuccess &&
                 launch_build_ptrec(npts, gpos, a.total_ext, ptrec, 0) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) { at3d_solver_destroy(sv); set_msg(errmsg, "device allocation failure"); return 4; }
    lap("uploads, records, tables");
    sv->gpos_d = gpos; sv->ptrec_d = ptrec; sv->transmin = (double)transmin;
    memset(&w.S, 0, sizeof(w.S));
    w.S.npts = npts; w.S.ncells = d->ncells; w.S.cellrec = cellrec; w.S.ptrec = ptrec;
    int dev = 0, nsm = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nst == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep3d_kernel<1, false>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep3d_kernel<3, false>, 256, 0);
    sv->blocks_resident = nsm * (per_sm > 0 ? per_sm : 1);
    // ordinates in flight together: all of a hemisphere (measured: the sweep in the reference order, i.e. the level pass,
    // takes 199 ms at 1.05 M points with all 177 in flight, 251 ms with 32, 487 ms with 8, 758 ms with 1)
    long g = nh;
    const char *genv = getenv("AT3D_SWEEP_GROUP");
    if (genv) g = atol(genv);
    w.group = (int)(g < 1 ? 1 : (g > nh ? nh : g));
    // dependency levels -> processing order (skipped for very large grids: the reference order is used as it is)
    const size_t nkeys = (size_t)npts * nang;
    const char *lenv = getenv("AT3D_SWEEP_LEVELS");
    const bool want_levels = lenv ? atoi(lenv) != 0 : true;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // key = ordinate << 48 | level << 24 | position: the sort must cover every ordinate bit (NMU=32 / NPHI=64 has ~1600
    // ordinates), so the end bit follows NANG; NANG is limited by the 16 bits above bit 48
    int key_end_bit = 49;
    while (key_end_bit < 64 && (1ull << (key_end_bit - 48)) < (unsigned long long)nang) key_end_bit++;
    if (want_levels && npts < (1 << 24) && nang < 65536 && nkeys * 32 < free_b / 2) {
        Arena T;
        int *level = T.alloc<int>(nkeys);
        unsigned long long *k0 = T.alloc<unsigned long long>(nkeys), *k1 = T.alloc<unsigned long long>(nkeys);
        int2 *plan = A.alloc<int2>(nkeys);
        unsigned char *nwalk = T.alloc<unsigned char>(nkeys);
        size_t tmpb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmpb, k0, k1, nkeys, 0, key_end_bit, 0);
        char *tmp = T.alloc<char>(tmpb);
        cudaError_t e = cudaSuccess;
        if (!level || !k0 || !k1 || !plan || !nwalk || !tmp) { at3d_solver_destroy(sv); set_msg(errmsg, "device allocation failure"); return 4; }
        const int nb = (int)((nkeys + 255) / 256);
        lap("level-pass allocations");
        cudaMemset(a.dofield, 0, (size_t)npts * nst * nang * sizeof(float));
        cudaMemset(nwalk, 0, nkeys);                               // the preset boundary points are never walked
        sweep3d_level_init_kernel<<<nb, 256>>>(npts, nang, w.bflag, level);
        w.level = level; w.nwalk = nwalk; w.plan = nullptr;
        for (int up = 0; up < 2; up++) {
            cudaMemsetAsync(w.ticket, 0, up == 0 ? 2 * sizeof(int) : sizeof(int), 0);
            sweep3d_kernel<1, true><<<w.nchunks * nh, 256>>>(w, up);
        }
        lap("level pass");
        sweep3d_key_kernel<<<nb, 256>>>(w, k0);
        cub::DeviceRadixSort::SortKeys(tmp, tmpb, k0, k1, nkeys, 0, key_end_bit, 0);
        sweep3d_plan_kernel<<<nb, 256>>>(w, k1, plan);
        int err = 0;
        e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(&err, w.err, sizeof(int), cudaMemcpyDeviceToHost);
        w.level = nullptr; w.nwalk = nullptr;
        if (e != cudaSuccess || err) {
            at3d_solver_destroy(sv);
            set_msg(errmsg, e != cudaSuccess ? "CUDA error in the level pass of at3d_solver_create: %s" : "BACK_INT_GRID3D walk failed in the level pass (code %s)",
                    e != cudaSuccess ? cudaGetErrorString(e) : (err == 3 ? "3: boundary without a valid face" : err == 4 ? "4: wait timed out" : "1/2: bad cell or SO<0"));
            return e != cudaSuccess ? 4 : 1;
        }
        w.plan = plan;
        lap("sort + plan records");
    }
    lap("release of the temporaries");
    *out = sv;
    return 0;
}

// PATH_INTEGRATION on device-resident SH arrays: shptr_d/src_d -> rad_d (addressed by rshptr_d), fluxes and BCRAD stay in
// the solver object.  Launches only (stream 0); *err_out is read back by the caller after a synchronisation.
static cudaError_t sv_path_integration_device(at3d_solver *sv, const int *shptr_d, const float *src_d, const int *rshptr_d, float *rad_d)
{
    const int npts = sv->npts, nst = sv->nst, nh = sv->nang / 2;
    PiArgs &a = sv->a;
    SwArgs &w = sv->w;
    cudaError_t e = cudaMemsetAsync(a.bcrad, 0, sv->nbc * sizeof(float), 0);
    if (e == cudaSuccess) e = tr_sh_to_do(sv->P, npts, shptr_d, src_d, a.dofield, 0);
    if (sv->ip) {
        // independent columns: all downward ordinates, the surface, all upward ordinates (in place on DOFIELD)
        const int nsw = (a.ncol * nh + 127) / 128, npbi = (npts + 255) / 256;
#define AT3D_SWEEP1D(NST)                                                      \
        pi_sweep_kernel<NST><<<nsw, 128>>>(a, 0);                              \
        pi_flux_kernel<NST><<<npbi, 256>>>(a, 0);                              \
        if (sv->lamb) pi_lambertian_kernel<<<(a.ncol + 127) / 128, 128>>>(a);  \
        else pi_brdf_kernel<NST><<<nsw, 128>>>(a);                             \
        pi_sweep_kernel<NST><<<nsw, 128>>>(a, 1);                              \
        pi_flux_kernel<NST><<<npbi, 256>>>(a, 1);
        if (nst == 1) { AT3D_SWEEP1D(1) } else { AT3D_SWEEP1D(3) }
#undef AT3D_SWEEP1D
        if (e == cudaSuccess) e = tr_do_to_sh(sv->P, npts, rshptr_d, a.dofield, rad_d, 0);
        return e;
    }
    const int npb = (npts + 255) / 256;
    const int ninit = (int)(((size_t)npts * nh + 255) / 256);
    const int ntb = (a.ntop * nh + 127) / 128, nbb = (a.nbot * nh + 127) / 128;
#define AT3D_SWEEP3D(NST)                                                                              \
    for (int up = 0; up < 2; up++) {                                                                   \
        cudaMemsetAsync(w.ticket, 0, up == 0 ? 2 * sizeof(int) : sizeof(int), 0);                      \
        sweep3d_init_kernel<NST><<<ninit, 256>>>(a, w.radf, sv->toppt, up);                            \
        if (up) {                                                                                      \
            if (sv->lamb) pi_lambertian_kernel<<<(a.nbot + 127) / 128, 128>>>(a);                      \
            else pi_brdf_kernel<NST><<<nbb, 128>>>(a);                                                 \
        }                                                                                              \
        sweep3d_boundary_kernel<NST><<<up ? nbb : ntb, 128>>>(a, w.radf, sv->toppt, up);               \
        sweep3d_kernel<NST, false><<<((npts - (w.plan ? w.npre[up] : 0) + 255) / 256) * nh, 256>>>(w, up);  \
        pi_flux_kernel<NST><<<npb, 256>>>(a, up);                                                      \
        if (!up && !sv->lamb) sweep3d_store_down_kernel<NST><<<nbb, 128>>>(a, w.radf);                 \
    }
    if (nst == 1) { AT3D_SWEEP3D(1) } else { AT3D_SWEEP3D(3) }
#undef AT3D_SWEEP3D
    if (e == cudaSuccess) e = tr_do_to_sh(sv->P, npts, rshptr_d, w.radf, rad_d, 0);
    return e;
}

static int sv_sweep_error(at3d_solver *sv, char *errmsg)
{
    int err = 0;
    if (sv->ip) return 0;
    cudaError_t e = cudaMemcpy(&err, sv->w.err, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in the 3-D sweep", cudaGetErrorString(e)); return 4; }
    if (err) {
        static const char *what[] = {"", "BACK_INT_GRID: ICELL=0", "BACK_INT_GRID: SO<0", "BACK_INT_GRID3D: INEXTCELL=0 without a valid face",
                                     "sweep wait timed out", "BACK_INT_GRID3D: RAD<0"};
        set_msg(errmsg, "%s", what[err < 6 ? err : 0]);
        return 1;
    }
    return 0;
}

// A new medium on the same grid (an optimisation step, at3d/medium.py:1813-1831 rebuilds the solver instead): the
// extinction, the direct beam and the surface parameters the object keeps are replaced; the topology records, SWEEPING_ORDER,
// the dependency levels and the sorted sweep plan stay -- with TRANSMIN >= 1 (the at3d default) the walks of BACK_INT_GRID
// end at the first valid face whatever the extinction, so the plan depends on the grid only.
extern "C" int at3d_solver_update_medium(at3d_solver *sv, const at3d_state_desc *d, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!sv || !d || !d->total_ext || !d->dirflux) { set_msg(errmsg, "null argument"); return 1; }
    PiArgs &a = sv->a;
    if (d->npts != sv->npts || d->nstokes != sv->nst || d->ntoppts != a.ntop || d->nbotpts != a.nbot || d->nsfcpar != a.nsfcpar) {
        set_msg(errmsg, "at3d_solver_update_medium: the state does not have the solver's grid");
        return 1;
    }
    if (!sv->ip && sv->transmin < 1.0) {
        set_msg(errmsg, "at3d_solver_update_medium: with TRANSMIN < 1 the sweep plan depends on the extinction; create a new solver");
        return 3;
    }
    const size_t npts = (size_t)sv->npts;
    cudaError_t e = cudaMemcpy((void *)a.total_ext, d->total_ext, npts * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void *)a.dirflux, d->dirflux, npts * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && d->sfcgridparms && a.sfcgridparms)
        e = cudaMemcpy((void *)a.sfcgridparms, d->sfcgridparms, (size_t)d->nsfcpar * a.nbot * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && d->skyrad && a.skyrad)
        e = cudaMemcpy((void *)a.skyrad, d->skyrad, (size_t)sv->nst * (d->nmu / 2) * d->nphi0max * sizeof(float), cudaMemcpyHostToDevice);
    a.gndalbedo = d->gndalbedo; a.gndtemp = d->gndtemp;
    if (e == cudaSuccess && sv->ptrec_d) {
        e = launch_build_ptrec((int)npts, sv->gpos_d, a.total_ext, sv->ptrec_d, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_solver_update_medium", cudaGetErrorString(e)); return 4; }
    return 0;
}

extern "C" int at3d_solver_path_integration(at3d_solver *sv, const int32_t *shptr, const float *source, const int32_t *rshptr,
                                            float *radiance, float *fluxes, float *bcrad, double *kernel_ms, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!sv || !shptr || !source || !rshptr || !radiance || !fluxes || !bcrad) { set_msg(errmsg, "null argument"); return 1; }
    const int npts = sv->npts, nst = sv->nst;
    PiArgs &a = sv->a;
    const size_t nsh = (size_t)nst * shptr[npts], nrad = (size_t)nst * rshptr[npts];
    Arena T;                                   // per-call: the SH arrays change length every iteration
    float *src_d = T.up(source, nsh), *rad_d = T.alloc<float>(nrad);
    if (!src_d || !rad_d) { set_msg(errmsg, "device allocation failure"); return 4; }
    cudaError_t e = cudaMemcpy(sv->shptr_d, shptr, ((size_t)npts + 1) * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(sv->rshptr_d, rshptr, ((size_t)npts + 1) * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_solver_path_integration", cudaGetErrorString(e)); return 4; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    e = sv_path_integration_device(sv, sv->shptr_d, src_d, sv->rshptr_d, rad_d);
    cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.0f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e == cudaSuccess) e = cudaMemcpy(radiance, rad_d, nrad * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(fluxes, a.fluxes, (size_t)2 * npts * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(bcrad, a.bcrad, sv->nbc * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in at3d_solver_path_integration", cudaGetErrorString(e)); return 4; }
    const int rc = sv_sweep_error(sv, errmsg);
    if (rc) return rc;
    if (kernel_ms) *kernel_ms = ms;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// SOLUTION_ITERATIONS on a fixed grid, device-resident (src/polarized/shdomsub1.f:445-822 without SPLIT_GRID):
// SOURCE, DELSOURCE, RADIANCE and their pointer arrays stay in HBM for the whole solve; per iteration only the four
// norms of COMPUTE_SOURCE, the two SH totals and the sweep's error flag come back to the host.
// ---------------------------------------------------------------------------------------------------------------------
struct RtArgs {
    int npts, ml, mm, nstleg, nleg, npart, nq, nstokes, interp_new, deltam, highorderrad;
    int ldp;                // leading dimension of the per-species point arrays (0: npts)
    float phasemax, shacc;
    const float *total_ext, *extinct, *albedo, *legen, *phaseinterpwt, *radiance;
    const int *iphase, *shptr, *rshptr_old, *lofj;
    size_t legen_size;      // elements of LEGEN
    int *first_zero;        // first point (0-based) whose old NR is 0 (NOTEND turns false there), npts if none
    int *nr;                // [npts+1] out
};

__global__ void rt_first_zero_kernel(RtArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.npts) return;
    if (a.rshptr_old[i + 1] - a.rshptr_old[i] == 0) atomicMin(a.first_zero, i);
}

// RADIANCE_TRUNCATION (shdomsub1.f:1615-1805), the adaptive branch: thread = grid point
__global__ void rt_adaptive_kernel(RtArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.npts) return;
    const int ml = a.ml, mm = a.mm, nq = a.nq;
    const size_t nlt = (size_t)a.nstleg * (a.nleg + 1);
    int lr;
    if (i < *a.first_zero) {
        const float ext = a.total_ext[i];
        const float rad0 = a.radiance[(size_t)a.nstokes * a.rshptr_old[i]];
        lr = 1;
        for (int l = 1; l <= ml; l++) {
            float rad = 0.0f;
            for (int ipa = 0; ipa < a.npart; ipa++) {
                const size_t po = (size_t)i + (size_t)(a.ldp ? a.ldp : a.npts) * ipa;
                const float w = ext == 0.0f ? 1.0f : a.extinct[po] / ext;
                if (w == 0.0f) continue;
                const int *iph = a.iphase + (size_t)nq * po;
                const float *pw = a.phaseinterpwt + (size_t)nq * po;
                float legent;
                if (!a.interp_new) {
                    legent = a.legen[nlt * (size_t)(iph[0] - 1) + (size_t)a.nstleg * l];
                } else {
                    if (pw[0] >= a.phasemax) legent = a.legen[nlt * (size_t)(iph[0] - 1) + (size_t)a.nstleg * l];
                    else {
                        legent = 0.0f;
                        for (int q = 0; q < nq; q++) {
                            if (pw[q] <= 1e-5f) continue;
                            legent = legent + a.legen[nlt * (size_t)(iph[q] - 1) + (size_t)a.nstleg * l] * pw[q];
                        }
                    }
                    if (a.deltam) {
                        // F(IPA) of the point; the reference indexes LEGEN(Q,ML+1,IPHASE(Q,.)) with the mixture index Q (:1672)
                        float f;
                        if (pw[0] >= a.phasemax) f = a.legen[nlt * (size_t)(iph[0] - 1) + (size_t)a.nstleg * (ml + 1)];
                        else {
                            f = 0.0f;
                            for (int q = 0; q < nq; q++) {
                                if (pw[q] <= 1e-5f) continue;
                                // (Q > NSTLEG runs into the following table entries; past the end of LEGEN the last
                                // element is read, as the host restatement and the oracle do)
                                size_t off = nlt * (size_t)(iph[q] - 1) + (size_t)a.nstleg * (ml + 1) + q;
                                if (off >= a.legen_size) off = a.legen_size - 1;
                                f = f + a.legen[off] * pw[q];
                            }
                        }
                        legent = legent * (1.0f / (1 - f));
                    }
                }
                rad = rad + w * a.albedo[po] * legent * rad0;
            }
            if (rad > a.shacc) lr = l;
        }
    } else {
        lr = ml;
    }
    int ns = a.shptr[i + 1] - a.shptr[i];
    if (ns < 1) ns = 1;
    const int ls = a.lofj[ns - 1];
    lr = min(lr, ls + ml / 8 + 2);
    if (a.highorderrad) lr = ml;
    a.nr[i] = lr <= mm ? lr * (lr + 1) + lr + 1 : (2 * mm + 1) * lr - (mm * (1 + (mm - 1))) + mm + 1;
}

// the FIXSH / out-of-memory branch (:1774-1802)
__global__ void rt_fixed_kernel(RtArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.npts) return;
    int nr = max(4, a.shptr[i + 1] - a.shptr[i]);
    if (a.highorderrad) nr = a.ml <= a.mm ? a.ml * (a.ml + 1) + a.ml + 1 : (2 * a.mm + 1) * a.ml - (a.mm * (1 + (a.mm - 1))) + a.mm + 1;
    a.nr[i] = nr;
}

__global__ void sv_iota_kernel(int n, int step, int *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = step * i;
}

// ACCELERATE_SOLUTION (shdomsub1.f:1807-1832): warp = grid point
__global__ void sv_accelerate_kernel(int npts, int nst, float accelpar, const int *shptr, const int *oshptr, float *source, const float *delsource)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= npts) return;
    const int is = shptr[i], isd = oshptr[i];
    const int nsc = min(shptr[i + 1] - is, oshptr[i + 1] - isd) * nst;
    float *s = source + (size_t)nst * is;
    const float *d = delsource + (size_t)nst * isd;
    for (int j = lane; j < nsc; j += 32) s[j] = s[j] + accelpar * d[j];
}

extern "C" int at3d_solver_solve(at3d_solver *sv, const at3d_state_desc *d, int maxiter, float solacc, float shacc, int accelflag,
                                 int highorderrad, int iterfixsh, int maxiv, int32_t *shptr, float *source, int32_t *rshptr,
                                 float *radiance, float *fluxes, float *bcrad, int32_t *iters_out, float *solcrit_out,
                                 double *ms_out /*[3]: PATH_INTEGRATION, COMPUTE_SOURCE, whole loop*/, char *errmsg)
{
    return at3d_solver_solve_from(sv, d, maxiter, solacc, shacc, accelflag, highorderrad, iterfixsh, maxiv, 0, shptr, source,
                                  rshptr, radiance, fluxes, bcrad, iters_out, solcrit_out, ms_out, errmsg);
}

// restore != 0: the iterations continue from the SHPTR / SOURCE / RSHPTR / RADIANCE passed in -- the solution of a nearby
// medium on the same grid, what RTE.load_solution + INIT_SOLUTION with INRADFLAG=.FALSE. set up (at3d/solver.py:2654-2666,
// shdomsub1.f:356-391): no first guess, no first COMPUTE_SOURCE, OSHPTR = SHPTR, DELSOURCE = 0, ITER = 0.
extern "C" int at3d_solver_solve_from(at3d_solver *sv, const at3d_state_desc *d, int maxiter, float solacc, float shacc,
                                      int accelflag, int highorderrad, int iterfixsh, int maxiv, int restore, int32_t *shptr,
                                      float *source, int32_t *rshptr, float *radiance, float *fluxes, float *bcrad,
                                      int32_t *iters_out, float *solcrit_out, double *ms_out, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!sv || !d || !shptr || !source || !rshptr || !radiance || !fluxes || !bcrad) { set_msg(errmsg, "null argument"); return 1; }
    if (d->npts != sv->npts || d->nstokes != sv->nst) { set_msg(errmsg, "at3d_solver_solve: desc does not match the solver object"); return 1; }
    const int npts = sv->npts, nst = sv->nst;
    const size_t nlt = (size_t)d->nstleg * (d->nleg + 1);
    if (nlt > 256) { set_msg(errmsg, "COMPUTE_SOURCE: Legendre table longer than 256 entries"); return 3; }
    const size_t maxir = (size_t)maxiv + npts;
    if ((size_t)4 * npts > maxir || maxiv < 1) { set_msg(errmsg, "MAXIV too small"); return 2; }
    const int nq = 8 * d->maxnmicro;
    Arena A;
    CsArgs a;
    memset(&a, 0, sizeof(a));
    a.npts = npts; a.nstokes = nst; a.nstleg = d->nstleg; a.nlm = d->nlm; a.ml = d->ml; a.mm = d->mm; a.nleg = d->nleg;
    a.npart = d->npart; a.nq = nq; a.srctype = d->srctype; a.deltam = d->deltam; a.interp_new = d->interp_new;
    a.newmethod = 1; a.accelflag = accelflag;
    a.phasemax = d->phasemax; a.secmu0 = 1.0f / fabsf(d->solarmu); a.srcmin = shacc;
    std::vector<int> lofj(d->nlm);
    {
        int j = 0;
        for (int l = 0; l <= d->ml; l++) {
            const int me = l < d->mm ? l : d->mm;
            for (int m = -me; m <= me; m++) { if (j < d->nlm) lofj[j] = l; j++; }
        }
        if (j != d->nlm) { set_msg(errmsg, "NLM inconsistent with ML, MM"); return 1; }
    }
    float albmax = 0.0f;
    for (size_t i = 0; i < (size_t)npts * d->npart; i++) albmax = d->albedo[i] > albmax ? d->albedo[i] : albmax;
    a.extinct = A.up(d->extinct, (size_t)npts * d->npart); a.albedo = A.up(d->albedo, (size_t)npts * d->npart);
    a.total_ext = sv->a.total_ext;
    a.legen = A.up(d->legen, nlt * d->numphase);
    a.iphase = A.up(d->iphase, (size_t)nq * npts * d->npart);
    a.phaseinterpwt = A.up(d->phaseinterpwt, (size_t)nq * npts * d->npart);
    a.dirflux = sv->a.dirflux;
    a.ylmsun = A.up(d->ylmsun, (size_t)d->nstleg * d->nlm);
    a.planck = d->planck ? A.up(d->planck, (size_t)npts * d->npart) : nullptr;
    a.lofj = A.up(lofj.data(), lofj.size());
    const int nblk = cs_grid_blocks(npts);
    size_t tmpb = cs_scan_bytes(npts);
    void *tmp = A.alloc<unsigned char>(tmpb);
    a.ns_new = A.alloc<int>((size_t)npts + 1);
    a.partials = A.alloc<double>((size_t)nblk * 4);
    a.bad = A.alloc<int>(2);
    double *sums = A.alloc<double>(4);
    a.mix_legent = A.alloc<float>((size_t)npts * nlt);       // mixed once: the optical properties do not change over the solve
    a.mix_ap = A.alloc<float2>((size_t)npts);
    if (!a.mix_legent || !a.mix_ap) { set_msg(errmsg, "device allocation failure"); return 4; }
    float *src[2] = {A.alloc<float>((size_t)nst * maxiv), A.alloc<float>((size_t)nst * maxiv)};
    float *dels = A.alloc<float>((size_t)nst * maxiv);
    float *rad = A.alloc<float>((size_t)nst * maxir);
    int *sh[2] = {A.alloc<int>((size_t)npts + 1), A.alloc<int>((size_t)npts + 1)};
    int *osh = A.alloc<int>((size_t)npts + 1);
    int *rsh[2] = {A.alloc<int>((size_t)npts + 2), A.alloc<int>((size_t)npts + 2)};
    int *nr = A.alloc<int>((size_t)npts + 1);
    if (!a.extinct || !a.albedo || !a.legen || !a.iphase || !a.phaseinterpwt || !a.ylmsun || !a.lofj || !tmp || !a.ns_new || !a.partials ||
        !a.bad || !sums || !src[0] || !src[1] || !dels || !rad || !sh[0] || !sh[1] || !osh || !rsh[0] || !rsh[1] || !nr || (d->planck && !a.planck)) {
        set_msg(errmsg, "at3d_solver_solve: NULL input array or device allocation failure"); return 4;
    }
    RtArgs rt;
    memset(&rt, 0, sizeof(rt));
    rt.npts = npts; rt.ml = d->ml; rt.mm = d->mm; rt.nstleg = d->nstleg; rt.nleg = d->nleg; rt.npart = d->npart; rt.nq = nq;
    rt.nstokes = nst; rt.interp_new = d->interp_new; rt.deltam = d->deltam; rt.highorderrad = highorderrad;
    rt.phasemax = d->phasemax; rt.shacc = shacc; rt.legen_size = nlt * (size_t)d->numphase;
    rt.total_ext = a.total_ext; rt.extinct = a.extinct; rt.albedo = a.albedo; rt.legen = a.legen; rt.phaseinterpwt = a.phaseinterpwt;
    rt.iphase = a.iphase; rt.lofj = a.lofj; rt.first_zero = a.bad + 1; rt.nr = nr; rt.radiance = rad;
    const int pb = (npts + 255) / 256;
    cudaEvent_t ev[4];
    for (auto &x : ev) cudaEventCreate(&x);
    cudaEventRecord(ev[0], 0);
    int cur = 0, rcur = 0, rc = 0, total_s = 0, iter = 0, fixsh = 0;
    int total_r = 4 * npts;
    size_t cap_new = (size_t)maxiv < (size_t)npts * d->nlm ? (size_t)maxiv : (size_t)npts * d->nlm;
    a.delsource_old = dels; a.delsource_new = dels;
    if (restore) {
        // the caller's solution becomes the current one
        total_s = shptr[npts]; total_r = rshptr[npts];
        if (total_s < 0 || total_r < 0 || total_s > maxiv || (size_t)total_r > maxir) {
            set_msg(errmsg, "at3d_solver_solve_from: the restored solution does not fit MAXIV"); rc = 2;
        } else {
            cudaMemcpyAsync(sh[cur], shptr, ((size_t)npts + 1) * sizeof(int), cudaMemcpyHostToDevice, 0);
            cudaMemcpyAsync(rsh[rcur], rshptr, ((size_t)npts + 1) * sizeof(int), cudaMemcpyHostToDevice, 0);
            cudaMemcpyAsync(rsh[rcur] + npts + 1, rsh[rcur] + npts, sizeof(int), cudaMemcpyDeviceToDevice, 0);
            cudaMemcpyAsync(src[cur], source, (size_t)nst * total_s * sizeof(float), cudaMemcpyHostToDevice, 0);
            cudaMemcpyAsync(rad, radiance, (size_t)nst * total_r * sizeof(float), cudaMemcpyHostToDevice, 0);
            cudaMemsetAsync(osh, 0, ((size_t)npts + 1) * sizeof(int), 0);
            // the mixed Legendre rows of the optical properties, which the first COMPUTE_SOURCE builds on a cold start
            a.first = 0; a.fixsh = 0;
            a.rshptr = rsh[rcur]; a.radiance = rad; a.shptr_old = sh[cur]; a.oshptr_old = osh; a.source_old = src[cur];
            if (cs_mix_points(a, 0, npts, 0) != cudaSuccess) { set_msg(errmsg, "CUDA error in cs_mix_points"); rc = 4; }
        }
    } else {
        // first guess: zero radiance with 4 terms per point, SOURCE from COMPUTE_SOURCE(FIRST=.TRUE.)
        sv_iota_kernel<<<(npts + 2 + 255) / 256, 256>>>(npts + 1, 4, rsh[rcur]);
        cudaMemcpyAsync(rsh[rcur] + npts + 1, rsh[rcur] + npts, sizeof(int), cudaMemcpyDeviceToDevice, 0);
        cudaMemsetAsync(rad, 0, (size_t)nst * 4 * npts * sizeof(float), 0);
        cudaMemsetAsync(sh[cur], 0, ((size_t)npts + 1) * sizeof(int), 0);
        cudaMemsetAsync(osh, 0, ((size_t)npts + 1) * sizeof(int), 0);
        a.first = 1; a.fixsh = 0;
        a.rshptr = rsh[rcur]; a.radiance = rad; a.shptr_old = sh[cur]; a.oshptr_old = osh; a.source_old = src[cur];
        rc = cs_device_step(a, nblk, tmp, tmpb, sh[1 - cur], sums, maxiv, cap_new, src[1 - cur], &total_s, false, errmsg);
        cur = 1 - cur;
    }
    if (!rc && accelflag) {
        cudaMemcpyAsync(osh, sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToDevice, 0);
        cudaMemsetAsync(dels, 0, (size_t)nst * total_s * sizeof(float), 0);
    }
    float solcrit = 1.0f, acc = 0.0f, deljdot = 0, deljold = 0, deljnew = 0, jnorm = 0;
    double ms_path = 0.0, ms_src = 0.0;
    while (!rc && iter < maxiter && solcrit > solacc) {
        iter++;
        // RADIANCE_TRUNCATION -> new RSHPTR (the old RADIANCE / RSHPTR are still in place)
        rt.shptr = sh[cur]; rt.rshptr_old = rsh[rcur];
        bool fixed = fixsh != 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            if (!fixed) {
                cudaMemcpyAsync(rt.first_zero, &npts, sizeof(int), cudaMemcpyHostToDevice, 0);
                rt_first_zero_kernel<<<pb, 256>>>(rt);
                rt_adaptive_kernel<<<pb, 256>>>(rt);
            } else {
                rt_fixed_kernel<<<pb, 256>>>(rt);
            }
            cudaMemsetAsync(nr + npts, 0, sizeof(int), 0);
            cub::DeviceScan::ExclusiveSum(tmp, tmpb, nr, rsh[1 - rcur], npts + 1);
            cudaMemcpy(&total_r, rsh[1 - rcur] + npts, sizeof(int), cudaMemcpyDeviceToHost);
            if ((size_t)total_r <= maxir) break;
            if (fixed) { set_msg(errmsg, "RADIANCE_TRUNCATION: Really out of memory for more radiance terms. Increase MAXIV."); rc = 2; break; }
            fixed = true;
        }
        if (rc) break;
        rcur = 1 - rcur;
        cudaMemcpyAsync(rsh[rcur] + npts + 1, rsh[rcur] + npts, sizeof(int), cudaMemcpyDeviceToDevice, 0);
        // PATH_INTEGRATION
        cudaEventRecord(ev[1], 0);
        cudaError_t e = sv_path_integration_device(sv, sh[cur], src[cur], rsh[rcur], rad);
        cudaEventRecord(ev[2], 0);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in PATH_INTEGRATION", cudaGetErrorString(e)); rc = 4; break; }
        if (solcrit < 0.001f || iter > iterfixsh) fixsh = 1;
        // COMPUTE_SOURCE
        a.first = 0; a.fixsh = fixsh;
        a.rshptr = rsh[rcur]; a.radiance = rad; a.shptr_old = sh[cur]; a.oshptr_old = osh; a.source_old = src[cur];
        int old_total = total_s;
        cap_new = fixsh ? (size_t)old_total : ((size_t)maxiv < (size_t)npts * d->nlm ? (size_t)maxiv : (size_t)npts * d->nlm);
        rc = cs_device_step(a, nblk, tmp, tmpb, sh[1 - cur], sums, maxiv, cap_new, src[1 - cur], &total_s, true, errmsg);
        cudaEventRecord(ev[3], 0);
        if (rc) break;
        if (accelflag) cudaMemcpyAsync(osh, sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToDevice, 0);   // OSHPTR = old SHPTR
        cur = 1 - cur;
        double hs[4];
        e = cudaMemcpy(hs, sums, sizeof(hs), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in the solution iterations", cudaGetErrorString(e)); rc = 4; break; }
        rc = sv_sweep_error(sv, errmsg);
        if (rc) break;
        float t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t1, ev[1], ev[2]); cudaEventElapsedTime(&t2, ev[2], ev[3]);
        ms_path += t1; ms_src += t2;
        deljdot = (float)hs[0]; deljold = (float)hs[1]; deljnew = (float)hs[2]; jnorm = (float)hs[3];
        // CALC_ACCEL_SOLCRIT (src/shdom_nompi.f:317-349)
        if (accelflag && acc == 0.0f && deljnew < deljold) {
            const float r = sqrtf(deljnew / deljold);
            const float theta = acosf(deljdot / sqrtf(deljold * deljnew));
            acc = (1 - r * cosf(theta) + powf(r, 1 + 0.5f * 3.14159f / theta)) / (1 + r * r - 2 * r * cosf(theta)) - 1.0f;
            acc = fminf(10.0f, fmaxf(0.0f, acc));
        } else {
            acc = 0.0f;
        }
        if (jnorm > 0.0f) solcrit = sqrtf(deljnew / jnorm);
        else if (deljnew == 0.0f) solcrit = 0.0f;
        if (acc > 0.0f) sv_accelerate_kernel<<<(int)(((size_t)npts * 32 + 255) / 256), 256>>>(npts, nst, acc, sh[cur], osh, src[cur], dels);
        if (albmax < solacc) solcrit = solacc;
    }
    float ms_all = 0.0f;
    cudaEventRecord(ev[3], 0);
    cudaError_t e = cudaEventSynchronize(ev[3]);
    if (e == cudaSuccess) cudaEventElapsedTime(&ms_all, ev[0], ev[3]);
    for (auto &x : ev) cudaEventDestroy(x);
    if (!rc && e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in the solution iterations", cudaGetErrorString(e)); rc = 4; }
    if (!rc) {
        e = cudaMemcpy(shptr, sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(rshptr, rsh[rcur], ((size_t)npts + 2) * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(source, src[cur], (size_t)nst * total_s * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(radiance, rad, (size_t)nst * total_r * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(fluxes, sv->a.fluxes, (size_t)2 * npts * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(bcrad, sv->a.bcrad, sv->nbc * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s copying the solution back", cudaGetErrorString(e)); rc = 4; }
    }
    if (iters_out) *iters_out = iter;
    if (solcrit_out) *solcrit_out = solcrit;
    if (ms_out) { ms_out[0] = ms_path; ms_out[1] = ms_src; ms_out[2] = ms_all; }
    return rc;
}

// =====================================================================================================================
// The adaptive solve: INIT_SOLUTION + SOLUTION_ITERATIONS with SPLIT_GRID (shdomsub1.f:113-822, :4703-5902).
// The loop of at3d_solver_solve with every array at its RTE._setup_memory capacity, the Eddington first guess
// (INIT_RADIANCE), and -- between iterations -- the cell splitting of at3d_adapt.cu.  When the grid has grown, the sweep
// structures (order, levels, plan records, boundary lists) are rebuilt for the new grid by at3d_solver_create.
// =====================================================================================================================
#include "at3d_adapt.h"

namespace {
struct AdaptCtx {
    const at3d_state_desc *d;
    const at3d_prop_desc *pg;
    at3d_adapt_io *io;
    AdaptGrid *G;
    CsArgs *cs;
    TpaArgs tpa;
    int npts_done, ncells_done, nst, nq, npart, maxig;
    float *gridpos_d; int *gridptr_d;
    float *extinct_d, *albedo_d, *planck_d, *temp_d, *total_ext_d, *pwt_d, *dirflux_d;
    int *iphase_d;
    const float *zlevels_d, *extdirp_d;
    int *flags_d;
    double beam_d[13]; int beam_i[5];
    int **sh, *cur, **rsh, *rcur, *osh, osh_end;
    float **src, *rad;
    DevBuf recs;
};

int adapt_interpolate_cb(void *arg, const std::vector<NewPointRec> &recs, char *errmsg)
{
    AdaptCtx &c = *(AdaptCtx *)arg;
    AdaptGrid &G = *c.G;
    const int p0 = c.npts_done, p1 = G.npts, np = p1 - p0, c0 = c.ncells_done, c1 = G.ncells, nst = c.nst;
    if (np != (int)recs.size()) { set_msg(errmsg, "SPLIT_GRID: new point bookkeeping mismatch"); return 1; }
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(cudaMemcpy(c.gridptr_d + (size_t)8 * c0, G.gridptr + (size_t)8 * c0, (size_t)8 * (c1 - c0) * sizeof(int), cudaMemcpyHostToDevice));
    c.ncells_done = c1;
    if (np == 0) { if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in SPLIT_GRID", cudaGetErrorString(e)); return 4; } return 0; }
    ok(cudaMemcpy(c.gridpos_d + (size_t)3 * p0, G.gridpos + (size_t)3 * p0, (size_t)3 * np * sizeof(float), cudaMemcpyHostToDevice));
    // medium properties of the new points from the property grid (TRILIN_INTERP_PROP + delta-M), Planck source, direct beam
    c.tpa.first = p0; c.tpa.count = np;
    ok(cudaMemsetAsync(c.flags_d, 0, 3 * sizeof(int), 0));
    ok(launch_tpa(c.tpa, 0));
    if (c.d->srctype != 'T')
        ok(launch_direct_points(c.beam_d, c.beam_i, c.d->bcflag, c.pg->npx, c.pg->npy, c.pg->npz, c.pg->xstart, c.pg->ystart, c.zlevels_d,
                                c.gridpos_d + (size_t)3 * p0, c.extdirp_d, c.d->solarflux, c.dirflux_d + p0, np, c.flags_d + 1, 0));
    // SH pointers of the new points (assigned on the host in creation order), zero-length DELSOURCE ranges
    ok(cudaMemcpyAsync(c.sh[*c.cur] + p0 + 1, G.shptr.data() + p0 + 1, (size_t)np * sizeof(int), cudaMemcpyHostToDevice, 0));
    G.rshptr[p1 + 1] = G.rshptr[p1];
    ok(cudaMemcpyAsync(c.rsh[*c.rcur] + p0 + 1, G.rshptr.data() + p0 + 1, (size_t)(np + 1) * sizeof(int), cudaMemcpyHostToDevice, 0));
    {
        std::vector<int> fill(np, c.osh_end);
        ok(cudaMemcpy(c.osh + p0 + 1, fill.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (c.recs.reserve(recs.size() * sizeof(NewPointRec)) != cudaSuccess) { set_msg(errmsg, "SPLIT_GRID: device allocation failure"); return 4; }
    ok(cudaMemcpy(c.recs.p, recs.data(), recs.size() * sizeof(NewPointRec), cudaMemcpyHostToDevice));
    // radiance = mean of the parents, source function from it
    CsArgs a = *c.cs;
    a.npts = p1;
    ok(cs_mix_points(a, p0, np, 0));
    ok(launch_interp_points(a, (const NewPointRec *)c.recs.p, np, c.rsh[*c.rcur], c.rad, c.src[*c.cur], 0));
    // host copies of the new points' properties (sweep set-up, the caller's arrays)
    int hf[3] = {0, 0, 0};
    ok(cudaMemcpy(hf, c.flags_d, sizeof(hf), cudaMemcpyDeviceToHost));
    ok(cudaMemcpy(c.io->total_ext + p0, c.total_ext_d + p0, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost));
    ok(cudaMemcpy(c.io->dirflux + p0, c.dirflux_d + p0, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost));
    if (c.io->temp) ok(cudaMemcpy(c.io->temp + p0, c.temp_d + p0, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost));
    for (int ipa = 0; ipa < c.npart; ipa++) {
        const size_t o = (size_t)c.maxig * ipa + p0;
        ok(cudaMemcpy(c.io->extinct + o, c.extinct_d + o, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost));
        ok(cudaMemcpy(c.io->albedo + o, c.albedo_d + o, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost));
        if (c.io->planck && c.planck_d) ok(cudaMemcpy(c.io->planck + o, c.planck_d + o, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost));
        ok(cudaMemcpy(c.io->iphase + (size_t)c.nq * o, c.iphase_d + (size_t)c.nq * o, (size_t)c.nq * np * sizeof(int), cudaMemcpyDeviceToHost));
        ok(cudaMemcpy(c.io->phaseinterpwt + (size_t)c.nq * o, c.pwt_d + (size_t)c.nq * o, (size_t)c.nq * np * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in INTERPOLATE_POINT", cudaGetErrorString(e)); return 4; }
    if (hf[0]) { set_msg(errmsg, hf[0] == 1 ? "TRILIN: Beyond X domain" : "TRILIN: Beyond Y domain"); return 1; }
    if (hf[2]) { set_msg(errmsg, "DIRECT_BEAM_PROP failed for a new grid point (code %d)", hf[2] & 7); return 1; }
    (void)nst;
    c.npts_done = p1;
    return 0;
}

float host_planck(float temp, int units, float wavelen)
{
    if (units == 'T') return temp;
    if (temp > 0.0f) return 1.1911e8f / (wavelen * wavelen * wavelen * wavelen * wavelen) / (expf(1.4388e4f / (wavelen * temp)) - 1);
    return 0.0f;
}

// SURFACE_PARM_INTERP (shdomsub1.f:2221-2275) on the host: a few parameters per bottom point
void surface_parm_interp(int nbot, const int *bcptr_bot, const float *gridpos, int srctype, int units, float wavelen, int nxsfc, int nysfc,
                         float delxsfc, float delysfc, int nsfcpar, const float *sfcparms, float *out)
{
    for (int ibc = 0; ibc < nbot; ibc++) {
        const int i = bcptr_bot[ibc];
        const float rx = gridpos[3 * (size_t)(i - 1)] / delxsfc, ry = gridpos[1 + 3 * (size_t)(i - 1)] / delysfc;
        const int ix = std::max(1, std::min(nxsfc, (int)rx + 1)), iy = std::max(1, std::min(nysfc, (int)ry + 1));
        const float u = fmaxf(0.0f, fminf(1.0f, rx - (ix - 1))), v = fmaxf(0.0f, fminf(1.0f, ry - (iy - 1)));
        auto P = [&](int j, int x, int y) { return sfcparms[j + (size_t)nsfcpar * ((x - 1) + (size_t)(nxsfc + 1) * (y - 1))]; };
        for (int j = 0; j < nsfcpar; j++)
            out[j + (size_t)nsfcpar * ibc] = (1 - u) * (1 - v) * P(j, ix, iy) + (1 - u) * v * P(j, ix, iy + 1)
                                             + u * (1 - v) * P(j, ix + 1, iy) + u * v * P(j, ix + 1, iy + 1);
        out[(size_t)nsfcpar * ibc] = srctype == 'S' ? 0.0f : host_planck(out[(size_t)nsfcpar * ibc], units, wavelen);
    }
}
}  // namespace

extern "C" int at3d_solve_adaptive(const at3d_state_desc *d, const at3d_prop_desc *pg, const float *wtmu, at3d_adapt_io *io,
                                   double *ms_out, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !pg || !wtmu || !io) { set_msg(errmsg, "null argument"); return 1; }
    if (!io->gridpos || !io->gridptr || !io->neighptr || !io->treeptr || !io->cellflags || !io->extinct || !io->albedo || !io->total_ext ||
        !io->dirflux || !io->fluxes || !io->iphase || !io->phaseinterpwt || !io->shptr || !io->rshptr || !io->source || !io->radiance ||
        !io->bcptr || !io->bcrad || !io->extdirp) { set_msg(errmsg, "at3d_solve_adaptive: null array in at3d_adapt_io"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if ((d->ipflag & 3) == 3) { set_msg(errmsg, "at3d_solve_adaptive: independent-pixel grids (IPFLAG=3) take at3d_solver_solve"); return 3; }
    if (d->npts != io->nbpts || d->ncells != io->nbcells) { set_msg(errmsg, "at3d_solve_adaptive: the solve must start from the base grid"); return 3; }
    if (d->sfctype0 == 'V' && io->splitacc > 0.0f && (!io->sfcparms || !io->sfcgridparms)) {
        set_msg(errmsg, "at3d_solve_adaptive: a variable surface needs SFCPARMS to place new bottom points"); return 1;
    }
    const int nst = d->nstokes, npart = d->npart, nq = 8 * d->maxnmicro, nlm = d->nlm;
    const int maxig = io->maxig, maxic = io->maxic, maxiv = io->maxiv;
    const size_t nlt = (size_t)d->nstleg * (d->nleg + 1), maxir = (size_t)maxiv + maxig;
    if (nlt > 256) { set_msg(errmsg, "COMPUTE_SOURCE: Legendre table longer than 256 entries"); return 3; }
    if (8 * d->maxnmicro > TPA_MAXQ) { set_msg(errmsg, "at3d_solve_adaptive: MAXNMICRO > %d", TPA_MAXQ / 8); return 3; }
    if (maxig < d->npts || maxic < d->ncells || (size_t)4 * d->npts > (size_t)maxiv) { set_msg(errmsg, "MAXIG / MAXIC / MAXIV too small"); return 2; }
    const bool lamb = d->sfctype1 == 'L';
    const double t_begin = wall_ms();
    double ms_split = 0.0;
    int rc = 0;
    // ---- MAKE_DIRECT on the base grid (also EXTDIRP and the beam constants for the new points) ----
    AdaptCtx ctx;
    memset(ctx.beam_d, 0, sizeof(ctx.beam_d)); memset(ctx.beam_i, 0, sizeof(ctx.beam_i));
    if (d->srctype != 'T') {
        rc = at3d_make_direct(d->npts, d->bcflag, d->ipflag, d->deltam, d->ml, d->nstleg, pg->nlegp, d->solarflux, d->solarmu, d->solaraz,
                              io->gridpos, pg->npx, pg->npy, pg->npz, pg->delx, pg->dely, pg->xstart, pg->ystart, pg->zlevels, pg->extinctp,
                              pg->albedop, pg->legenp, pg->numphase, pg->iphasep, pg->phasewtp, pg->maxnmicro, npart, pg->nzckd, pg->zckd,
                              pg->gasabs, io->extdirp, io->dirflux, ctx.beam_d, ctx.beam_i, errmsg);
        if (rc) return rc;
    } else {
        for (int i = 0; i < d->npts; i++) io->dirflux[i] = 0.0f;
    }
    AdaptGrid G;
    G.maxig = maxig; G.maxic = maxic; G.maxiv = maxiv; G.maxido = io->maxido; G.npts = d->npts; G.ncells = d->ncells;
    G.gridptr = io->gridptr; G.neighptr = io->neighptr; G.treeptr = io->treeptr; G.cellflags = io->cellflags; G.gridpos = io->gridpos;
    G.shptr.assign((size_t)maxig + 2, 0); G.rshptr.assign((size_t)maxig + 3, 0);
    // ---- device arrays at their capacities ----
    Arena A;
    const size_t maxpg = (size_t)pg->npx * pg->npy * pg->npz;
    float *gridpos_d = A.alloc<float>((size_t)3 * maxig), *total_ext_d = A.alloc<float>(maxig), *dirflux_d = A.alloc<float>(maxig);
    float *temp_d = A.alloc<float>(maxig);
    float *extinct_d = A.alloc<float>((size_t)maxig * npart), *albedo_d = A.alloc<float>((size_t)maxig * npart);
    float *planck_d = (io->planck && d->srctype != 'S') ? A.alloc<float>((size_t)maxig * npart) : nullptr;
    int *iphase_d = A.alloc<int>((size_t)nq * maxig * npart);
    float *pwt_d = A.alloc<float>((size_t)nq * maxig * npart);
    int *gridptr_d = A.alloc<int>((size_t)8 * maxic);
    std::vector<float> ftab(d->numphase > 0 ? d->numphase : 1, 0.0f);
    for (int i = 0; i < d->numphase; i++) ftab[i] = d->legen[nlt * i + (size_t)d->nstleg * (d->ml + 1 <= d->nleg ? d->ml + 1 : d->nleg)];
    std::vector<int> lofj(nlm);
    {
        int j = 0;
        for (int l = 0; l <= d->ml; l++) { const int me = l < d->mm ? l : d->mm; for (int m = -me; m <= me; m++) { if (j < nlm) lofj[j] = l; j++; } }
        if (j != nlm) { set_msg(errmsg, "NLM inconsistent with ML, MM"); return 1; }
    }
    CsArgs a;
    memset(&a, 0, sizeof(a));
    a.npts = G.npts; a.ldp = maxig; a.nstokes = nst; a.nstleg = d->nstleg; a.nlm = nlm; a.ml = d->ml; a.mm = d->mm; a.nleg = d->nleg;
    a.npart = npart; a.nq = nq; a.srctype = d->srctype; a.deltam = d->deltam; a.interp_new = d->interp_new;
    a.newmethod = 1; a.accelflag = io->accelflag;
    a.phasemax = d->phasemax; a.secmu0 = 1.0f / fabsf(d->solarmu); a.srcmin = io->shacc;
    a.extinct = extinct_d; a.albedo = albedo_d; a.total_ext = total_ext_d; a.iphase = iphase_d; a.phaseinterpwt = pwt_d; a.dirflux = dirflux_d;
    a.planck = planck_d;
    a.legen = A.up(d->legen, nlt * d->numphase);
    a.ylmsun = A.up(d->ylmsun, (size_t)d->nstleg * nlm);
    a.lofj = A.up(lofj.data(), lofj.size());
    const int nblk_max = cs_grid_blocks(maxig);
    size_t tmpb = cs_scan_bytes(maxig + 1);
    void *tmp = A.alloc<unsigned char>(tmpb);
    a.ns_new = A.alloc<int>((size_t)maxig + 1);
    a.partials = A.alloc<double>((size_t)nblk_max * 4);
    a.bad = A.alloc<int>(2);
    double *sums = A.alloc<double>(4);
    a.mix_legent = A.alloc<float>((size_t)maxig * nlt);
    a.mix_ap = A.alloc<float2>((size_t)maxig);
    float *src[2] = {A.alloc<float>((size_t)nst * maxiv), A.alloc<float>((size_t)nst * maxiv)};
    float *dels = A.alloc<float>((size_t)nst * maxiv);
    float *rad = A.alloc<float>((size_t)nst * maxir);
    int *sh[2] = {A.alloc<int>((size_t)maxig + 2), A.alloc<int>((size_t)maxig + 2)};
    int *osh = A.alloc<int>((size_t)maxig + 2);
    int *rsh[2] = {A.alloc<int>((size_t)maxig + 3), A.alloc<int>((size_t)maxig + 3)};
    int *nr = A.alloc<int>((size_t)maxig + 2);
    int *flags_d = A.alloc<int>(3);
    const float *zgrid_d = A.up(d->zgrid, d->nz);
    // property grid on the device (TRILIN_INTERP_PROP / DIRECT_BEAM_PROP of the new points)
    TpaArgs &tp = ctx.tpa;
    memset(&tp, 0, sizeof(tp));
    tp.ld = maxig; tp.npart = npart; tp.mnm = pg->maxnmicro; tp.npx = pg->npx; tp.npy = pg->npy; tp.npz = pg->npz; tp.ml = d->ml;
    tp.deltam = d->deltam; tp.interp_new = d->interp_new; tp.nzckd = pg->nzckd; tp.srctype = d->srctype; tp.units = d->units;
    tp.delx = pg->delx; tp.dely = pg->dely; tp.xstart = pg->xstart; tp.ystart = pg->ystart; tp.phasemax = d->phasemax; tp.wavelen = d->wavelen;
    tpa_extmin(pg->zlevels, pg->npz, &tp.extmin, &tp.scatmin);
    tp.gridpos = gridpos_d; tp.zlevels = A.up(pg->zlevels, pg->npz);
    tp.tempp = pg->tempp ? A.up(pg->tempp, maxpg) : nullptr;
    tp.extinctp = A.up(pg->extinctp, maxpg * npart); tp.albedop = A.up(pg->albedop, maxpg * npart);
    tp.iphasep = A.up(pg->iphasep, (size_t)pg->maxnmicro * maxpg * npart); tp.phasewtp = A.up(pg->phasewtp, (size_t)pg->maxnmicro * maxpg * npart);
    tp.ftab = A.up(ftab.data(), ftab.size());
    tp.zckd = pg->nzckd > 0 ? A.up(pg->zckd, pg->nzckd) : nullptr; tp.gasabs = pg->nzckd > 0 ? A.up(pg->gasabs, pg->nzckd) : nullptr;
    tp.extinct = extinct_d; tp.albedo = albedo_d; tp.total_ext = total_ext_d; tp.phaseinterpwt = pwt_d; tp.temp = temp_d; tp.planck = planck_d;
    tp.iphase = iphase_d; tp.bad = flags_d;
    const float *extdirp_d = A.up(io->extdirp, maxpg);
    if (!gridpos_d || !total_ext_d || !dirflux_d || !temp_d || !extinct_d || !albedo_d || !iphase_d || !pwt_d || !gridptr_d || !a.legen ||
        !a.ylmsun || !a.lofj || !tmp || !a.ns_new || !a.partials || !a.bad || !sums || !a.mix_legent || !a.mix_ap || !src[0] || !src[1] ||
        !dels || !rad || !sh[0] || !sh[1] || !osh || !rsh[0] || !rsh[1] || !nr || !flags_d || !zgrid_d || !tp.zlevels || !tp.extinctp ||
        !tp.albedop || !tp.iphasep || !tp.phasewtp || !tp.ftab || !extdirp_d || (io->planck && d->srctype != 'S' && !planck_d)) {
        set_msg(errmsg, "at3d_solve_adaptive: device allocation failure"); return 4;
    }
    {
        const int n0 = G.npts;
        cudaMemcpy(gridpos_d, io->gridpos, (size_t)3 * n0 * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(gridptr_d, io->gridptr, (size_t)8 * G.ncells * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(total_ext_d, io->total_ext, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(dirflux_d, io->dirflux, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
        if (io->temp) cudaMemcpy(temp_d, io->temp, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
        else cudaMemset(temp_d, 0, (size_t)maxig * sizeof(float));
        for (int ipa = 0; ipa < npart; ipa++) {
            const size_t o = (size_t)maxig * ipa;
            cudaMemcpy(extinct_d + o, io->extinct + o, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
            cudaMemcpy(albedo_d + o, io->albedo + o, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
            if (planck_d) cudaMemcpy(planck_d + o, io->planck + o, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
            cudaMemcpy(iphase_d + (size_t)nq * o, io->iphase + (size_t)nq * o, (size_t)nq * n0 * sizeof(int), cudaMemcpyHostToDevice);
            cudaMemcpy(pwt_d + (size_t)nq * o, io->phaseinterpwt + (size_t)nq * o, (size_t)nq * n0 * sizeof(float), cudaMemcpyHostToDevice);
        }
    }
    float albmax = 0.0f;
    for (int ipa = 0; ipa < npart; ipa++)
        for (int i = 0; i < G.npts; i++) albmax = std::max(albmax, io->albedo[i + (size_t)maxig * ipa]);
    RtArgs rt;
    memset(&rt, 0, sizeof(rt));
    rt.npts = G.npts; rt.ldp = maxig; rt.ml = d->ml; rt.mm = d->mm; rt.nstleg = d->nstleg; rt.nleg = d->nleg; rt.npart = npart; rt.nq = nq;
    rt.nstokes = nst; rt.interp_new = d->interp_new; rt.deltam = d->deltam; rt.highorderrad = io->highorderrad;
    rt.phasemax = d->phasemax; rt.shacc = io->shacc; rt.legen_size = nlt * (size_t)d->numphase;
    rt.total_ext = total_ext_d; rt.extinct = extinct_d; rt.albedo = albedo_d; rt.legen = a.legen; rt.phaseinterpwt = pwt_d;
    rt.iphase = iphase_d; rt.lofj = a.lofj; rt.first_zero = a.bad + 1; rt.nr = nr; rt.radiance = rad;
    cudaEvent_t ev[4];
    for (auto &x : ev) cudaEventCreate(&x);
    cudaEventRecord(ev[0], 0);
    // ---- first guess: INIT_RADIANCE (Eddington) or zero, 4 terms per point; SOURCE from COMPUTE_SOURCE(FIRST=.TRUE.) ----
    int cur = 0, rcur = 0, total_s = 0, iter = 0, fixsh = 0, npts = G.npts, oldnpts = 0;
    int total_r = 4 * npts;
    sv_iota_kernel<<<(npts + 2 + 255) / 256, 256>>>(npts + 1, 4, rsh[rcur]);
    cudaMemcpyAsync(rsh[rcur] + npts + 1, rsh[rcur] + npts, sizeof(int), cudaMemcpyDeviceToDevice, 0);
    cudaMemsetAsync(rad, 0, (size_t)nst * 4 * npts * sizeof(float), 0);
    cudaMemsetAsync(sh[cur], 0, ((size_t)maxig + 2) * sizeof(int), 0);
    cudaMemsetAsync(osh, 0, ((size_t)maxig + 2) * sizeof(int), 0);
    if (io->inradflag) {
        float skyradalb = 0.0f;
        for (int imu = 1; imu <= d->nmu / 2; imu++)
            for (int iphi = 1; iphi <= d->nphi0[imu - 1]; iphi++)
                skyradalb = skyradalb + fabsf(d->mu[imu - 1]) * d->wtdo[(imu - 1) + (size_t)d->nmu * (iphi - 1)]
                            * d->skyrad[(size_t)nst * ((imu - 1) + (size_t)(d->nmu / 2) * (iphi - 1))];
        rc = adapt_init_radiance(d, maxig, npts / d->nz, zgrid_d, extinct_d, albedo_d, total_ext_d, temp_d, a.legen, iphase_d, pwt_d,
                                 skyradalb, 0.0f, rad, errmsg);
    }
    a.first = 1; a.fixsh = 0;
    a.rshptr = rsh[rcur]; a.radiance = rad; a.shptr_old = sh[cur]; a.oshptr_old = osh; a.source_old = src[cur];
    a.delsource_old = dels; a.delsource_new = dels;
    size_t cap_new = std::min((size_t)maxiv, (size_t)npts * nlm);
    int nblk = cs_grid_blocks(npts);
    if (!rc) rc = cs_device_step(a, nblk, tmp, tmpb, sh[1 - cur], sums, maxiv, cap_new, src[1 - cur], &total_s, false, errmsg);
    cur = 1 - cur;
    int osh_end = 0;
    if (!rc && io->accelflag) {
        cudaMemcpyAsync(osh, sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToDevice, 0);
        cudaMemsetAsync(dels, 0, (size_t)nst * total_s * sizeof(float), 0);
        osh_end = total_s;
    }
    // ---- the callback context of SPLIT_GRID ----
    ctx.d = d; ctx.pg = pg; ctx.io = io; ctx.G = &G; ctx.cs = &a; ctx.npts_done = npts; ctx.ncells_done = G.ncells; ctx.nst = nst; ctx.nq = nq;
    ctx.npart = npart; ctx.maxig = maxig; ctx.gridpos_d = gridpos_d; ctx.gridptr_d = gridptr_d; ctx.extinct_d = extinct_d;
    ctx.albedo_d = albedo_d; ctx.planck_d = planck_d; ctx.temp_d = temp_d; ctx.total_ext_d = total_ext_d; ctx.pwt_d = pwt_d;
    ctx.dirflux_d = dirflux_d; ctx.iphase_d = iphase_d; ctx.zlevels_d = tp.zlevels; ctx.extdirp_d = extdirp_d; ctx.flags_d = flags_d;
    ctx.sh = sh; ctx.cur = &cur; ctx.rsh = rsh; ctx.rcur = &rcur; ctx.osh = osh; ctx.src = src; ctx.rad = rad;
    at3d_solver *sv = nullptr;
    at3d_state_desc dd = *d;
    std::vector<float> sfcgp;
    const float endadaptsol = 0.001f, startadaptsol = 0.1f, splitacc = io->splitacc, solacc = io->solacc;
    const float adaptrange = startadaptsol / (3.0f * endadaptsol);
    float cursplitacc = splitacc * adaptrange, startsplitacc = cursplitacc, solcrit = 1.0f, avgsolcrit = solcrit, splitcrit = 0.0f;
    float acc = 0.0f, deljdot = 0, deljold = 0, deljnew = 0, jnorm = 0;
    bool splittesting = true, outofmem = false, mix_ready = true;     // the first COMPUTE_SOURCE mixed the base points
    int nsplit_calls = 0;
    double ms_path = 0.0, ms_src = 0.0;
    while (!rc && iter < io->maxiter && (solcrit > solacc || (splitcrit > splitacc && cursplitacc > splitacc && !outofmem))) {
        iter++;
        if (splitacc > 0.0f) {
            avgsolcrit = sqrtf(avgsolcrit * solcrit);
            const bool dosplit = solcrit <= startadaptsol && (solcrit > endadaptsol || cursplitacc > splitacc) && !outofmem;
            const float beta = logf(startsplitacc / splitacc) / logf(adaptrange);
            cursplitacc = fminf(cursplitacc, fmaxf(splitacc, splitacc * powf(avgsolcrit / (3.0f * endadaptsol), beta)));
            if (solcrit <= endadaptsol) cursplitacc = splitacc;
            if (splittesting) {
                const double t0 = wall_ms();
                cudaMemcpy(G.shptr.data(), sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToHost);
                cudaMemcpy(G.rshptr.data(), rsh[rcur], ((size_t)npts + 2) * sizeof(int), cudaMemcpyDeviceToHost);
                AdaptDev D = {gridptr_d, gridpos_d, total_ext_d, sh[cur], src[cur], nst};
                ctx.osh_end = osh_end;
                rc = G.split_grid(D, dosplit, outofmem, cursplitacc, splitcrit, d->nphi0max, nlm, adapt_interpolate_cb, &ctx, errmsg);
                if (rc) break;
                nsplit_calls++;
                if (G.npts != npts) {
                    npts = G.npts;
                    total_s = G.shptr[npts]; total_r = G.rshptr[npts];
                }
                if (solcrit > startadaptsol) startsplitacc = splitcrit;
                ms_split += wall_ms() - t0;
            }
            if (solcrit <= endadaptsol) splittesting = false;
        }
        a.npts = npts; rt.npts = npts;
        nblk = cs_grid_blocks(npts);
        const int pb = (npts + 255) / 256;
        // RADIANCE_TRUNCATION -> new RSHPTR
        rt.shptr = sh[cur]; rt.rshptr_old = rsh[rcur];
        bool fixed = fixsh != 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            if (!fixed) {
                cudaMemcpyAsync(rt.first_zero, &npts, sizeof(int), cudaMemcpyHostToDevice, 0);
                rt_first_zero_kernel<<<pb, 256>>>(rt);
                rt_adaptive_kernel<<<pb, 256>>>(rt);
            } else {
                rt_fixed_kernel<<<pb, 256>>>(rt);
            }
            cudaMemsetAsync(nr + npts, 0, sizeof(int), 0);
            cub::DeviceScan::ExclusiveSum(tmp, tmpb, nr, rsh[1 - rcur], npts + 1);
            cudaMemcpy(&total_r, rsh[1 - rcur] + npts, sizeof(int), cudaMemcpyDeviceToHost);
            if ((size_t)total_r <= maxir) break;
            if (fixed) { set_msg(errmsg, "RADIANCE_TRUNCATION: Really out of memory for more radiance terms. Increase MAXIV."); rc = 2; break; }
            fixed = true;
        }
        if (rc) break;
        rcur = 1 - rcur;
        cudaMemcpyAsync(rsh[rcur] + npts + 1, rsh[rcur] + npts, sizeof(int), cudaMemcpyDeviceToDevice, 0);
        // the sweep structures of the current grid (SWEEPING_ORDER, BOUNDARY_PNTS, SURFACE_PARM_INTERP when NPTS changed)
        if (npts != oldnpts) {
            const double t0 = wall_ms();
            int ntop = 0, nbot = 0;
            if (G.boundary_points(d->nang, lamb, io->maxnbc, io->maxbcrad, d->zgrid[0], d->zgrid[d->nz - 1], io->bcptr, &ntop, &nbot)) {
                set_msg(errmsg, "BOUNDARY_PNTS: MAXNBC exceeded"); rc = 1; break;
            }
            dd.npts = npts; dd.ncells = G.ncells; dd.gridptr = io->gridptr; dd.neighptr = io->neighptr; dd.treeptr = io->treeptr;
            dd.cellflags = io->cellflags; dd.gridpos = io->gridpos; dd.total_ext = io->total_ext; dd.dirflux = io->dirflux;
            dd.bcptr = io->bcptr; dd.maxnbc = io->maxnbc; dd.ntoppts = ntop; dd.nbotpts = nbot; dd.sfcgridrad = nullptr;
            if (d->sfctype0 == 'V' && io->sfcparms) {
                surface_parm_interp(nbot, io->bcptr + io->maxnbc, io->gridpos, d->srctype, d->units, d->wavelen, io->nxsfc, io->nysfc,
                                    io->delxsfc, io->delysfc, d->nsfcpar, io->sfcparms, io->sfcgridparms);
                dd.sfcgridparms = io->sfcgridparms;
            } else if (io->sfcgridparms) dd.sfcgridparms = io->sfcgridparms;
            else { sfcgp.assign((size_t)std::max(1, d->nsfcpar) * std::max(1, nbot), 0.0f); dd.sfcgridparms = sfcgp.data(); }
            if (sv) at3d_solver_destroy(sv);
            sv = nullptr;
            rc = at3d_solver_create(&dd, wtmu, io->transmin, &sv, errmsg);
            if (rc) break;
            io->ntoppts = ntop; io->nbotpts = nbot;
            if (oldnpts != 0) mix_ready = false;      // the mixed tables of the new points exist already (cs_mix_points), but keep it simple
            mix_ready = true;
            oldnpts = npts;
            ms_split += wall_ms() - t0;
        }
        // PATH_INTEGRATION
        cudaEventRecord(ev[1], 0);
        cudaError_t e = sv_path_integration_device(sv, sh[cur], src[cur], rsh[rcur], rad);
        cudaEventRecord(ev[2], 0);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in PATH_INTEGRATION", cudaGetErrorString(e)); rc = 4; break; }
        if (solcrit < endadaptsol || iter > io->iterfixsh) fixsh = 1;
        // COMPUTE_SOURCE
        a.first = 0; a.fixsh = fixsh;
        a.rshptr = rsh[rcur]; a.radiance = rad; a.shptr_old = sh[cur]; a.oshptr_old = osh; a.source_old = src[cur];
        const int old_total = total_s;
        cap_new = fixsh ? (size_t)old_total : std::min((size_t)maxiv, (size_t)npts * nlm);
        rc = cs_device_step(a, nblk, tmp, tmpb, sh[1 - cur], sums, maxiv, cap_new, src[1 - cur], &total_s, mix_ready, errmsg);
        cudaEventRecord(ev[3], 0);
        if (rc) break;
        if (io->accelflag) { cudaMemcpyAsync(osh, sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToDevice, 0); osh_end = old_total; }
        cur = 1 - cur;
        double hs[4];
        e = cudaMemcpy(hs, sums, sizeof(hs), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in the solution iterations", cudaGetErrorString(e)); rc = 4; break; }
        rc = sv_sweep_error(sv, errmsg);
        if (rc) break;
        float t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t1, ev[1], ev[2]); cudaEventElapsedTime(&t2, ev[2], ev[3]);
        ms_path += t1; ms_src += t2;
        deljdot = (float)hs[0]; deljold = (float)hs[1]; deljnew = (float)hs[2]; jnorm = (float)hs[3];
        if (io->accelflag && acc == 0.0f && deljnew < deljold) {
            const float r = sqrtf(deljnew / deljold);
            const float theta = acosf(deljdot / sqrtf(deljold * deljnew));
            acc = (1 - r * cosf(theta) + powf(r, 1 + 0.5f * 3.14159f / theta)) / (1 + r * r - 2 * r * cosf(theta)) - 1.0f;
            acc = fminf(10.0f, fmaxf(0.0f, acc));
        } else {
            acc = 0.0f;
        }
        if (jnorm > 0.0f) solcrit = sqrtf(deljnew / jnorm);
        else if (deljnew == 0.0f) solcrit = 0.0f;
        if (acc > 0.0f) sv_accelerate_kernel<<<(int)(((size_t)npts * 32 + 255) / 256), 256>>>(npts, nst, acc, sh[cur], osh, src[cur], dels);
        if (albmax < solacc) solcrit = solacc;
        if (getenv("AT3D_SOLVER_VERBOSE"))
            fprintf(stderr, "  %4d %8.3f %10.3E %8d %8.2f %6.3f\n", iter, log10f(fmaxf(solcrit, 1.0e-20f)), splitcrit, npts,
                    (float)total_s / npts, (float)total_s / (npts * (float)nlm));
    }
    float ms_all = 0.0f;
    cudaEventRecord(ev[3], 0);
    cudaError_t e = cudaEventSynchronize(ev[3]);
    if (e == cudaSuccess) cudaEventElapsedTime(&ms_all, ev[0], ev[3]);
    for (auto &x : ev) cudaEventDestroy(x);
    if (!rc && e != cudaSuccess) { set_msg(errmsg, "CUDA error %s in the solution iterations", cudaGetErrorString(e)); rc = 4; }
    if (!rc && sv) {
        e = cudaMemcpy(io->shptr, sh[cur], ((size_t)npts + 1) * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(io->rshptr, rsh[rcur], ((size_t)npts + 2) * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(io->source, src[cur], (size_t)nst * total_s * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(io->radiance, rad, (size_t)nst * total_r * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(io->fluxes, sv->a.fluxes, (size_t)2 * npts * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(io->bcrad, sv->a.bcrad, std::min(sv->nbc, (size_t)nst * io->maxbcrad) * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s copying the solution back", cudaGetErrorString(e)); rc = 4; }
    }
    if (sv) at3d_solver_destroy(sv);
    ctx.recs.release();
    io->npts = npts; io->ncells = G.ncells; io->iters = iter; io->solcrit = solcrit; io->splitcrit = splitcrit; io->nsplit_calls = nsplit_calls;
    if (ms_out) { ms_out[0] = ms_path; ms_out[1] = ms_src; ms_out[2] = ms_split; ms_out[3] = wall_ms() - t_begin; (void)ms_all; }
    return rc;
}
