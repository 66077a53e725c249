// at3d_surface.cu -- general (non-Lambertian) bottom boundary of RENDER (sm_100a).
// Replaces, for rays that end on the surface, the non-Lambertian branch of FIND_BOUNDARY_RADIANCE
// (src/polarized/shdomsub2.f:2748-2863) with VARIABLE_BRDF_SURFACE (src/polarized/shdomsub1.f:2597-2669).
//
// The reference evaluates, per ray and for each of the 4 boundary points of the exit face, the BRDF for the
// NANG/2 stored downwelling ordinates plus the direct beam (4 x 178 evaluations at NMU=16, NPHI=32) inside the
// ray march.  Here the march kernels only leave a 64-byte SurfHit per ray; this kernel then gives every hit a
// whole warp (lanes over the ordinates, warp-shuffle reduction), so the march keeps its occupancy and the
// BRDF work is not serialised on one lane.
#include "at3d_ray.cuh"
#include "at3d_surface.cuh"
#include "at3d_host.h"

// bilinear interpolation of the 4 face-point radiances (REAL u, v as in FIND_BOUNDARY_RADIANCE) and the final sum
template <int NST, typename OUTA>
__device__ __forceinline__ void finish_hit(const DevState &S, const SurfHit *hits, int iray, const float (&x)[4],
                                           const float (&y)[4], const float (&rad)[4][NST], OUTA *out)
{
    const double xb = hits[iray].xb, yb = hits[iray].yb, tr = hits[iray].transmit;
    float u, v;
    if (x[1] - x[0] > 0.0f) u = (float)((xb - x[0]) / (x[1] - x[0])); else u = 0.0f;
    if (y[2] - y[0] > 0.0f) v = (float)((yb - y[0]) / (y[2] - y[0])); else v = 0.0f;
#pragma unroll
    for (int k = 0; k < NST; k++) {
        const float radbnd = (1 - u) * (1 - v) * rad[0][k] + u * (1 - v) * rad[1][k]
                             + (1 - u) * v * rad[2][k] + u * v * rad[3][k];
        out[k + NST * (size_t)iray] = (OUTA)(hits[iray].rad[k] + tr * radbnd);
    }
    if (S.counts) atomicAdd(&S.counts[6], 1ull);
}

template <int NST, typename OUTA>
__global__ void __launch_bounds__(256)
surface_kernel(DevState S, int nrays, const SurfHit *hits, const double *cammu, const double *camphi,
               OUTA *out, RayErr *err)
{
    const int gf[6][4] = {{1,3,5,7},{2,4,6,8},{1,2,5,6},{3,4,7,8},{1,2,3,4},{5,6,7,8}};
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nh = S.nang / 2;
    const float opi = 1.0f / acosf(-1.0f);
    for (int iray = warp; iray < nrays; iray += nwarps) {
        const int kface = hits[iray].kface;
        if (kface == 0) continue;
        const int icell = hits[iray].icell;
        const float mu2 = (float)__ldg(&cammu[iray]), phi2 = (float)__ldg(&camphi[iray]);
        float x[4], y[4], rad[4][NST];
        bool bad = false;
        for (int j = 0; j < 4; j++) {
            const int ip = cell_gp(S, icell, gf[kface - 1][j]);
            x[j] = pt_coord(S, ip, 1);
            y[j] = pt_coord(S, ip, 2);
            const int ibc = dev_bc_search(S.bcptr + S.maxnbc, S.nbotpts, ip);
            if (!ibc) { bad = true; break; }
            const float *parms = S.sfcgridparms + (size_t)S.nsfcpar * (ibc - 1);
            const float planck = __ldg(&parms[0]);
            float acc[NST];
#pragma unroll
            for (int k = 0; k < NST; k++) acc[k] = 0.0f;
            float reflect[16];
            // integrate over the incident discrete ordinates (shdomsub1.f:2647-2666)
            for (int jang = lane; jang < nh; jang += 32) {
                dev_surface_brdf(S.sfctype1, parms + 1, S.wavelen, mu2, phi2, __ldg(&S.ord_mu[jang]),
                                 __ldg(&S.ord_phi[jang]), NST, reflect);
                const float w = __ldg(&S.ord_w[jang]);
                const float *down = S.bcrad + (size_t)NST * (S.ntoppts + (ibc - 1) + (size_t)S.nbotpts * (jang + 1));
#pragma unroll
                for (int k1 = 0; k1 < NST; k1++) {
                    const float d = __ldg(&down[k1]);
#pragma unroll
                    for (int k = 0; k < NST; k++) acc[k] = acc[k] + w * reflect[k + 4 * k1] * d;
                }
                acc[0] = acc[0] + w * (1 - reflect[0]) * planck;
#pragma unroll
                for (int k = 1; k < NST; k++) acc[k] = acc[k] - w * reflect[k] * planck;
            }
#pragma unroll
            for (int k = 0; k < NST; k++) acc[k] = warp_sum(acc[k]);
            // reflection of the direct beam (shdomsub1.f:2636-2644)
            if (S.srctype != 'T') {
                dev_surface_brdf(S.sfctype1, parms + 1, S.wavelen, mu2, phi2, S.solarmu, S.solaraz, NST, reflect);
                const float df = __ldg(&S.dirflux[ip - 1]);
#pragma unroll
                for (int k = 0; k < NST; k++) acc[k] = acc[k] + opi * reflect[k] * df;
            }
#pragma unroll
            for (int k = 0; k < NST; k++) rad[j][k] = acc[k];
            if (S.sfcgridrad || S.srctype == 'T') rad[j][0] = dev_surface_emission(S, ibc, mu2, phi2) + acc[0];
        }
        if (bad) { if (lane == 0) set_err(err, 3, iray); continue; }
        if (lane == 0) finish_hit<NST, OUTA>(S, hits, iray, x, y, rad, out);
    }
}

// Ocean surfaces (SFCTYPE 'VO', NSTOKES=1): the direction-dependent part of ocean_brdf_sw (OceanGeom) is the same for
// the 4 face points of a hit and the parameter-dependent part (OceanPoint) is the same for all ordinates, so the
// ordinate loop is outermost: one OceanGeom per (ray, ordinate), four cheap dev_ocean_eval's.
template <typename OUTA>
__global__ void __launch_bounds__(256)
surface_ocean_kernel(DevState S, int nrays, const SurfHit *hits, const double *cammu, const double *camphi,
                     OUTA *out, RayErr *err)
{
    const int gf[6][4] = {{1,3,5,7},{2,4,6,8},{1,2,5,6},{3,4,7,8},{1,2,3,4},{5,6,7,8}};
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nh = S.nang / 2;
    const float opi = 1.0f / acosf(-1.0f);
    // incident-direction part of the BRDF geometry: once per block for the stored ordinates and the sun
    extern __shared__ __align__(16) unsigned char surf_sm[];
    OceanInc *inc_s = (OceanInc *)surf_sm;           // [nh + 1]
    for (int jang = threadIdx.x; jang <= nh; jang += blockDim.x) {
        if (jang < nh) dev_ocean_inc(-__ldg(&S.ord_mu[jang]), __ldg(&S.ord_phi[jang]), inc_s[jang]);
        else dev_ocean_inc(-S.solarmu, S.solaraz, inc_s[nh]);
    }
    __syncthreads();
    for (int iray = warp; iray < nrays; iray += nwarps) {
        const int kface = hits[iray].kface;
        if (kface == 0) continue;
        const int icell = hits[iray].icell;
        const float mu2 = (float)__ldg(&cammu[iray]), phi2 = (float)__ldg(&camphi[iray]);
        OceanView view;
        dev_ocean_view(mu2, view);
        float x[4], y[4], rad[4][1], planck[4], acc[4];
        int ibcs[4], ips[4];
        OceanPoint pt[4];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int ip = cell_gp(S, icell, gf[kface - 1][j]);
            ips[j] = ip;
            x[j] = pt_coord(S, ip, 1);
            y[j] = pt_coord(S, ip, 2);
            const int ibc = dev_bc_search(S.bcptr + S.maxnbc, S.nbotpts, ip);
            ibcs[j] = ibc;
            if (!ibc) { bad = true; continue; }
            const float *parms = S.sfcgridparms + (size_t)S.nsfcpar * (ibc - 1);
            planck[j] = __ldg(&parms[0]);
            // SURFACE_BRDF 'O': ocean_brdf_sw(REFPARMS(1), -1., REFPARMS(2), WAVELEN, -MU1, MU2, PHI1-PHI2, PHI1)
            dev_ocean_point(__ldg(&parms[1]), -1.f, __ldg(&parms[2]), S.wavelen, pt[j]);
            acc[j] = 0.0f;
        }
        if (bad) { if (lane == 0) set_err(err, 3, iray); continue; }
        for (int jang = lane; jang < nh; jang += 32) {
            const float phi1 = __ldg(&S.ord_phi[jang]);
            OceanGeom g;
            dev_ocean_pair(inc_s[jang], view, phi1 - phi2, g);
            const float w = __ldg(&S.ord_w[jang]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float r = dev_ocean_eval(pt[j], g);
                const float d = __ldg(&S.bcrad[S.ntoppts + (ibcs[j] - 1) + (size_t)S.nbotpts * (jang + 1)]);
                acc[j] = acc[j] + w * r * d;
                acc[j] = acc[j] + w * (1 - r) * planck[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc[j] = warp_sum(acc[j]);
        if (S.srctype != 'T') {
            OceanGeom g;
            dev_ocean_pair(inc_s[nh], view, S.solaraz - phi2, g);
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc[j] = acc[j] + opi * dev_ocean_eval(pt[j], g) * __ldg(&S.dirflux[ips[j] - 1]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            rad[j][0] = acc[j];
            if (S.sfcgridrad || S.srctype == 'T') rad[j][0] = dev_surface_emission(S, ibcs[j], mu2, phi2) + acc[j];
        }
        if (lane == 0) finish_hit<1, OUTA>(S, hits, iray, x, y, rad, out);
    }
}

cudaError_t launch_surface(const DevState &S, int nrays, const SurfHit *hits, const double *cammu,
                           const double *camphi, float *out, RayErr *err, cudaStream_t stream)
{
    if (nrays <= 0) return cudaSuccess;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long want = ((long)nrays + 7) / 8;
    const long cap = (long)nsm * 8;
    const int nb = (int)(want < cap ? want : cap);
    if (S.nstokes == 1 && S.sfctype1 == 'O')
        surface_ocean_kernel<float><<<nb, 256, (size_t)(S.nang / 2 + 1) * sizeof(OceanInc), stream>>>(S, nrays, hits, cammu, camphi, out, err);
    else if (S.nstokes == 1) surface_kernel<1, float><<<nb, 256, 0, stream>>>(S, nrays, hits, cammu, camphi, out, err);
    else surface_kernel<3, float><<<nb, 256, 0, stream>>>(S, nrays, hits, cammu, camphi, out, err);
    return cudaGetLastError();
}
