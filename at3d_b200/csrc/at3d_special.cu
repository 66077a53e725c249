// at3d_special.cu -- special-function tables on the device (sm_100a):
//   at3d_ylmall                  replaces YLMALL / YLMALL_UNPOL (src/polarized/shdomsub2.f:4244-4539)
//   at3d_precompute_phase_check  replaces PRECOMPUTE_PHASE_CHECK[_GRAD] (shdomsub4.f:2388-2585)
// Both produce the reference's own array layouts (Fortran order) in HOST memory.
#include "at3d_mem.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include "at3d_host.h"

// One thread per azimuthal order m: streams the Wigner recurrences over l and writes YR(:,j) in the
// reference layout YR[nstleg,nlm].  Same recurrences as warp_ylmall (at3d_device.cuh).
__global__ void ylmall_kernel(int transpose, float mu, float phi, int ml, int mm, int nstleg, float *yr)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > mm) return;
    const double x = (double)mu;
    const double pi = 3.14159265358979323846;
    const double fct = 1.0 / sqrt(2.0 * pi);
    double cosm = 1.0, sinm = 0.0;
    if (m > 0) { cosm = (double)cosf((float)m * phi); sinm = (double)sinf((float)m * phi); }
#define YR(i, j0) yr[((i) - 1) + (size_t)nstleg * (j0)]
    if (nstleg == 1) {
        double dprev = 0.0, dcur = (m == 0) ? 1.0 : dev_dm_m10_n0(x, m);
        for (int n = m; n <= ml; n++) {
            double t = sqrt(n + 0.5) * dcur;
            t = fct * t;
            YR(1, sh_index(n, m, mm)) = (float)((cosm - sinm) * t);
            YR(1, sh_index(n, -m, mm)) = (float)((cosm + sinm) * t);
            double dnext;
            if (m == 0) dnext = (n == 0) ? x : ((2 * n + 1) * x * dcur - n * dprev) / (n + 1);
            else dnext = ((2 * n + 1) * x * dcur - sqrt((double)(n * n - m * m)) * dprev)
                         / sqrt((double)((n + 1) * (n + 1) - m * m));
            dprev = dcur; dcur = dnext;
        }
        return;
    }
    const double sign = transpose ? -1.0 : 1.0;
    const double xp = x, xm = -x;
    const int n0 = m > 2 ? m : 2;
    double d0p = 0.0, d0c = 0.0, pp = 0.0, pc = 0.0, qp = 0.0, qc = 0.0;
    for (int n = 0; n <= ml; n++) {
        if (n < m) d0c = 0.0;
        else if (n == m) d0c = (m == 0) ? 1.0 : dev_dmm1_n0(xp, m, 0);
        if (n < n0) { pc = 0.0; qc = 0.0; }
        else if (n == n0 && ml >= 2) { pc = dev_dmm1_n0(xp, m, 2); qc = dev_dmm1_n0(xm, m, 2); }
        if (n >= m) {
            const double dm0 = sqrt(n + 0.5) * d0c;
            double dm2m = (((n + m) & 1) ? -1.0 : 1.0) * qc;
            const double dm2p = sqrt(n + 0.5) * pc;
            dm2m = sqrt(n + 0.5) * dm2m;
            const double p1 = fct * dm0;
            const double p2 = -0.5 * fct * (dm2p + dm2m);
            const double p3 = -0.5 * fct * (dm2p - dm2m);
            const int jp = sh_index(n, m, mm);
            if (m == 0) {
                YR(1, jp) = (float)p1;
                if (nstleg == 6) {
                    YR(2, jp) = (float)p2; YR(3, jp) = (float)p2; YR(4, jp) = (float)p1;
                    YR(5, jp) = (float)p3; YR(6, jp) = (float)p3;
                }
            } else {
                const int jn = sh_index(n, -m, mm);
                YR(1, jp) = (float)(p1 * cosm - p1 * sinm);
                YR(1, jn) = (float)(p1 * sinm + p1 * cosm);
                if (nstleg == 6) {
                    YR(2, jp) = (float)(p2 * cosm - p2 * sinm);
                    YR(3, jp) = (float)(p2 * cosm + p2 * sinm);
                    YR(4, jp) = (float)(p1 * cosm + p1 * sinm);
                    YR(5, jp) = (float)(p3 * cosm - sign * p3 * sinm);
                    YR(6, jp) = (float)(p3 * cosm + sign * p3 * sinm);
                    YR(2, jn) = (float)(p2 * sinm + p2 * cosm);
                    YR(3, jn) = (float)(p2 * sinm - p2 * cosm);
                    YR(4, jn) = (float)(p1 * sinm - p1 * cosm);
                    YR(5, jn) = (float)(p3 * sinm + sign * p3 * cosm);
                    YR(6, jn) = (float)(p3 * sinm - sign * p3 * cosm);
                }
            }
        }
        if (n >= m && n < ml) {
            double dnext;
            if (m == 0) {
                if (n == 0) dnext = xp;
                else {
                    const double fact1 = (double)(2 * n + 1) * xp / (double)(n + 1);
                    const double fact2 = (double)n / (double)(n + 1);
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
            d0p = d0c; d0c = dnext;
        }
        if (n >= n0 && n < ml) {
            const double factp = (double)(n * (n + 1)) * xp - (double)(2 * m);
            const double factm = (double)(n * (n + 1)) * xm - (double)(2 * m);
            double fact1 = 1.0 / sqrt((double)((n + 1) * (n + 1) - m * m));
            fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - 4));
            fact1 = fact1 * (double)(2 * n + 1) / (double)n;
            double fact2 = sqrt((double)(n * n - m * m)) * sqrt((double)(n * n - 4));
            fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - m * m));
            fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - 4));
            fact2 = fact2 * (double)(n + 1) / (double)n;
            const double pn = factp * fact1 * pc - fact2 * pp;
            const double qn = factm * fact1 * qc - fact2 * qp;
            pp = pc; pc = pn; qp = qc; qc = qn;
        }
    }
#undef YR
}

static void set_msg2(char *errmsg, const char *msg)
{
    if (errmsg) { strncpy(errmsg, msg, AT3D_ERRMSG_LEN - 1); errmsg[AT3D_ERRMSG_LEN - 1] = 0; }
}

extern "C" int at3d_ylmall(int transpose, float mu, float phi, int ml, int mm, int nstleg, float *yr,
                           char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!yr || ml < 0 || mm < 0 || mm > ml || !(nstleg == 1 || nstleg == 6)) { set_msg2(errmsg, "at3d_ylmall: bad argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg2(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    int nlm = 0;
    for (int l = 0; l <= ml; l++) nlm += 2 * (l < mm ? l : mm) + 1;
    float *d = nullptr;
    const size_t nb = (size_t)nstleg * nlm * sizeof(float);
    if (at3d_malloc((void **)&d, nb) != cudaSuccess) { set_msg2(errmsg, "at3d_ylmall: cudaMalloc failed"); return 4; }
    cudaMemset(d, 0, nb);
    ylmall_kernel<<<(mm + 32) / 32, 32>>>(transpose, mu, phi, ml, mm, nstleg, d);
    cudaError_t e = cudaMemcpy(yr, d, nb, cudaMemcpyDeviceToHost);
    at3d_free(d);
    if (e != cudaSuccess) { set_msg2(errmsg, cudaGetErrorString(e)); return 4; }
    return 0;
}

// PRECOMPUTE_PHASE_CHECK[_GRAD]: one thread per (scattering angle j, phase function iph).
// WIGNERFCT(x, nleg, 0, 0) and WIGNERFCT(x, nleg, 2, 0) recurrences (shdomsub2.f:4604-4645) are
// streamed in double; sums in the reference's order.
__global__ void phase_check_kernel(int nscatangle, int numphase, int nstphase, int nstokes, int nstleg,
                                   int nleg, const float *legen, float *phasetab, int negcheck,
                                   int scale_by_2l1, int *bad)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nscatangle * numphase) return;
    const int iph = t % numphase;      // 0-based
    const int j = t / numphase;        // 0-based
    const double pi = acos(-1.0);
    const double ofourpi = 1.0 / (4.0 * pi);
    const double x = cos(pi * (double)j / (nscatangle - 1));
    const float *lg = legen + (size_t)nstleg * (nleg + 1) * iph;
    // m=0,m1=0: Legendre polynomials
    double a1 = 0.0, dprev = 0.0, dcur = 1.0;
    for (int l = 0; l <= nleg; l++) {
        const double u = scale_by_2l1 ? (double)(lg[(size_t)nstleg * l] / (float)(2 * l + 1)) : (double)lg[(size_t)nstleg * l];
        a1 = a1 + (2.0 * l + 1.0) * u * dcur;
        double dnext;
        if (l == 0) dnext = x;
        else {
            const double fact1 = (double)(2 * l + 1) * x / (double)(l + 1);
            const double fact2 = (double)l / (double)(l + 1);
            dnext = fact1 * dcur - fact2 * dprev;
        }
        dprev = dcur; dcur = dnext;
    }
    if (negcheck && a1 <= 0.0) { atomicCAS(bad, 0, 1 + t); return; }
    phasetab[(size_t)nstphase * (iph + (size_t)numphase * j)] = (float)(a1 * ofourpi);
    if (nstokes > 1) {
        // m=2,m1=0
        double b1 = 0.0;
        dprev = 0.0; dcur = 0.0;
        for (int l = 0; l <= nleg; l++) {
            if (l == 2) dcur = dev_dmm1_n0(x, 2, 0);
            const double u = scale_by_2l1 ? (double)(lg[(size_t)nstleg * l + 4] / (float)(2 * l + 1)) : (double)lg[(size_t)nstleg * l + 4];
            b1 = b1 - (2.0 * l + 1.0) * u * dcur;
            if (l >= 2) {
                const int n = l;
                double fact1 = (double)(n * (n + 1)) * x - 0.0;
                fact1 = fact1 / sqrt((double)((n + 1) * (n + 1) - 4));
                fact1 = fact1 / sqrt((double)((n + 1) * (n + 1)));
                fact1 = fact1 * (double)(2 * n + 1) / (double)n;
                double fact2 = sqrt((double)(n * n - 4)) * sqrt((double)(n * n));
                fact2 = fact2 / sqrt((double)((n + 1) * (n + 1) - 4));
                fact2 = fact2 / sqrt((double)((n + 1) * (n + 1)));
                fact2 = fact2 * (double)(n + 1) / (double)n;
                const double dnext = fact1 * dcur - fact2 * dprev;
                dprev = dcur; dcur = dnext;
            }
        }
        phasetab[(size_t)nstphase * (iph + (size_t)numphase * j) + 1] = (float)(b1 * ofourpi);
    }
}

extern "C" int at3d_precompute_phase_check(int nscatangle, int numphase, int nstphase, int nstokes, int ml,
                                           int nlm, int nstleg, int nleg, const float *legen, float *phasetab,
                                           int deltam, int negcheck, int grad, char *errmsg)
{
    (void)ml; (void)nlm; (void)deltam;   // all DELTAM branches are identical (shdomsub4.f:2438-2445)
    if (errmsg) errmsg[0] = 0;
    if (!legen || !phasetab || nscatangle < 2 || numphase < 1) { set_msg2(errmsg, "at3d_precompute_phase_check: bad argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg2(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    float *dl = nullptr, *dt = nullptr; int *bad = nullptr;
    const size_t nl = (size_t)nstleg * (nleg + 1) * numphase, nt = (size_t)nstphase * numphase * nscatangle;
    int rc = 0;
    if (at3d_malloc((void **)&dl, nl * sizeof(float)) != cudaSuccess || at3d_malloc((void **)&dt, nt * sizeof(float)) != cudaSuccess ||
        at3d_malloc((void **)&bad, sizeof(int)) != cudaSuccess) { set_msg2(errmsg, "cudaMalloc failed"); rc = 4; }
    if (!rc) {
        cudaMemcpy(dl, legen, nl * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemset(dt, 0, nt * sizeof(float));
        cudaMemset(bad, 0, sizeof(int));
        const int n = nscatangle * numphase;
        phase_check_kernel<<<(n + 127) / 128, 128>>>(nscatangle, numphase, nstphase, nstokes, nstleg, nleg, dl, dt,
                                                      negcheck, grad ? 0 : 1, bad);
        int hbad = 0;
        if (cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { set_msg2(errmsg, "CUDA error in phase_check_kernel"); rc = 4; }
        else if (hbad) {
            if (errmsg) snprintf(errmsg, AT3D_ERRMSG_LEN, "PRECOMPUTE_PHASE_CHECK%s: negative phase function for tabulated phase function: IPH %d J %d",
                                 grad ? "_GRAD" : "", (hbad - 1) % numphase + 1, (hbad - 1) / numphase + 1);
            rc = 1;
        } else if (cudaMemcpy(phasetab, dt, nt * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { set_msg2(errmsg, "copy back failed"); rc = 4; }
    }
    at3d_free(dl); at3d_free(dt); at3d_free(bad);
    return rc;
}

// ------------------------------------------------------------------------------------------
// average_subpixel_rays (src/util.f90:484-518): ray -> pixel segmented sum over the sorted
// pixel_index (0-based pixel numbers).  One thread per pixel sums its contiguous run in order in
// double; like the reference the last ray is left out of the runs and added to the last pixel, and
// the first ray is taken to belong to pixel 0.
// ------------------------------------------------------------------------------------------
__global__ void average_subpixel_kernel(int npixels, int nrays, int nstokes, const float *ws,
                                        const int *pixel_index, float *obs)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npixels) return;
    // first ray index whose pixel number is >= p  (ray 0 counts as pixel 0)
    auto lower = [&](int v) {
        int lo = 0, hi = nrays;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            const int pv = mid == 0 ? 0 : pixel_index[mid];
            if (pv < v) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int r0 = lower(p);
    int r1 = lower(p + 1);
    if (r1 > nrays - 1) r1 = nrays - 1;          // the loop of the reference never consumes the last ray
    for (int k = 0; k < nstokes; k++) {
        double t = 0.0;
        for (int r = r0; r < r1; r++) t = t + ws[k + nstokes * (size_t)r];
        float o = (float)t;
        if (p == npixels - 1) o = o + ws[k + nstokes * (size_t)(nrays - 1)];
        obs[k + nstokes * (size_t)p] = o;
    }
}

extern "C" int at3d_average_subpixel_rays(int npixels, int nrays, int nstokes, const float *weighted_stokes,
                                          const int32_t *pixel_index, float *observables, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!weighted_stokes || !pixel_index || !observables || npixels < 1 || nrays < 1 || nstokes < 1) { set_msg2(errmsg, "at3d_average_subpixel_rays: bad argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg2(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    float *w = nullptr, *o = nullptr; int *pi = nullptr;
    int rc = 0;
    if (at3d_malloc((void **)&w, (size_t)nstokes * nrays * sizeof(float)) != cudaSuccess ||
        at3d_malloc((void **)&o, (size_t)nstokes * npixels * sizeof(float)) != cudaSuccess ||
        at3d_malloc((void **)&pi, (size_t)nrays * sizeof(int)) != cudaSuccess) { set_msg2(errmsg, "cudaMalloc failed"); rc = 4; }
    if (!rc) {
        cudaMemcpy(w, weighted_stokes, (size_t)nstokes * nrays * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(pi, pixel_index, (size_t)nrays * sizeof(int), cudaMemcpyHostToDevice);
        average_subpixel_kernel<<<(npixels + 127) / 128, 128>>>(npixels, nrays, nstokes, w, pi, o);
        if (cudaMemcpy(observables, o, (size_t)nstokes * npixels * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { set_msg2(errmsg, "CUDA error in average_subpixel_kernel"); rc = 4; }
    }
    at3d_free(w); at3d_free(o); at3d_free(pi);
    return rc;
}

// ------------------------------------------------------------------------------------------
// UPDATE_COSTFUNCTION (src/polarized/shdomsub4.f:13-91): cost and gradient update of one pixel from
// its ray gradient RAYGRAD_PIXEL[nstokes,maxpg,numder] (Jacobian path).  The per-Stokes weights are
// scalars computed on the host; the kernel is the axpy over maxpg*numder.
// ------------------------------------------------------------------------------------------
__global__ void update_cost_kernel(size_t n, int nstokes, const double *rg, double *gradout, double w0, double w1, double w2, double w3)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double g = gradout[t] + w0 * rg[(size_t)nstokes * t];
    if (nstokes > 1) g = g + w1 * rg[(size_t)nstokes * t + 1];
    if (nstokes > 2) g = g + w2 * rg[(size_t)nstokes * t + 2];
    if (nstokes > 3) g = g + w3 * rg[(size_t)nstokes * t + 3];
    gradout[t] = g;
}

extern "C" int at3d_update_costfunction(const double *stokesout, const double *raygrad_pixel,
                                        double *gradout, double *cost, const double *uncertainties,
                                        int costfunc_ll, int nstokes, int maxpg, int numder,
                                        const double *measurement, int nuncertainty, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!stokesout || !raygrad_pixel || !gradout || !cost || !uncertainties || !measurement) { set_msg2(errmsg, "at3d_update_costfunction: null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg2(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    const int nu = nuncertainty;
#define UNC(a, b) uncertainties[((a) - 1) + nu * ((b) - 1)]
    double w[4] = {0.0, 0.0, 0.0, 0.0};
    if (!costfunc_ll) {
        if (nstokes > 4 || nu < nstokes) { set_msg2(errmsg, "at3d_update_costfunction: NSTOKES<=4 and NUNCERTAINTY>=NSTOKES required"); return 3; }
        for (int i = 1; i <= nstokes; i++) {
            const double pe = stokesout[i - 1] - measurement[i - 1];
            for (int j = 1; j <= nstokes; j++) {
                cost[0] = cost[0] + 0.5 * UNC(i, j) * (pe * pe);
                w[i - 1] = w[i - 1] + UNC(i, j) * pe;      // gradient weight of RAYGRAD_PIXEL(i,:,:)
            }
        }
    } else {
        const double raderror = log(stokesout[0]) - log(measurement[0]);
        cost[0] = cost[0] + 0.5 * (raderror * raderror * UNC(1, 1));
        w[0] = raderror * UNC(1, 1) / stokesout[0];
        if (nstokes > 1) {
            const double q = stokesout[1], u = stokesout[2];
            const double dolp1 = sqrt(q * q + u * u) / stokesout[0];
            const double dolp2 = sqrt(measurement[1] * measurement[1] + measurement[2] * measurement[2]) / measurement[0];
            const double dolperr = log(dolp1) - log(dolp2);
            cost[0] = cost[0] + 0.5 * (dolperr * dolperr * UNC(2, 2));
            w[1] = dolperr * UNC(2, 2) * q / (q * q + u * u);
            w[2] = dolperr * UNC(2, 2) * u / (q * q + u * u);
        }
    }
#undef UNC
    const size_t n = (size_t)maxpg * numder;
    double *rg = nullptr, *go = nullptr;
    int rc = 0;
    if (at3d_malloc((void **)&rg, n * nstokes * sizeof(double)) != cudaSuccess || at3d_malloc((void **)&go, n * sizeof(double)) != cudaSuccess) { set_msg2(errmsg, "cudaMalloc failed"); rc = 4; }
    if (!rc) {
        cudaMemcpy(rg, raygrad_pixel, n * nstokes * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(go, gradout, n * sizeof(double), cudaMemcpyHostToDevice);
        update_cost_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, nstokes, rg, go, w[0], w[1], w[2], w[3]);
        if (cudaMemcpy(gradout, go, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { set_msg2(errmsg, "CUDA error in update_cost_kernel"); rc = 4; }
    }
    at3d_free(rg); at3d_free(go);
    return rc;
}
