// at3d_transform.cu -- spherical-harmonic <-> discrete-ordinate transforms on the device (sm_100a).
// Replaces SH_TO_DO[_UNPOL] / DO_TO_SH[_UNPOL] (src/polarized/shdomsub1.f:2789-3260 of the AT3D reference) with the
// coefficient tables of MAKE_SH_DO_COEF / PLMALL (src/polarized/shdomsub2.f:1146-1220, 4650-4752), for ALL zenith
// angles in one launch (the reference transforms one zenith angle at a time inside PATH_INTEGRATION to save memory).
//
// The transform is kept in the reference's factorised form -- Legendre/Wigner sum per azimuthal mode m
// (NLM x NMU multiply-adds per point) followed by the azimuthal Fourier sum ((2MM+1) x NANG) -- which needs 6.5x
// fewer flops than the dense [NPTS x NLM].[NLM x NANG] product (15 k instead of 90 k multiply-adds per point at
// NMU=16, NPHI=32).  Its inner dimensions (<= 16 degrees l per mode, <= 31 modes per ordinate) are too short and too
// ragged for tcgen05 tiles, and TF32 inputs would not meet the 1e-4 parity bar without the 3x split; the kernels use
// FP32 FMA with both stages fused through shared memory, so HBM sees each SH block and each ordinate value once.
// A block owns a tile of 32 grid points; lane = point, so every shared-memory access is conflict free (points are
// the fastest index) and the discrete-ordinate field DOFIELD(NPTS, NSTOKES, NANG) is written/read in 128-byte rows.
#include "at3d_mem.h"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "at3d_host.h"

#define TR_TP 32          // points per tile (= warp width)
#ifndef TR_THREADS
#define TR_THREADS 1024
#endif
#define TR_WARPS (TR_THREADS / 32)

static void set_msg(char *errmsg, const char *fmt, ...)
{
    if (!errmsg) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, AT3D_ERRMSG_LEN, fmt, ap);
    va_end(ap);
}

struct TrArgs {
    int npts, nst, nstleg, ml, mm, nlm, nmu, nang, ntask;
    const int *shptr;         // [npts+1] offsets into the SH array
    const float *sh;          // SH array (NSTOKES, *) interleaved  (input of sh_to_do / output of do_to_sh)
    float *sh_out;
    float *dofield;           // DOFIELD(NPTS, NSTOKES, NANG)
    const float *cmu;         // sh_to_do: CMU1 as [comp][j][16] (zenith angle fastest); do_to_sh: CMU2 as [comp][m][imu][16] (degree fastest)
    const float *az;          // sh_to_do: per zenith angle [m][NPHI0 rounded up to 8]; do_to_sh: [iang][32] (m fastest, x DELPHI)
    const int *azoff;         // [nmu] offset of each zenith angle's block in az (sh_to_do)
    const int2 *tasks;        // sh_to_do stage B: (imu, first azimuth of a group of 8)
    const int *me_of;         // [nmu]: min(NPHI0/2-1, MM) >= 0
    const int *ang0;          // [nmu+1] first ordinate of each zenith angle
    int azsize;
};

// PLMALL (shdomsub2.f:4650-4752) for every zenith angle: prc[(q-1) + 6*((j-1) + nlm*imu)].  Thread = (imu, m>=0).
__global__ void plmall_kernel(int nmu, const float *mu, int ml, int mm, int nlm, int transpose, float *prc)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nmu * (mm + 1)) return;
    const int imu = t / (mm + 1), m = t % (mm + 1);
    const double xp = (double)mu[imu], xm = -xp;
    const double pi = 3.14159265358979323846;
    const double fct = 1.0 / sqrt(2.0 * pi);
    const double sign = transpose ? -1.0 : 1.0;
    float *P = prc + (size_t)6 * nlm * imu;
    const int n0 = m > 2 ? m : 2;
    // WIGNERFCT02P2M_NORMALIZED recurrences (shdomsub2.f:4363-4448), streamed over n
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
            P[0 + 6 * jp] = (float)p1; P[1 + 6 * jp] = (float)p2; P[2 + 6 * jp] = (float)p2;
            P[3 + 6 * jp] = (float)p1; P[4 + 6 * jp] = (float)p3; P[5 + 6 * jp] = (float)p3;
            if (m > 0) {
                const int jn = sh_index(n, -m, mm);
                P[0 + 6 * jn] = (float)p1; P[1 + 6 * jn] = (float)p2; P[2 + 6 * jn] = (float)(-p2);
                P[3 + 6 * jn] = (float)(-p1); P[4 + 6 * jn] = (float)(sign * p3); P[5 + 6 * jn] = (float)(-sign * p3);
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
}

// re-pack PRC(6,NLM) per angle: layout 0 = [comp][j][16] for SH_TO_DO, layout 1 = [comp][m][imu][16] (x WTMU) for
// DO_TO_SH; unused slots are zero
__global__ void pack_cmu_kernel(int nmu, int nlm, int ncomp, int ml, int mm, int layout, const float *prc,
                                const float *wtmu, float *cmu)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nmu * nlm * ncomp) return;
    const int j = t % nlm, imu = (t / nlm) % nmu, c = t / (nlm * nmu);
    float v = prc[c + 6 * (j + (size_t)nlm * imu)];
    if (layout == 0) {
        cmu[imu + 16 * (j + (size_t)nlm * c)] = v;
    } else {
        // l and m of j (inverse of sh_index)
        int l = 0, m = 0, jj = 0;
        for (l = 0; l <= ml; l++) {
            const int me = l < mm ? l : mm;
            if (j < jj + 2 * me + 1) { m = j - jj - me; break; }
            jj += 2 * me + 1;
        }
        const int am = m < 0 ? -m : m, nm = 2 * mm + 1;
        cmu[(l - am) + 16 * (imu + (size_t)nmu * ((m + mm) + (size_t)nm * c))] = v * wtmu[imu];
    }
}

// terms of the Stokes coupling (SH_TO_DO: shdomsub1.f:2853-2866, DO_TO_SH: :3144-3153), 0-based:
// sh_to_do: discrete-ordinate plane n is built from (table component, SH plane) terms;
// do_to_sh: SH plane o is built from (table component, discrete-ordinate plane) terms
__device__ __forceinline__ int tr_nterms(int n) { return n == 0 ? 1 : 2; }
__device__ __forceinline__ void tr_term(int n, int t, int &comp, int &plane)
{
    if (n == 0) { comp = 0; plane = 0; }
    else if (n == 1) { comp = t == 0 ? 1 : 4; plane = t == 0 ? 1 : 2; }     // CMU(2), CMU(5)
    else { comp = t == 0 ? 5 : 2; plane = t == 0 ? 1 : 2; }                 // CMU(6), CMU(3)
}

// SUMCS <-> SUMUV (shdomsub1.f:2871-2890 and :3121-3136), in place on uv_s[imu][m][point]; stokes3 = third Stokes plane
__device__ __forceinline__ void tr_combine(float *uv_s, int nmu, int mm, bool to_cs, bool stokes3, int warp, int lane)
{
    const int nm = 2 * mm + 1;
    for (int task = warp; task < nmu * mm; task += TR_WARPS) {
        const int imu = task / mm, m = task % mm + 1;
        float *pu = &uv_s[(imu * nm + (mm + m)) * TR_TP + lane], *nu = &uv_s[(imu * nm + (mm - m)) * TR_TP + lane];
        const float vp = *pu, vn = *nu;
        if (to_cs) {
            if (stokes3) { *nu = vp - vn; *pu = vp + vn; }
            else { *pu = vp + vn; *nu = vn - vp; }
        } else {
            if (stokes3) { *pu = vp + vn; *nu = vp - vn; }
            else { *pu = vp - vn; *nu = vp + vn; }
        }
    }
}

// SH_TO_DO for all ordinates.  Shared memory: sh_s[nlm][33] | cmu_s[nlm][16] | uv_s[nmu][nm][32] | az_s[azsize]
__global__ void __launch_bounds__(TR_THREADS)
sh_to_do_kernel(TrArgs a, int ntiles)
{
    extern __shared__ __align__(16) float tr_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mm = a.mm, nm = 2 * mm + 1, nlm = a.nlm, nmu = a.nmu;
    float *sh_s = tr_sm, *cmu_s = sh_s + (size_t)nlm * 33, *uv_s = cmu_s + (size_t)nlm * 16;
    float *az_s = uv_s + (size_t)nmu * nm * TR_TP;
    for (int i = threadIdx.x; i < a.azsize; i += TR_THREADS) az_s[i] = a.az[i];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * TR_TP;
        for (int n = 0; n < a.nst; n++) {
            __syncthreads();
            for (int i = threadIdx.x; i < nmu * nm * TR_TP; i += TR_THREADS) uv_s[i] = 0.0f;
            for (int t = 0; t < tr_nterms(n); t++) {
                int comp, plane;
                tr_term(n, t, comp, plane);
                __syncthreads();
                // stage the SH plane of the tile, transposed to [j][point], and the coefficient table
                for (int q = warp; q < TR_TP; q += TR_WARPS) {
                    const int p = p0 + q;
                    int is = 0, ns = 0;
                    if (p < a.npts) { is = a.shptr[p]; ns = a.shptr[p + 1] - is; }
                    for (int j = lane; j < nlm; j += 32)
                        sh_s[j * 33 + q] = j < ns ? __ldg(&a.sh[plane + (size_t)a.nst * (is + j)]) : 0.0f;
                }
                for (int i = threadIdx.x; i < nlm * 16; i += TR_THREADS) cmu_s[i] = __ldg(&a.cmu[i + (size_t)nlm * 16 * comp]);
                __syncthreads();
                // stage A: SUMUV(imu, m) += CMU1(comp, j, imu) * SH(plane, j) over the degrees l of mode m;
                // 16 zenith angles per SH value: 1 + 4 shared-memory loads for 16 FMAs
                for (int mi = warp; mi < nm; mi += TR_WARPS) {
                    const int m = mi - mm, am = m < 0 ? -m : m;
                    float acc[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) acc[i] = 0.0f;
                    for (int l = am; l <= a.ml; l++) {
                        const int j = sh_index(l, m, mm);
                        const float v = sh_s[j * 33 + lane];
                        const float4 *c4 = (const float4 *)(cmu_s + j * 16);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float4 c = c4[q];
                            acc[4 * q] = fmaf(c.x, v, acc[4 * q]); acc[4 * q + 1] = fmaf(c.y, v, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(c.z, v, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(c.w, v, acc[4 * q + 3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (i < nmu) uv_s[(i * nm + mi) * TR_TP + lane] += acc[i];
                }
            }
            __syncthreads();
            tr_combine(uv_s, nmu, mm, true, n == 2, warp, lane);
            __syncthreads();
            // stage B: azimuthal sums for 8 ordinates of one zenith angle at a time: 1 + 2 loads for 8 FMAs;
            // every ordinate is a 128-byte row of DOFIELD
            for (int task = warp; task < a.ntask; task += TR_WARPS) {
                const int2 tk = a.tasks[task];
                const int imu = tk.x, k0 = tk.y, me = a.me_of[imu];
                const int nphi0 = a.ang0[imu + 1] - a.ang0[imu], np8 = (nphi0 + 7) & ~7;
                const float *azb = az_s + a.azoff[imu] + k0;
                const float *cs = uv_s + (size_t)imu * nm * TR_TP + lane;
                float acc[8];
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = 0.0f;
                for (int mi = mm - me; mi <= mm + me; mi++) {
                    const float c = cs[mi * TR_TP];
                    const float4 a0 = *(const float4 *)(azb + mi * np8), a1 = *(const float4 *)(azb + mi * np8 + 4);
                    acc[0] = fmaf(a0.x, c, acc[0]); acc[1] = fmaf(a0.y, c, acc[1]);
                    acc[2] = fmaf(a0.z, c, acc[2]); acc[3] = fmaf(a0.w, c, acc[3]);
                    acc[4] = fmaf(a1.x, c, acc[4]); acc[5] = fmaf(a1.y, c, acc[5]);
                    acc[6] = fmaf(a1.z, c, acc[6]); acc[7] = fmaf(a1.w, c, acc[7]);
                }
                if (p0 + lane < a.npts) {
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (k0 + k < nphi0)
                            a.dofield[(p0 + lane) + (size_t)a.npts * (n + (size_t)a.nst * (a.ang0[imu] + k0 + k))] = acc[k];
                }
            }
        }
    }
}

// ---- NSTOKES=1 variant of sh_to_do_kernel: the SH tile of the NEXT tile is fetched with cp.async into a second buffer
// while the current one is transformed, the (single) coefficient table is loaded once per block, and SUMUV is written,
// not accumulated.  Shared memory: sh_s[2][nlm][33] | cmu_s[nlm][16] | uv_s[nmu][nm][32] | az_s[azsize]
__device__ __forceinline__ void tr_cp_async4(float *smem, const float *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}

__device__ __forceinline__ void tr_stage_tile_async(const TrArgs &a, int tile, float *buf, int warp, int lane)
{
    const int p0 = tile * TR_TP;
    for (int q = warp; q < TR_TP; q += TR_WARPS) {
        const int p = p0 + q;
        int is = 0, ns = 0;
        if (p < a.npts) { is = a.shptr[p]; ns = a.shptr[p + 1] - is; }
        for (int j = lane; j < a.nlm; j += 32) {
            if (j < ns) tr_cp_async4(&buf[j * 33 + q], &a.sh[is + j]);
            else buf[j * 33 + q] = 0.0f;
        }
    }
    asm volatile("cp.async.commit_group;");
}

__global__ void __launch_bounds__(TR_THREADS)
sh_to_do_kernel_s1(TrArgs a, int ntiles)
{
    extern __shared__ __align__(16) float tr_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mm = a.mm, nm = 2 * mm + 1, nlm = a.nlm, nmu = a.nmu;
    float *shb0 = tr_sm, *shb1 = shb0 + (size_t)nlm * 33, *cmu_s = shb1 + (size_t)nlm * 33;
    float *uv_s = cmu_s + (size_t)nlm * 16, *az_s = uv_s + (size_t)nmu * nm * TR_TP;
    for (int i = threadIdx.x; i < a.azsize; i += TR_THREADS) az_s[i] = a.az[i];
    for (int i = threadIdx.x; i < nlm * 16; i += TR_THREADS) cmu_s[i] = __ldg(&a.cmu[i]);
    int cur = 0;
    if ((int)blockIdx.x < ntiles) tr_stage_tile_async(a, blockIdx.x, shb0, warp, lane);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * TR_TP;
        float *sh_s = cur ? shb1 : shb0;
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();                       // the tile is in shared memory; stage B of the previous tile is finished
        if (tile + (int)gridDim.x < ntiles) tr_stage_tile_async(a, tile + gridDim.x, cur ? shb0 : shb1, warp, lane);
        // stage A (see sh_to_do_kernel)
        for (int mi = warp; mi < nm; mi += TR_WARPS) {
            const int m = mi - mm, am = m < 0 ? -m : m;
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = 0.0f;
            for (int l = am; l <= a.ml; l++) {
                const int j = sh_index(l, m, mm);
                const float v = sh_s[j * 33 + lane];
                const float4 *c4 = (const float4 *)(cmu_s + j * 16);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 c = c4[q];
                    acc[4 * q] = fmaf(c.x, v, acc[4 * q]); acc[4 * q + 1] = fmaf(c.y, v, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(c.z, v, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(c.w, v, acc[4 * q + 3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (i < nmu) uv_s[(i * nm + mi) * TR_TP + lane] = acc[i];
        }
        __syncthreads();
        tr_combine(uv_s, nmu, mm, true, false, warp, lane);
        __syncthreads();
        // stage B
        for (int task = warp; task < a.ntask; task += TR_WARPS) {
            const int2 tk = a.tasks[task];
            const int imu = tk.x, k0 = tk.y, me = a.me_of[imu];
            const int nphi0 = a.ang0[imu + 1] - a.ang0[imu], np8 = (nphi0 + 7) & ~7;
            const float *azb = az_s + a.azoff[imu] + k0;
            const float *cs = uv_s + (size_t)imu * nm * TR_TP + lane;
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = 0.0f;
            for (int mi = mm - me; mi <= mm + me; mi++) {
                const float c = cs[mi * TR_TP];
                const float4 a0 = *(const float4 *)(azb + mi * np8), a1 = *(const float4 *)(azb + mi * np8 + 4);
                acc[0] = fmaf(a0.x, c, acc[0]); acc[1] = fmaf(a0.y, c, acc[1]);
                acc[2] = fmaf(a0.z, c, acc[2]); acc[3] = fmaf(a0.w, c, acc[3]);
                acc[4] = fmaf(a1.x, c, acc[4]); acc[5] = fmaf(a1.y, c, acc[5]);
                acc[6] = fmaf(a1.z, c, acc[6]); acc[7] = fmaf(a1.w, c, acc[7]);
            }
            if (p0 + lane < a.npts) {
                float *dst = a.dofield + (p0 + lane) + (size_t)a.npts * (a.ang0[imu] + k0);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k0 + k < nphi0) dst[(size_t)a.npts * k] = acc[k];
            }
        }
        cur ^= 1;
    }
}

// DO_TO_SH summed over all zenith angles (OUTDATA is set, not accumulated).
// Shared memory: sh_s[nlm][33] | cmu_s[nm][nmu][16] | uv_s[nmu][nm][32] | in_s[nang][32] | az_s[nang][32]
__global__ void __launch_bounds__(TR_THREADS)
do_to_sh_kernel(TrArgs a, int ntiles)
{
    extern __shared__ __align__(16) float tr_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mm = a.mm, nm = 2 * mm + 1, nlm = a.nlm, nmu = a.nmu;
    float *sh_s = tr_sm, *cmu_s = sh_s + (size_t)nlm * 33, *uv_s = cmu_s + (size_t)nm * nmu * 16;
    float *in_s = uv_s + (size_t)nmu * nm * TR_TP, *az_s = in_s + (size_t)a.nang * TR_TP;
    for (int i = threadIdx.x; i < a.nang * 32; i += TR_THREADS) az_s[i] = a.az[i];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * TR_TP;
        for (int o = 0; o < a.nst; o++) {
            __syncthreads();
            for (int i = threadIdx.x; i < nlm * 33; i += TR_THREADS) sh_s[i] = 0.0f;
            for (int t = 0; t < tr_nterms(o); t++) {
                int comp, n;
                tr_term(o, t, comp, n);
                if (o == 2) comp = t == 0 ? 5 : 2;       // U <- CMU2(6)*SUMUV(2) + CMU2(3)*SUMUV(3)
                if (o == 1) comp = t == 0 ? 1 : 4;       // Q <- CMU2(2)*SUMUV(2) + CMU2(5)*SUMUV(3)
                __syncthreads();
                for (int i = threadIdx.x; i < a.nang * TR_TP; i += TR_THREADS) {
                    const int iang = i / TR_TP, q = i % TR_TP;
                    in_s[i] = (p0 + q < a.npts) ? __ldg(&a.dofield[(p0 + q) + (size_t)a.npts * (n + (size_t)a.nst * iang)]) : 0.0f;
                }
                for (int i = threadIdx.x; i < nm * nmu * 16; i += TR_THREADS)
                    cmu_s[i] = __ldg(&a.cmu[i + (size_t)nm * nmu * 16 * comp]);
                __syncthreads();
                // stage B': SUMCS(m) = sum over the azimuths of CPHI2 * INDATA for 8 modes at a time (az carries DELPHI)
                for (int task = warp; task < nmu * 4; task += TR_WARPS) {
                    const int imu = task >> 2, mi0 = (task & 3) * 8, me = a.me_of[imu];
                    float acc[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[k] = 0.0f;
                    for (int iang = a.ang0[imu]; iang < a.ang0[imu + 1]; iang++) {
                        const float x = in_s[iang * TR_TP + lane];
                        const float4 a0 = *(const float4 *)(az_s + iang * 32 + mi0), a1 = *(const float4 *)(az_s + iang * 32 + mi0 + 4);
                        acc[0] = fmaf(a0.x, x, acc[0]); acc[1] = fmaf(a0.y, x, acc[1]);
                        acc[2] = fmaf(a0.z, x, acc[2]); acc[3] = fmaf(a0.w, x, acc[3]);
                        acc[4] = fmaf(a1.x, x, acc[4]); acc[5] = fmaf(a1.y, x, acc[5]);
                        acc[6] = fmaf(a1.z, x, acc[6]); acc[7] = fmaf(a1.w, x, acc[7]);
                    }
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int mi = mi0 + k, m = mi - mm;
                        if (mi < nm) uv_s[(imu * nm + mi) * TR_TP + lane] = (m >= -me && m <= me) ? acc[k] : 0.0f;
                    }
                }
                __syncthreads();
                tr_combine(uv_s, nmu, mm, false, n == 2, warp, lane);
                __syncthreads();
                // stage A': OUTDATA(j) += CMU2(comp, imu, j) * SUMUV(m(j)) summed over the zenith angles, all degrees l of
                // one mode m per task: 1 + 4 loads for 16 FMAs
                for (int mi = warp; mi < nm; mi += TR_WARPS) {
                    const int m = mi - mm, am = m < 0 ? -m : m;
                    float acc[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) acc[i] = 0.0f;
                    for (int imu = 0; imu < nmu; imu++) {
                        const float u = uv_s[(imu * nm + mi) * TR_TP + lane];
                        const float4 *c4 = (const float4 *)(cmu_s + (mi * nmu + imu) * 16);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float4 c = c4[q];
                            acc[4 * q] = fmaf(c.x, u, acc[4 * q]); acc[4 * q + 1] = fmaf(c.y, u, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(c.z, u, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(c.w, u, acc[4 * q + 3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (am + i <= a.ml) sh_s[sh_index(am + i, m, mm) * 33 + lane] += acc[i];
                }
            }
            __syncthreads();
            // write the truncated SH blocks (RSHPTR), lanes over j
            for (int q = warp; q < TR_TP; q += TR_WARPS) {
                const int p = p0 + q;
                if (p >= a.npts) continue;
                const int is = a.shptr[p], ns = a.shptr[p + 1] - is;
                for (int j = lane; j < ns; j += 32) a.sh_out[o + (size_t)a.nst * (is + j)] = sh_s[j * 33 + q];
            }
        }
    }
}

// ---- host side: a plan holds the coefficient tables of one angular resolution on the device ----
struct TrPlan {
    std::vector<void *> ptrs;
    TrArgs fwd, bwd;            // table pointers of the two directions (point arrays are filled per launch)
    size_t smem_fwd = 0, smem_bwd = 0;
    int nang = 0, nsm = 148;
    // tensor-core SH_TO_DO (sh_to_do_tc_kernel): pre-split, pre-swizzled basis tiles; 0 chunks = not available
    const unsigned char *tc_b = nullptr;
    int tc_kch = 0, tc_n1 = 0, tc_n2 = 0 /*ordinates of the two halves*/, tc_ring = 0;
    const unsigned char *tcb_b = nullptr;      // DO_TO_SH (do_to_sh_tc_kernel)
    int tcb_kch = 0, tcb_nn = 0, tcb_ring = 0;
    size_t tcb_smem = 0;
    size_t tc_smem = 0;
    ~TrPlan() { for (void *p : ptrs) at3d_free(p); }
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

static int tr_tc_build(TrPlan *P, char *errmsg);
static int tr_tc_build_back(TrPlan *P, char *errmsg);
void tr_plan_destroy(TrPlan *p) { delete p; }
int tr_plan_nang(const TrPlan *p) { return p->nang; }

// builds CMU1/CMU2 (PLMALL on the device), the azimuthal tables of both directions and the stage-B task list
int tr_plan_create(int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max, const int32_t *nphi0,
                   const float *mu, const float *phi, const float *wtmu, TrPlan **out, char *errmsg)
{
    *out = nullptr;
    if (!(nstokes == 1 || nstokes == 3) || (nstokes == 1) != (nstleg == 1)) { set_msg(errmsg, "NSTOKES must be 1 (NSTLEG=1) or 3 (NSTLEG=6)"); return 3; }
    if (nlm != (2 * mm + 1) * (ml + 1) - mm * (mm + 1)) { set_msg(errmsg, "inconsistent NLM/ML/MM"); return 1; }
    if (nmu > 16 || ml > 15) { set_msg(errmsg, "the tiled SH/DO transforms support NMU <= 16 (ML <= 15)"); return 3; }
    const int nm = 2 * mm + 1;
    int nang = 0, azsize = 0;
    std::vector<int> me_of(nmu), ang0(nmu + 1), azoff(nmu);
    std::vector<int2> tasks;
    for (int i = 0; i < nmu; i++) {
        ang0[i] = nang;
        nang += nphi0[i];
        int me = nphi0[i] / 2 - 1; if (me > mm) me = mm; if (me < 0) me = 0;
        me_of[i] = me;
        azoff[i] = azsize;
        azsize += nm * ((nphi0[i] + 7) & ~7);
        for (int k0 = 0; k0 < nphi0[i]; k0 += 8) tasks.push_back(make_int2(i, k0));
    }
    ang0[nmu] = nang;
    // azimuthal basis.  FFTFLAG (MAKE_ANGLE_SET, shdomsub2.f:1131): the reference uses FFTPACK on the exact angles
    // 2 pi k/N there, and REAL COS(M*PHI(I,K)) tables otherwise; DO_TO_SH carries DELPHI = WTDO/WTMU
    std::vector<float> azf((size_t)azsize, 0.0f), azb((size_t)nang * 32, 0.0f);
    const int mmax = nphi0max / 2 - 1 > 0 ? nphi0max / 2 - 1 : 0;
    for (int i = 0, ia = 0; i < nmu; i++) {
        const bool fft = nphi0[i] > 14 || mmax > 15;
        const float delphi = 2.0f * acosf(-1.0f) / nphi0[i];
        const int np8 = (nphi0[i] + 7) & ~7;
        for (int k = 0; k < nphi0[i]; k++, ia++)
            for (int m = -mm; m <= mm; m++) {
                double v;
                if (m == 0) v = 1.0;
                else if (fft) {
                    const double ang = 2.0 * acos(-1.0) * (double)((abs(m) * k) % nphi0[i]) / (double)nphi0[i];
                    v = m > 0 ? cos(ang) : sin(ang);
                } else {
                    const float ph = phi[i + (size_t)nmu * k];
                    v = m > 0 ? (double)cosf(m * ph) : (double)sinf(-m * ph);
                }
                azf[(size_t)azoff[i] + (size_t)(m + mm) * np8 + k] = (float)v;
                azb[(size_t)ia * 32 + (m + mm)] = (float)v * delphi;
            }
    }
    TrPlan *P = new TrPlan();
    const int ncomp = nstleg;
    float *mu_d = P->up(mu, nmu), *wt_d = P->up(wtmu, nmu);
    const size_t n1 = (size_t)ncomp * nlm * 16, n2 = (size_t)ncomp * nm * nmu * 16;
    float *prc = P->alloc<float>((size_t)6 * nlm * nmu), *cmu1 = P->alloc<float>(n1), *cmu2 = P->alloc<float>(n2);
    TrArgs a;
    memset(&a, 0, sizeof(a));
    a.nst = nstokes; a.nstleg = nstleg; a.ml = ml; a.mm = mm; a.nlm = nlm; a.nmu = nmu; a.nang = nang;
    a.ntask = (int)tasks.size();
    a.azoff = P->up(azoff.data(), azoff.size());
    a.tasks = P->up(tasks.data(), tasks.size());
    a.me_of = P->up(me_of.data(), me_of.size()); a.ang0 = P->up(ang0.data(), ang0.size());
    const float *azf_d = P->up(azf.data(), azf.size()), *azb_d = P->up(azb.data(), azb.size());
    if (!mu_d || !wt_d || !prc || !cmu1 || !cmu2 || !a.azoff || !a.tasks || !a.me_of || !a.ang0 || !azf_d || !azb_d) {
        delete P; set_msg(errmsg, "device allocation failure"); return 4;
    }
    cudaMemset(cmu1, 0, n1 * sizeof(float)); cudaMemset(cmu2, 0, n2 * sizeof(float));
    for (int dir = 0; dir < 2; dir++) {
        cudaMemset(prc, 0, (size_t)6 * nlm * nmu * sizeof(float));
        plmall_kernel<<<(nmu * (mm + 1) + 127) / 128, 128>>>(nmu, mu_d, ml, mm, nlm, dir, prc);
        pack_cmu_kernel<<<(nmu * nlm * ncomp + 255) / 256, 256>>>(nmu, nlm, ncomp, ml, mm, dir, prc, wt_d, dir ? cmu2 : cmu1);
    }
    P->fwd = a; P->fwd.cmu = cmu1; P->fwd.az = azf_d; P->fwd.azsize = azsize;
    P->bwd = a; P->bwd.cmu = cmu2; P->bwd.az = azb_d; P->bwd.azsize = nang * 32;
    P->smem_fwd = ((size_t)nlm * 33 + (size_t)nlm * 16 + (size_t)nmu * nm * TR_TP + (size_t)azsize) * sizeof(float);
    P->smem_bwd = ((size_t)nlm * 33 + (size_t)nm * nmu * 16 + (size_t)nmu * nm * TR_TP + (size_t)2 * nang * TR_TP) * sizeof(float);
    if (P->smem_fwd + (nstokes == 1 ? (size_t)nlm * 33 * sizeof(float) : 0) > 227 * 1024 || P->smem_bwd > 227 * 1024) {
        delete P; set_msg(errmsg, "angular resolution too high for the shared-memory tiles"); return 3;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&P->nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(sh_to_do_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_fwd);
    cudaFuncSetAttribute(sh_to_do_kernel_s1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(P->smem_fwd + (size_t)nlm * 33 * sizeof(float)));
    cudaFuncSetAttribute(do_to_sh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_bwd);
    P->nang = nang;
    if (cudaDeviceSynchronize() != cudaSuccess) { delete P; set_msg(errmsg, "CUDA error building the SH/DO tables"); return 4; }
    {
        int rc = tr_tc_build(P, errmsg);
        if (!rc) rc = tr_tc_build_back(P, errmsg);
        if (rc) { delete P; return rc; }
    }
    *out = P;
    return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// SH_TO_DO on the 5th-generation tensor cores (NSTOKES=1): DOFIELD[128 points x NANG] = SH[128 x NLM] . Y[NLM x NANG] as
// one tcgen05.mma chain per tile of 128 grid points, accumulators in tensor memory, 3xTF32 (north_star: "tensor cores
// only if 3xTF32 ... meets tolerance"): every FP32 operand is split a = hi + lo with hi, lo in TF32 and the product is
// hi.hi + lo.hi + hi.lo accumulated in FP32 (the dropped lo.lo term is 2^-22 relative), so the result has FP32 accuracy.
//
//  * The dense form needs 6x the multiply-adds of the factorised Legendre + Fourier form of the FP32 kernels (and 3x
//    more for the split), but the tensor pipe has ~27x the FP32-FMA rate: measured numbers in DESIGN.md 3.6.
//  * B = the basis Y(j, ordinate), built once per plan by running the FP32 kernel on the NLM unit vectors (so both
//    variants use the same basis values), split into hi / lo, cut into K chunks of 32 (one 128-byte swizzle row) and
//    two ordinate chunks (N <= 256 per MMA), stored in global memory ALREADY in the canonical K-major SWIZZLE_128B
//    shared-memory layout: one elected thread streams a tile with cp.async.bulk (TMA) into place.
//  * A = the ragged SH rows (adaptive truncation, CSR): 128 threads stage a [128 x 32] chunk, zero-filled beyond NS(point),
//    split it and write hi / lo in the same swizzled layout (16-byte chunk c of row r at c ^ (r & 7)).
//  * One thread issues the MMAs (M=128, N=N1 | N2, K=8 per instruction, 24 per K chunk); tcgen05.commit releases the
//    A stage and each B half as soon as the MMAs reading them retire, so the TMA of the next chunk's first half runs
//    under the MMAs of this chunk's second half.
//  * Epilogue: each of the four staging warps owns 32 TMEM lanes (= 32 consecutive grid points), tcgen05.ld 16 columns
//    at a time and writes DOFIELD(NPTS, 1, NANG) in 128-byte rows (points fastest).
// ------------------------------------------------------------------------------------------------------------------
#define TC_BM 128
#define TC_BK 32
#define TC_PWARPS 8                // staging warps
#define TC_THREADS ((TC_PWARPS + 2 + 4) * 32)

__device__ __forceinline__ unsigned tc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(unsigned long long *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(unsigned long long *b)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(tc_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(unsigned long long *b, unsigned bytes)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(tc_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(unsigned long long *b, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TC_DONE_%=;\n\t"
        "bra TC_WAIT_%=;\n\t"
        "TC_DONE_%=:\n\t}" ::"r"(tc_smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start address >> 4, SBO = 1024 B (8 rows of
// 128 B), version 1, layout type 2
__device__ __forceinline__ unsigned long long tc_desc(unsigned smem_addr)
{
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                            unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ unsigned tc_tf32(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// Work item = (tile of 128 grid points, ordinate half): each half (<= 256 ordinates, one MMA wide) has its own accumulator
// stage in tensor memory (2 x n0 <= 512 columns), so the epilogue of an item drains under the MMAs of the next one.  The
// grid is even (one CTA per SM), so a CTA always works on the same half and streams only that half's basis tiles.
struct TcArgs {
    int npts, nang, nlm, kch, n0, n1, ring, nhalves;   // n0 >= n1 ordinates per half (multiples of 16), ring slots in shared memory
    const int *shptr;
    const float *sh;
    float *dofield;
    const unsigned char *bpack;                        // [half][kch] tiles of 2*n0*128 bytes (hi | lo)
};
#define TC_RING_MAX 4

__global__ void __launch_bounds__(TC_THREADS, 1) sh_to_do_tc_kernel(TcArgs a, int nitems)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    // carve: A stages (hi | lo) x 2, the ring of basis tiles, barriers
    unsigned char *base = (unsigned char *)(((size_t)tc_smem + 1023) & ~(size_t)1023);
    unsigned char *A0 = base, *A1 = base + 2 * TC_BM * 128;
    unsigned char *Bring = base + 4 * TC_BM * 128;
    const unsigned slotb = 2u * (unsigned)a.n0 * 128u;
    unsigned long long *bars = (unsigned long long *)(Bring + (size_t)a.ring * slotb);
    unsigned long long *a_full = bars, *a_empty = bars + 2, *b_full = bars + 4, *b_empty = bars + 4 + TC_RING_MAX;
    unsigned long long *t_full = bars + 4 + 2 * TC_RING_MAX, *t_empty = t_full + 2;
    unsigned *tmem_slot = (unsigned *)(t_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            tc_mbar_init(&a_full[i], TC_PWARPS * 32); tc_mbar_init(&a_empty[i], 1);
            tc_mbar_init(&t_full[i], 1); tc_mbar_init(&t_empty[i], 128);
        }
        for (int i = 0; i < TC_RING_MAX; i++) { tc_mbar_init(&b_full[i], 1); tc_mbar_init(&b_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_PWARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;
    // item -> (tile, half): consecutive items are the two halves of a tile
    const int nh = a.nhalves;

    if (warp < TC_PWARPS) {
        // ===== A staging (256 threads: 8 lanes x 16 bytes per row, 4 rows per thread and chunk) =====
        const int t = threadIdx.x, c = t & 7, r0 = t >> 3;
        unsigned g = 0;                                               // chunks staged so far (stage = g & 1)
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item / nh;
            int off[4], ns[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int p = tile * TC_BM + r0 + 32 * i;
                if (p < a.npts) { off[i] = __ldg(&a.shptr[p]); ns[i] = __ldg(&a.shptr[p + 1]) - off[i]; }
                else { off[i] = 0; ns[i] = 0; }
            }
            // two chunks of values are in flight while a third is converted and written
            float va[4][4], vb[4][4];
            auto fetch = [&](float (&v)[4][4], int kc) {
                const int j0 = kc * TC_BK + c * 4;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float *src = a.sh + off[i] + j0;
#pragma unroll
                    for (int e = 0; e < 4; e++) v[i][e] = (j0 + e < ns[i]) ? __ldg(src + e) : 0.0f;
                }
            };
            auto stage = [&](const float (&v)[4][4]) {
                const unsigned st = g & 1;
                uint4 h[4], l[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    h[i].x = tc_tf32(v[i][0]); h[i].y = tc_tf32(v[i][1]); h[i].z = tc_tf32(v[i][2]); h[i].w = tc_tf32(v[i][3]);
                    l[i].x = tc_tf32(v[i][0] - __uint_as_float(h[i].x)); l[i].y = tc_tf32(v[i][1] - __uint_as_float(h[i].y));
                    l[i].z = tc_tf32(v[i][2] - __uint_as_float(h[i].z)); l[i].w = tc_tf32(v[i][3] - __uint_as_float(h[i].w));
                }
                tc_mbar_wait(&a_empty[st], ((g >> 1) & 1) ^ 1);
                unsigned char *hi = st ? A1 : A0, *lo = hi + TC_BM * 128;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int row = r0 + 32 * i;
                    const unsigned o = (unsigned)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
                    *(uint4 *)(hi + o) = h[i];
                    *(uint4 *)(lo + o) = l[i];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (MMA)
                tc_mbar_arrive(&a_full[st]);
                g++;
            };
            fetch(va, 0);
            if (a.kch > 1) fetch(vb, 1);
            for (int kc = 0; kc < a.kch; kc += 2) {
                stage(va);
                if (kc + 2 < a.kch) fetch(va, kc + 2);
                if (kc + 1 < a.kch) {
                    stage(vb);
                    if (kc + 3 < a.kch) fetch(vb, kc + 3);
                }
            }
        }
    } else if (warp == TC_PWARPS) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            unsigned g = 0, it = 0, rb = 0;                            // rb: basis tiles consumed (ring slot = rb % ring)
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
                const int half = item % nh;
                const int n = half == 0 ? a.n0 : a.n1;
                const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(TC_BM >> 4) << 24);
                const unsigned ts = it & 1;                            // accumulator stage
                tc_mbar_wait(&t_empty[ts], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d = tmem + ts * (unsigned)a.n0;
                for (int kc = 0; kc < a.kch; kc++, g++, rb++) {
                    const unsigned st = g & 1;
                    const unsigned slot = rb % (unsigned)a.ring;
                    tc_mbar_wait(&a_full[st], (g >> 1) & 1);
                    tc_mbar_wait(&b_full[slot], (rb / (unsigned)a.ring) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned ahi = tc_smem_u32(st ? A1 : A0), alo = ahi + TC_BM * 128;
                    const unsigned bhi = tc_smem_u32(Bring + (size_t)slot * slotb), blo = bhi + (unsigned)a.n0 * 128u;
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; k++) {
                        const unsigned ko = (unsigned)k * 32;                      // 8 TF32 = 32 bytes along K inside the swizzle row
                        tc_mma_tf32(d, tc_desc(ahi + ko), tc_desc(bhi + ko), idesc, (kc | k) != 0);
                        tc_mma_tf32(d, tc_desc(alo + ko), tc_desc(bhi + ko), idesc, 1u);
                        tc_mma_tf32(d, tc_desc(ahi + ko), tc_desc(blo + ko), idesc, 1u);
                    }
                    tc_commit(&b_empty[slot]);                                     // this basis tile may be overwritten
                    tc_commit(&a_empty[st]);
                }
                tc_commit(&t_full[ts]);
            }
        }
        __syncwarp();
    } else if (warp == TC_PWARPS + 1) {
        // ===== basis loader: one thread streams the pre-swizzled tiles with cp.async.bulk, `ring` tiles ahead =====
        if (lane == 0) {
            unsigned rb = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int half = item % nh;
                for (int kc = 0; kc < a.kch; kc++, rb++) {
                    const unsigned slot = rb % (unsigned)a.ring;
                    tc_mbar_wait(&b_empty[slot], ((rb / (unsigned)a.ring) & 1) ^ 1);
                    tc_mbar_expect_tx(&b_full[slot], slotb);
                    tc_bulk_g2s(Bring + (size_t)slot * slotb, a.bpack + ((size_t)half * a.kch + kc) * slotb, slotb, &b_full[slot]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps (the last four): TMEM lanes 32*(warp & 3) .. +31 are grid points tile*128 + 32*(warp & 3) + lane =====
        const int q4 = warp & 3;
        unsigned it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const int tile = item / nh, half = item % nh;
            const int ncols = half == 0 ? a.n0 : a.n1, col0 = half == 0 ? 0 : a.n0;
            const unsigned ts = it & 1;
            tc_mbar_wait(&t_full[ts], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int p = tile * TC_BM + 32 * q4 + lane;
            float *outp = a.dofield + (p < a.npts ? p : 0) + (size_t)col0 * (size_t)a.npts;
            const size_t np_ = (size_t)a.npts;
            const int nvalid = min(a.nang - col0, ncols);                          // ordinates of this half that exist
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                unsigned w[32];
                const unsigned taddr = tmem + ((unsigned)(32 * q4) << 16) + ts * (unsigned)a.n0 + (unsigned)c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                               "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                             : "r"(taddr) : "memory");
                if (c0 + 16 < ncols)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
                                   "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                                 : "r"(taddr + 16u) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p < a.npts) {
                    float *o = outp + (size_t)c0 * np_;
                    if (c0 + 32 <= nvalid) {
#pragma unroll
                        for (int e = 0; e < 32; e++) { __stcs(o, __uint_as_float(w[e])); o += np_; }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; e++) { if (c0 + e < nvalid) __stcs(o, __uint_as_float(w[e])); o += np_; }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_mbar_arrive(&t_empty[ts]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_PWARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------------------------
// DO_TO_SH on the tensor cores (NSTOKES=1): SH[128 points x NLM] = DOFIELD[128 x NANG] . W[NANG x NLM], the same
// machinery as sh_to_do_tc_kernel with the roles of the two index sets exchanged.  K = ordinates in chunks of 32, N = all
// NLM coefficients in one MMA (<= 256), two accumulator stages of NLM columns in tensor memory.  The A operand comes from
// DOFIELD(NPTS,1,NANG) (points fastest): a staging thread owns one grid point (row) and 16 ordinates of the chunk, so
// every global load of a warp is one 128-byte row, and writes 16-byte pieces of its row into the K-major SWIZZLE_128B
// layout.  The epilogue writes each point's coefficients to its ragged RADIANCE row (j < NR(point)).
// ------------------------------------------------------------------------------------------------------------------
struct TcBackArgs {
    int npts, nang, nlm, kch, nn, ring;        // nn: NLM rounded up to 16 (MMA N), kch: ordinate chunks of 32
    const int *rshptr;
    const float *dofield;
    float *sh_out;
    const unsigned char *bpack;                // [kch] tiles of 2*nn*128 bytes (hi | lo), rows = coefficient j, k = ordinate
};

__global__ void __launch_bounds__(TC_THREADS, 1) do_to_sh_tc_kernel(TcBackArgs a, int nitems)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *base = (unsigned char *)(((size_t)tc_smem + 1023) & ~(size_t)1023);
    unsigned char *A0 = base, *A1 = base + 2 * TC_BM * 128;
    unsigned char *Bring = base + 4 * TC_BM * 128;
    const unsigned slotb = 2u * (unsigned)a.nn * 128u;
    unsigned long long *bars = (unsigned long long *)(Bring + (size_t)a.ring * slotb);
    unsigned long long *a_full = bars, *a_empty = bars + 2, *b_full = bars + 4, *b_empty = bars + 4 + TC_RING_MAX;
    unsigned long long *t_full = bars + 4 + 2 * TC_RING_MAX, *t_empty = t_full + 2;
    unsigned *tmem_slot = (unsigned *)(t_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            tc_mbar_init(&a_full[i], TC_PWARPS * 32); tc_mbar_init(&a_empty[i], 1);
            tc_mbar_init(&t_full[i], 1); tc_mbar_init(&t_empty[i], 128);
        }
        for (int i = 0; i < TC_RING_MAX; i++) { tc_mbar_init(&b_full[i], 1); tc_mbar_init(&b_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_PWARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    if (warp < TC_PWARPS) {
        // ===== A staging: thread = (row = t & 127, ordinates 16*(t >> 7) .. +15 of the chunk) =====
        const int t = threadIdx.x, row = t & 127, kh = t >> 7;
        unsigned g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int p = item * TC_BM + row;
            const bool pin = p < a.npts;
            const float *col = a.dofield + (pin ? p : 0);
            float va[16], vb[16];
            auto fetch = [&](float (&v)[16], int kc) {
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int ia = kc * TC_BK + kh * 16 + e;
                    v[e] = (pin && ia < a.nang) ? __ldg(col + (size_t)ia * a.npts) : 0.0f;
                }
            };
            auto stage = [&](const float (&v)[16]) {
                const unsigned st = g & 1;
                uint4 h[4], l[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    h[q].x = tc_tf32(v[4 * q]); h[q].y = tc_tf32(v[4 * q + 1]); h[q].z = tc_tf32(v[4 * q + 2]); h[q].w = tc_tf32(v[4 * q + 3]);
                    l[q].x = tc_tf32(v[4 * q] - __uint_as_float(h[q].x)); l[q].y = tc_tf32(v[4 * q + 1] - __uint_as_float(h[q].y));
                    l[q].z = tc_tf32(v[4 * q + 2] - __uint_as_float(h[q].z)); l[q].w = tc_tf32(v[4 * q + 3] - __uint_as_float(h[q].w));
                }
                tc_mbar_wait(&a_empty[st], ((g >> 1) & 1) ^ 1);
                unsigned char *hi = st ? A1 : A0, *lo = hi + TC_BM * 128;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = kh * 4 + q;
                    const unsigned o = (unsigned)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
                    *(uint4 *)(hi + o) = h[q];
                    *(uint4 *)(lo + o) = l[q];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_mbar_arrive(&a_full[st]);
                g++;
            };
            fetch(va, 0);
            if (a.kch > 1) fetch(vb, 1);
            for (int kc = 0; kc < a.kch; kc += 2) {
                stage(va);
                if (kc + 2 < a.kch) fetch(va, kc + 2);
                if (kc + 1 < a.kch) {
                    stage(vb);
                    if (kc + 3 < a.kch) fetch(vb, kc + 3);
                }
            }
        }
    } else if (warp == TC_PWARPS) {
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(a.nn >> 3) << 17) | ((unsigned)(TC_BM >> 4) << 24);
            unsigned g = 0, it = 0, rb = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
                const unsigned ts = it & 1;
                tc_mbar_wait(&t_empty[ts], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d = tmem + ts * (unsigned)a.nn;
                for (int kc = 0; kc < a.kch; kc++, g++, rb++) {
                    const unsigned st = g & 1;
                    const unsigned slot = rb % (unsigned)a.ring;
                    tc_mbar_wait(&a_full[st], (g >> 1) & 1);
                    tc_mbar_wait(&b_full[slot], (rb / (unsigned)a.ring) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned ahi = tc_smem_u32(st ? A1 : A0), alo = ahi + TC_BM * 128;
                    const unsigned bhi = tc_smem_u32(Bring + (size_t)slot * slotb), blo = bhi + (unsigned)a.nn * 128u;
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; k++) {
                        const unsigned ko = (unsigned)k * 32;
                        tc_mma_tf32(d, tc_desc(ahi + ko), tc_desc(bhi + ko), idesc, (kc | k) != 0);
                        tc_mma_tf32(d, tc_desc(alo + ko), tc_desc(bhi + ko), idesc, 1u);
                        tc_mma_tf32(d, tc_desc(ahi + ko), tc_desc(blo + ko), idesc, 1u);
                    }
                    tc_commit(&b_empty[slot]);
                    tc_commit(&a_empty[st]);
                }
                tc_commit(&t_full[ts]);
            }
        }
        __syncwarp();
    } else if (warp == TC_PWARPS + 1) {
        if (lane == 0) {
            unsigned rb = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                for (int kc = 0; kc < a.kch; kc++, rb++) {
                    const unsigned slot = rb % (unsigned)a.ring;
                    tc_mbar_wait(&b_empty[slot], ((rb / (unsigned)a.ring) & 1) ^ 1);
                    tc_mbar_expect_tx(&b_full[slot], slotb);
                    tc_bulk_g2s(Bring + (size_t)slot * slotb, a.bpack + (size_t)kc * slotb, slotb, &b_full[slot]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: lane = grid point, its coefficients go to its ragged row =====
        const int q4 = warp & 3;
        unsigned it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
            const unsigned ts = it & 1;
            const int p = item * TC_BM + 32 * q4 + lane;
            int off = 0, nr = 0;
            if (p < a.npts) { off = __ldg(&a.rshptr[p]); nr = __ldg(&a.rshptr[p + 1]) - off; }
            if (nr > a.nlm) nr = a.nlm;
            float *outp = a.sh_out + off;
            tc_mbar_wait(&t_full[ts], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c0 = 0; c0 < a.nn; c0 += 32) {
                unsigned w[32];
                const unsigned taddr = tmem + ((unsigned)(32 * q4) << 16) + ts * (unsigned)a.nn + (unsigned)c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                               "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                             : "r"(taddr) : "memory");
                if (c0 + 16 < a.nn)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
                                   "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                                 : "r"(taddr + 16u) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int lim = min(nr, a.nn) - c0;
                if (lim >= 32) {
#pragma unroll
                    for (int e = 0; e < 32; e++) outp[c0 + e] = __uint_as_float(w[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e++) if (e < lim) outp[c0 + e] = __uint_as_float(w[e]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_mbar_arrive(&t_empty[ts]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_PWARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// basis tiles of the tensor-core variant (once per plan): Y(j, ordinate) from the FP32 kernel applied to the unit
// vectors, split into TF32 hi / lo and laid out as the kernel's shared-memory tiles
static unsigned tc_host_tf32(float x)
{
    unsigned u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return u;
    u += 0x1000u;                                  // cvt.rna: round to nearest, ties away from zero, 13 low bits dropped
    return u & 0xFFFFE000u;
}

static int tr_tc_build(TrPlan *P, char *errmsg)
{
    const TrArgs &f = P->fwd;
    if (f.nst != 1) return 0;
    const int nlm = f.nlm, nang = P->nang;
    const int ntot = (nang + 15) & ~15;
    if (ntot > 512 || nlm > 512) return 0;
    // two ordinate halves (one accumulator stage each); a single half when everything fits 16 columns
    int n0 = ntot, n1 = 0;
    if (ntot > 16) { n0 = ((ntot / 2) + 15) & ~15; n1 = ntot - n0; }
    if (n0 > 256) return 0;
    const int nhalves = n1 > 0 ? 2 : 1;
    const int kch = (nlm + TC_BK - 1) / TC_BK;
    const size_t slotb = 2 * (size_t)n0 * 128;
    int ring = (int)(((size_t)227 * 1024 - 1024 - 4 * TC_BM * 128 - 512) / slotb);
    if (ring > TC_RING_MAX) ring = TC_RING_MAX;
    if (ring < 2) return 0;
    const size_t smem = 1024 + 4 * TC_BM * 128 + (size_t)ring * slotb + 512;
    // unit vectors through the FP32 kernel
    std::vector<int> ptr(nlm + 1);
    for (int i = 0; i <= nlm; i++) ptr[i] = i * nlm;
    std::vector<float> eye((size_t)nlm * nlm, 0.0f);
    for (int i = 0; i < nlm; i++) eye[(size_t)i * nlm + i] = 1.0f;
    int *ptr_d = nullptr; float *eye_d = nullptr, *y_d = nullptr;
    if (at3d_malloc(&ptr_d, sizeof(int) * (nlm + 1)) != cudaSuccess || at3d_malloc(&eye_d, sizeof(float) * eye.size()) != cudaSuccess ||
        at3d_malloc(&y_d, sizeof(float) * (size_t)nlm * nang) != cudaSuccess) { set_msg(errmsg, "device allocation failure"); return 4; }
    cudaMemcpy(ptr_d, ptr.data(), sizeof(int) * (nlm + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(eye_d, eye.data(), sizeof(float) * eye.size(), cudaMemcpyHostToDevice);
    TrArgs a = f;
    a.npts = nlm; a.shptr = ptr_d; a.sh = eye_d; a.dofield = y_d;
    const int ntiles = (nlm + TR_TP - 1) / TR_TP;
    sh_to_do_kernel_s1<<<ntiles < P->nsm ? ntiles : P->nsm, TR_THREADS, P->smem_fwd + (size_t)a.nlm * 33 * sizeof(float)>>>(a, ntiles);
    std::vector<float> y((size_t)nlm * nang);                       // y[j + nlm*ia]
    cudaError_t e = cudaMemcpy(y.data(), y_d, sizeof(float) * y.size(), cudaMemcpyDeviceToHost);
    at3d_free(ptr_d); at3d_free(eye_d); at3d_free(y_d);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error building the tensor-core basis (%s)", cudaGetErrorString(e)); return 4; }
    std::vector<unsigned char> pack((size_t)nhalves * kch * slotb, 0);
    for (int half = 0; half < nhalves; half++)
        for (int kc = 0; kc < kch; kc++) {
            unsigned char *hi = pack.data() + ((size_t)half * kch + kc) * slotb, *lo = hi + (size_t)n0 * 128;
            const int nn = half == 0 ? n0 : n1, col0 = half == 0 ? 0 : n0;
            for (int n = 0; n < nn; n++)
                for (int k = 0; k < TC_BK; k++) {
                    const int j = kc * TC_BK + k, ia = col0 + n;
                    const float v = (j < nlm && ia < nang) ? y[(size_t)j + (size_t)nlm * ia] : 0.0f;
                    const unsigned h = tc_host_tf32(v);
                    float hf; memcpy(&hf, &h, 4);
                    const unsigned l = tc_host_tf32(v - hf);
                    const size_t o = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((k >> 2) ^ (n & 7)) << 4) + (size_t)(k & 3) * 4;
                    memcpy(hi + o, &h, 4);
                    memcpy(lo + o, &l, 4);
                }
        }
    unsigned char *b_d = P->up(pack.data(), pack.size());
    if (!b_d) { set_msg(errmsg, "device allocation failure"); return 4; }
    if (cudaFuncSetAttribute(sh_to_do_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    P->tc_b = b_d; P->tc_kch = kch; P->tc_n1 = n0; P->tc_n2 = n1; P->tc_ring = ring; P->tc_smem = smem;
    return 0;
}

// basis tiles of the tensor-core DO_TO_SH: W(ordinate, j) from the FP32 kernel applied to the NANG unit ordinate fields
static int tr_tc_build_back(TrPlan *P, char *errmsg)
{
    const TrArgs &f = P->bwd;
    if (f.nst != 1) return 0;
    const int nlm = f.nlm, nang = P->nang;
    const int nn = (nlm + 15) & ~15;
    if (nn > 256) return 0;
    const int kch = (nang + TC_BK - 1) / TC_BK;
    const size_t slotb = 2 * (size_t)nn * 128;
    int ring = (int)(((size_t)227 * 1024 - 1024 - 4 * TC_BM * 128 - 512) / slotb);
    if (ring > TC_RING_MAX) ring = TC_RING_MAX;
    if (ring < 2) return 0;
    const size_t smem = 1024 + 4 * TC_BM * 128 + (size_t)ring * slotb + 512;
    // unit ordinate fields: point p has DOFIELD(p, 1, ia) = (ia == p), full-length rows
    std::vector<int> ptr(nang + 1);
    for (int i = 0; i <= nang; i++) ptr[i] = i * nlm;
    std::vector<float> eye((size_t)nang * nang, 0.0f);
    for (int i = 0; i < nang; i++) eye[(size_t)i * nang + i] = 1.0f;
    int *ptr_d = nullptr; float *eye_d = nullptr, *w_d = nullptr;
    if (at3d_malloc(&ptr_d, sizeof(int) * (nang + 1)) != cudaSuccess || at3d_malloc(&eye_d, sizeof(float) * eye.size()) != cudaSuccess ||
        at3d_malloc(&w_d, sizeof(float) * (size_t)nlm * nang) != cudaSuccess) { set_msg(errmsg, "device allocation failure"); return 4; }
    cudaMemcpy(ptr_d, ptr.data(), sizeof(int) * (nang + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(eye_d, eye.data(), sizeof(float) * eye.size(), cudaMemcpyHostToDevice);
    cudaMemset(w_d, 0, sizeof(float) * (size_t)nlm * nang);
    TrArgs a = f;
    a.npts = nang; a.shptr = ptr_d; a.sh_out = w_d; a.dofield = eye_d;
    const int ntiles = (nang + TR_TP - 1) / TR_TP;
    do_to_sh_kernel<<<ntiles < P->nsm ? ntiles : P->nsm, TR_THREADS, P->smem_bwd>>>(a, ntiles);
    std::vector<float> w((size_t)nlm * nang);                       // w[j + nlm*ia]
    cudaError_t e = cudaMemcpy(w.data(), w_d, sizeof(float) * w.size(), cudaMemcpyDeviceToHost);
    at3d_free(ptr_d); at3d_free(eye_d); at3d_free(w_d);
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error building the tensor-core basis (%s)", cudaGetErrorString(e)); return 4; }
    std::vector<unsigned char> pack((size_t)kch * slotb, 0);
    for (int kc = 0; kc < kch; kc++) {
        unsigned char *hi = pack.data() + (size_t)kc * slotb, *lo = hi + (size_t)nn * 128;
        for (int n = 0; n < nlm; n++)
            for (int k = 0; k < TC_BK; k++) {
                const int ia = kc * TC_BK + k;
                const float v = ia < nang ? w[(size_t)n + (size_t)nlm * ia] : 0.0f;
                const unsigned h = tc_host_tf32(v);
                float hf; memcpy(&hf, &h, 4);
                const unsigned l = tc_host_tf32(v - hf);
                const size_t o = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((k >> 2) ^ (n & 7)) << 4) + (size_t)(k & 3) * 4;
                memcpy(hi + o, &h, 4);
                memcpy(lo + o, &l, 4);
            }
    }
    unsigned char *b_d = P->up(pack.data(), pack.size());
    if (!b_d) { set_msg(errmsg, "device allocation failure"); return 4; }
    if (cudaFuncSetAttribute(do_to_sh_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    P->tcb_b = b_d; P->tcb_kch = kch; P->tcb_nn = nn; P->tcb_ring = ring; P->tcb_smem = smem;
    return 0;
}

static int tr_variant()
{
    // 0: FP32 FMA kernels, 1: tensor cores (3xTF32) where available.  AT3D_B200_TRANSFORM=fp32|tc
    const char *e = getenv("AT3D_B200_TRANSFORM");
    if (e && !strcmp(e, "tc")) return 1;
    return 0;
}

cudaError_t tr_sh_to_do_tc(const TrPlan *P, int npts, const int *shptr_d, const float *sh_d, float *do_d, cudaStream_t st)
{
    TcArgs a;
    a.npts = npts; a.nang = P->nang; a.nlm = P->fwd.nlm; a.kch = P->tc_kch; a.n0 = P->tc_n1; a.n1 = P->tc_n2;
    a.ring = P->tc_ring; a.nhalves = P->tc_n2 > 0 ? 2 : 1;
    a.shptr = shptr_d; a.sh = sh_d; a.dofield = do_d; a.bpack = P->tc_b;
    const int nitems = ((npts + TC_BM - 1) / TC_BM) * a.nhalves;
    int grid = nitems < P->nsm ? nitems : P->nsm;
    if (a.nhalves == 2) grid &= ~1;              // an even grid: a CTA keeps its ordinate half (and its basis tiles in L2)
    if (grid < 1) grid = 1;
    sh_to_do_tc_kernel<<<grid, TC_THREADS, P->tc_smem, st>>>(a, nitems);
    return cudaGetLastError();
}
int tr_plan_has_tc(const TrPlan *P) { return P->tc_kch > 0; }

// device-resident launches: SH array (NSTOKES,*) at shptr offsets <-> DOFIELD(NPTS,NSTOKES,NANG)
cudaError_t tr_sh_to_do(const TrPlan *P, int npts, const int *shptr_d, const float *sh_d, float *do_d, cudaStream_t st)
{
    if (P->tc_kch > 0 && tr_variant() == 1) return tr_sh_to_do_tc(P, npts, shptr_d, sh_d, do_d, st);
    TrArgs a = P->fwd;
    a.npts = npts; a.shptr = shptr_d; a.sh = sh_d; a.dofield = do_d;
    const int ntiles = (npts + TR_TP - 1) / TR_TP;
    if (a.nst == 1)
        sh_to_do_kernel_s1<<<ntiles < P->nsm ? ntiles : P->nsm, TR_THREADS, P->smem_fwd + (size_t)a.nlm * 33 * sizeof(float), st>>>(a, ntiles);
    else
        sh_to_do_kernel<<<ntiles < P->nsm ? ntiles : P->nsm, TR_THREADS, P->smem_fwd, st>>>(a, ntiles);
    return cudaGetLastError();
}

cudaError_t tr_do_to_sh(const TrPlan *P, int npts, const int *rshptr_d, const float *do_d, float *sh_d, cudaStream_t st)
{
    if (P->tcb_kch > 0 && tr_variant() == 1) {
        TcBackArgs b;
        b.npts = npts; b.nang = P->nang; b.nlm = P->bwd.nlm; b.kch = P->tcb_kch; b.nn = P->tcb_nn; b.ring = P->tcb_ring;
        b.rshptr = rshptr_d; b.dofield = do_d; b.sh_out = sh_d; b.bpack = P->tcb_b;
        const int nitems = (npts + TC_BM - 1) / TC_BM;
        do_to_sh_tc_kernel<<<nitems < P->nsm ? nitems : P->nsm, TC_THREADS, P->tcb_smem, st>>>(b, nitems);
        return cudaGetLastError();
    }
    TrArgs a = P->bwd;
    a.npts = npts; a.shptr = rshptr_d; a.sh_out = sh_d; a.dofield = (float *)do_d;
    const int ntiles = (npts + TR_TP - 1) / TR_TP;
    // (a cp.async double-buffered variant like sh_to_do_kernel_s1 was measured slower here: the second input buffer
    // only fits if the azimuthal table leaves shared memory, 3.8 ms vs 3.4 ms at 1 M points)
    do_to_sh_kernel<<<ntiles < P->nsm ? ntiles : P->nsm, TR_THREADS, P->smem_bwd, st>>>(a, ntiles);
    return cudaGetLastError();
}

// direction: 0 = SH_TO_DO, 1 = DO_TO_SH; host buffers
static int transform(int direction, int npts, int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max,
                     const int32_t *nphi0, const float *mu, const float *phi, const float *wtmu,
                     const int32_t *ptr, float *sh, float *dofield, double *kernel_ms, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!nphi0 || !mu || !phi || !wtmu || !ptr || !sh || !dofield) { set_msg(errmsg, "null argument"); return 1; }
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    TrPlan *P = nullptr;
    int rc = tr_plan_create(nstokes, nstleg, ml, mm, nlm, nmu, nphi0max, nphi0, mu, phi, wtmu, &P, errmsg);
    if (rc) return rc;
    const size_t nsh = (size_t)nstokes * ptr[npts], ndo = (size_t)npts * nstokes * P->nang;
    const int *ptr_d = P->up(ptr, (size_t)npts + 1);
    float *sh_d = direction == 0 ? P->up(sh, nsh) : P->alloc<float>(nsh);
    float *do_d = direction == 1 ? P->up(dofield, ndo) : P->alloc<float>(ndo);
    if (!ptr_d || !sh_d || !do_d) { delete P; set_msg(errmsg, "device allocation failure"); return 4; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    cudaError_t e = direction == 0 ? tr_sh_to_do(P, npts, ptr_d, sh_d, do_d, 0) : tr_do_to_sh(P, npts, ptr_d, do_d, sh_d, 0);
    cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.0f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) { delete P; set_msg(errmsg, "CUDA error %s in the SH/DO transform", cudaGetErrorString(e)); return 4; }
    if (kernel_ms) *kernel_ms = ms;
    if (direction == 0) e = cudaMemcpy(dofield, do_d, ndo * sizeof(float), cudaMemcpyDeviceToHost);
    else e = cudaMemcpy(sh, sh_d, nsh * sizeof(float), cudaMemcpyDeviceToHost);
    delete P;
    if (e != cudaSuccess) { set_msg(errmsg, "CUDA error %s copying the result", cudaGetErrorString(e)); return 4; }
    return 0;
}

extern "C" int at3d_sh_to_do(int npts, int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max,
                             const int32_t *nphi0, const float *mu, const float *phi, const float *wtmu,
                             const int32_t *shptr, const float *indata, float *dofield, double *kernel_ms, char *errmsg)
{
    return transform(0, npts, nstokes, nstleg, ml, mm, nlm, nmu, nphi0max, nphi0, mu, phi, wtmu, shptr,
                     (float *)indata, dofield, kernel_ms, errmsg);
}

extern "C" int at3d_do_to_sh(int npts, int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max,
                             const int32_t *nphi0, const float *mu, const float *phi, const float *wtmu,
                             const int32_t *rshptr, const float *dofield, float *outdata, double *kernel_ms, char *errmsg)
{
    return transform(1, npts, nstokes, nstleg, ml, mm, nlm, nmu, nphi0max, nphi0, mu, phi, wtmu, rshptr,
                     outdata, (float *)dofield, kernel_ms, errmsg);
}
