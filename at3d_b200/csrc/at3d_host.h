// at3d_host.h -- host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <string>
#include "at3d_device.cuh"
#include "../../include/at3d_b200.h"

struct RayErr;

// device buffer that grows on demand and is reused between calls
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// gradient-side device arrays (at3d_state_attach_gradient)
struct GradDev {
    int attached = 0;
    int maxpg = 0, numder = 0, dnumphase = 0, deriv_maxnmicro = 0, longest_path_pts = 0;
    int maxnmicro = 0;
    const int *partder = nullptr, *doexact = nullptr;
    const float *dext = nullptr, *dalb = nullptr, *dextm = nullptr, *dalbm = nullptr, *dfj = nullptr;
    const float *optinterpwt = nullptr;
    const int *interpptr = nullptr;
    const float *dleg = nullptr, *dphasetab = nullptr, *dphasewtp = nullptr, *phasewtp = nullptr;
    const int *diphasep = nullptr, *iphasep = nullptr;
    const float *extinctp = nullptr, *albedop = nullptr;
    // transposed direct-beam path lists (CSR by property point): GRADOUT(ib) -= DEXTM*sum(DPATH*BW)
    const int *dbt_rowptr = nullptr;   // [maxpg+1]
    const int *dbt_col = nullptr;      // RTE grid point (1-based)
    const float *dbt_val = nullptr;    // DPATH
    long long dbt_nnz = 0;
};

struct at3d_state {
    DevState S;
    GradDev G;
    std::vector<void *> owned;      // device allocations freed at destroy
    size_t bytes = 0;
    int device = 0;
    // reusable per-call buffers
    DevBuf rays, out, trace, misc, slabs, err;
    float *bcrad_dev = nullptr;
    int nbcrad = 0;
};

size_t render_smem_bytes(const DevState &S);
cudaError_t launch_render(const DevState &S, int nrays, const float *camx, const float *camy,
                          const float *camz, const double *cammu, const double *camphi,
                          float *out_f32, double *out_f64, int mode,
                          int correctinterpolate, int singlescatter, int nosurface, int maxsub,
                          int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub,
                          RayErr *err, cudaStream_t stream);
cudaError_t launch_build_cellrec(int ncells, const int *gridptr, const int *neighptr, const int *treeptr,
                                 const short *cellflags, int4 *cellrec, cudaStream_t s);
cudaError_t launch_build_ptrec(int npts, const float *gridpos, const float *total_ext, float4 *ptrec, cudaStream_t s);
cudaError_t launch_lambertian_boundary(const DevState &S, const float *fluxes, float *bcrad, cudaStream_t s);
cudaError_t launch_prep_sh(const DevState &S, int tms, const int *shptr, const float *sh_in,
                           const int2 *rec, float *sh_out, int *sscount, int2 *ssent, cudaStream_t s);
