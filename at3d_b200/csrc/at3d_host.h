// at3d_host.h -- host-side declarations shared by the .cu translation units.
#pragma once
#include "at3d_mem.h"
#include <cuda_runtime.h>
#include <vector>
#include <mutex>
#include <string>
#include "at3d_device.cuh"
#include "../../include/at3d_b200.h"

struct RayErr;

// cudaFuncSetAttribute + occupancy query of a kernel for (device, block size, dynamic shared memory): asked once and
// remembered at the call site -- driver queries are not free on a box where several processes (or nvidia-smi) talk to the
// driver at the same time, and these sit between the launches of every step
struct KernelFit {
    const void *fn = nullptr;
    int dev = -1, threads = 0;
    size_t smem = 0;
    int per_sm = 1, nsm = 148;
};
template <class K>
inline void kernel_fit(KernelFit &c, K kernel, int threads, size_t smem)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (c.fn == (const void *)kernel && c.dev == dev && c.threads == threads && c.smem == smem) return;
    cudaDeviceGetAttribute(&c.nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, kernel, threads, smem);
    if (c.per_sm < 1) c.per_sm = 1;
    c.fn = (const void *)kernel; c.dev = dev; c.threads = threads; c.smem = smem;
}

// device buffer that grows on demand and is reused between calls.  These are the large streaming buffers (source stream,
// visit records, pairs, ray staging): plain cudaMalloc blocks (contiguous, large pages), parked in a process-wide cache
// when their owner goes away and taken from it by the next owner (at3d_capi.cu).
cudaError_t at3d_big_take(void **p, size_t bytes, size_t *cap);
void at3d_big_park(void *p, size_t cap);

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) at3d_big_park(p, cap);
        p = nullptr; cap = 0;
        return at3d_big_take(&p, bytes, &cap);
    }
    void release() { if (p) at3d_big_park(p, cap); p = nullptr; cap = 0; }
};

struct at3d_state {
    DevState S;
    DevGrad G;
    int grad_attached = 0;
    std::vector<void *> owned;      // device allocations freed at destroy
    std::vector<void *> grad_owned; // derivative tables (replaced by every attach_gradient)
    size_t bytes = 0;
    int device = 0;
    // reusable per-call buffers
    DevBuf rays, out, trace, misc, slabs, err, pix, work, recs, hits, viewsrc, pairs;
    RayGeom geom;                   // host copy of the per-ray setup constants
    RayPack *packs_h = nullptr;     // pinned host staging of the per-ray packs (grows on demand)
    size_t packs_cap = 0;
    float *bcrad_dev = nullptr;
    unsigned long long *counts_dev = nullptr;
    std::vector<long long> recoff_h; // per-ray visit-record offsets of the last gradient call (host copy)
    void *hpin = nullptr;            // small pinned buffer for the asynchronous read-backs of the gradient call
    double src_gb_limit = 0.0;       // size limit of the source-stream pool (from the free memory, once per state)
    std::vector<int> nr_h;          // radiance SH length per point (host copy, for the gradient tables)
    int *ray_counter = nullptr;     // work counter of the persistent ray kernels
    int nbcrad = 0;
    std::mutex mu;                  // the per-call buffers above are shared: calls on one state are serialised (the
                                    // reference calls RENDER from several joblib threads on slices of the rays)
    double gw_rec_per_ray = 0.0;    // visit records per ray seen by the single-walk derivative pass (sizes its pool)
    int view_min_rays = 256;        // shortest run of equal-direction rays that gets a pre-evaluated view source (0: off)
};

size_t render_smem_bytes(const DevState &S);
cudaError_t launch_forward(const DevState &S, int nrays, const float *camx, const float *camy,
                           const float *camz, const double *cammu, const double *camphi,
                           const RayPack *packs, float *out_f32, double *out_f64, double *out_tot, int modes,
                           int correctinterpolate, int singlescatter, int nosurface, int maxsub,
                           int *trace_cells, int trace_cap, int *trace_n, int *trace_nsub,
                           RayErr *err, int *ray_counter, int *npt_out, cudaStream_t stream);
cudaError_t launch_build_cellrec(int ncells, const int *gridptr, const int *neighptr, const int *treeptr,
                                 const short *cellflags, int4 *cellrec, cudaStream_t s);
cudaError_t launch_build_ptrec(int npts, const float *gridpos, const float *total_ext, float4 *ptrec, cudaStream_t s);
cudaError_t launch_build_ptsrc(int npts, int kmax, const int2 *srcrec, const int *sscount, const int2 *ssent,
                               int ncells, const int4 *cellrec, const float4 *ptrec, int4 *ptsrc, cudaStream_t s);
cudaError_t launch_lambertian_boundary(const DevState &S, const float *fluxes, float *bcrad, cudaStream_t s);
void view_segments(const at3d_state *st, const at3d_rays *rays, std::vector<size_t> &seg_start,
                   std::vector<size_t> &seg_len, std::vector<char> &seg_view);
cudaError_t launch_view_source(const DevState &S, const RayPack &pk, double mu2, double phi2, int singlescatter,
                               float *viewsrc, cudaStream_t stream);
int tray_block_threads(const DevState &S);
cudaError_t launch_surface(const DevState &S, int nrays, const SurfHit *hits, const double *cammu,
                           const double *camphi, float *out, RayErr *err, cudaStream_t stream);
cudaError_t launch_prep_sh(const DevState &S, int tms, const int *shptr, const float *sh_in,
                           const int2 *rec, float *sh_out, int *sscount, int2 *ssent, cudaStream_t s);

// SH <-> discrete-ordinate transforms on device-resident arrays (at3d_transform.cu)
struct TrPlan;
int tr_plan_create(int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max, const int32_t *nphi0,
                   const float *mu, const float *phi, const float *wtmu, TrPlan **out, char *errmsg);
void tr_plan_destroy(TrPlan *p);
int tr_plan_nang(const TrPlan *p);
cudaError_t tr_sh_to_do(const TrPlan *P, int npts, const int *shptr_d, const float *sh_d, float *do_d, cudaStream_t st);
cudaError_t tr_do_to_sh(const TrPlan *P, int npts, const int *rshptr_d, const float *do_d, float *sh_d, cudaStream_t st);

// ---- COMPUTE_SOURCE on device-resident arrays (at3d_source.cu), shared with the solution iterations (at3d_solver.cu) ----
struct CsArgs {
    int npts, nstokes, nstleg, nlm, ml, mm, nleg, npart, nq, srctype, deltam, interp_new, newmethod;
    int ldp;                  // leading dimension of the per-species point arrays (0: npts)
    int first, accelflag, fixsh;
    float phasemax, secmu0, srcmin;
    const float *extinct, *albedo, *total_ext, *legen, *phaseinterpwt, *dirflux, *radiance, *ylmsun, *planck;
    const int *iphase, *rshptr, *lofj;
    const int *shptr_old, *oshptr_old;
    const float *source_old, *delsource_old;
    int *ns_new;              // [npts]
    double *partials;         // [nblocks,4]
    int *bad;                 // NR>NLM flag
    // mixed Legendre table of every point (cs_mix_kernel; the optical properties only): [npts][NSTLEG*(NLEG+1)], and
    // (albedo, planck) per point
    float *mix_legent;
    float2 *mix_ap;
    // second kernel
    const int *shptr_new;
    float *source_new, *delsource_new;
};

size_t cs_scan_bytes(int npts);
int cs_grid_blocks(int npts);
int cs_device_step(CsArgs &a, int nblk, void *scan_tmp, size_t tmpb, int *shptr_new, double *sums, int maxiv, size_t cap_new,
                   float *source_new, int *total_new_out, bool mix_ready, char *errmsg);

// ---- TRILIN_INTERP_PROP on the device (at3d_prep.cu), shared with the adaptive solve (at3d_adapt.cu) ----
#define TPA_MAXQ 32
struct TpaArgs {
    int first, count, ld, npart, mnm, npx, npy, npz, ml, deltam, interp_new, nzckd, srctype, units;
    int prepare_prop;         // 1: TOTAL_EXT in the operation order of PREPARE_PROP (base grid), 0: of INTERPOLATE_POINT
    float delx, dely, xstart, ystart, phasemax, wavelen;
    double extmin, scatmin;
    const float *gridpos, *zlevels, *tempp, *extinctp, *albedop, *ftab, *zckd, *gasabs;   // ftab[numphase] = LEGEN(1,ML+1,.)
    const int *iphasep;
    const float *phasewtp;
    float *extinct, *albedo, *total_ext, *phaseinterpwt, *temp, *planck;
    int *iphase, *bad;
};

cudaError_t launch_tpa(const TpaArgs &a, cudaStream_t s);
void tpa_extmin(const float *zlevels, int npz, double *extmin, double *scatmin);
cudaError_t launch_direct_points(const double *out_d, const int *out_i, int bcflag, int npx, int npy, int npz,
                                 float xstart, float ystart, const float *zlevels_d, const float *gridpos_d,
                                 const float *extdirp_d, float solarflux, float *dirflux_d, int count, int *flags_d,
                                 cudaStream_t s);

// ---- new grid points of SPLIT_GRID (INTERPOLATE_POINT, shdomsub1.f:5109-5211): radiance = mean of the two parents,
// source function from it.  rec[k] = {ip1, ip2, ip, ir, nr, is, ns} (0-based points, SH offsets / lengths). ----
struct NewPointRec { int ip1, ip2, ip, ir, nr, is, ns, pad; };
cudaError_t cs_mix_points(CsArgs a, int first, int count, cudaStream_t s);
cudaError_t launch_interp_points(const CsArgs &a, const NewPointRec *rec_d, int count, const int *rshptr_d, float *radiance,
                                 float *source, cudaStream_t s);
