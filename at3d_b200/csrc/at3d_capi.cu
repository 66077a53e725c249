// at3d_capi.cu -- C-ABI of libat3d_b200.so (include/at3d_b200.h): state residency and RENDER.
#include "at3d_mem.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <mutex>
#include <vector>
#include <unordered_set>
#include "at3d_host.h"
#include "at3d_ray.cuh"

static void set_msg(char *errmsg, const char *fmt, ...)
{
    if (!errmsg) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errmsg, AT3D_ERRMSG_LEN, fmt, ap);
    va_end(ap);
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_msg(errmsg, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__,   \
                    __LINE__, #expr);                                                          \
            return 4;                                                                          \
        }                                                                                      \
    } while (0)

extern "C" const char *at3d_b200_version(void) { return "at3d_b200 0.1 (sm_100a)"; }

// ---- device memory (at3d_mem.h).  Default: cudaMalloc / cudaFree.  With memory reuse switched on
// (at3d_set_memory_reuse(1), or AT3D_B200_POOL_GB > 0 in the environment) the state arrays, derivative tables, solver
// objects and per-call arenas come from the driver's stream-ordered pool with a release threshold, and the large
// streaming buffers (DevBuf) are parked in a process-wide cache between owners. ----
static bool g_pool_ready[64];
static std::mutex g_pool_mu;
static std::unordered_set<void *> g_pooled;          // pointers that came from cudaMallocAsync
static double g_pool_gb = -1.0;                      // < 0: not initialised yet

static double pool_gb()
{
    if (g_pool_gb < 0.0) {
        double gb = 0.0;
        if (const char *e = getenv("AT3D_B200_POOL_GB")) gb = atof(e);
        g_pool_gb = gb > 0.0 ? gb : 0.0;
    }
    return g_pool_gb;
}

extern "C" int at3d_set_memory_reuse(int on)
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    const int was = pool_gb() > 0.0 ? 1 : 0;
    if (on) { if (g_pool_gb <= 0.0) g_pool_gb = 64.0; }
    else g_pool_gb = 0.0;
    for (int i = 0; i < 64; i++) g_pool_ready[i] = false;      // thresholds are set again on the next allocation
    return was;
}

static bool pool_setup(int dev)
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    const double gb = pool_gb();
    if (gb <= 0.0 || dev < 0 || dev >= 64) return false;
    if (g_pool_ready[dev]) return true;
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
    if (!supported) return false;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    unsigned long long thr = (unsigned long long)(gb * 1073741824.0);
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    g_pool_ready[dev] = true;
    return true;
}

cudaError_t at3d_pool_alloc(void **p, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (!pool_setup(dev)) return cudaMalloc(p, bytes ? bytes : 1);
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t)0);
    if (e == cudaErrorMemoryAllocation) {
        // what the pool keeps may be what is missing: give it back and try once more
        cudaGetLastError();
        cudaDeviceSynchronize();
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t)0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);      // valid on every stream from here on
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lock(g_pool_mu); g_pooled.insert(*p); }
    return e;
}

cudaError_t at3d_pool_free(void *p)
{
    if (!p) return cudaSuccess;
    bool pooled;
    { std::lock_guard<std::mutex> lock(g_pool_mu); pooled = g_pooled.erase(p) > 0; }
    if (!pooled) return cudaFree(p);
    cudaError_t e = cudaDeviceSynchronize();                               // cudaFree semantics: nothing uses p any more
    if (e != cudaSuccess) return e;
    return cudaFreeAsync(p, (cudaStream_t)0);
}

// ---- the cache of large buffers (DevBuf): cudaMalloc blocks parked between owners ----
struct BigBlock { void *p; size_t cap; int dev; };
static std::vector<BigBlock> g_big;
static size_t g_big_bytes = 0;

static size_t big_limit()
{
    const double gb = pool_gb();
    return gb > 0.0 ? (size_t)(gb * 1073741824.0) : 0;
}

cudaError_t at3d_big_take(void **p, size_t bytes, size_t *cap)
{
    *p = nullptr; *cap = 0;
    if (bytes == 0) bytes = 1;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        int best = -1;
        for (int i = 0; i < (int)g_big.size(); i++)
            if (g_big[i].dev == dev && g_big[i].cap >= bytes && (best < 0 || g_big[i].cap < g_big[best].cap)) best = i;
        // a parked block serves requests down to a quarter of its size (a 16 GB block is not spent on 1 MB)
        if (best >= 0 && g_big[best].cap / 4 <= bytes + (1 << 20)) {
            *p = g_big[best].p; *cap = g_big[best].cap;
            g_big_bytes -= g_big[best].cap;
            g_big.erase(g_big.begin() + best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        // what is parked (here and in the driver's pool) may be what is missing
        cudaGetLastError();
        at3d_trim_memory();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

void at3d_big_park(void *p, size_t cap)
{
    if (!p) return;
    cudaDeviceSynchronize();                          // cudaFree semantics: nothing in flight uses the block any more
    std::vector<void *> drop;
    cudaPointerAttributes attr;
    int dev = 0;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) dev = attr.device; else cudaGetLastError();
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        const size_t limit = big_limit();
        if (cap > limit) drop.push_back(p);
        else {
            g_big.push_back({p, cap, dev});
            g_big_bytes += cap;
            while (g_big_bytes > limit && !g_big.empty()) {       // oldest first
                drop.push_back(g_big.front().p);
                g_big_bytes -= g_big.front().cap;
                g_big.erase(g_big.begin());
            }
        }
    }
    for (void *q : drop) cudaFree(q);
}

// one parked pinned staging buffer (the per-ray setup records of host rays, 128 B per ray): a state that is destroyed
// leaves it for the next one instead of cudaFreeHost / cudaMallocHost (~15 ms for the 46 MB of BASELINE configs[1])
static void *g_pinned_p = nullptr;
static size_t g_pinned_bytes = 0;

static void pinned_park(void *p, size_t bytes)
{
    if (!p) return;
    void *drop = p;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (bytes > g_pinned_bytes) { drop = g_pinned_p; g_pinned_p = p; g_pinned_bytes = bytes; }
    }
    if (drop) cudaFreeHost(drop);
}

static cudaError_t pinned_take(void **p, size_t bytes, size_t *got)
{
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (g_pinned_p && g_pinned_bytes >= bytes) {
            *p = g_pinned_p; *got = g_pinned_bytes;
            g_pinned_p = nullptr; g_pinned_bytes = 0;
            return cudaSuccess;
        }
    }
    *got = bytes;
    return cudaMallocHost(p, bytes);
}

extern "C" int at3d_trim_memory(void)
{
    {
        void *drop = nullptr;
        { std::lock_guard<std::mutex> lock(g_pool_mu); drop = g_pinned_p; g_pinned_p = nullptr; g_pinned_bytes = 0; }
        if (drop) cudaFreeHost(drop);
        std::vector<BigBlock> big;
        { std::lock_guard<std::mutex> lock(g_pool_mu); big.swap(g_big); g_big_bytes = 0; }
        for (auto &b : big) cudaFree(b.p);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceSynchronize() != cudaSuccess) return 4;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return cudaMemPoolTrimTo(pool, 0) == cudaSuccess ? 0 : 4;
}

extern "C" int at3d_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int at3d_set_device(int device)
{
    return cudaSetDevice(device) == cudaSuccess ? 0 : 4;
}

template <typename T>
static int upload(at3d_state *st, const T *host, size_t n, const T **dev, char *errmsg)
{
    *dev = nullptr;
    if (!host || n == 0) return 0;
    void *p = nullptr;
    CUDA_TRY(at3d_malloc(&p, n * sizeof(T)));
    st->owned.push_back(p);
    st->bytes += n * sizeof(T);
    CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T *)p;
    return 0;
}

template <typename T>
static int dalloc(at3d_state *st, size_t n, T **dev, char *errmsg)
{
    void *p = nullptr;
    if (n == 0) n = 1;
    CUDA_TRY(at3d_malloc(&p, n * sizeof(T)));
    st->owned.push_back(p);
    st->bytes += n * sizeof(T);
    *dev = (T *)p;
    return 0;
}

// planar padded SH block offsets: point block = nstokes planes of AT3D_SHPAD(ns) floats
static size_t make_sh_records(const int32_t *shptr, int npts, int nstokes, std::vector<int2> &rec)
{
    rec.resize(npts);
    size_t off = 0;
    for (int i = 0; i < npts; i++) {
        int ns = shptr[i + 1] - shptr[i];
        int nsp = AT3D_SHPAD(ns);
        rec[i] = make_int2((int)off, ns);
        off += (size_t)nstokes * nsp;
    }
    return off;
}

static int prep_sh_array(at3d_state *st, int tms, const int32_t *shptr_h, const float *sh_h,
                         const int2 **rec_out, const float **sh_out, int *sscount, int2 *ssent, char *errmsg)
{
    const DevState &S = st->S;
    std::vector<int2> rec;
    size_t total = make_sh_records(shptr_h, S.npts, S.nstokes, rec);
    if (total >= (size_t)1 << 31) { set_msg(errmsg, "SH array too large for 32-bit offsets"); return 2; }
    const int2 *rec_d; float *out_d;
    int rc = upload(st, rec.data(), rec.size(), &rec_d, errmsg); if (rc) return rc;
    rc = dalloc(st, total + 4, &out_d, errmsg); if (rc) return rc;
    // staging copies of the reference-layout arrays (freed after the re-layout)
    int *shptr_d = nullptr; float *in_d = nullptr;
    size_t nin = (size_t)S.nstokes * (size_t)shptr_h[S.npts];
    CUDA_TRY(at3d_malloc((void **)&shptr_d, (S.npts + 1) * sizeof(int)));
    CUDA_TRY(at3d_malloc((void **)&in_d, (nin + 1) * sizeof(float)));
    CUDA_TRY(cudaMemcpy(shptr_d, shptr_h, (S.npts + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(in_d, sh_h, nin * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(launch_prep_sh(S, tms, shptr_d, in_d, rec_d, out_d, sscount, ssent, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    at3d_free(shptr_d); at3d_free(in_d);
    *rec_out = rec_d; *sh_out = out_d;
    return 0;
}

extern "C" int at3d_state_destroy(at3d_state *st)
{
    if (!st) return 0;
    for (void *p : st->owned) at3d_free(p);
    for (void *p : st->grad_owned) at3d_free(p);
    st->pix.release(); st->work.release();
    st->hits.release(); st->viewsrc.release();
    pinned_park(st->packs_h, st->packs_cap * sizeof(RayPack));
    st->packs_h = nullptr; st->packs_cap = 0;
    if (st->hpin) cudaFreeHost(st->hpin);
    st->rays.release(); st->out.release(); st->trace.release(); st->misc.release();
    st->slabs.release(); st->err.release(); st->recs.release(); st->pairs.release();
    delete st;
    return 0;
}

extern "C" int64_t at3d_state_bytes(const at3d_state *st) { return st ? (int64_t)st->bytes : 0; }

extern "C" int at3d_state_create(const at3d_state_desc *d, at3d_state **out, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!d || !out) { set_msg(errmsg, "null argument"); return 1; }
    *out = nullptr;
    if (at3d_device_count() < 1) { set_msg(errmsg, "no CUDA device: at3d_b200 has no CPU fallback"); return 4; }
    if (!(d->srctype == 'S' || d->srctype == 'T' || d->srctype == 'B')) { set_msg(errmsg, "SRCTYPE must be S, T or B"); return 1; }
    if (d->srctype != 'S' && d->units == 'B') { set_msg(errmsg, "band-integrated Planck units (UNITS='B') are not implemented"); return 3; }
    if (!(d->nstokes == 1 || d->nstokes == 3)) { set_msg(errmsg, "NSTOKES must be 1 or 3"); return 3; }
    if ((d->nstokes == 1) != (d->nstleg == 1)) { set_msg(errmsg, "NSTLEG must be 1 for NSTOKES=1 and 6 otherwise"); return 3; }
    {
        const int t = d->sfctype1;
        if (!(t == 'L' || t == 'W' || t == 'D' || t == 'O' || t == 'R' || t == 'M')) { set_msg(errmsg, "SURFACE_BRDF: Unknown BRDF type"); return 1; }
        if (t != 'L' && d->sfctype0 != 'V') { set_msg(errmsg, "general BRDF surfaces are variable surfaces (SFCTYPE 'V?')"); return 1; }
        if ((t == 'O' || t == 'R' || t == 'M') && d->nstokes > 1) { set_msg(errmsg, "Ocean, RPV and RossLi BRDFs are only for the unpolarized case"); return 1; }
        if (t != 'L' && (!d->wtdo || !d->bcrad || !d->sfcgridparms)) { set_msg(errmsg, "general BRDF surfaces need WTDO, SFCGRIDPARMS and the stored downwelling BCRAD"); return 1; }
    }
    if (d->numphase < 1) { set_msg(errmsg, "NUMPHASE=0 is not supported."); return 1; }
    at3d_state *st = new at3d_state();
    if (const char *e = getenv("AT3D_VIEW_MIN_RAYS")) st->view_min_rays = atoi(e);   // developer knob (0 disables view sources)
    cudaGetDevice(&st->device);
    DevState &S = st->S;
    memset(&S, 0, sizeof(S));
    S.nstokes = d->nstokes; S.nstleg = d->nstleg; S.nx = d->nx; S.ny = d->ny; S.nz = d->nz;
    S.npts = d->npts; S.ncells = d->ncells; S.ml = d->ml; S.mm = d->mm; S.nlm = d->nlm;
    S.nlmp = AT3D_SHPAD(d->nlm); S.nleg = d->nleg; S.numphase = d->numphase; S.npart = d->npart;
    S.maxnmicro = d->maxnmicro; S.nq = 8 * d->maxnmicro; S.bcflag = d->bcflag; S.ipflag = d->ipflag;
    S.nmu = d->nmu; S.nphi0max = d->nphi0max; S.maxnbc = d->maxnbc; S.ntoppts = d->ntoppts;
    S.nbotpts = d->nbotpts; S.nsfcpar = d->nsfcpar; S.nscatangle = d->nscatangle; S.nstphase = d->nstphase;
    S.deltam = d->deltam; S.srctype = d->srctype; S.sfctype0 = d->sfctype0; S.sfctype1 = d->sfctype1;
    S.interp_new = d->interp_new; S.kmax = d->npart * 8 * d->maxnmicro;
    S.ny_comp = d->nstokes == 1 ? 1 : 5;
    S.solarmu = d->solarmu; S.solaraz = d->solaraz; S.gndalbedo = d->gndalbedo; S.phasemax = d->phasemax;
    S.tautol = d->tautol; S.transcut = d->transcut;
    S.nang = d->nang; S.units = d->units; S.wavelen = d->wavelen; S.gndtemp = d->gndtemp;
    st->geom.solarmu = d->solarmu; st->geom.solaraz = d->solaraz; st->geom.ztop = d->zgrid[d->nz - 1];
    st->geom.zbot = d->zgrid[0]; st->geom.nscatangle = d->nscatangle; st->geom.srctype = d->srctype;
    st->geom.deltam = d->deltam; st->geom.nstokes = d->nstokes;
    int rc = 0;
#define UP(field, count) if (!rc) rc = upload(st, d->field, (size_t)(count), &S.field, errmsg)
    const int nxg = (d->bcflag & 5) ? d->nx : d->nx + 1;
    const int nyg = (d->bcflag & 10) ? d->ny : d->ny + 1;
    UP(xgrid, nxg); UP(ygrid, nyg); UP(zgrid, d->nz);
    UP(bcptr, (size_t)d->maxnbc * 2);
    UP(nphi0, d->nmu); UP(mu, d->nmu); UP(phi, (size_t)d->nmu * d->nphi0max);
    UP(skyrad, (size_t)d->nstokes * (d->nmu / 2) * d->nphi0max);
    UP(phasetab, (size_t)d->nstphase * d->numphase * d->nscatangle);
    UP(extinct, (size_t)d->npts * d->npart); UP(albedo, (size_t)d->npts * d->npart);
    UP(dirflux, d->npts);
    UP(legen, (size_t)d->nstleg * (d->nleg + 1) * d->numphase);
    UP(iphase, (size_t)S.nq * d->npts * d->npart); UP(phaseinterpwt, (size_t)S.nq * d->npts * d->npart);
    UP(ylmsun, (size_t)d->nstleg * d->nlm);
    UP(sfcgridparms, (size_t)d->nsfcpar * d->nbotpts);
    if (d->srctype != 'S' && d->temp) UP(temp, d->npts);        // thermal component of the gradient (shdomsub4.f:1792-1799)
#undef UP
    if (rc) { at3d_state_destroy(st); return rc; }
    // LOFJ (shdomsub1.f:1057-1065)
    {
        std::vector<int> lofj(d->nlm);
        int j = 0;
        for (int l = 0; l <= d->ml; l++) {
            int me = l < d->mm ? l : d->mm;
            for (int m = -me; m <= me; m++) { if (j < d->nlm) lofj[j] = l; j++; }
        }
        if (j != d->nlm) { set_msg(errmsg, "NLM=%d inconsistent with ML=%d MM=%d", d->nlm, d->ml, d->mm); at3d_state_destroy(st); return 1; }
        rc = upload(st, lofj.data(), lofj.size(), &S.lofj, errmsg);
        if (rc) { at3d_state_destroy(st); return rc; }
    }
    // ordinate tables of the surface kernels: downward ordinates in VARIABLE_BRDF_SURFACE order with
    // W = OPI*ABS(MU)*WTDO (shdomsub1.f:2652), upward ordinates with the row of SFCGRIDRAD that the
    // "surface emission hack" of FIND_BOUNDARY_RADIANCE reads for them (shdomsub2.f:2832-2846)
    if (d->sfctype1 != 'L' || d->sfcgridrad || d->srctype != 'S') {
        const int nh = d->nang / 2, nmu = d->nmu;
        std::vector<float> omu(nh), ophi(nh), ow(nh), umu(nh), uphi(nh);
        std::vector<int> usrc(nh, -1);
        const float opi = 1.0f / acosf(-1.0f);
        int q = 0;
        for (int jmu = 1; jmu <= nmu / 2; jmu++)
            for (int jphi = 1; jphi <= d->nphi0[jmu - 1]; jphi++) {
                if (q >= nh) break;
                omu[q] = d->mu[jmu - 1]; ophi[q] = d->phi[(jmu - 1) + nmu * (jphi - 1)];
                ow[q] = d->wtdo ? opi * fabsf(d->mu[jmu - 1]) * d->wtdo[(jmu - 1) + nmu * (jphi - 1)] : 0.0f;
                q++;
            }
        std::vector<int> first(nmu / 2 + 1, 0);      // IANG before row I of SFCRAD_TEMP
        for (int i = 1; i <= nmu / 2; i++) first[i] = first[i - 1] + d->nphi0[i - 1];
        q = 0;
        for (int i = nmu / 2 + 1; i <= nmu; i++)
            for (int j = 1; j <= d->nphi0[i - 1]; j++) {
                if (q >= nh) break;
                umu[q] = d->mu[i - 1]; uphi[q] = d->phi[(i - 1) + nmu * (j - 1)];
                const int is = i - nmu / 2;           // SKYRAD(:,I-NMU/2,J): filled only for J <= NPHI0(I-NMU/2)
                usrc[q] = (j <= d->nphi0[is - 1]) ? first[is - 1] + j : -1;
                q++;
            }
        rc = upload(st, omu.data(), omu.size(), &S.ord_mu, errmsg);
        if (!rc) rc = upload(st, ophi.data(), ophi.size(), &S.ord_phi, errmsg);
        if (!rc) rc = upload(st, ow.data(), ow.size(), &S.ord_w, errmsg);
        if (!rc) rc = upload(st, umu.data(), umu.size(), &S.up_mu, errmsg);
        if (!rc) rc = upload(st, uphi.data(), uphi.size(), &S.up_phi, errmsg);
        if (!rc) rc = upload(st, usrc.data(), usrc.size(), &S.up_src, errmsg);
        if (!rc && d->sfcgridrad) {
            const size_t n = (size_t)(nh + 1) * d->nbotpts;
            bool nonzero = false;
            for (size_t i = 0; i < n && !nonzero; i++) nonzero = d->sfcgridrad[i] != 0.0f;
            if (nonzero) rc = upload(st, d->sfcgridrad, n, &S.sfcgridrad, errmsg);
        }
        if (rc) { at3d_state_destroy(st); return rc; }
    }
    // cell and point records
    {
        const int *gp, *np, *tp; const short *cf; const float *gpos, *text;
        at3d_state tmp;   // staging allocations, freed right after
        rc = upload(&tmp, d->gridptr, (size_t)8 * d->ncells, &gp, errmsg);
        if (!rc) rc = upload(&tmp, d->neighptr, (size_t)6 * d->ncells, &np, errmsg);
        if (!rc) rc = upload(&tmp, d->treeptr, (size_t)2 * d->ncells, &tp, errmsg);
        if (!rc) rc = upload(&tmp, (const short *)d->cellflags, (size_t)d->ncells, &cf, errmsg);
        if (!rc) rc = upload(&tmp, d->gridpos, (size_t)3 * d->npts, &gpos, errmsg);
        if (!rc) rc = upload(&tmp, d->total_ext, (size_t)d->npts, &text, errmsg);
        int4 *cellrec = nullptr; float4 *ptrec = nullptr;
        if (!rc) rc = dalloc(st, (size_t)4 * d->ncells, &cellrec, errmsg);
        if (!rc) rc = dalloc(st, (size_t)d->npts, &ptrec, errmsg);
        if (!rc && launch_build_cellrec(d->ncells, gp, np, tp, cf, cellrec, 0) != cudaSuccess) { set_msg(errmsg, "cellrec launch failed"); rc = 4; }
        if (!rc && launch_build_ptrec(d->npts, gpos, text, ptrec, 0) != cudaSuccess) { set_msg(errmsg, "ptrec launch failed"); rc = 4; }
        cudaDeviceSynchronize();
        for (void *p : tmp.owned) at3d_free(p);
        tmp.owned.clear();
        if (rc) { at3d_state_destroy(st); return rc; }
        S.cellrec = cellrec; S.ptrec = ptrec;
    }
    // boundary radiances: copy BCRAD, then the Lambertian bottom boundary (RENDER, shdomsub4.f:201-209)
    {
        st->nbcrad = d->nstokes * (d->ntoppts + d->nbotpts * (d->sfctype1 == 'L' ? 1 : 1 + d->nang / 2));
        float *bc = nullptr;
        rc = dalloc(st, (size_t)st->nbcrad, &bc, errmsg);
        if (!rc && d->bcrad) { if (cudaMemcpy(bc, d->bcrad, st->nbcrad * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) rc = 4; }
        const float *fl = nullptr;
        at3d_state tmp;
        if (!rc) rc = upload(&tmp, d->fluxes, (size_t)2 * d->npts, &fl, errmsg);
        S.bcrad = bc; st->bcrad_dev = bc;
        if (!rc && launch_lambertian_boundary(S, fl, bc, 0) != cudaSuccess) { set_msg(errmsg, "boundary launch failed"); rc = 4; }
        cudaDeviceSynchronize();
        for (void *p : tmp.owned) at3d_free(p);
        tmp.owned.clear();
        if (!rc && d->bcrad) cudaMemcpy(d->bcrad, bc, st->nbcrad * sizeof(float), cudaMemcpyDeviceToHost);
        if (rc) { at3d_state_destroy(st); return rc; }
    }
    // spherical-harmonic source: TMS-corrected planar blocks + single-scatter lists
    {
        int *sscount = nullptr; int2 *ssent = nullptr;
        rc = dalloc(st, (size_t)d->npts, &sscount, errmsg);
        if (!rc) rc = dalloc(st, (size_t)d->npts * S.kmax, &ssent, errmsg);
        if (!rc) { cudaMemset(sscount, 0, (size_t)d->npts * sizeof(int)); }
        S.sscount = sscount; S.ssent = ssent;
        const int tms = (d->srctype != 'T' && d->deltam) ? 1 : 0;
        if (!rc) rc = prep_sh_array(st, tms, d->shptr, d->source, &S.srcrec, &S.shsrc, sscount, ssent, errmsg);
        if (!rc && d->rshptr && d->radiance) {
            rc = prep_sh_array(st, 0, d->rshptr, d->radiance, &S.radrec, &S.shrad, nullptr, nullptr, errmsg);
            st->nr_h.resize(d->npts);
            for (int i = 0; i < d->npts; i++) st->nr_h[i] = d->rshptr[i + 1] - d->rshptr[i];
        }
        if (!rc) {
            int4 *ptsrc = nullptr;
            rc = dalloc(st, (size_t)d->npts, &ptsrc, errmsg);
            if (!rc && launch_build_ptsrc(d->npts, S.kmax, S.srcrec, sscount, ssent, d->ncells, S.cellrec, S.ptrec, ptsrc, 0) != cudaSuccess) { set_msg(errmsg, "ptsrc launch failed"); rc = 4; }
            cudaDeviceSynchronize();
            S.ptsrc = ptsrc;
        }
        if (rc) { at3d_state_destroy(st); return rc; }
    }
    {
        unsigned long long *c = nullptr;
        rc = dalloc(st, 8, &c, errmsg);
        if (rc) { at3d_state_destroy(st); return rc; }
        cudaMemset(c, 0, 8 * sizeof(unsigned long long));
        st->counts_dev = c; S.counts = c;
        int *rcnt = nullptr;
        rc = dalloc(st, 4, &rcnt, errmsg);
        if (rc) { at3d_state_destroy(st); return rc; }
        st->ray_counter = rcnt;
    }
    *out = st;
    return 0;
}

extern "C" int at3d_state_get_counts(at3d_state *st, int64_t *counts, char *errmsg)
{
    if (!st || !counts) { set_msg(errmsg, "null argument"); return 1; }
    CUDA_TRY(cudaMemcpy(counts, st->counts_dev, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int at3d_state_get_bcrad(at3d_state *st, float *bcrad_host, char *errmsg)
{
    if (!st || !bcrad_host) { set_msg(errmsg, "null argument"); return 1; }
    CUDA_TRY(cudaMemcpy(bcrad_host, st->bcrad_dev, st->nbcrad * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

const char *ray_err_text(int code)
{
    switch (code) {
    case 1: return "INTEGRATE_1RAY: SO<0";
    case 2: return "RENDER: Level below domain";
    case 3: return "FIND_BOUNDARY_RADIANCE: Not at boundary";
    case 4: return "ADJOINT_INTEGRATE_1RAY: The maximum number of subgrid intervals for calculation of the radiance along the ray path has been exceeded.";
    case 5: return "DIRECT_BEAM_AND_PATHS_PROP: the walk toward the sun left the property grid";
    default: return "unknown ray error";
    }
}

int check_ray_err(at3d_state *st, cudaStream_t stream, char *errmsg)
{
    RayErr h;
    CUDA_TRY(cudaMemcpyAsync(&h, st->err.p, sizeof(RayErr), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (h.code) { set_msg(errmsg, "%s (ray %d)", ray_err_text(h.code), h.ray); return 1; }
    return 0;
}

// Stage the ray arrays on the device when they are host arrays.  For host rays the per-ray setup
// (make_ray_pack) is evaluated here with the host libm -- the reference's own -- so that the walk is
// bit-exact; device-resident rays get it evaluated in the kernel (packs = nullptr).
int stage_rays(at3d_state *st, const at3d_rays *rays, cudaStream_t stream, const float **camx,
               const float **camy, const float **camz, const double **cammu, const double **camphi,
               const RayPack **packs, char *errmsg)
{
    const size_t n = rays->nrays;
    *packs = nullptr;
    if (rays->memspace == AT3D_MEM_DEVICE) {
        *camx = rays->camx; *camy = rays->camy; *camz = rays->camz; *cammu = rays->cammu; *camphi = rays->camphi;
        *packs = (const RayPack *)rays->packs;              // host-libm setup records of at3d_make_ray_packs, or NULL
        return 0;
    }
    const size_t nb = n * (sizeof(RayPack) + 2 * sizeof(double)) + 64;
    CUDA_TRY(st->rays.reserve(nb));
    RayPack *dpk = (RayPack *)st->rays.p;
    double *dmu = (double *)(dpk + n), *dphi = dmu + n;
    if (n > st->packs_cap) {
        // pinned, so that the copy below is a true asynchronous DMA
        pinned_park(st->packs_h, st->packs_cap * sizeof(RayPack));
        st->packs_h = nullptr; st->packs_cap = 0;
        size_t got = 0;
        CUDA_TRY(pinned_take((void **)&st->packs_h, n * sizeof(RayPack), &got));
        st->packs_cap = got / sizeof(RayPack);
    }
    RayPack *hp = st->packs_h;
    const RayGeom g = st->geom;
    const float *hx = rays->camx, *hy = rays->camy, *hz = rays->camz;
    const double *hmu = rays->cammu, *hphi = rays->camphi;
    CUDA_TRY(cudaMemcpyAsync(dmu, rays->cammu, n * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(dphi, rays->camphi, n * sizeof(double), cudaMemcpyHostToDevice, stream));
    // in chunks, so that the DMA of one chunk overlaps the host libm work on the next
    const size_t chunk = 65536;
    for (size_t c0 = 0; c0 < n; c0 += chunk) {
        const long long c1 = (long long)(c0 + chunk < n ? c0 + chunk : n);
#pragma omp parallel for schedule(static) if (n > 4096)
        for (long long i = (long long)c0; i < c1; i++)
            make_ray_pack(g, (double)hx[i], (double)hy[i], (double)hz[i], hmu[i], hphi[i], hp[i]);
        CUDA_TRY(cudaMemcpyAsync(dpk + c0, hp + c0, (size_t)(c1 - (long long)c0) * sizeof(RayPack), cudaMemcpyHostToDevice, stream));
    }
    // cammu/camphi are pageable caller memory: staged when the calls return.  packs_h is pinned: its DMA completes
    // before this call returns (every entry point synchronises the stream before returning results), and it is only
    // rewritten by the next call on this state
    *camx = nullptr; *camy = nullptr; *camz = nullptr; *cammu = dmu; *camphi = dphi; *packs = dpk;
    return 0;
}

extern "C" int64_t at3d_ray_pack_bytes(void) { return (int64_t)sizeof(RayPack); }

extern "C" int at3d_make_ray_packs(at3d_state *st, const at3d_rays *rays, void *packs_out, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!st || !rays || !packs_out) { set_msg(errmsg, "null argument"); return 1; }
    if (rays->memspace != AT3D_MEM_HOST) { set_msg(errmsg, "at3d_make_ray_packs takes host ray arrays"); return 1; }
    const long long n = rays->nrays;
    const RayGeom g = st->geom;
    RayPack *hp = (RayPack *)packs_out;
    const float *hx = rays->camx, *hy = rays->camy, *hz = rays->camz;
    const double *hmu = rays->cammu, *hphi = rays->camphi;
#pragma omp parallel for schedule(static) if (n > 4096)
    for (long long i = 0; i < n; i++)
        make_ray_pack(g, (double)hx[i], (double)hy[i], (double)hz[i], hmu[i], hphi[i], hp[i]);
    return 0;
}

// Runs of rays with one common direction (orthographic views; host ray arrays only) of at least view_min_rays rays
// get a pre-evaluated view source; everything in between is merged into generic segments.
void view_segments(const at3d_state *st, const at3d_rays *rays, std::vector<size_t> &seg_start,
                   std::vector<size_t> &seg_len, std::vector<char> &seg_view)
{
    const size_t n = rays->nrays;
    seg_start.clear(); seg_len.clear(); seg_view.clear();
    if (rays->memspace == AT3D_MEM_HOST && st->view_min_rays > 0) {
        const double *hmu = rays->cammu, *hphi = rays->camphi;
        size_t i = 0;
        while (i < n) {
            size_t j = i + 1;
            while (j < n && hmu[j] == hmu[i] && hphi[j] == hphi[i]) j++;
            const bool view = (j - i) >= (size_t)st->view_min_rays;
            if (!seg_view.empty() && !view && !seg_view.back()) seg_len.back() += j - i;
            else { seg_start.push_back(i); seg_len.push_back(j - i); seg_view.push_back(view ? 1 : 0); }
            i = j;
        }
    } else if (n > 0) {
        seg_start.push_back(0); seg_len.push_back(n); seg_view.push_back(0);
    }
}

extern "C" int at3d_render(at3d_state *st, const at3d_rays *rays, float *stokes,
                           int correctinterpolate, int singlescatter, int nosurface,
                           const at3d_trace *trace, void *cuda_stream, double *kernel_ms, char *errmsg)
{
    if (errmsg) errmsg[0] = 0;
    if (!st || !rays || !stokes) { set_msg(errmsg, "null argument"); return 1; }
    std::lock_guard<std::mutex> lock(st->mu);
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const size_t n = rays->nrays;
    const int nst = st->S.nstokes;
    if (n == 0) return 0;
    const float *camx, *camy, *camz; const double *cammu, *camphi; const RayPack *packs;
    int rc = stage_rays(st, rays, stream, &camx, &camy, &camz, &cammu, &camphi, &packs, errmsg);
    if (rc) return rc;
    const bool host = rays->memspace == AT3D_MEM_HOST;
    float *out_d = stokes;
    if (host) { CUDA_TRY(st->out.reserve(n * nst * sizeof(float))); out_d = (float *)st->out.p; }
    int *tc = nullptr, *tn = nullptr, *ts = nullptr; int tcap = 0;
    if (trace && trace->cells) {
        tcap = trace->max_per_ray;
        if (host) {
            CUDA_TRY(st->trace.reserve(((size_t)tcap * n + 2 * n) * sizeof(int)));
            tc = (int *)st->trace.p; tn = tc + (size_t)tcap * n; ts = tn + n;
            CUDA_TRY(cudaMemsetAsync(tc, 0, ((size_t)tcap * n + 2 * n) * sizeof(int), stream));
        } else { tc = trace->cells; tn = trace->ncells; ts = trace->nsub; }
    }
    CUDA_TRY(st->err.reserve(sizeof(RayErr)));
    CUDA_TRY(cudaMemsetAsync(st->err.p, 0, sizeof(RayErr), stream));
    CUDA_TRY(cudaMemsetAsync(st->counts_dev, 0, 8 * sizeof(unsigned long long), stream));
    DevState S = st->S;
    const bool general_brdf = S.sfctype1 != 'L' && !nosurface;
    if (general_brdf) {
        CUDA_TRY(st->hits.reserve(n * sizeof(SurfHit)));
        CUDA_TRY(cudaMemsetAsync(st->hits.p, 0, n * sizeof(SurfHit), stream));
        S.surfhits = st->hits.p;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, stream); }
    // Orthographic views: runs of rays with one common direction (host ray arrays, NSTOKES=1).  Their SRCEXT is
    // evaluated once per grid point (view_source_kernel) instead of once per (ray, corner); same arithmetic per point.
    std::vector<size_t> seg_start, seg_len;
    std::vector<char> seg_view;
    view_segments(st, rays, seg_start, seg_len, seg_view);
    for (size_t sgi = 0; sgi < seg_start.size(); sgi++) {
        const size_t s0 = seg_start[sgi], sn = seg_len[sgi];
        DevState Sg = S;
        if (seg_view[sgi]) {
            CUDA_TRY(st->viewsrc.reserve((size_t)S.npts * nst * sizeof(float)));
            CUDA_TRY(launch_view_source(S, st->packs_h[s0], rays->cammu[s0], rays->camphi[s0], singlescatter,
                                        (float *)st->viewsrc.p, stream));
            Sg.viewsrc = (const float *)st->viewsrc.p;
        }
        if (Sg.surfhits) Sg.surfhits = (SurfHit *)Sg.surfhits + s0;
        Sg.ray_base = (int)s0;
        CUDA_TRY(launch_forward(Sg, (int)sn, camx ? camx + s0 : nullptr, camy ? camy + s0 : nullptr,
                                camz ? camz + s0 : nullptr, cammu + s0, camphi + s0, packs ? packs + s0 : nullptr,
                                out_d + (size_t)nst * s0, nullptr, nullptr, 1,
                                correctinterpolate, singlescatter, nosurface, 0, tc ? tc + (size_t)tcap * s0 : nullptr, tcap,
                                tn ? tn + s0 : nullptr, ts ? ts + s0 : nullptr,
                                (RayErr *)st->err.p, st->ray_counter, nullptr, stream));
    }
    if (general_brdf)
        CUDA_TRY(launch_surface(S, (int)n, (const SurfHit *)st->hits.p, cammu, camphi, out_d,
                                (RayErr *)st->err.p, stream));
    if (kernel_ms) cudaEventRecord(e1, stream);
    if (host) {
        CUDA_TRY(cudaMemcpyAsync(stokes, out_d, n * nst * sizeof(float), cudaMemcpyDeviceToHost, stream));
        if (tc) {
            CUDA_TRY(cudaMemcpyAsync(trace->cells, tc, (size_t)tcap * n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(trace->ncells, tn, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(trace->nsub, ts, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
    }
    rc = check_ray_err(st, stream, errmsg);
    if (kernel_ms) {
        float ms = 0.0f;
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        *kernel_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return rc;
}
