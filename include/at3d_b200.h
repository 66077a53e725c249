/*
 * at3d_b200.h -- C-ABI of the B200-native SHDOM hot path (libat3d_b200.so).
 *
 * Drop-in boundary for the f2py extension `at3d.core` of CloudTomography/AT3D, for the hot path
 * only.  Each entry point names the reference routine it replaces (paths relative to the AT3D
 * checkout).  Plain pointers and sizes, no torch types, `int` return code (0 ok, 1 generic error,
 * 2 out of spherical-harmonic memory, 3 unsupported configuration, 4 CUDA error) plus a
 * caller-supplied `char errmsg[600]` -- the reference's IERR/ERRMSG convention
 * (at3d/checks.py:300-311).  The library never aborts and never frees caller memory.
 *
 * Array layout is the reference's own: Fortran (column-major) order, 1-based index CONTENTS,
 * REAL=float, DOUBLE PRECISION=double, INTEGER=int32, INTEGER*2=int16.
 *
 * There is NO CPU fallback: every compute entry point launches sm_100a kernels and returns
 * code 4 if no CUDA device is usable.
 */
#ifndef AT3D_B200_H
#define AT3D_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT3D_ERRMSG_LEN 600

/* Where the pointers of a call live. */
#define AT3D_MEM_HOST   0
#define AT3D_MEM_DEVICE 1

/*
 * The solved SHDOM state that RENDER / LEVISAPPROX_GRADIENT read (all read-only except bcrad).
 * Field names are the Fortran dummy-argument names of RENDER (src/polarized/shdomsub4.f:93-109)
 * and LEVISAPPROX_GRADIENT (shdomsub4.f:288-317), i.e. the keyword names at3d passes at
 * at3d/solver.py:681-759 and at3d/gradient.py:262-398.  Pointers are HOST pointers.
 */
typedef struct at3d_state_desc {
    int32_t nstokes, nstleg, nx, ny, nz, npts, ncells;
    int32_t ml, mm, nlm, nleg, numphase, npart, maxnmicro;
    int32_t bcflag, ipflag;
    int32_t nmu, nphi0max, nang;
    int32_t maxnbc, ntoppts, nbotpts, nsfcpar;
    int32_t nscatangle, nstphase;
    int32_t deltam;             /* LOGICAL */
    int32_t srctype;            /* 'S' | 'T' | 'B'  (gradient entry points: 'S' only, others -> code 3) */
    int32_t units;              /* 'R' | 'T' | 'B' */
    int32_t sfctype0, sfctype1; /* SFCTYPE(1:1), SFCTYPE(2:2): FL, VL, VW, VD, VO, VR, VM (gradient: FL, VL) */
    int32_t interp_new;         /* INTERPMETHOD(2:2) == 'N' */
    float solarmu, solaraz, solarflux, wavelen, gndtemp, gndalbedo, phasemax;
    float waveno0, waveno1;
    double tautol, transcut;
    const int32_t *gridptr;     /* [8,ncells]  */
    const int32_t *neighptr;    /* [6,ncells]  */
    const int32_t *treeptr;     /* [2,ncells]  */
    const int16_t *cellflags;   /* [ncells]    */
    const float *xgrid, *ygrid, *zgrid;   /* [nx+1] (periodic) or [nx] (open), [ny+1]|[ny], [nz] */
    const float *gridpos;       /* [3,npts]    */
    const float *extinct;       /* [npts,npart] */
    const float *albedo;        /* [npts,npart] */
    const float *total_ext;     /* [npts] */
    const float *legen;         /* [nstleg,0:nleg,numphase] */
    const int32_t *iphase;      /* [8*maxnmicro,npts,npart] */
    const float *phaseinterpwt; /* [8*maxnmicro,npts,npart] */
    const float *dirflux;       /* [npts] */
    const float *fluxes;        /* [2,npts] */
    const int32_t *shptr;       /* [npts+1] */
    const float *source;        /* [nstokes, shptr[npts]] */
    const int32_t *rshptr;      /* [npts+2]  (gradient only; may be NULL for render) */
    const float *radiance;      /* [nstokes, rshptr[npts]] (gradient only) */
    const float *ylmsun;        /* [nstleg,nlm] */
    const float *phasetab;      /* [nstphase,numphase,nscatangle] */
    const float *planck;        /* [npts,npart] (compute_source only; may be NULL) */
    const float *temp;          /* unused (thermal) */
    const int32_t *nphi0;       /* [nmu] */
    const float *mu;            /* [nmu] */
    const float *phi;           /* [nmu,nphi0max] */
    const float *wtdo;          /* [nmu,nphi0max] */
    const float *skyrad;        /* [nstokes,nmu/2,nphi0max] */
    const int32_t *bcptr;       /* [maxnbc,2] */
    float *bcrad;               /* [nstokes, ntoppts+nbotpts(...)]; bottom part rewritten like RENDER does */
    const float *sfcgridparms;  /* [nsfcpar,nbotpts] */
    const float *sfcgridrad;    /* [nang/2+1, nbotpts] surface emission (may be null / zero) */
} at3d_state_desc;

/* Sensor rays: CAMX..CAMPHI of RENDER (shdomsub4.f:164-166).  memspace says where they live. */
typedef struct at3d_rays {
    int32_t nrays;
    int32_t memspace;           /* AT3D_MEM_HOST | AT3D_MEM_DEVICE */
    const float *camx, *camy, *camz;
    const double *cammu, *camphi;
    const void *packs;          /* optional, AT3D_MEM_DEVICE only: device array of nrays per-ray setup records written by
                                   at3d_make_ray_packs (evaluated with the HOST libm, the reference's own, so that the walk
                                   of device-resident rays is bit-exact too); NULL: the kernels evaluate the setup themselves
                                   with the CUDA math library (values agree, cell sequences may differ in rare ties) */
} at3d_rays;


/* The extra inputs of LEVISAPPROX_GRADIENT (shdomsub4.f:299-317); HOST pointers except the
 * per-ray / per-pixel arrays, which follow rays->memspace. */
typedef struct at3d_grad_desc {
    int32_t npix, maxpg, numder, dnumphase, deriv_maxnmicro, longest_path_pts;
    int32_t nuncertainty, maxsubgridints, exact_single_scatter, singlescatter;
    int32_t costfunc_ll;        /* COSTFUNC: 0 'L2', 1 'LL' */
    double extmin, scatmin;
    const int32_t *partder;     /* [numder] */
    const int32_t *doexact;     /* [numder] */
    const float *measurements;      /* [nstokes,npix]      (rays->memspace) */
    const double *uncertainties;    /* [nunc,nunc,npix]    (rays->memspace) */
    const int32_t *rays_per_pixel;  /* [npix]              (rays->memspace) */
    const double *ray_weights;      /* [nrays]             (rays->memspace) */
    const double *stokes_weights;   /* [nstokes,npix]      (rays->memspace) */
    const float *dext, *dalb;       /* [maxpg,numder] */
    const float *dextm;             /* [maxpg,numder] */
    const float *dalbm, *dfj;       /* [8,npts,numder] */
    const float *optinterpwt;       /* [8,npts] */
    const int32_t *interpptr;       /* [8,npts] */
    const float *dleg;              /* [nstleg,0:nleg,dnumphase] */
    const float *dphasetab;         /* [nstphase,dnumphase,nscatangle] */
    const int32_t *diphasep;        /* [deriv_maxnmicro,maxpg,numder] */
    const float *dphasewtp;         /* [deriv_maxnmicro,maxpg,numder] */
    const int32_t *iphasep;         /* [maxnmicro,maxpg,npart] */
    const float *phasewtp;          /* [maxnmicro,maxpg,npart] */
    const float *extinctp, *albedop;/* [maxpg,npart] */
    const float *dtemp;             /* unused (thermal) */
    const float *dpath;             /* [longest_path_pts,npts]; NULL (with dptr) selects the streaming direct-beam term */
    const int32_t *dptr;            /* [longest_path_pts,npts] */
    /* Streaming direct-beam derivative (exact_single_scatter with dpath == dptr == NULL): instead of reading the dense
     * lists of MAKE_DIRECT_DERIVATIVE (shdomsub5.f:1553-2004; LONGEST_PATH_PTS x NPTS entries, ~125 GB at 6.5 M points),
     * the gradient call walks from every grid point with a non-zero beam weight toward the sun through the property grid
     * and accumulates the same terms in the same order (COMPUTE_DIRECT_BEAM_DERIV_ADJOINT, shdomsub4.f:4117-4143).
     * beam_d / beam_i are the constants at3d_make_direct returned (out_d[13], out_i[5]). */
    int32_t beam_npx, beam_npy, beam_npz;
    float beam_xstart, beam_ystart;
    const float *beam_zlevels;      /* [beam_npz] property-grid levels */
    const double *beam_d;           /* [13] */
    const int32_t *beam_i;          /* [5]  */
} at3d_grad_desc;

/* Optional per-ray trace of the visited cells (parity tests of the bit-exact indexing). */
typedef struct at3d_trace {
    int32_t max_per_ray;
    int32_t *cells;             /* [max_per_ray,nrays] (rays->memspace) */
    int32_t *ncells;            /* [nrays] */
    int32_t *nsub;              /* [nrays] number of sub-intervals integrated */
} at3d_trace;

typedef struct at3d_state at3d_state;   /* opaque: the state resident in HBM */

/* ---- library / device ---- */
const char *at3d_b200_version(void);
int at3d_device_count(void);
int at3d_set_device(int device);
/* Memory reuse for loops that build and drop states (an inversion evaluates a new medium per step; the reference rebuilds
 * its solvers, at3d/medium.py:1813-1831).  Off (default): cudaMalloc / cudaFree.  On: state arrays, derivative tables,
 * solver objects and per-call arenas come from the CUDA driver's stream-ordered pool, which keeps up to AT3D_B200_POOL_GB
 * (default 64) of freed memory reserved, and the large streaming buffers are parked between owners; a destroyed state
 * hands its memory to the next one.  Returns the previous setting.  AT3D_B200_POOL_GB > 0 in the environment switches it
 * on from the start.  at3d_trim_memory returns everything that is kept to the system. */
int at3d_set_memory_reuse(int on);
int at3d_trim_memory(void);

/* ---- state residency (replaces the per-call array marshalling of f2py) ---- */
int at3d_state_create(const at3d_state_desc *desc, at3d_state **out, char *errmsg);
int at3d_state_attach_gradient(at3d_state *st, const at3d_grad_desc *g, char *errmsg);
int at3d_state_destroy(at3d_state *st);
int64_t at3d_state_bytes(const at3d_state *st);   /* HBM bytes held */
/* RENDER returns BCRAD as an in/out argument (at3d/solver.py:747): copy it back, [nstokes,ntoppts+nbotpts]
 * (general BRDF surfaces: [nstokes, ntoppts + nbotpts*(1+nang/2)], the stored downwelling radiances included) */
int at3d_state_get_bcrad(at3d_state *st, float *bcrad_host, char *errmsg);
/* Work counters of the last at3d_render / at3d_levisapprox_gradient (adjoint pass) on this state, for
 * roofline accounting: [0] cells visited, [1] grid points evaluated, [2] sum of NS over them,
 * [3] sum of NR (gradient), [4] sub-intervals integrated, [5] rays marched,
 * [6] rays that ended on a general-BRDF surface (at3d_render), [7] reserved. */
int at3d_state_get_counts(at3d_state *st, int64_t *counts /*[8]*/, char *errmsg);

/* ---- a7: YLMALL (shdomsub2.f:4244) and PRECOMPUTE_PHASE_CHECK[_GRAD] (shdomsub4.f:2388,2493) ---- */
int at3d_ylmall(int transpose, float mu, float phi, int ml, int mm, int nstleg, float *yr /*host*/,
                char *errmsg);
int at3d_precompute_phase_check(int nscatangle, int numphase, int nstphase, int nstokes, int ml,
                                int nlm, int nstleg, int nleg, const float *legen, float *phasetab,
                                int deltam, int negcheck, int grad /*0: LEGEN/(2l+1), 1: DLEG*/,
                                char *errmsg);

/* ---- a1: COMPUTE_SOURCE (shdomsub1.f:967).  All pointers HOST; in/out arrays as the reference:
 *      shptr, source, oshptr, delsource are updated in place. ---- */
int at3d_compute_source(const at3d_state_desc *desc, int fixsh, float shacc, int maxiv,
                        int first, int accelflag, int newmethod,
                        int32_t *shptr, float *source, int32_t *oshptr, float *delsource,
                        float *deljdot, float *deljold, float *deljnew, float *jnorm,
                        double *kernel_ms /*optional: device time of the kernels*/, char *errmsg);

/* ---- a1 on device-resident arrays (what the solution iterations of a GPU-resident solver call every iteration,
 *      shdomsub1.f:600-640): every array below is a DEVICE pointer with the reference's layout.  SOURCE / SHPTR are
 *      double-buffered by the caller (source_new may alias source_old only with fixsh); DELSOURCE is rewritten at the
 *      OLD SHPTR offsets and delsource_new may alias delsource_old.  Only the four norms (DELJDOT, DELJOLD, DELJNEW,
 *      JNORM), the new SHPTR(NPTS+1) and the error flag come back to the host.  The mixed Legendre rows of the points
 *      (NEWMETHOD, shdomsub1.f:1089-1141) are kept between calls on the same property arrays: pass properties_changed=1
 *      after the optical properties were modified in place. ---- */
typedef struct {
    int32_t npts, nstokes, nstleg, nlm, ml, mm, nleg, npart, maxnmicro, numphase, deltam, interp_new;
    char srctype;                 /* 'S', 'T' or 'B' */
    float phasemax, solarmu;
    const float *extinct, *albedo;            /* [npts, npart] */
    const float *total_ext;                   /* [npts] */
    const float *legen;                       /* [nstleg, nleg+1, numphase] */
    const int32_t *iphase;                    /* [8*maxnmicro, npts, npart] */
    const float *phaseinterpwt;               /* [8*maxnmicro, npts, npart] */
    const float *dirflux;                     /* [npts] */
    const int32_t *rshptr;                    /* [npts+1] */
    const float *radiance;                    /* [nstokes, rshptr[npts]] */
    const float *ylmsun;                      /* [nstleg, nlm] */
    const float *planck;                      /* [npts, npart] or NULL */
} at3d_cs_device_desc;

int at3d_compute_source_device(const at3d_cs_device_desc *desc, int fixsh, float shacc, int64_t maxiv, int first,
                               int accelflag, const int32_t *shptr_old, const float *source_old,
                               const int32_t *oshptr_old, const float *delsource_old, float *delsource_new,
                               int32_t *shptr_new, float *source_new, int64_t source_new_capacity /*entries per Stokes component*/,
                               int properties_changed, float *norms /*host [4]*/, int32_t *total_new /*host*/,
                               double *kernel_ms /*optional*/, char *errmsg);

/* per-ray setup records of host rays (direction cosines, clipped entry point, scattering-angle interpolation; what RENDER
 * computes at the top of its ray loop, shdomsub4.f:213-236): packs_out is a HOST array of nrays * at3d_ray_pack_bytes()
 * bytes that the caller uploads next to its device-resident ray arrays (at3d_rays.packs). */
int64_t at3d_ray_pack_bytes(void);
int at3d_make_ray_packs(at3d_state *st, const at3d_rays *rays_host, void *packs_out, char *errmsg);

/* ---- a2/a3/a4/a5/a6: RENDER (shdomsub4.f:93) ---- */
int at3d_render(at3d_state *st, const at3d_rays *rays, float *stokes /*[nstokes,nrays], rays->memspace*/,
                int correctinterpolate, int singlescatter, int nosurface,
                const at3d_trace *trace /*optional*/, void *cuda_stream /*optional*/,
                double *kernel_ms /*optional*/, char *errmsg);

/* ---- a8..a14: LEVISAPPROX_GRADIENT, MAKEJACOBIAN=.FALSE. (shdomsub4.f:288) ----
 * gradout [maxpg,numder] f64, cost [1] f64, stokesout [nstokes,npix] f32 follow rays->memspace. */
int at3d_levisapprox_gradient(at3d_state *st, const at3d_rays *rays, const at3d_grad_desc *g,
                              double *gradout, double *cost, float *stokesout,
                              const at3d_trace *trace /*optional*/, void *cuda_stream /*optional*/,
                              double *kernel_ms /*optional [8]: forward, derivative pass (weights + apply + pair sums), beam, total, weights, pair sort + sums; [6..7] reserved*/,
                              char *errmsg);

/* ---- a10: LEVISAPPROX_GRADIENT with MAKEJACOBIAN=.TRUE. (shdomsub4.f:536-631, GRAD_INTEGRATE_1RAY :811) ----
 * As at3d_levisapprox_gradient, plus JACOBIAN(NSTOKES,NUMDER,NUM_JACOBIAN_PTS,NPIX) f32 at the property points
 * JACOBIANPTR(NUM_JACOBIAN_PTS) (1-based); both follow rays->memspace.  One derivative pass per pixel and
 * Stokes component (the reference's slow path as well). */
int at3d_levisapprox_gradient_jacobian(at3d_state *st, const at3d_rays *rays, const at3d_grad_desc *g,
                                       double *gradout, double *cost, float *stokesout,
                                       int num_jacobian_pts, const int32_t *jacobianptr, float *jacobian,
                                       void *cuda_stream /*optional*/, char *errmsg);

/* ---- a15: PREPARE_DERIV_INTERPS (shdomsub4.f:2917); HOST pointers ---- */
int at3d_prepare_deriv_interps(const at3d_state_desc *desc, int npx, int npy, int npz, int maxpg,
                               float delx, float dely, float xstart, float ystart,
                               const float *zlevels, const at3d_grad_desc *g,
                               float *optinterpwt, int32_t *interpptr,
                               float *dalbm, float *dextm, float *dfj, char *errmsg);

/* ---- a14 precompute: MAKE_DIRECT (src/polarized/shdomsub2.f:393, at3d/solver.py:2827-2871); HOST pointers.
 *      out_d[13] = CX,CY,CZ,CXINV,CYINV,CZINV,EPSS,EPSZ,XDOMAIN,YDOMAIN,UNIFORMZLEV,DELXD,DELYD
 *      out_i[5]  = IPDIRECT,DI,DJ,DK,LONGEST_PATH_PTS ---- */
int at3d_make_direct(int npts, int bcflag, int ipflag, int deltam, int ml, int nstleg, int nlegp,
                     float solarflux, float solarmu, float solaraz, const float *gridpos,
                     int npx, int npy, int npz, float delx, float dely, float xstart, float ystart,
                     const float *zlevels, const float *extinctp, const float *albedop,
                     const float *legenp /*[nstleg,0:nlegp,numphase]*/, int numphase,
                     const int32_t *iphasep, const float *phasewtp, int maxnmicro, int npart,
                     int nzckd, const float *zckd, const float *gasabs,
                     float *extdirp /*[maxpg]*/, float *dirflux /*[npts]*/, double *out_d, int32_t *out_i,
                     char *errmsg);

/* ---- a14 precompute: MAKE_DIRECT_DERIVATIVE (src/shdomsub5.f:1553); HOST pointers ---- */
int at3d_make_direct_derivative(int npts, int bcflag, int npx, int npy, int npz,
                                float delx, float dely, float xstart, float ystart,
                                const float *gridpos, const float *zlevels,
                                int ipdirect, int di, int dj, int dk,
                                double cx, double cy, double cz,
                                double cxinv, double cyinv, double czinv,
                                double epss, double epsz, double xdomain, double ydomain,
                                double uniformzlev, double delxd, double delyd,
                                float *dpath, int32_t *dptr, int longest_path_pts, char *errmsg);

/* ---- a16: average_subpixel_rays (src/util.f90:484); HOST pointers ---- */
int at3d_average_subpixel_rays(int npixels, int nrays, int nstokes, const float *weighted_stokes,
                               const int32_t *pixel_index, float *observables, char *errmsg);

/* ---- a9: UPDATE_COSTFUNCTION (shdomsub4.f:13); HOST pointers ---- */
int at3d_update_costfunction(const double *stokesout, const double *raygrad_pixel,
                             double *gradout, double *cost, const double *uncertainties,
                             int costfunc_ll, int nstokes, int maxpg, int numder,
                             const double *measurement, int nuncertainty, char *errmsg);

/* ---- next: the SH <-> discrete-ordinate transforms of PATH_INTEGRATION ----
 * SH_TO_DO / DO_TO_SH (src/polarized/shdomsub1.f:2789-3260) with MAKE_SH_DO_COEF's tables (shdomsub2.f:1146-1220), for
 * ALL zenith angles at once.  dofield is DOFIELD(NPTS, NSTOKES, NANG) (Fortran order), ordinate IANG = (IMU, IPHI) in
 * PATH_INTEGRATION order; indata/outdata are (NSTOKES, *) SH arrays addressed by SHPTR / RSHPTR.  mu[nmu], phi[nmu,
 * nphi0max], wtmu[nmu], nphi0[nmu] come from MAKE_ANGLE_SET.  at3d_do_to_sh returns the sum over all zenith angles
 * (the reference accumulates one angle per call into a zeroed RADIANCE, shdomsub1.f:2043-2046). Host pointers. */
int at3d_sh_to_do(int npts, int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max,
                  const int32_t *nphi0, const float *mu, const float *phi, const float *wtmu,
                  const int32_t *shptr, const float *indata, float *dofield, double *kernel_ms, char *errmsg);
int at3d_do_to_sh(int npts, int nstokes, int nstleg, int ml, int mm, int nlm, int nmu, int nphi0max,
                  const int32_t *nphi0, const float *mu, const float *phi, const float *wtmu,
                  const int32_t *rshptr, const float *dofield, float *outdata, double *kernel_ms, char *errmsg);

/* PATH_INTEGRATION (src/polarized/shdomsub1.f:1836-2167) for independent-pixel grids (IPFLAG=3, unsplit base grid):
 * SH_TO_DO of SOURCE, the BACK_INT_GRID1D sweeps of all ordinates with the top / Lambertian / general-BRDF boundary
 * conditions, hemispheric fluxes, DO_TO_SH into RADIANCE (addressed by RSHPTR).  bcrad is BCRAD (in: unused; out: top
 * radiances, upwelling bottom radiances and, for general BRDF surfaces, the stored downwelling radiances); fluxes is
 * FLUXES(2,NPTS).  Host pointers; desc supplies the grid, TOTAL_EXT, DIRFLUX, the ordinate set and the surface. */
int at3d_path_integration_ip(const at3d_state_desc *desc, const float *wtmu, const int32_t *shptr, const float *source,
                             const int32_t *rshptr, float *radiance, float *fluxes, float *bcrad,
                             double *kernel_ms, char *errmsg);

/* ---- f2: TRANSFER_PA_TO_GRID: TRILIN_INTERP_PROP (src/polarized/shdom90.f90:17-346) + the delta-M scaling of
 * PREPARE_PROP (src/polarized/shdomsub2.f:479-608), INTERPMETHOD 'ON' (at3d/solver.py:2405-2447).  HOST pointers.
 * Property grid: extinctp/albedop[maxpg,npart], iphasep/phasewtp[maxnmicro,maxpg,npart]; ftab[numphase] =
 * LEGEN(1,ML+1,.) before the delta-M subtraction (needed when deltam).  Outputs on the RTE grid: extinct/albedo
 * [npts,npart], total_ext[npts], iphase/phaseinterpwt[8*maxnmicro,npts,npart]. ---- */
int at3d_transfer_pa_to_grid(int npts, const float *gridpos, int npx, int npy, int npz, float delx, float dely,
                             float xstart, float ystart, const float *zlevels, int npart, int maxnmicro,
                             const float *extinctp, const float *albedop, const int32_t *iphasep,
                             const float *phasewtp, int numphase, const float *ftab, int ml, int deltam,
                             float phasemax, float *extinct, float *albedo, float *total_ext, int32_t *iphase,
                             float *phaseinterpwt, char *errmsg);

/* ---- f1: PATH_INTEGRATION on 3-D grids (fixed grid, base or already split) ----
 * Replaces PATH_INTEGRATION (src/polarized/shdomsub1.f:1836-2167) with BACK_INT_GRID3D[_UNPOL] (:3354-4036) in the
 * order of SWEEPING_ORDER (:3261-3352), for IPFLAG 0 or 1 and periodic or open boundaries; IPFLAG=2 (independent pixels
 * in Y, e.g. ny = 1) takes BACK_INT_GRID2D (:4039-4293), IPFLAG=3 the independent columns of at3d_path_integration_ip.  The solver object holds
 * what does not change over the solution iterations on a fixed grid (topology, sweep order, ordinate geometry, the
 * SH <-> ordinate transform tables, the two discrete-ordinate fields); transmin is TRANSMIN (at3d default 1.0).
 * at3d_solver_path_integration has the argument meaning of at3d_path_integration_ip. */
typedef struct at3d_solver at3d_solver;
/* SWEEPING_ORDER (:3261-3352) alone, on the host: sweepord is SWEEPORD(NPTS,NOCT), NOCT = 8 (4 for IPFLAG=2), entries
 * cell<<3 | corner-1; no device needed. */
int at3d_sweeping_order(const at3d_state_desc *desc, int32_t *sweepord, char *errmsg);
int at3d_solver_create(const at3d_state_desc *desc, const float *wtmu, float transmin, at3d_solver **out, char *errmsg);
/* A new medium on the solver's grid (what an optimisation step changes; the reference rebuilds its solver objects,
 * at3d/medium.py:1813-1831): TOTAL_EXT, DIRFLUX, SFCGRIDPARMS, SKYRAD, GNDALBEDO / GNDTEMP of `desc` replace the ones the
 * object holds; topology, sweep order and the sorted sweep plan are kept (TRANSMIN >= 1 only: code 3 otherwise).  The
 * other optical arrays are read from the desc passed to at3d_solver_solve. */
int at3d_solver_update_medium(at3d_solver *sv, const at3d_state_desc *desc, char *errmsg);
int at3d_solver_path_integration(at3d_solver *sv, const int32_t *shptr, const float *source, const int32_t *rshptr,
                                 float *radiance, float *fluxes, float *bcrad, double *kernel_ms, char *errmsg);
/* SOLUTION_ITERATIONS on a fixed grid (src/polarized/shdomsub1.f:445-822 without SPLIT_GRID; at3d/solver.py:279
 * RTE.solve with split_accuracy=0), device-resident: RADIANCE_TRUNCATION (:1615), PATH_INTEGRATION, COMPUTE_SOURCE
 * (:967), CALC_ACCEL_SOLCRIT (src/shdom_nompi.f:317), ACCELERATE_SOLUTION (:1807) loop in HBM; the first guess is a zero
 * radiance field.  desc supplies the optical properties of the grid the solver object was created for.  Outputs (host):
 * shptr[npts+1], source[nstokes,maxiv], rshptr[npts+2], radiance[nstokes,maxiv+npts], fluxes[2,npts], bcrad; iterations
 * done, final SOLCRIT, ms[3] = CUDA-event time in PATH_INTEGRATION, in COMPUTE_SOURCE, of the whole loop.
 * Returns 2 (the reference's IERR=2) when MAXIV is too small. */
int at3d_solver_solve(at3d_solver *sv, const at3d_state_desc *desc, int maxiter, float solacc, float shacc, int accelflag,
                      int highorderrad, int iterfixsh, int maxiv, int32_t *shptr, float *source, int32_t *rshptr,
                      float *radiance, float *fluxes, float *bcrad, int32_t *iters, float *solcrit, double *ms, char *errmsg);
/* The same loop continued from a solution (restore != 0): shptr / source / rshptr / radiance hold, on entry, the solution of a
 * nearby medium on the same grid -- what RTE.load_solution + INIT_SOLUTION with INRADFLAG=.FALSE. set up in the reference
 * (at3d/solver.py:2654-2666, shdomsub1.f:356-391: no first guess, OSHPTR = SHPTR, DELSOURCE = 0) and what an optimisation
 * loop does between its steps.  restore = 0 is at3d_solver_solve. */
int at3d_solver_solve_from(at3d_solver *sv, const at3d_state_desc *desc, int maxiter, float solacc, float shacc, int accelflag,
                           int highorderrad, int iterfixsh, int maxiv, int restore, int32_t *shptr, float *source,
                           int32_t *rshptr, float *radiance, float *fluxes, float *bcrad, int32_t *iters, float *solcrit,
                           double *ms, char *errmsg);
int at3d_solver_destroy(at3d_solver *sv);

/* ---- f3: the adaptive solve: INIT_SOLUTION + SOLUTION_ITERATIONS with SPLIT_GRID ----
 * Replaces INIT_SOLUTION (src/polarized/shdomsub1.f:113-443: MAKE_DIRECT, INIT_RADIANCE + EDDRTF shdomsub2.f:614-1057,
 * first COMPUTE_SOURCE, BOUNDARY_PNTS) and SOLUTION_ITERATIONS (:445-822) including SPLIT_GRID, DIVIDE_CELL,
 * INTERPOLATE_POINT, GRID_SMOOTH_TEST (:4703-5902) -- RTE.solve of at3d/solver.py:279 with the default
 * split_accuracy > 0.  The SH arrays stay in HBM over the whole solve; the cell tree is split on the host between
 * iterations (see csrc/at3d_adapt.cu), its criterion evaluated on the GPU.
 *
 * at3d_prop_desc: the property grid (ShdomPropertyArrays, at3d/solver.py:25-104).
 * at3d_adapt_io: the arrays SOLUTION_ITERATIONS has intent(in,out), with the capacities of RTE._setup_memory
 * (at3d/solver.py:2286-2322); HOST pointers; point arrays [maxig] or [maxig,npart] (leading dimension maxig).  On entry
 * they hold the base grid and the optical properties on it (TRANSFER_PA_TO_GRID); on return the split grid and the
 * solution, with npts / ncells / ntoppts / nbotpts / iters / solcrit / splitcrit filled.
 * desc: scalars and constant tables (LEGEN, YLMSUN, ordinates, SKYRAD, XGRID/YGRID/ZGRID, SFCTYPE ...); its grid and
 * point arrays are ignored in favour of io's. */
typedef struct {
    int32_t npx, npy, npz, numphase, nlegp, maxnmicro, npart, nzckd, nstleg;
    float delx, dely, xstart, ystart;
    const float *zlevels;      /* [npz] */
    const float *tempp;        /* [maxpg] or NULL */
    const float *extinctp;     /* [maxpg,npart] */
    const float *albedop;      /* [maxpg,npart] */
    const float *legenp;       /* [nstleg,0:nlegp,numphase] */
    const int32_t *iphasep;    /* [maxnmicro,maxpg,npart] */
    const float *phasewtp;     /* [maxnmicro,maxpg,npart] */
    const float *zckd, *gasabs;/* [nzckd] */
} at3d_prop_desc;

typedef struct {
    int32_t maxig, maxic, maxiv, maxido, maxnbc, maxbcrad, nbpts, nbcells;
    int32_t maxiter, accelflag, highorderrad, iterfixsh, inradflag;
    float solacc, splitacc, shacc, transmin;
    int32_t nxsfc, nysfc;              /* variable surfaces: SFCPARMS[nsfcpar, nxsfc+1, nysfc+1] */
    float delxsfc, delysfc;
    const float *sfcparms;
    float *gridpos;                    /* [3,maxig] */
    int32_t *gridptr, *neighptr, *treeptr;   /* [8|6|2, maxic] */
    int16_t *cellflags;                /* [maxic] */
    float *temp, *planck, *extinct, *albedo, *total_ext, *dirflux, *fluxes;   /* fluxes [2,maxig] */
    int32_t *iphase;                   /* [8*maxnmicro, maxig, npart] */
    float *phaseinterpwt;
    int32_t *shptr, *rshptr;           /* [maxig+1], [maxig+2] */
    float *source, *radiance;          /* [nstokes,maxiv], [nstokes,maxiv+maxig] */
    int32_t *bcptr;                    /* [maxnbc,2] */
    float *bcrad;                      /* [nstokes,maxbcrad] */
    float *sfcgridparms;               /* [nsfcpar,maxnbc] (variable surfaces) */
    float *extdirp;                    /* [maxpg] out */
    int32_t npts, ncells, ntoppts, nbotpts, iters, nsplit_calls;
    float solcrit, splitcrit;
} at3d_adapt_io;

int at3d_solve_adaptive(const at3d_state_desc *desc, const at3d_prop_desc *prop, const float *wtmu, at3d_adapt_io *io,
                        double *ms /*[4]: PATH_INTEGRATION, COMPUTE_SOURCE, SPLIT_GRID + sweep set-up, whole loop*/, char *errmsg);

#ifdef __cplusplus
}
#endif
#endif
