/*
 * shdom_oracle.h -- CPU parity oracle for the AT3D / polarized-SHDOM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in at3d_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / reported CPU baseline.
 *
 * It is a plain-C restatement (gcc -O2 -ffp-contract=off, REAL=float,
 * DOUBLE PRECISION=double, INTEGER=int, INTEGER*2=short) of these reference routines
 * (paths relative to /root/reference):
 *   src/polarized/shdomsub2.f:4244-4645  YLMALL / YLMALL_UNPOL / WIGNERFCT*
 *   src/polarized/shdomsub4.f:2388-2585  PRECOMPUTE_PHASE_CHECK[_GRAD]
 *   src/polarized/shdomsub2.f:4043-4206  LOCATE_GRID_CELL
 *   src/polarized/shdomsub1.f:4470-4522  NEXT_CELL
 *   src/polarized/shdomsub2.f:2311-2743  INTEGRATE_1RAY
 *   src/polarized/shdomsub2.f:2748-2863  FIND_BOUNDARY_RADIANCE
 *   src/polarized/shdomsub2.f:2868-3192  COMPUTE_SOURCE_1CELL[_UNPOL]
 *   src/polarized/shdomsub2.f:3277-3314  ROTATE_POL_PLANE
 *   src/polarized/shdomsub1.f:823-1611   CALC_SOURCE_PNT[_UNPOL], COMPUTE_SOURCE
 *   src/polarized/shdomsub1.f:2336-2529  COMPUTE_TOP_RADIANCES, *_LAMBERTIAN_BOUNDARY
 *   src/polarized/shdomsub4.f:13-809     UPDATE_COSTFUNCTION, RENDER, LEVISAPPROX_GRADIENT
 *   src/polarized/shdomsub4.f:1546-2042  COMPUTE_SOURCE_GRAD_1CELL
 *   src/polarized/shdomsub4.f:2151-2347  FIND_BOUNDARY_RADIANCE_GRAD (Lambertian)
 *   src/polarized/shdomsub4.f:2836-3169  COMPUTE_SOURCE_DIRECTION, PREPARE_DERIV_INTERPS
 *   src/polarized/shdomsub4.f:3223-4143  ADJOINT_INTEGRATE_1RAY and adjoint helpers
 *   src/shdomsub5.f:1497-2004            GET_INTERP_KERNEL, MAKE_DIRECT_DERIVATIVE
 *   src/util.f90:484-518                 average_subpixel_rays
 *   src/polarized/shdomsub1.f:2597-2669, shdomsub2.f:1222-1699, src/ocean_brdf.f
 *                                        VARIABLE_BRDF_SURFACE, SURFACE_BRDF and its models (oracle_surface.c)
 *   src/polarized/shdomsub1.f:445-4700 (parts), src/shdom_nompi.f:317-349
 *                                        SOLUTION_ITERATIONS / PATH_INTEGRATION (oracle_solver.c)
 *   src/polarized/shdomsub1.f:4703-5902, shdomsub2.f:614-1057, :1924-1987
 *                                        SPLIT_GRID and helpers, INIT_RADIANCE, EDDRTF, INTERP_RADIANCE (oracle_adapt.c)
 *   src/polarized/shdom90.f90:17-346, shdomsub2.f:313-612, shdomsub1.f:19-112
 *                                        TRILIN_INTERP_PROP, INTERP_GRID, PREPARE_PROP, TRANSFER_PA_TO_GRID (oracle_prop.c)
 *
 * All arrays are in the reference's own layout: Fortran (column-major) order,
 * 1-based index CONTENTS (GRIDPTR, NEIGHPTR, IPHASE, BCPTR, INTERPPTR ... hold
 * 1-based indices; SHPTR/RSHPTR hold 0-based offsets exactly as in Fortran).
 *
 * Parity pinning status: see oracle/README.md.
 */
#ifndef SHDOM_ORACLE_H
#define SHDOM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    /* ---- dimensions / flags ---- */
    int nstokes, nstleg, nx, ny, nz, npts, ncells;
    int ml, mm, nlm, nleg, numphase, npart, maxnmicro;
    int bcflag, ipflag;
    int nmu, nphi0max, nang;
    int maxnbc, ntoppts, nbotpts, nsfcpar;
    int nscatangle, nstphase;
    int deltam;               /* LOGICAL */
    int srctype;              /* 'S','T','B' as int */
    int units;                /* 'R','T','B' */
    int sfctype0, sfctype1;   /* e.g. 'F','L' */
    int interp_new;           /* INTERPMETHOD(2:2)=='N' -> 1, 'O' -> 0 */
    float solarmu, solaraz, solarflux, wavelen, gndtemp, gndalbedo, phasemax;
    float waveno0, waveno1;
    double tautol, transcut;
    /* ---- grid ---- */
    const int *gridptr;       /* [8,ncells] */
    const int *neighptr;      /* [6,ncells] */
    const int *treeptr;       /* [2,ncells] */
    const short *cellflags;   /* [ncells]   */
    const float *xgrid, *ygrid, *zgrid;
    const float *gridpos;     /* [3,npts] */
    /* ---- optics ---- */
    const float *extinct;     /* [npts,npart] */
    const float *albedo;      /* [npts,npart] */
    const float *total_ext;   /* [npts] */
    const float *legen;       /* [nstleg,0:nleg,numphase] */
    const int *iphase;        /* [8*maxnmicro,npts,npart] */
    const float *phaseinterpwt;
    const float *dirflux;     /* [npts] */
    const float *fluxes;      /* [2,npts] */
    const int *shptr;         /* [npts+1] */
    const float *source;      /* [nstokes,*] */
    const int *rshptr;        /* [npts+2] */
    const float *radiance;    /* [nstokes,*] */
    const float *ylmsun;      /* [nstleg,nlm] */
    const float *phasetab;    /* [nstphase,numphase,nscatangle] */
    const float *planck;      /* [npts,npart] (COMPUTE_SOURCE only) */
    const float *temp;        /* [npts] (thermal gradient only) */
    /* ---- discrete ordinates (sky radiance interpolation) ---- */
    const int *nphi0;         /* [nmu] */
    const float *mu;          /* [nmu] */
    const float *phi;         /* [nmu,nphi0max] */
    const float *wtdo;        /* [nmu,nphi0max] */
    const float *skyrad;      /* [nstokes,nmu/2,nphi0max] */
    /* ---- boundaries ---- */
    const int *bcptr;         /* [maxnbc,2] */
    float *bcrad;             /* [nstokes,*] mutated like the reference */
    const float *sfcgridparms;/* [nsfcpar,nbotpts] */
    const float *sfcgridrad;  /* [nang/2+1,*] */
} oracle_state;

typedef struct {
    int nrays;
    const float *camx, *camy, *camz;      /* REAL (f2py downcast) */
    const double *cammu, *camphi;
} oracle_rays;

typedef struct {
    /* extra inputs of LEVISAPPROX_GRADIENT (shdomsub4.f:288-317) */
    int npix, maxpg, numder, dnumphase, deriv_maxnmicro, longest_path_pts;
    int nuncertainty, maxsubgridints, exact_single_scatter, singlescatter;
    int costfunc_ll;          /* 0 -> 'L2', 1 -> 'LL' */
    double extmin, scatmin;
    const int *partder;       /* [numder] */
    const int *doexact;       /* [numder] */
    const float *measurements;     /* [nstokes,npix] */
    const double *uncertainties;   /* [nunc,nunc,npix] */
    const int *rays_per_pixel;     /* [npix] */
    const double *ray_weights;     /* [nrays] */
    const double *stokes_weights;  /* [nstokes,npix] */
    const float *dext, *dalb;      /* [maxpg,numder] */
    const float *dextm;            /* [maxpg,numder] */
    const float *dalbm, *dfj;      /* [8,npts,numder] */
    const float *optinterpwt;      /* [8,npts] */
    const int *interpptr;          /* [8,npts] */
    const float *dleg;             /* [nstleg,0:nleg,dnumphase] */
    const float *dphasetab;        /* [nstphase,dnumphase,nscatangle] */
    const int *diphasep;           /* [deriv_maxnmicro,maxpg,numder] */
    const float *dphasewtp;        /* [deriv_maxnmicro,maxpg,numder] */
    const int *iphasep;            /* [maxnmicro,maxpg,npart] */
    const float *phasewtp;         /* [maxnmicro,maxpg,npart] */
    const float *extinctp, *albedop; /* [maxpg,npart] */
    const float *dtemp;            /* [maxpg,numder] */
    const float *dpath;            /* [longest_path_pts,npts] */
    const int *dptr;               /* [longest_path_pts,npts] */
} oracle_grad_in;

/* optional per-ray trace of visited cells (bit-exact indexing checks) */
typedef struct {
    int max_per_ray;      /* capacity per ray                       */
    int *cells;           /* [max_per_ray, nrays] visited ICELL     */
    int *ncells;          /* [nrays] number of cells visited        */
    int *nsub;            /* [nrays] total number of sub-intervals  */
} oracle_trace;

/* ---- special functions ---- */
void oracle_ylmall(int transpose, float mu, float phi, int ml, int mm, int nstleg, float *yr);
int  oracle_precompute_phase_check(int nscatangle, int numphase, int nstphase, int nstokes,
                                   int ml, int nstleg, int nleg, const float *legen,
                                   float *phasetab, int deltam, int negcheck, char *errmsg);
int  oracle_precompute_phase_check_grad(int nscatangle, int dnumphase, int nstphase, int nstokes,
                                   int ml, int nstleg, int nleg, const float *dleg,
                                   float *dphasetab, int deltam, int negcheck, char *errmsg);

/* ---- grid ---- */
int  oracle_locate_grid_cell(const oracle_state *st, double *x0, double *y0, double *z0);

/* ---- a1: COMPUTE_SOURCE ---- */
int  oracle_compute_source(const oracle_state *st, int fixsh, float shacc, int maxiv,
                           int first, int accelflag, int newmethod,
                           int *shptr, float *source, int *oshptr, float *delsource,
                           float *deljdot, float *deljold, float *deljnew, float *jnorm,
                           char *errmsg);

/* ---- a2/a3: RENDER ---- */
int  oracle_render(const oracle_state *st, const oracle_rays *rays, float *stokes,
                   int correctinterpolate, int singlescatter, int nosurface,
                   oracle_trace *trace, int nthreads, char *errmsg);

/* ---- a8..a14: LEVISAPPROX_GRADIENT, default adjoint ("double sweep") path ---- */
int  oracle_levisapprox_gradient(const oracle_state *st, const oracle_rays *rays,
                                 const oracle_grad_in *g, double *gradout /*[maxpg,numder]*/,
                                 double *cost, float *stokesout /*[nstokes,npix]*/,
                                 oracle_trace *trace, int nthreads, char *errmsg);
/* LEVISAPPROX_GRADIENT, MAKEJACOBIAN=.TRUE. (GRAD_INTEGRATE_1RAY shdomsub4.f:811-1544, COMPUTE_RADIANCE_DERIVATIVE
 * :2588-2662, COMPUTE_DIRECT_BEAM_DERIV :2778-2834).  jacobian: float[nstokes, numder, num_jacobian_pts, npix]. */
int  oracle_levisapprox_jacobian(const oracle_state *st, const oracle_rays *rays,
                                 const oracle_grad_in *g, int num_jacobian_pts, const int *jacobianptr,
                                 double *gradout, double *cost, float *stokesout, float *jacobian,
                                 char *errmsg);

/* ---- fixed-grid solution iterations (oracle_solver.c; IPFLAG=3 sweeps), used to reproduce the reference's
 * SHDOM verification outputs.  shptr[npts+1], source[nstokes,maxiv], rshptr[npts+2], radiance[nstokes,maxiv+npts],
 * fluxes[2,npts], bcrad[nstokes, ntop + nbot*(1 or 1+nang/2)] are outputs. ---- */
int  oracle_solve_fixed_grid(const oracle_state *st, const float *wtmu, int maxiter, float solacc, float shacc,
                             int accelflag, int highorderrad, int iterfixsh, int maxiv,
                             int *shptr, float *source, int *rshptr, float *radiance, float *fluxes, float *bcrad,
                             int *iters_out, float *solcrit_out, char *errmsg);
/* restore != 0: continue from the SHPTR / SOURCE / RSHPTR / RADIANCE passed in (INIT_SOLUTION with INRADFLAG=.FALSE.) */
int  oracle_solve_fixed_grid_from(const oracle_state *st, const float *wtmu, int maxiter, float solacc, float shacc,
                                  int accelflag, int highorderrad, int iterfixsh, int maxiv, int restore,
                                  int *shptr, float *source, int *rshptr, float *radiance, float *fluxes, float *bcrad,
                                  int *iters_out, float *solcrit_out, char *errmsg);
int  oracle_sweeping_order(const oracle_state *st, int *sweepord /*[npts,8]*/);
void oracle_set_transmin(float transmin);   /* TRANSMIN of the 3-D sweeps (default 1.0) */
int  oracle_path_integration_once(const oracle_state *st, const float *wtmu, const int *shptr, const float *source,
                                  const int *rshptr, float *radiance, float *fluxes, float *bcrad, char *errmsg);
int  oracle_surface_brdf(int sfctype, const float *refparms, float wavelen, float mu2, float phi2,
                         float mu1, float phi1, int nstokes, float *reflect /*REFLECT(4,4)*/);

/* ---- property grid (ShdomPropertyArrays of at3d/solver.py:25-104) and the adaptive solve ---- */
typedef struct {
    int npx, npy, npz, numphase, nlegp, maxnmicro, npart, nzckd, nstleg;
    float delx, dely, xstart, ystart;
    const float *zlevels;     /* [npz] */
    const float *tempp;       /* [maxpg] or NULL */
    const float *extinctp;    /* [maxpg,npart] */
    const float *albedop;     /* [maxpg,npart] */
    const float *legenp;      /* [nstleg,0:nlegp,numphase] */
    int *iphasep;             /* [maxnmicro,maxpg,npart] (sorted in place by the 'O' interpolation mode) */
    float *phasewtp;          /* [maxnmicro,maxpg,npart] */
    const float *zckd, *gasabs; /* [nzckd] */
} oracle_prop;

/* mutable arrays of SOLUTION_ITERATIONS; point arrays have leading dimension maxig */
typedef struct {
    int maxig, maxic, maxiv, maxido;
    int npts, ncells, accelflag;
    int *gridptr, *neighptr, *treeptr;
    short *cellflags;
    float *gridpos;
    float *temp, *planck, *extinct, *albedo, *total_ext, *dirflux;
    int *iphase;
    float *phaseinterpwt;
    int *shptr, *rshptr, *oshptr;
    float *source, *radiance;
    const oracle_prop *pg;
    double extmin, scatmin;
    double beam_d[13];
    int beam_i[5];
    const float *extdirp;
} oracle_adapt;

void oracle_ssort(float *x, int *y, int n, int kflag);
void oracle_prop_extmin(const oracle_prop *pg, double *extmin, double *scatmin);
int  oracle_trilin_interp_prop(const oracle_prop *pg, int ipa, float x, float y, float z, int interp_new,
                               double extmin, double scatmin, float *temp, float *extinct, float *albedo,
                               int *iphase, float *phaseinterpwt, float *kg, char *errmsg);
float oracle_deltam_f(const float *legen, int nstleg, int nleg, int ml, const int *iphase, const float *pwt, int nq,
                      int interp_new, float phasemax);
int  oracle_transfer_pa_to_grid(const oracle_prop *pg, int npts, const float *gridpos, int ml, int nleg, int deltam,
                                int interp_new, float phasemax, int srctype, int units, const float *waveno,
                                float wavelen, float *temp, float *planck, float *extinct, float *albedo, float *legen,
                                int *iphase, float *phaseinterpwt, float *total_ext, double *extmin,
                                double *scatmin, float *albmax, char *errmsg);
void oracle_point_source(const oracle_state *st, int ld, int i, float *sourcet);
int  oracle_direct_beam_point(const double *out_d, const int *out_i, int bcflag, int npx, int npy, int npz,
                              float xstart, float ystart, const float *zlevels, const float *extdirp,
                              float solarflux, float x, float y, float z, float *dirflux, char *errmsg);
void oracle_cell_split_test(const oracle_adapt *a, int nstokes, int icell, float *adaptcrit, float *maxadapt, int *idir);
int  oracle_divide_cell(oracle_adapt *a, int icell, int idir, int newpoints[4][3]);
int  oracle_split_grid(oracle_adapt *a, const oracle_state *cst, int dosplit, int *outofmem, float cursplitacc,
                       float *splitcrit, char *errmsg);
int  oracle_boundary_pnts(int npts, int nang, int lambertian, int maxnbc, int maxbcrad, float zbot, float ztop,
                          const float *gridpos, int *ntoppts, int *nbotpts, int *bcptr);
int  oracle_eddrtf(int nlayer, const float *optdepths, const float *albedos, const float *asymmetries,
                   const float *temps, int deltam, int srctype, float solarflux, float solarmu, float gndtemp,
                   float gndemis, float skyrad, int units, const float *waveno, float wavelen, float *fluxes,
                   float surface_flux);
int  oracle_init_radiance(const oracle_state *st, int ld, int nxy, int nz, const float *extinct, const float *albedo,
                          const float *total_ext, const float *temp, const int *iphase, const float *phaseinterpwt,
                          float skyrad, float surface_flux, int *rshptr, float *radiance);
int  oracle_interp_radiance(int nstokes, int oldnpts, int nbcells, int ncells, const int *treeptr, const int *gridptr,
                            const float *gridpos, int *rshptr, float *radiance);
/* INIT_SOLUTION + SOLUTION_ITERATIONS with adaptive cell splitting (shdomsub1.f:113-822).  Every array of `st` that
 * grows (grid, optics, SH arrays, fluxes, dirflux, bcptr, bcrad) must have the capacity given by the max* arguments,
 * point arrays with leading dimension maxig; npts / ncells / ntoppts / nbotpts of `st` are updated.
 * sizes[0..3] out: shptr[npts], rshptr[npts] totals, iterations, oldnpts. */
int  oracle_solve_adaptive(oracle_state *st, oracle_prop *pg, const float *wtmu, float *temp,
                           int maxig, int maxic, int maxiv, int maxido, int maxbcrad, int nbpts, int nbcells,
                           int maxiter, float solacc, float splitacc, float shacc, int accelflag, int highorderrad,
                           int iterfixsh, int inradflag, float *extdirp, int *iters_out, float *solcrit_out,
                           float *splitcrit_out, char *errmsg);

/* ---- helpers on the path ---- */
int  oracle_update_costfunction(const double *stokesout, const double *raygrad_pixel,
                                double *gradout, double *cost, const double *uncertainties,
                                int costfunc_ll, int nstokes, int maxpg, int numder,
                                const double *measurement, int nuncertainty);
int  oracle_prepare_deriv_interps(const oracle_state *st, int npx, int npy, int npz, int maxpg,
                                  float delx, float dely, float xstart, float ystart,
                                  const float *zlevels, const oracle_grad_in *g,
                                  float *optinterpwt, int *interpptr,
                                  float *dalbm, float *dextm, float *dfj, char *errmsg);
void oracle_average_subpixel_rays(int npixels, int nrays, int nstokes, const float *weighted_stokes,
                                  const int *pixel_index, float *observables);
int  oracle_make_direct_derivative(int npts, int bcflag, int npx, int npy, int npz,
                                   float delx, float dely, float xstart, float ystart,
                                   const float *gridpos, const float *zlevels,
                                   int ipdirect, int di, int dj, int dk,
                                   double cx, double cy, double cz,
                                   double cxinv, double cyinv, double czinv,
                                   double epss, double epsz, double xdomain, double ydomain,
                                   double uniformzlev, double delxd, double delyd,
                                   float *dpath, int *dptr, int longest_path_pts, char *errmsg);

int  oracle_make_direct(int npts, int bcflag, int ipflag, int deltam, int ml, int nstleg, int nlegp,
                        float solarflux, float solarmu, float solaraz, const float *gridpos,
                        int npx, int npy, int npz, float delx, float dely, float xstart, float ystart,
                        const float *zlevels, const float *extinctp, const float *albedop,
                        const float *legenp, const int *iphasep, const float *phasewtp,
                        int maxnmicro, int npart, int nzckd, const float *zckd, const float *gasabs,
                        float *extdirp, float *dirflux, double *out_d /*[13]*/, int *out_i /*[5]*/,
                        char *errmsg);

#ifdef __cplusplus
}
#endif
#endif
