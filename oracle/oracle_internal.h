/* oracle_internal.h -- TEST INFRASTRUCTURE (see shdom_oracle.h).  Fortran-style accessors. */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H
#include <stddef.h>
#include "shdom_oracle.h"

#define BTEST(x, b) ((((int)(x)) >> (b)) & 1)
#define IBITS2(x) ((((int)(x)) >> 2) & 3)
#define IMAX3(a, b, c) ((a) > (b) ? ((a) > (c) ? (a) : (c)) : ((b) > (c) ? (b) : (c)))

#define GRIDPTR(st, n, ic) ((st)->gridptr[((n) - 1) + 8 * (size_t)((ic) - 1)])
#define NEIGHPTR(st, n, ic) ((st)->neighptr[((n) - 1) + 6 * (size_t)((ic) - 1)])
#define TREEPTR(st, n, ic) ((st)->treeptr[((n) - 1) + 2 * (size_t)((ic) - 1)])
#define CELLFLAGS(st, ic) ((st)->cellflags[(ic) - 1])
#define GRIDPOS(st, i, ip) ((st)->gridpos[((i) - 1) + 3 * (size_t)((ip) - 1)])
#define SOURCE(st, k, j) ((st)->source[((k) - 1) + (size_t)(st)->nstokes * ((j) - 1)])
#define RADIANCE(st, k, j) ((st)->radiance[((k) - 1) + (size_t)(st)->nstokes * ((j) - 1)])

typedef struct {
    double cx, cy, cz, cxinv, cyinv, czinv, cosscat;
    int bitx, bity, bitz, ioct;
    float xm, ym;
} ray_dir;

typedef struct {
    float *ylmdir;     /* [nstleg,nlm] */
    float *singscat;   /* [nstokes,numphase] */
    float *dsingscat;  /* [nstokes,dnumphase] */
    float *legent;     /* scratch for Legendre tables */
} ray_scratch;

int  oracle_next_cell(const oracle_state *st, double xe, double ye, double ze,
                      int iface, int jface, int icell);
void oracle_rotate_pol_plane(int nstokes, double cosscat, float solarmu, float mu,
                             float delphi, float *scatvect);
void oracle_ray_setup(const oracle_state *st, double mu2, double phi2, ray_dir *rd,
                      float *ylmdir, float *singscat,
                      const float *dphasetab, int dnumphase, float *dsingscat);
int  oracle_bc_search(const int *bcptr_col, int n, int ip);
float oracle_sky_radiance(const oracle_state *st, float mu, float phi);
void oracle_lambertian_boundary(const oracle_state *st, float *bcrad);
void oracle_compute_top_radiances(const oracle_state *st, const float *skyrad, int imu, int iphi,
                                  float mu, float phi, int flag, float *out);
float oracle_planck_function(float temp, int units, const float *waveno, float wavelen);
int  oracle_surface_brdf(int sfctype, const float *refparms, float wavelen, float mu2, float phi2,
                         float mu1, float phi1, int nstokes, float *reflect);
int  oracle_variable_brdf_surface(const oracle_state *st, int ibeg, int iend, float mu2, float phi2,
                                  float *bcrad_bot);
void oracle_donethis(const oracle_state *st, int iface, int *donethis);
int  oracle_integrate_1ray(const oracle_state *st, float *bcrad, float skyrad_top,
                           double mu2, double phi2, double x0, double y0, double z0,
                           double *transmit_io, double *radiance,
                           int correctinterpolate, int singlescatter, int nosurface,
                           ray_scratch *sc, int *trace_cells, int trace_cap, int *trace_n,
                           int *nsub_out, char *errmsg);
ray_scratch *oracle_scratch_new(const oracle_state *st, int dnumphase);
void oracle_scratch_free(ray_scratch *sc);
int  oracle_ray_start(const oracle_state *st, double mu2, double phi2,
                      double *x0, double *y0, double *z0, int *ierr);
#endif
