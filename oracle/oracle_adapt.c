/*
 * oracle_adapt.c -- TEST INFRASTRUCTURE ONLY (see shdom_oracle.h).
 * The adaptive grid of SHDOM and the Eddington first guess, restated from (paths relative to /root/reference):
 *   src/polarized/shdomsub1.f:4703-4932  SPLIT_GRID
 *   src/polarized/shdomsub1.f:4937-5283  INTERPOLATE_POINT
 *   src/polarized/shdomsub1.f:5289-5364  DIVIDE_CELL
 *   src/polarized/shdomsub1.f:5370-5458  MATCH_NEIGHBOR_FACE
 *   src/polarized/shdomsub1.f:5462-5506  INHERIT_NEIGHBOR
 *   src/polarized/shdomsub1.f:5514-5596  NEW_GRID_POINTS
 *   src/polarized/shdomsub1.f:5600-5697  MATCH_GRID_POINT
 *   src/polarized/shdomsub1.f:5703-5791  CELL_SPLIT_TEST
 *   src/polarized/shdomsub1.f:5796-5902  GRID_SMOOTH_TEST
 *   src/polarized/shdomsub1.f:2173-2215  BOUNDARY_PNTS
 *   src/polarized/shdomsub2.f:614-762    INIT_RADIANCE
 *   src/polarized/shdomsub2.f:765-995    EDDRTF
 *   src/polarized/shdomsub2.f:999-1057   TRIDIAG
 *   src/polarized/shdomsub2.f:1924-1987  INTERP_RADIANCE
 * Pinned by the reference's rico32x36x26w672 SHDOM verification outputs (tests/test_shdom_adaptive.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

#define AGRIDPTR(a, n, ic) ((a)->gridptr[((n) - 1) + 8 * (size_t)((ic) - 1)])
#define ANEIGHPTR(a, n, ic) ((a)->neighptr[((n) - 1) + 6 * (size_t)((ic) - 1)])
#define ATREEPTR(a, n, ic) ((a)->treeptr[((n) - 1) + 2 * (size_t)((ic) - 1)])
#define ACELLFLAGS(a, ic) ((a)->cellflags[(ic) - 1])
#define AGRIDPOS(a, i, ip) ((a)->gridpos[((i) - 1) + 3 * (size_t)((ip) - 1)])

static const int GRIDCORNER[3][4][2] = {
    {{1, 2}, {3, 4}, {5, 6}, {7, 8}}, {{1, 3}, {2, 4}, {5, 7}, {6, 8}}, {{1, 5}, {2, 6}, {3, 7}, {4, 8}}};
static const int OPPFACE[6] = {2, 1, 4, 3, 6, 5};

/* CELL_SPLIT_TEST  shdomsub1.f:5703-5791 (EXTINCT is TOTAL_EXT at the call sites) */
void oracle_cell_split_test(const oracle_adapt *a, int nstokes, int icell, float *adaptcrit, float *maxadapt, int *idir)
{
    const float c0 = 0.282095f;
    int id, ie, j;
    *maxadapt = -1.0f;
    for (id = 1; id <= 3; id++) {
        float sum1 = 0.0f;
        int num = 0;
        for (ie = 1; ie <= 4; ie++) {
            const int ip1 = AGRIDPTR(a, GRIDCORNER[id - 1][ie - 1][0], icell);
            const int ip2 = AGRIDPTR(a, GRIDCORNER[id - 1][ie - 1][1], icell);
            if (ip1 != ip2) {
                const int is1 = a->shptr[ip1 - 1], is2 = a->shptr[ip2 - 1];
                const int ns1 = a->shptr[ip1] - is1, ns2 = a->shptr[ip2] - is2;
                const int ns = ns1 < ns2 ? ns1 : ns2;
                const float e1 = a->total_ext[ip1 - 1], e2 = a->total_ext[ip2 - 1];
                float jay = 0.0f, ext, tau, split;
                num = num + 1;
                for (j = 1; j <= ns; j++) {
                    const float d = e2 * a->source[(size_t)nstokes * (is2 + j - 1)]
                                    - e1 * a->source[(size_t)nstokes * (is1 + j - 1)];
                    jay = jay + d * d;
                }
                for (j = ns + 1; j <= ns1; j++) {
                    const float d = e1 * a->source[(size_t)nstokes * (is1 + j - 1)];
                    jay = jay + d * d;
                }
                for (j = ns + 1; j <= ns2; j++) {
                    const float d = e2 * a->source[(size_t)nstokes * (is2 + j - 1)];
                    jay = jay + d * d;
                }
                ext = 0.5f * (e1 + e2);
                if (ext > 0.0f) jay = c0 * sqrtf(jay) / ext;
                else jay = 0.0f;
                tau = fabsf(ext * (AGRIDPOS(a, id, ip2) - AGRIDPOS(a, id, ip1)));
                split = fabsf(jay) * (1 - expf(-tau));
                sum1 = sum1 + split;
            }
        }
        if (num > 0) adaptcrit[id - 1] = sum1 / num;
        else adaptcrit[id - 1] = 0.0f;
        if (adaptcrit[id - 1] > *maxadapt) {
            *maxadapt = adaptcrit[id - 1];
            *idir = id;
        }
    }
}

/* NEXT_CELL on the mutable grid (shdomsub1.f:4470-4522) */
static int next_cell(const oracle_adapt *a, double xe, double ye, double ze, int iface, int jface, int icell)
{
    int inext = ANEIGHPTR(a, iface, icell);
    if (inext < 0) {
        int ic = -inext;
        while (ATREEPTR(a, 2, ic) > 0) {
            const int dir = IBITS2(ACELLFLAGS(a, ic));
            const int ic1 = ATREEPTR(a, 2, ic);
            if (dir == jface) {
                ic = ic1 + 1 - ((iface - 1) % 2);
            } else {
                ic = ic1;
                if (dir == 1) { if (xe > AGRIDPOS(a, 1, AGRIDPTR(a, 8, ic1))) ic = ic + 1; }
                else if (dir == 2) { if (ye > AGRIDPOS(a, 2, AGRIDPTR(a, 8, ic1))) ic = ic + 1; }
                else { if (ze > AGRIDPOS(a, 3, AGRIDPTR(a, 8, ic1))) ic = ic + 1; }
            }
        }
        inext = ic;
    }
    return inext;
}

/* GRID_SMOOTH_TEST  shdomsub1.f:5796-5902 */
static void grid_smooth_test(const oracle_adapt *a, int icell, int *idir_out)
{
    static const int FACECORNER[6][4] = {{1, 3, 5, 7}, {2, 4, 6, 8}, {1, 2, 5, 6}, {3, 4, 7, 8}, {1, 2, 3, 4}, {5, 6, 7, 8}};
    static const int EDGECORNER[3][2] = {{1, 2}, {1, 3}, {1, 5}};
    int idir = 0, id, j, i;
    float sizeratio = 1.0f, curgridsize[3], gridsize[2], gridsizeinv, ratio;
    int incell[2], dirsplit[2];
    for (id = 1; id <= 3; id++) {
        const int ip1 = AGRIDPTR(a, EDGECORNER[id - 1][0], icell);
        const int ip2 = AGRIDPTR(a, EDGECORNER[id - 1][1], icell);
        curgridsize[id - 1] = 1.0e20f;
        if (ip1 != ip2) {
            curgridsize[id - 1] = fabsf(AGRIDPOS(a, id, ip1) - AGRIDPOS(a, id, ip2));
            gridsizeinv = 1.0f / curgridsize[id - 1];
            for (j = 1; j <= 2; j++) {
                const int iface = 2 * (id - 1) + j;
                double xe = 0.0, ye = 0.0, ze = 0.0;
                int in;
                dirsplit[j - 1] = 0;
                for (i = 1; i <= 4; i++) {
                    const int ip = AGRIDPTR(a, FACECORNER[iface - 1][i - 1], icell);
                    xe = xe + AGRIDPOS(a, 1, ip) * 0.25f;
                    ye = ye + AGRIDPOS(a, 2, ip) * 0.25f;
                    ze = ze + AGRIDPOS(a, 3, ip) * 0.25f;
                }
                incell[j - 1] = next_cell(a, xe, ye, ze, iface, id, icell);
                if (incell[j - 1] == 0) {
                    gridsize[j - 1] = curgridsize[id - 1];
                } else {
                    gridsize[j - 1] = fabsf(AGRIDPOS(a, id, AGRIDPTR(a, 1, incell[j - 1]))
                                            - AGRIDPOS(a, id, AGRIDPTR(a, 8, incell[j - 1])));
                    if (ATREEPTR(a, 1, incell[j - 1]) == 0) gridsize[j - 1] = curgridsize[id - 1];
                    if (id <= 2 && BTEST(ACELLFLAGS(a, incell[j - 1]), id - 1)) gridsize[j - 1] = curgridsize[id - 1];
                }
                in = ANEIGHPTR(a, iface, icell);
                if (in < 0) dirsplit[j - 1] = IBITS2(ACELLFLAGS(a, abs(in)));
            }
            if (gridsize[0] * gridsizeinv < 0.75f && gridsize[1] * gridsizeinv < 0.75f) idir = id;
            if (dirsplit[0] > 0 && dirsplit[0] != id && dirsplit[0] == dirsplit[1]) idir = dirsplit[0];
            ratio = fminf(gridsize[0] * gridsizeinv, gridsize[1] * gridsizeinv);
            if (ratio < 0.4f && ratio < sizeratio) {
                int isum = 0;
                for (i = 1; i <= 8; i++) {
                    const int ip = AGRIDPTR(a, i, icell);
                    isum = isum + a->shptr[ip] - a->shptr[ip - 1];
                }
                if (isum > 0) {
                    idir = id;
                    sizeratio = ratio;
                }
            }
        }
    }
    *idir_out = idir;
}

/* MATCH_GRID_POINT  shdomsub1.f:5600-5697 */
static int match_grid_point(const oracle_adapt *a, float xp, float yp, float zp, int icell, int iface)
{
    static const int GRIDFACE[6][4] = {{1, 3, 5, 7}, {2, 4, 6, 8}, {1, 2, 5, 6}, {3, 4, 7, 8}, {1, 2, 3, 4}, {5, 6, 7, 8}};
    int cellstack[64], sp = 0;
    const int idir = (iface + 1) / 2, kface = OPPFACE[iface - 1];
    int ic = abs(ANEIGHPTR(a, iface, icell)), i, dir, ic1;
    if (ic == 0) return 0;
    for (;;) {
        while (ATREEPTR(a, 2, ic) == 0) {
            for (i = 1; i <= 4; i++) {
                const int ipt = AGRIDPTR(a, GRIDFACE[kface - 1][i - 1], ic);
                if (xp == AGRIDPOS(a, 1, ipt) && yp == AGRIDPOS(a, 2, ipt) && zp == AGRIDPOS(a, 3, ipt)) return ipt;
            }
            if (sp == 0) return 0;
            ic = cellstack[sp - 1];
            sp = sp - 1;
        }
        dir = IBITS2(ACELLFLAGS(a, ic));
        ic1 = ATREEPTR(a, 2, ic);
        if (dir == idir) {
            ic = ic1 + 1 - ((iface - 1) % 2);
        } else {
            const float p = dir == 1 ? xp : (dir == 2 ? yp : zp);
            const float s = AGRIDPOS(a, dir, AGRIDPTR(a, 8, ic1));
            ic = ic1;
            if (p == s) {
                if (sp >= 63) return 0;      /* MAXSTACK exceeded: the reference STOPs */
                cellstack[sp] = ic + 1;
                sp = sp + 1;
            } else if (p > s) {
                ic = ic + 1;
            }
        }
    }
}

/* NEW_GRID_POINTS  shdomsub1.f:5514-5596 */
static void new_grid_points(oracle_adapt *a, int idir, int icell, int newcell, int newpoints[4][3])
{
    static const int FACEGRID[3][4][2] = {
        {{3, 5}, {4, 5}, {3, 6}, {4, 6}}, {{1, 5}, {2, 5}, {1, 6}, {2, 6}}, {{1, 3}, {2, 3}, {1, 4}, {2, 4}}};
    int i, k;
    for (k = 1; k <= 8; k++) {
        AGRIDPTR(a, k, newcell) = AGRIDPTR(a, k, icell);
        AGRIDPTR(a, k, newcell + 1) = AGRIDPTR(a, k, icell);
    }
    for (i = 1; i <= 4; i++) {
        const int i1 = GRIDCORNER[idir - 1][i - 1][0], i2 = GRIDCORNER[idir - 1][i - 1][1];
        const int ip1 = AGRIDPTR(a, i1, icell), ip2 = AGRIDPTR(a, i2, icell);
        const float xp = (AGRIDPOS(a, 1, ip1) + AGRIDPOS(a, 1, ip2)) / 2;
        const float yp = (AGRIDPOS(a, 2, ip1) + AGRIDPOS(a, 2, ip2)) / 2;
        const float zp = (AGRIDPOS(a, 3, ip1) + AGRIDPOS(a, 3, ip2)) / 2;
        const int iface1 = FACEGRID[idir - 1][i - 1][0], iface2 = FACEGRID[idir - 1][i - 1][1];
        int ipmatch = match_grid_point(a, xp, yp, zp, icell, iface1);
        if (ipmatch == 0) ipmatch = match_grid_point(a, xp, yp, zp, icell, iface2);
        if (ipmatch == 0) {
            const int icell2 = abs(ANEIGHPTR(a, iface1, icell));
            if (icell2 > 0) ipmatch = match_grid_point(a, xp, yp, zp, icell2, iface2);
        }
        if (ipmatch == 0) {
            a->npts = a->npts + 1;
            AGRIDPTR(a, i2, newcell) = a->npts;
            AGRIDPTR(a, i1, newcell + 1) = a->npts;
            AGRIDPOS(a, 1, a->npts) = xp;
            AGRIDPOS(a, 2, a->npts) = yp;
            AGRIDPOS(a, 3, a->npts) = zp;
            newpoints[i - 1][0] = ip1;
            newpoints[i - 1][1] = ip2;
            newpoints[i - 1][2] = a->npts;
        } else {
            AGRIDPTR(a, i2, newcell) = ipmatch;
            AGRIDPTR(a, i1, newcell + 1) = ipmatch;
            newpoints[i - 1][2] = 0;
        }
    }
}

/* INHERIT_NEIGHBOR  shdomsub1.f:5462-5506 */
static void inherit_neighbor(oracle_adapt *a, int icell, int iface, int in)
{
    const int jface = (iface + 1) / 2;
    int ic = icell, sp = 0, stack[50], done = 0;
    while (!done) {
        ANEIGHPTR(a, iface, ic) = in;
        if (ATREEPTR(a, 2, ic) == 0) {
            if (sp == 0) done = 1;
            else { ic = stack[sp - 1]; sp = sp - 1; }
        } else {
            const int dir = IBITS2(ACELLFLAGS(a, ic));
            if (dir == jface) {
                ic = ATREEPTR(a, 2, ic) + ((iface - 1) % 2);
            } else {
                if (sp >= 50) return;
                stack[sp] = ATREEPTR(a, 2, ic) + 1;
                sp = sp + 1;
                ic = ATREEPTR(a, 2, ic);
            }
        }
    }
}

/* MATCH_NEIGHBOR_FACE  shdomsub1.f:5370-5458 */
static void match_neighbor_face(oracle_adapt *a, int iface, int ic)
{
    int in = abs(ANEIGHPTR(a, iface, ic)), inn, jface, ic1, ic8, in1, in8, dir, dir1, dir2, done = 0;
    float pos[4];
    if (in == 0) return;
    jface = (iface + 1) / 2;
    ic1 = AGRIDPTR(a, 1, ic);
    ic8 = AGRIDPTR(a, 8, ic);
    dir1 = (jface - 1 + 1) % 3 + 1;
    dir2 = (jface - 1 + 2) % 3 + 1;
    pos[1] = pos[2] = pos[3] = 0.0f;
    pos[dir1] = (AGRIDPOS(a, dir1, ic1) + AGRIDPOS(a, dir1, ic8)) / 2;
    pos[dir2] = (AGRIDPOS(a, dir2, ic1) + AGRIDPOS(a, dir2, ic8)) / 2;
    while (!done && ic != in) {
        in1 = AGRIDPTR(a, 1, in);
        in8 = AGRIDPTR(a, 8, in);
        if (AGRIDPOS(a, dir1, ic1) >= AGRIDPOS(a, dir1, in1) && AGRIDPOS(a, dir1, ic8) <= AGRIDPOS(a, dir1, in8) &&
            AGRIDPOS(a, dir2, ic1) >= AGRIDPOS(a, dir2, in1) && AGRIDPOS(a, dir2, ic8) <= AGRIDPOS(a, dir2, in8)) {
            if (ATREEPTR(a, 2, in) == 0) ANEIGHPTR(a, iface, ic) = in;
            else ANEIGHPTR(a, iface, ic) = -in;
        }
        if (AGRIDPOS(a, dir1, in1) >= AGRIDPOS(a, dir1, ic1) && AGRIDPOS(a, dir1, in8) <= AGRIDPOS(a, dir1, ic8) &&
            AGRIDPOS(a, dir2, in1) >= AGRIDPOS(a, dir2, ic1) && AGRIDPOS(a, dir2, in8) <= AGRIDPOS(a, dir2, ic8)) {
            inherit_neighbor(a, in, OPPFACE[iface - 1], ic);
        } else {
            ANEIGHPTR(a, OPPFACE[iface - 1], in) = -abs(ANEIGHPTR(a, OPPFACE[iface - 1], in));
        }
        if (ATREEPTR(a, 2, in) == 0) {
            done = 1;
        } else {
            dir = IBITS2(ACELLFLAGS(a, in));
            inn = ATREEPTR(a, 2, in);
            if (dir == jface) {
                in = inn + 1 - ((iface - 1) % 2);
            } else {
                if (pos[dir] > AGRIDPOS(a, dir, AGRIDPTR(a, 8, inn))) in = inn + 1;
                else in = inn;
            }
        }
    }
}

/* DIVIDE_CELL  shdomsub1.f:5289-5364 */
int oracle_divide_cell(oracle_adapt *a, int icell, int idir, int newpoints[4][3])
{
    int newcell, iface, i;
    if (ATREEPTR(a, 2, icell) != 0) return 1;
    newcell = a->ncells + 1;
    a->ncells = a->ncells + 2;
    ATREEPTR(a, 2, icell) = newcell;
    ATREEPTR(a, 1, newcell) = icell;
    ATREEPTR(a, 2, newcell) = 0;
    ATREEPTR(a, 1, newcell + 1) = icell;
    ATREEPTR(a, 2, newcell + 1) = 0;
    ACELLFLAGS(a, icell) = (short)(ACELLFLAGS(a, icell) | (idir << 2));
    ACELLFLAGS(a, newcell) = 0;
    ACELLFLAGS(a, newcell + 1) = 0;
    for (i = 0; i <= 1; i++)
        if (BTEST(ACELLFLAGS(a, icell), i)) {
            ACELLFLAGS(a, newcell) = (short)(ACELLFLAGS(a, newcell) | (1 << i));
            ACELLFLAGS(a, newcell + 1) = (short)(ACELLFLAGS(a, newcell + 1) | (1 << i));
        }
    new_grid_points(a, idir, icell, newcell, newpoints);
    for (iface = 1; iface <= 6; iface++) {
        if (ANEIGHPTR(a, iface, icell) == icell) {
            ANEIGHPTR(a, iface, newcell) = newcell;
            ANEIGHPTR(a, iface, newcell + 1) = newcell + 1;
        } else if (iface == 2 * idir) {
            ANEIGHPTR(a, iface, newcell) = newcell + 1;
            ANEIGHPTR(a, iface, newcell + 1) = ANEIGHPTR(a, iface, icell);
        } else if (iface == 2 * idir - 1) {
            ANEIGHPTR(a, iface, newcell + 1) = newcell;
            ANEIGHPTR(a, iface, newcell) = ANEIGHPTR(a, iface, icell);
        } else {
            ANEIGHPTR(a, iface, newcell) = ANEIGHPTR(a, iface, icell);
            ANEIGHPTR(a, iface, newcell + 1) = ANEIGHPTR(a, iface, icell);
        }
        match_neighbor_face(a, iface, newcell);
        match_neighbor_face(a, iface, newcell + 1);
    }
    return 0;
}

/* INTERPOLATE_POINT  shdomsub1.f:4937-5283.  The point arrays of `a` have leading dimension a->maxig. */
static int interpolate_point(oracle_adapt *a, const oracle_state *cst, int newpoints[4][3], char *errmsg)
{
    const int nstokes = cst->nstokes, npart = cst->npart, nq = 8 * cst->maxnmicro, ld = a->maxig, ml = cst->ml;
    const oracle_prop *pg = a->pg;
    oracle_state st = *cst;
    float *sourcet = (float *)malloc(sizeof(float) * nstokes * cst->nlm);
    int i, ipa, j, k, ierr = 0;
    /* a view of the big arrays (leading dimension ld) for the source evaluation */
    st.extinct = a->extinct; st.albedo = a->albedo; st.total_ext = a->total_ext; st.planck = a->planck;
    st.iphase = a->iphase; st.phaseinterpwt = a->phaseinterpwt; st.dirflux = a->dirflux;
    st.rshptr = a->rshptr; st.radiance = a->radiance;
    for (i = 1; i <= 4 && !ierr; i++) {
        if (newpoints[i - 1][2] > 0) {
            const int ip1 = newpoints[i - 1][0], ip2 = newpoints[i - 1][1], ip = newpoints[i - 1][2];
            const float x = AGRIDPOS(a, 1, ip), y = AGRIDPOS(a, 2, ip), z = AGRIDPOS(a, 3, ip);
            int ir1, ir2, nr1, nr2, nr, ir, ns1, ns2, ns, is;
            float kg, f;
            a->total_ext[ip - 1] = 0.0f;
            for (ipa = 1; ipa <= npart; ipa++) {
                const size_t o = (ip - 1) + (size_t)ld * (ipa - 1);
                ierr = oracle_trilin_interp_prop(pg, ipa, x, y, z, cst->interp_new, a->extmin, a->scatmin,
                                                 &a->temp[ip - 1], &a->extinct[o], &a->albedo[o], &a->iphase[nq * o],
                                                 &a->phaseinterpwt[nq * o], &kg, errmsg);
                if (ierr) break;
                if (cst->deltam) {
                    f = oracle_deltam_f(cst->legen, cst->nstleg, cst->nleg, ml, &a->iphase[nq * o],
                                        &a->phaseinterpwt[nq * o], nq, cst->interp_new, cst->phasemax);
                    a->extinct[o] = (1.0f - a->albedo[o] * f) * a->extinct[o];
                    a->albedo[o] = (1.0f - f) * a->albedo[o] / (1.0f - a->albedo[o] * f);
                }
                if (ipa == 1) a->total_ext[ip - 1] = a->total_ext[ip - 1] + kg;
                a->total_ext[ip - 1] = a->total_ext[ip - 1] + a->extinct[o];
                if (cst->srctype != 'S') {
                    const float wn[2] = {cst->waveno0, cst->waveno1};
                    const float bb = oracle_planck_function(a->temp[ip - 1], cst->units, wn, cst->wavelen);
                    a->planck[o] = (1.0f - a->albedo[o]) * bb;
                }
            }
            if (ierr) break;
            if (cst->srctype != 'T') {
                ierr = oracle_direct_beam_point(a->beam_d, a->beam_i, cst->bcflag, pg->npx, pg->npy, pg->npz,
                                                pg->xstart, pg->ystart, pg->zlevels, a->extdirp, cst->solarflux,
                                                x, y, z, &a->dirflux[ip - 1], errmsg);
                if (ierr) break;
            }
            ir1 = a->rshptr[ip1 - 1]; ir2 = a->rshptr[ip2 - 1];
            nr1 = a->rshptr[ip1] - ir1; nr2 = a->rshptr[ip2] - ir2;
            nr = nr1 > nr2 ? nr1 : nr2;
            ir = a->rshptr[ip - 1];
            a->rshptr[ip] = ir + nr;
            for (j = 1; j <= nr; j++)
                for (k = 0; k < nstokes; k++) {
                    const float r1 = j <= nr1 ? a->radiance[k + (size_t)nstokes * (ir1 + j - 1)] : 0.0f;
                    const float r2 = j <= nr2 ? a->radiance[k + (size_t)nstokes * (ir2 + j - 1)] : 0.0f;
                    a->radiance[k + (size_t)nstokes * (ir + j - 1)] = 0.5f * (r1 + r2);
                }
            ns1 = a->shptr[ip1] - a->shptr[ip1 - 1];
            ns2 = a->shptr[ip2] - a->shptr[ip2 - 1];
            ns = ns1 > ns2 ? ns1 : ns2;
            is = a->shptr[ip - 1];
            a->shptr[ip] = is + ns;
            if (a->accelflag) a->oshptr[ip] = a->oshptr[ip - 1];
            st.npts = a->npts;
            oracle_point_source(&st, ld, ip, sourcet);
            for (j = 1; j <= ns; j++)
                for (k = 0; k < nstokes; k++)
                    a->source[k + (size_t)nstokes * (is + j - 1)] = sourcet[k + (size_t)nstokes * (j - 1)];
        }
    }
    free(sourcet);
    return ierr;
}

/* SPLIT_GRID  shdomsub1.f:4703-4932 */
int oracle_split_grid(oracle_adapt *a, const oracle_state *cst, int dosplit, int *outofmem, float cursplitacc,
                      float *splitcrit, char *errmsg)
{
    const int nstokes = cst->nstokes, nphi0max = cst->nphi0max, nlm = cst->nlm;
    const int maxic = a->maxic, maxig = a->maxig, maxido = a->maxido, maxiv = a->maxiv;
    float adapt[3], crit;
    int icell, idir, ierr = 0;
    if (dosplit) {
        float *adaptcrit = (float *)malloc(sizeof(float) * ((size_t)maxic + 2));
        int *adaptind = (int *)malloc(sizeof(int) * ((size_t)maxic + 2));
        int icell1 = 1;
        int outofmem0 = *outofmem;
        while (icell1 <= a->ncells && !ierr) {
            int n = 0, i, maxcells, maxpts, maxwork, maxsh, maxrad, newpoints[4][3];
            const float frac = 0.03f;
            for (icell = icell1; icell <= a->ncells; icell++)
                if (ATREEPTR(a, 2, icell) == 0) {
                    oracle_cell_split_test(a, nstokes, icell, adapt, &adaptcrit[n], &idir);
                    adaptind[n] = 4 * icell + idir;
                    n = n + 1;
                }
            icell1 = a->ncells + 1;
            adaptcrit[n] = 0.0f;
            if (n > 0) oracle_ssort(adaptcrit, adaptind, n, -2);
            maxcells = (int)(maxic - frac * (maxic - a->ncells) - 2);
            maxpts = (int)(maxig - frac * (maxig - a->npts) - 4);
            maxwork = (int)(maxido - frac * (maxido - nphi0max * a->npts) - 4 * nphi0max);
            maxsh = (int)(maxiv - frac * (maxiv - a->shptr[a->npts]) - 4 * nlm);
            maxrad = (int)(maxiv + maxig - frac * (maxiv + maxig - a->rshptr[a->npts]) - 4 * nlm);
            outofmem0 = *outofmem;
            i = 1;
            while (i <= n && adaptcrit[i - 1] > cursplitacc && !outofmem0) {
                if (a->ncells > maxcells || a->npts > maxpts || a->npts * nphi0max > maxwork ||
                    a->shptr[a->npts] > maxsh || a->rshptr[a->npts] > maxrad) {
                    outofmem0 = 1;
                } else {
                    icell = adaptind[i - 1] / 4;
                    idir = adaptind[i - 1] & 3;
                    if (oracle_divide_cell(a, icell, idir, newpoints)) {
                        if (errmsg) snprintf(errmsg, 600, "DIVIDE_CELL: Cannot divide already split cell.");
                        ierr = 1; break;
                    }
                    ierr = interpolate_point(a, cst, newpoints, errmsg);
                    if (ierr) break;
                }
                i = i + 1;
            }
            if (ierr) break;
            maxcells = maxic - 2;
            maxpts = maxig - 4;
            maxwork = maxido - 4 * nphi0max;
            maxsh = maxiv - 4 * nlm;
            maxrad = maxiv + maxig - 4 * nlm;
            while (!*outofmem && i <= n) {
                if (a->ncells > maxcells || a->npts > maxpts || a->npts * nphi0max > maxwork ||
                    a->shptr[a->npts] > maxsh || a->rshptr[a->npts] > maxrad) {
                    *outofmem = 1;
                } else {
                    icell = adaptind[i - 1] / 4;
                    grid_smooth_test(a, icell, &idir);
                    if (idir > 0) {
                        if (oracle_divide_cell(a, icell, idir, newpoints)) {
                            if (errmsg) snprintf(errmsg, 600, "DIVIDE_CELL: Cannot divide already split cell.");
                            ierr = 1; break;
                        }
                        ierr = interpolate_point(a, cst, newpoints, errmsg);
                        if (ierr) break;
                    }
                }
                i = i + 1;
            }
        }
        if (outofmem0) *outofmem = 1;
        free(adaptcrit); free(adaptind);
        if (ierr) return ierr;
    }
    *splitcrit = 0.0f;
    for (icell = 1; icell <= a->ncells; icell++)
        if (ATREEPTR(a, 2, icell) == 0) {
            oracle_cell_split_test(a, nstokes, icell, adapt, &crit, &idir);
            if (crit > *splitcrit) *splitcrit = crit;
        }
    return 0;
}

/* BOUNDARY_PNTS  shdomsub1.f:2173-2215 */
int oracle_boundary_pnts(int npts, int nang, int lambertian, int maxnbc, int maxbcrad, float zbot, float ztop,
                         const float *gridpos, int *ntoppts, int *nbotpts, int *bcptr)
{
    const int na = lambertian ? 1 : nang / 2 + 1;
    int i, it = 0, ib = 0;
    for (i = 1; i <= npts; i++) {
        if (gridpos[2 + 3 * (size_t)(i - 1)] >= ztop) {
            it = it + 1;
            if (it > maxnbc || it > maxbcrad) return 1;
            bcptr[it - 1] = i;
        }
        if (gridpos[2 + 3 * (size_t)(i - 1)] <= zbot) {
            ib = ib + 1;
            if (ib > maxnbc || it + ib * na > maxbcrad) return 1;
            bcptr[maxnbc + ib - 1] = i;
        }
    }
    *ntoppts = it;
    *nbotpts = ib;
    return 0;
}

/* TRIDIAG  shdomsub2.f:999-1057 (1-based arrays) */
static int tridiag(int n, double *lower, double *diag, double *upper, double *rhs)
{
    int k, kb;
    double t;
    if (n == 1) {
        if (diag[1] == 0.0) return 1;
        rhs[1] = rhs[1] / diag[1];
    }
    lower[1] = diag[1];
    diag[1] = upper[1];
    upper[1] = 0.0;
    upper[n] = 0.0;
    for (k = 1; k <= n - 1; k++) {
        if (fabs(lower[k + 1]) >= fabs(lower[k])) {
            t = lower[k + 1]; lower[k + 1] = lower[k]; lower[k] = t;
            t = diag[k + 1]; diag[k + 1] = diag[k]; diag[k] = t;
            t = upper[k + 1]; upper[k + 1] = upper[k]; upper[k] = t;
            t = rhs[k + 1]; rhs[k + 1] = rhs[k]; rhs[k] = t;
        }
        if (lower[k] == 0.0) return 1;
        t = -lower[k + 1] / lower[k];
        lower[k + 1] = diag[k + 1] + t * diag[k];
        diag[k + 1] = upper[k + 1] + t * upper[k];
        upper[k + 1] = 0.0;
        rhs[k + 1] = rhs[k + 1] + t * rhs[k];
    }
    if (lower[n] == 0.0) return 1;
    rhs[n] = rhs[n] / lower[n];
    rhs[n - 1] = (rhs[n - 1] - diag[n - 1] * rhs[n]) / lower[n - 1];
    for (kb = 1; kb <= n - 2; kb++) {
        k = n - 2 - kb + 1;
        rhs[k] = (rhs[k] - diag[k] * rhs[k + 1] - upper[k] * rhs[k + 2]) / lower[k];
    }
    return 0;
}

/* EDDRTF  shdomsub2.f:765-995.  fluxes[3,nlayer+1]. */
int oracle_eddrtf(int nlayer, const float *optdepths, const float *albedos, const float *asymmetries,
                  const float *temps, int deltam, int srctype, float solarflux, float solarmu, float gndtemp,
                  float gndemis, float skyrad, int units, const float *waveno, float wavelen, float *fluxes,
                  float surface_flux)
{
    const double pi = (double)3.1415926535f;      /* PARAMETER (PI=3.1415926535): a REAL literal */
    const int n = 2 * nlayer + 2;
    double *lower = (double *)calloc(n + 2, sizeof(double)), *upper = (double *)calloc(n + 2, sizeof(double));
    double *diag = (double *)calloc(n + 2, sizeof(double)), *rhs = (double *)calloc(n + 2, sizeof(double));
    double deltau, g, omega, f, lambda = 0, r = 0, t = 0, d = 0, cp, cm, a, b, x1 = 0, x2 = 0;
    double reflect, trans, sourcep, sourcem, radp1p, radp1m, radp2p, radp2m;
    double mu0, skyflux, gndflux, planck1 = 0, planck2, c, tau, exlp = 0, exlm = 0, v, ds, b1, b2, solpp, solpm;
    float bbrad;
    int l, i, ierr = 0;
#define FLUXES(k, l) fluxes[((k) - 1) + 3 * ((l) - 1)]
    if (srctype == 'T') {
        bbrad = oracle_planck_function(temps[0], units, waveno, wavelen);
        planck1 = pi * bbrad;
    }
    mu0 = fabsf(solarmu);
    tau = 0.0;
    i = 2;
    for (l = 1; l <= nlayer; l++) {
        deltau = optdepths[l - 1];
        if (deltau < 0.0) { ierr = 1; goto done; }
        if (deltau == 0.0) {
            trans = 1.0; reflect = 0.0; sourcep = 0.0; sourcem = 0.0;
        } else {
            omega = albedos[l - 1];
            g = asymmetries[l - 1];
            if (deltam) {
                f = g * g;
                deltau = (1 - omega * f) * deltau;
                omega = (1 - f) * omega / (1 - omega * f);
                g = (g - f) / (1 - f);
            }
            r = (1.0 - omega * (4.0 - 3.0 * g)) / 4.0;
            t = (7.0 - omega * (4.0 + 3.0 * g)) / 4.0;
            lambda = sqrt(3.0 * (1.0 - omega) * (1.0 - omega * g));
            if (lambda == 0.0) {
                d = 1.0 / (1.0 + t * deltau);
                trans = d;
                reflect = -r * deltau * d;
            } else {
                x1 = -r;
                x2 = lambda + t;
                exlp = exp(fmin(lambda * deltau, 75.0));
                exlm = 1.0 / exlp;
                trans = 2. * lambda / (x2 * exlp + (lambda - t) * exlm);
                reflect = x1 * (exlp - exlm) * trans / (2. * lambda);
                d = 1.0 / (x2 * x2 * exlp - x1 * x1 * exlm);
            }
            if (srctype == 'T') {
                bbrad = oracle_planck_function(temps[l], units, waveno, wavelen);
                planck2 = pi * bbrad;
                v = 2.0 * (planck2 - planck1) / (3.0 * (1. - omega * g) * deltau);
                radp1p = -v + planck1;
                radp2m = v + planck2;
                radp2p = -v + planck2;
                radp1m = v + planck1;
                if (lambda == 0.0) {
                    a = (r * deltau * radp1p - radp2m) * d;
                    b = -(r * radp1p + t * radp2m) * d;
                    sourcep = (b - t * (a + b * deltau)) / r + radp2p;
                    sourcem = a + radp1m;
                } else {
                    cp = (x1 * exlm * radp1p - x2 * radp2m) * d;
                    cm = (-x2 * exlp * radp1p + x1 * radp2m) * d;
                    sourcep = x1 * cp * exlp + x2 * cm * exlm + radp2p;
                    sourcem = x2 * cp + x1 * cm + radp1m;
                }
                planck1 = planck2;
                FLUXES(3, l) = 0.0f;
            } else {
                FLUXES(3, l) = (float)(solarflux * exp(-tau / mu0));
                ds = 1.0 / (lambda * lambda - 1.0 / (mu0 * mu0));
                b1 = 0.5 * omega * (solarflux / mu0) * exp(-tau / mu0) * ds;
                b2 = 0.5 * omega * (solarflux / mu0) * exp(-(tau + deltau) / mu0) * ds;
                solpp = 1.0 + 1.5 * g * mu0;
                solpm = -1.0 + 1.5 * g * mu0;
                radp1p = ((t + 1.0 / mu0) * solpp + r * solpm) * b1;
                radp2m = ((-t + 1.0 / mu0) * solpm - r * solpp) * b2;
                radp2p = ((t + 1.0 / mu0) * solpp + r * solpm) * b2;
                radp1m = ((-t + 1.0 / mu0) * solpm - r * solpp) * b1;
                if (lambda == 0.0) {
                    a = (r * deltau * radp1p - radp2m) * d;
                    b = -(r * radp1p + t * radp2m) * d;
                    sourcep = (b - t * (a + b * deltau)) / r + radp2p;
                    sourcem = a + radp1m;
                } else {
                    cp = (x1 * exlm * radp1p - x2 * radp2m) * d;
                    cm = (-x2 * exlp * radp1p + x1 * radp2m) * d;
                    sourcep = x1 * cp * exlp + x2 * cm * exlm + radp2p;
                    sourcem = x2 * cp + x1 * cm + radp1m;
                }
                tau = tau + deltau;
            }
        }
        diag[i] = -reflect;
        diag[i + 1] = -reflect;
        lower[i] = 1.0;
        lower[i + 1] = -trans;
        upper[i] = -trans;
        upper[i + 1] = 1.0;
        rhs[i] = sourcem;
        rhs[i + 1] = sourcep;
        i = i + 2;
    }
    if (srctype == 'S') {
        FLUXES(3, nlayer + 1) = (float)(solarflux * exp(-tau / mu0));
        gndflux = (1.0f - gndemis) * solarflux * exp(-tau / mu0);
        skyflux = pi * skyrad;
    } else {
        const float zero = 0.0f;
        (void)zero;
        FLUXES(3, nlayer + 1) = 0.0f;
        bbrad = oracle_planck_function(gndtemp, units, waveno, wavelen);
        gndflux = pi * bbrad * gndemis;
        bbrad = oracle_planck_function(skyrad, units, waveno, wavelen);
        skyflux = pi * bbrad;
    }
    gndflux = gndflux + surface_flux;
    rhs[1] = skyflux;
    diag[1] = 0.0;
    upper[1] = 1.0;
    diag[n] = -(1.0f - gndemis);
    lower[n] = 1.0;
    rhs[n] = gndflux;
    if (tridiag(n, lower, diag, upper, rhs)) { ierr = 2; goto done; }
    if (units == 'T') c = 1.0 / pi; else c = 1.0;
    i = 1;
    for (l = 1; l <= nlayer + 1; l++) {
        FLUXES(1, l) = (float)(c * rhs[i]);
        FLUXES(2, l) = (float)(c * rhs[i + 1]);
        i = i + 2;
    }
#undef FLUXES
done:
    free(lower); free(upper); free(diag); free(rhs);
    return ierr;
}

/* INIT_RADIANCE  shdomsub2.f:614-762: Eddington first guess on the NXY columns of the base grid.
 * The point arrays have leading dimension ld (species stride); nbpts = nxy*nz points, z fastest. */
int oracle_init_radiance(const oracle_state *st, int ld, int nxy, int nz, const float *extinct, const float *albedo,
                         const float *total_ext, const float *temp, const int *iphase, const float *phaseinterpwt,
                         float skyrad, float surface_flux, int *rshptr, float *radiance)
{
    const int nstokes = st->nstokes, nstleg = st->nstleg, nleg = st->nleg, npart = st->npart, ml = st->ml;
    const int nq = 8 * st->maxnmicro, nlayer = nz - 1;
    float *optdepths = (float *)calloc(nz + 1, sizeof(float)), *albedos = (float *)calloc(nz + 1, sizeof(float));
    float *asymmetries = (float *)calloc(nz + 1, sizeof(float)), *temps = (float *)calloc(nz + 1, sizeof(float));
    float *fluxes = (float *)calloc(3 * (nz + 1), sizeof(float));
    float *f0 = (float *)calloc(npart, sizeof(float)), *f1 = (float *)calloc(npart, sizeof(float));
    float *lt0 = (float *)calloc(npart, sizeof(float)), *lt1 = (float *)calloc(npart, sizeof(float));
    const float pi = acosf(-1.0f);
    const float c0 = sqrtf(1.0f / pi), c1 = sqrtf(3.0f / (4 * pi));
    const float wn[2] = {st->waveno0, st->waveno1};
    float gndemis;
    int i, iz, ir, l, q, ipa, k, ierr = 0;
#define PT(iz, i) (((iz) - 1) + (size_t)nz * ((i) - 1))
#define EXT(iz, i, ipa) extinct[PT(iz, i) + (size_t)ld * (ipa)]
#define ALB(iz, i, ipa) albedo[PT(iz, i) + (size_t)ld * (ipa)]
#define IPH(q, iz, i, ipa) iphase[(q) + (size_t)nq * (PT(iz, i) + (size_t)ld * (ipa))]
#define PWT(q, iz, i, ipa) phaseinterpwt[(q) + (size_t)nq * (PT(iz, i) + (size_t)ld * (ipa))]
#define LEG1(l, iph) st->legen[nstleg * ((l) + (size_t)(nleg + 1) * ((iph) - 1))]
    ir = 0;
    rshptr[0] = 0;
    for (i = 1; i <= nxy && !ierr; i++) {
        for (iz = 1; iz <= nlayer; iz++) {
            float ext0, ext1, scat0 = 0.0f, scat1 = 0.0f, g0 = 0.0f, g1 = 0.0f;
            l = nz - iz;
            ext0 = total_ext[PT(iz, i)];
            ext1 = total_ext[PT(iz + 1, i)];
            for (ipa = 0; ipa < npart; ipa++) {
                scat0 = scat0 + ALB(iz, i, ipa) * EXT(iz, i, ipa);
                scat1 = scat1 + ALB(iz + 1, i, ipa) * EXT(iz + 1, i, ipa);
            }
            optdepths[l - 1] = (st->zgrid[iz] - st->zgrid[iz - 1]) * (ext0 + ext1) / 2;
            if (ext0 + ext1 > 0.0f) albedos[l - 1] = (scat0 + scat1) / (ext0 + ext1);
            else albedos[l - 1] = 0.0f;
            if (!st->interp_new) {
                for (ipa = 0; ipa < npart; ipa++) {
                    g0 = g0 + ALB(iz, i, ipa) * EXT(iz, i, ipa) * LEG1(1, IPH(0, iz, i, ipa));
                    g1 = g1 + ALB(iz + 1, i, ipa) * EXT(iz + 1, i, ipa) * LEG1(1, IPH(0, iz + 1, i, ipa));
                }
            } else {
                /* the dominant-table test reads species 1 for all species (shdomsub2.f:682,691) */
                if (PWT(0, iz, i, 0) >= st->phasemax) {
                    for (ipa = 0; ipa < npart; ipa++) lt0[ipa] = LEG1(1, IPH(0, iz, i, ipa));
                } else {
                    for (ipa = 0; ipa < npart; ipa++) lt0[ipa] = 0.0f;
                    for (q = 0; q < nq; q++)
                        for (ipa = 0; ipa < npart; ipa++)
                            lt0[ipa] = lt0[ipa] + LEG1(1, IPH(q, iz, i, ipa)) * PWT(q, iz, i, ipa);
                }
                if (PWT(0, iz + 1, i, 0) >= st->phasemax) {
                    for (ipa = 0; ipa < npart; ipa++) lt1[ipa] = LEG1(1, IPH(0, iz + 1, i, ipa));
                } else {
                    for (ipa = 0; ipa < npart; ipa++) lt1[ipa] = 0.0f;
                    for (q = 0; q < nq; q++)
                        for (ipa = 0; ipa < npart; ipa++)
                            lt1[ipa] = lt1[ipa] + LEG1(1, IPH(q, iz + 1, i, ipa)) * PWT(q, iz + 1, i, ipa);
                }
                if (st->deltam) {
                    if (PWT(0, iz, i, 0) >= st->phasemax) {
                        for (ipa = 0; ipa < npart; ipa++) f0[ipa] = LEG1(ml + 1, IPH(0, iz, i, ipa));
                    } else {
                        for (ipa = 0; ipa < npart; ipa++) f0[ipa] = 0.0f;
                        for (q = 0; q < nq; q++)
                            for (ipa = 0; ipa < npart; ipa++)
                                f0[ipa] = f0[ipa] + LEG1(ml + 1, IPH(q, iz, i, ipa)) * PWT(q, iz, i, ipa);
                    }
                    if (PWT(0, iz + 1, i, 0) >= st->phasemax) {
                        for (ipa = 0; ipa < npart; ipa++) f1[ipa] = LEG1(ml + 1, IPH(0, iz + 1, i, ipa));
                    } else {
                        for (ipa = 0; ipa < npart; ipa++) f1[ipa] = 0.0f;
                        for (q = 0; q < nq; q++)
                            for (ipa = 0; ipa < npart; ipa++)
                                f1[ipa] = f1[ipa] + LEG1(ml + 1, IPH(q, iz + 1, i, ipa)) * PWT(q, iz + 1, i, ipa);
                    }
                    for (ipa = 0; ipa < npart; ipa++) {
                        lt0[ipa] = lt0[ipa] / (1 - f0[ipa]);
                        lt1[ipa] = lt1[ipa] / (1 - f1[ipa]);
                    }
                }
                for (ipa = 0; ipa < npart; ipa++) {
                    g0 = g0 + ALB(iz, i, ipa) * EXT(iz, i, ipa) * lt0[ipa];
                    g1 = g1 + ALB(iz + 1, i, ipa) * EXT(iz + 1, i, ipa) * lt1[ipa];
                }
            }
            if (scat0 + scat1 > 0.0f) asymmetries[l - 1] = (g0 + g1) / (scat0 + scat1);
            else asymmetries[l - 1] = 0.0f;
            temps[l] = temp ? temp[PT(iz, i)] : 0.0f;
        }
        temps[0] = temp ? temp[PT(nz, i)] : 0.0f;
        gndemis = 1.0f - st->gndalbedo;
        ierr = oracle_eddrtf(nlayer, optdepths, albedos, asymmetries, temps, 0, st->srctype, st->solarflux,
                             st->solarmu, st->gndtemp, gndemis, skyrad, st->units, wn, st->wavelen, fluxes,
                             surface_flux);
        if (ierr) break;
        for (iz = 1; iz <= nz; iz++) {
            l = nz + 1 - iz;
            for (k = 0; k < 4 * nstokes; k++) radiance[(size_t)nstokes * ir + k] = 0.0f;
            radiance[(size_t)nstokes * ir] = c0 * (fluxes[3 * (l - 1)] + fluxes[1 + 3 * (l - 1)]);
            radiance[(size_t)nstokes * (ir + 2)] = c1 * (fluxes[3 * (l - 1)] - fluxes[1 + 3 * (l - 1)]);
            ir = ir + 4;
            rshptr[iz + nz * (i - 1)] = ir;
        }
    }
#undef PT
#undef EXT
#undef ALB
#undef IPH
#undef PWT
#undef LEG1
    free(optdepths); free(albedos); free(asymmetries); free(temps); free(fluxes);
    free(f0); free(f1); free(lt0); free(lt1);
    return ierr;
}

/* INTERP_RADIANCE  shdomsub2.f:1924-1987 */
int oracle_interp_radiance(int nstokes, int oldnpts, int nbcells, int ncells, const int *treeptr, const int *gridptr,
                           const float *gridpos, int *rshptr, float *radiance)
{
    int icell, ic, idir, i, j, k;
    if (oldnpts <= 0) return 0;
    for (icell = nbcells + 1; icell <= ncells; icell++)
        for (ic = 1; ic <= 8; ic++) {
            const int ip = gridptr[(ic - 1) + 8 * (size_t)(icell - 1)];
            if (ip > oldnpts) {
                const int iparent = treeptr[2 * (size_t)(icell - 1)];
                int ip1 = 0, ip2 = 0, found = 0, ir1, ir2, nr1, nr2, nr, ir;
                for (idir = 1; idir <= 3 && !found; idir++) {
                    const int dir1 = (idir - 1 + 1) % 3 + 1, dir2 = (idir - 1 + 2) % 3 + 1;
                    for (i = 1; i <= 4; i++) {
                        ip1 = gridptr[(GRIDCORNER[idir - 1][i - 1][0] - 1) + 8 * (size_t)(iparent - 1)];
                        ip2 = gridptr[(GRIDCORNER[idir - 1][i - 1][1] - 1) + 8 * (size_t)(iparent - 1)];
                        if (gridpos[(dir1 - 1) + 3 * (size_t)(ip1 - 1)] == gridpos[(dir1 - 1) + 3 * (size_t)(ip - 1)] &&
                            gridpos[(dir2 - 1) + 3 * (size_t)(ip1 - 1)] == gridpos[(dir2 - 1) + 3 * (size_t)(ip - 1)]) {
                            found = 1;
                            break;
                        }
                    }
                }
                if (!found) return 1;
                ir1 = rshptr[ip1 - 1]; ir2 = rshptr[ip2 - 1];
                nr1 = rshptr[ip1] - ir1; nr2 = rshptr[ip2] - ir2;
                nr = nr1 > nr2 ? nr1 : nr2;
                ir = rshptr[ip - 1];
                rshptr[ip] = ir + nr;
                for (j = 1; j <= nr; j++)
                    for (k = 0; k < nstokes; k++) {
                        const float r1 = j <= nr1 ? radiance[k + (size_t)nstokes * (ir1 + j - 1)] : 0.0f;
                        const float r2 = j <= nr2 ? radiance[k + (size_t)nstokes * (ir2 + j - 1)] : 0.0f;
                        radiance[k + (size_t)nstokes * (ir + j - 1)] = 0.5f * (r1 + r2);
                    }
            }
        }
    return 0;
}
