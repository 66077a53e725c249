// at3d_adapt.cu -- the adaptive grid of the SHDOM solution iterations and the Eddington first guess.
//
// Replaces (paths relative to the AT3D checkout):
//   SPLIT_GRID        src/polarized/shdomsub1.f:4703-4932   (driver: batches, memory limits, grid smoothing)
//   CELL_SPLIT_TEST   :5703-5791     DIVIDE_CELL :5289-5364      NEW_GRID_POINTS :5514-5596
//   MATCH_GRID_POINT  :5600-5697     MATCH_NEIGHBOR_FACE :5370-5458   INHERIT_NEIGHBOR :5462-5506
//   GRID_SMOOTH_TEST  :5796-5902     SSORT src/polarized/shdomsub2.f:4961-5244
//   INTERPOLATE_POINT :4937-5283     (property interpolation / direct beam / radiance / source of new points)
//   INIT_RADIANCE + EDDRTF + TRIDIAG  src/polarized/shdomsub2.f:614-1057
//
// Division of labour.  The splitting criterion is a sum over the SH source of the 8 corner points of every leaf cell --
// data-parallel and the only part that touches the big arrays: `edge_terms_kernel` evaluates the 12 edge terms of every
// listed cell on the GPU (sequential REAL sums per edge, i.e. the reference's rounding).  The tree surgery is inherently
// serial and touches a few integers per cell: it runs on the host on the caller's GRIDPTR/NEIGHPTR/TREEPTR/CELLFLAGS/
// GRIDPOS arrays (the reference's layout, 1-based contents), in the reference's visiting order (SSORT's order of equal
// keys included) so that cells and points get the reference's numbers.  1-EXP(-TAU) is evaluated on the host with the
// reference's libm so that the decisions ADAPTCRIT > CURSPLITACC agree bit for bit given the same SOURCE.  All new
// points of a batch have parents that existed before the batch, so their properties (TRILIN_INTERP_PROP, direct beam),
// radiance and source are evaluated together by kernels after the host pass.
#include "at3d_mem.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include <omp.h>
#include "at3d_host.h"
#include "at3d_adapt.h"

namespace {

const int EDGE_CORNER[3][4][2] = {
    {{1, 2}, {3, 4}, {5, 6}, {7, 8}}, {{1, 3}, {2, 4}, {5, 7}, {6, 8}}, {{1, 5}, {2, 6}, {3, 7}, {4, 8}}};
const int OPPOSITE[6] = {2, 1, 4, 3, 6, 5};
const int FACE_POINTS[6][4] = {{1, 3, 5, 7}, {2, 4, 6, 8}, {1, 2, 5, 6}, {3, 4, 7, 8}, {1, 2, 3, 4}, {5, 6, 7, 8}};

// ---- the 12 edge terms of the listed cells: thread = (cell, edge) ----
// out[12*k + e] = (C0*SQRT(JAY)/EXT, TAU) of edge e = 4*(ID-1) + (IE-1) of cell cells[k]; x < 0 marks IP1 == IP2.
__global__ void edge_terms_kernel(int n, const int *cells, const int *gridptr, const float *gridpos, const float *total_ext,
                                  const int *shptr, const float *source, int nst, float2 *out)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * 12) return;
    const int k = (int)(t / 12), e = (int)(t % 12), id = e / 4, ie = e % 4;
    static const int c1[12] = {1, 3, 5, 7, 1, 2, 5, 6, 1, 2, 3, 4}, c2[12] = {2, 4, 6, 8, 3, 4, 7, 8, 5, 6, 7, 8};
    (void)ie;
    const int icell = cells[k];
    const int ip1 = gridptr[(c1[e] - 1) + 8 * (size_t)(icell - 1)], ip2 = gridptr[(c2[e] - 1) + 8 * (size_t)(icell - 1)];
    if (ip1 == ip2) { out[t] = make_float2(-1.0f, 0.0f); return; }
    const int is1 = shptr[ip1 - 1], is2 = shptr[ip2 - 1];
    const int ns1 = shptr[ip1] - is1, ns2 = shptr[ip2] - is2, ns = ns1 < ns2 ? ns1 : ns2;
    const float e1 = total_ext[ip1 - 1], e2 = total_ext[ip2 - 1];
    const float *s1 = source + (size_t)nst * is1, *s2 = source + (size_t)nst * is2;
    float jay = 0.0f;
    for (int j = 0; j < ns; j++) { const float d = e2 * s2[(size_t)nst * j] - e1 * s1[(size_t)nst * j]; jay = jay + d * d; }
    for (int j = ns; j < ns1; j++) { const float d = e1 * s1[(size_t)nst * j]; jay = jay + d * d; }
    for (int j = ns; j < ns2; j++) { const float d = e2 * s2[(size_t)nst * j]; jay = jay + d * d; }
    const float ext = 0.5f * (e1 + e2);
    if (ext > 0.0f) jay = 0.282095f * sqrtf(jay) / ext; else jay = 0.0f;
    const float tau = fabsf(ext * (gridpos[id + 3 * (size_t)(ip2 - 1)] - gridpos[id + 3 * (size_t)(ip1 - 1)]));
    out[t] = make_float2(fabsf(jay), tau);
}

// SSORT with KFLAG = -2 on the host (the order of equal keys decides the numbering of the smoothing splits)
void ssort_desc(std::vector<float> &xv, std::vector<int> &yv, int n)
{
    if (n < 1) return;
    float *x = xv.data() - 1;
    int *y = yv.data() - 1;
    float r = 0.375f, t, tt;
    int ty, tty, i = 1, j = n, k, l, m = 1, ij, il[64], iu[64];
    for (k = 1; k <= n; k++) x[k] = -x[k];
    enum { PICK, PART, POP, SMALL } st = PICK;
    for (;;) {
        if (st == PICK) {
            if (i == j) { st = POP; continue; }
            if (r <= 0.5898437f) r = r + 3.90625e-2f; else r = r - 0.21875f;
            st = PART;
        }
        if (st == PART) {
            k = i;
            ij = i + (int)((j - i) * r);
            t = x[ij]; ty = y[ij];
            if (x[i] > t) { x[ij] = x[i]; x[i] = t; t = x[ij]; y[ij] = y[i]; y[i] = ty; ty = y[ij]; }
            l = j;
            if (x[j] < t) {
                x[ij] = x[j]; x[j] = t; t = x[ij]; y[ij] = y[j]; y[j] = ty; ty = y[ij];
                if (x[i] > t) { x[ij] = x[i]; x[i] = t; t = x[ij]; y[ij] = y[i]; y[i] = ty; ty = y[ij]; }
            }
            for (;;) {
                do { l--; } while (x[l] > t);
                do { k++; } while (x[k] < t);
                if (k > l) break;
                tt = x[l]; x[l] = x[k]; x[k] = tt; tty = y[l]; y[l] = y[k]; y[k] = tty;
            }
            if (l - i > j - k) { il[m] = i; iu[m] = l; i = k; m++; }
            else { il[m] = k; iu[m] = j; j = l; m++; }
            st = SMALL;
        }
        if (st == POP) {
            m--;
            if (m == 0) break;
            i = il[m]; j = iu[m];
            st = SMALL;
        }
        if (st == SMALL) {
            if (j - i >= 1) { st = PART; continue; }
            if (i == 1) { st = PICK; continue; }
            i--;
            for (;;) {
                i++;
                if (i == j) break;
                t = x[i + 1]; ty = y[i + 1];
                if (x[i] <= t) continue;
                k = i;
                do { x[k + 1] = x[k]; y[k + 1] = y[k]; k--; } while (t < x[k]);
                x[k + 1] = t; y[k + 1] = ty;
            }
            st = POP;
        }
    }
    for (k = 1; k <= n; k++) x[k] = -x[k];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// the host view of the cell tree
// ---------------------------------------------------------------------------------------------------------------------
int AdaptGrid::split_dir(int ic) const { return (cellflags[ic - 1] >> 2) & 3; }

// NEXT_CELL (shdomsub1.f:4470-4522) on the host arrays
int AdaptGrid::next_cell(double xe, double ye, double ze, int iface, int jface, int icell) const
{
    int inext = nb(iface, icell);
    if (inext >= 0) return inext;
    int ic = -inext;
    while (tree(2, ic) > 0) {
        const int dir = split_dir(ic), ic1 = tree(2, ic);
        if (dir == jface) ic = ic1 + 1 - ((iface - 1) % 2);
        else {
            const double p = dir == 1 ? xe : (dir == 2 ? ye : ze);
            ic = ic1 + (p > pos(dir, gp(8, ic1)) ? 1 : 0);
        }
    }
    return ic;
}

// MATCH_GRID_POINT: an existing grid point at (xp,yp,zp) among the face points of the leaf cells across face `iface`
int AdaptGrid::match_grid_point(float xp, float yp, float zp, int icell, int iface) const
{
    const int idir = (iface + 1) / 2, kface = OPPOSITE[iface - 1];
    int ic = std::abs(nb(iface, icell));
    if (ic == 0) return 0;
    int stack[64], sp = 0;
    for (;;) {
        while (tree(2, ic) == 0) {
            for (int i = 0; i < 4; i++) {
                const int ipt = gp(FACE_POINTS[kface - 1][i], ic);
                if (xp == pos(1, ipt) && yp == pos(2, ipt) && zp == pos(3, ipt)) return ipt;
            }
            if (sp == 0) return 0;
            ic = stack[--sp];
        }
        const int dir = split_dir(ic), ic1 = tree(2, ic);
        if (dir == idir) { ic = ic1 + 1 - ((iface - 1) % 2); continue; }
        const float p = dir == 1 ? xp : (dir == 2 ? yp : zp), line = pos(dir, gp(8, ic1));
        ic = ic1;
        if (p == line) { if (sp < 63) stack[sp++] = ic1 + 1; }      // on the split line: both children
        else if (p > line) ic = ic1 + 1;
    }
}

void AdaptGrid::inherit_neighbor(int icell, int iface, int in)
{
    const int jface = (iface + 1) / 2;
    int stack[50], sp = 0, ic = icell;
    for (;;) {
        nb(iface, ic) = in;
        if (tree(2, ic) == 0) {
            if (sp == 0) return;
            ic = stack[--sp];
        } else if (split_dir(ic) == jface) {
            ic = tree(2, ic) + ((iface - 1) % 2);
        } else {
            if (sp >= 50) return;
            stack[sp++] = tree(2, ic) + 1;
            ic = tree(2, ic);
        }
    }
}

void AdaptGrid::match_neighbor_face(int iface, int ic)
{
    int in = std::abs(nb(iface, ic));
    if (in == 0) return;
    const int jface = (iface + 1) / 2, ic1 = gp(1, ic), ic8 = gp(8, ic);
    const int dir1 = jface % 3 + 1, dir2 = (jface + 1) % 3 + 1;
    float centre[4] = {0, 0, 0, 0};
    centre[dir1] = (pos(dir1, ic1) + pos(dir1, ic8)) / 2;
    centre[dir2] = (pos(dir2, ic1) + pos(dir2, ic8)) / 2;
    bool done = false;
    while (!done && ic != in) {
        const int in1 = gp(1, in), in8 = gp(8, in);
        // the new face lies inside the neighbour's face: point at it (negative if it has children)
        if (pos(dir1, ic1) >= pos(dir1, in1) && pos(dir1, ic8) <= pos(dir1, in8) &&
            pos(dir2, ic1) >= pos(dir2, in1) && pos(dir2, ic8) <= pos(dir2, in8))
            nb(iface, ic) = tree(2, in) == 0 ? in : -in;
        // the neighbour's face lies inside the new face: it (and its children on that face) now border the new cell
        if (pos(dir1, in1) >= pos(dir1, ic1) && pos(dir1, in8) <= pos(dir1, ic8) &&
            pos(dir2, in1) >= pos(dir2, ic1) && pos(dir2, in8) <= pos(dir2, ic8))
            inherit_neighbor(in, OPPOSITE[iface - 1], ic);
        else
            nb(OPPOSITE[iface - 1], in) = -std::abs(nb(OPPOSITE[iface - 1], in));
        if (tree(2, in) == 0) done = true;
        else {
            const int dir = split_dir(in), inn = tree(2, in);
            if (dir == jface) in = inn + 1 - ((iface - 1) % 2);
            else in = inn + (centre[dir] > pos(dir, gp(8, inn)) ? 1 : 0);
        }
    }
}

// DIVIDE_CELL + NEW_GRID_POINTS: returns false if the cell is not a leaf.  newpts[i] = {parent 1, parent 2, new point or 0}
bool AdaptGrid::divide_cell(int icell, int idir, int newpts[4][3])
{
    static const int FACEGRID[3][4][2] = {
        {{3, 5}, {4, 5}, {3, 6}, {4, 6}}, {{1, 5}, {2, 5}, {1, 6}, {2, 6}}, {{1, 3}, {2, 3}, {1, 4}, {2, 4}}};
    if (tree(2, icell) != 0) return false;
    const int c0 = ncells + 1, c1 = ncells + 2;
    ncells += 2;
    tree(2, icell) = c0;
    tree(1, c0) = icell; tree(2, c0) = 0;
    tree(1, c1) = icell; tree(2, c1) = 0;
    const short inherited = cellflags[icell - 1] & 3;
    cellflags[icell - 1] = (short)(cellflags[icell - 1] | (idir << 2));
    cellflags[c0 - 1] = inherited;
    cellflags[c1 - 1] = inherited;
    for (int k = 1; k <= 8; k++) { gp(k, c0) = gp(k, icell); gp(k, c1) = gp(k, icell); }
    for (int i = 0; i < 4; i++) {
        const int i1 = EDGE_CORNER[idir - 1][i][0], i2 = EDGE_CORNER[idir - 1][i][1];
        const int ip1 = gp(i1, icell), ip2 = gp(i2, icell);
        const float xp = (pos(1, ip1) + pos(1, ip2)) / 2, yp = (pos(2, ip1) + pos(2, ip2)) / 2, zp = (pos(3, ip1) + pos(3, ip2)) / 2;
        const int f1 = FACEGRID[idir - 1][i][0], f2 = FACEGRID[idir - 1][i][1];
        int ipm = match_grid_point(xp, yp, zp, icell, f1);
        if (!ipm) ipm = match_grid_point(xp, yp, zp, icell, f2);
        if (!ipm) {
            const int across = std::abs(nb(f1, icell));
            if (across > 0) ipm = match_grid_point(xp, yp, zp, across, f2);
        }
        if (!ipm) {
            npts++;
            ipm = npts;
            pos(1, npts) = xp; pos(2, npts) = yp; pos(3, npts) = zp;
            newpts[i][0] = ip1; newpts[i][1] = ip2; newpts[i][2] = npts;
        } else {
            newpts[i][2] = 0;
        }
        gp(i2, c0) = ipm;
        gp(i1, c1) = ipm;
    }
    for (int iface = 1; iface <= 6; iface++) {
        const int pn = nb(iface, icell);
        if (pn == icell) { nb(iface, c0) = c0; nb(iface, c1) = c1; }
        else if (iface == 2 * idir) { nb(iface, c0) = c1; nb(iface, c1) = pn; }
        else if (iface == 2 * idir - 1) { nb(iface, c1) = c0; nb(iface, c0) = pn; }
        else { nb(iface, c0) = pn; nb(iface, c1) = pn; }
        match_neighbor_face(iface, c0);
        match_neighbor_face(iface, c1);
    }
    return true;
}

// GRID_SMOOTH_TEST: direction (1..3) in which the cell should be split to keep the grid smooth, 0 for none
int AdaptGrid::grid_smooth_test(int icell) const
{
    static const int EDGE0[3][2] = {{1, 2}, {1, 3}, {1, 5}};
    int idir = 0;
    float sizeratio = 1.0f;
    for (int id = 1; id <= 3; id++) {
        const int ip1 = gp(EDGE0[id - 1][0], icell), ip2 = gp(EDGE0[id - 1][1], icell);
        if (ip1 == ip2) continue;
        const float cur = fabsf(pos(id, ip1) - pos(id, ip2)), inv = 1.0f / cur;
        float gs[2];
        int dsplit[2];
        for (int j = 0; j < 2; j++) {
            const int iface = 2 * (id - 1) + j + 1;
            double xe = 0.0, ye = 0.0, ze = 0.0;
            for (int i = 0; i < 4; i++) {
                const int ip = gp(FACE_POINTS[iface - 1][i], icell);
                xe = xe + pos(1, ip) * 0.25f; ye = ye + pos(2, ip) * 0.25f; ze = ze + pos(3, ip) * 0.25f;
            }
            const int in = next_cell(xe, ye, ze, iface, id, icell);
            if (in == 0) gs[j] = cur;
            else {
                gs[j] = fabsf(pos(id, gp(1, in)) - pos(id, gp(8, in)));
                if (tree(1, in) == 0) gs[j] = cur;                          // a base cell
                if (id <= 2 && ((cellflags[in - 1] >> (id - 1)) & 1)) gs[j] = cur;   // independent-pixel cell
            }
            const int raw = nb(iface, icell);
            dsplit[j] = raw < 0 ? split_dir(-raw) : 0;
        }
        if (gs[0] * inv < 0.75f && gs[1] * inv < 0.75f) idir = id;
        if (dsplit[0] > 0 && dsplit[0] != id && dsplit[0] == dsplit[1]) idir = dsplit[0];
        const float ratio = fminf(gs[0] * inv, gs[1] * inv);
        if (ratio < 0.4f && ratio < sizeratio) {
            int isum = 0;
            for (int i = 1; i <= 8; i++) { const int ip = gp(i, icell); isum += shptr[ip] - shptr[ip - 1]; }
            if (isum > 0) { idir = id; sizeratio = ratio; }
        }
    }
    return idir;
}

// criterion of one cell from its 12 edge terms (the tail of CELL_SPLIT_TEST): max over the directions, first one wins
static inline void cell_criterion(const float2 *t, float *crit, int *dir)
{
    float best = -1.0f;
    int bdir = 1;
    for (int id = 0; id < 3; id++) {
        float sum1 = 0.0f;
        int num = 0;
        for (int ie = 0; ie < 4; ie++) {
            const float2 v = t[4 * id + ie];
            if (v.x < 0.0f) continue;
            num++;
            sum1 = sum1 + v.x * (1 - expf(-v.y));
        }
        const float c = num > 0 ? sum1 / num : 0.0f;
        if (c > best) { best = c; bdir = id + 1; }
    }
    *crit = best; *dir = bdir;
}

// criteria of the listed cells: GPU edge terms + host tail
int AdaptGrid::evaluate(const std::vector<int> &cells, const AdaptDev &D, std::vector<float> &crit, std::vector<int> &dir, char *errmsg)
{
    const int n = (int)cells.size();
    crit.resize(n); dir.resize(n);
    if (n == 0) return 0;
    if (work_cells.reserve((size_t)n * sizeof(int)) != cudaSuccess || work_terms.reserve((size_t)n * 12 * sizeof(float2)) != cudaSuccess) {
        if (errmsg) snprintf(errmsg, 600, "SPLIT_GRID: device allocation failure"); return 4;
    }
    cudaMemcpy(work_cells.p, cells.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice);
    const size_t nt = (size_t)n * 12;
    edge_terms_kernel<<<(unsigned)((nt + 255) / 256), 256>>>(n, (const int *)work_cells.p, D.gridptr, D.gridpos, D.total_ext, D.shptr,
                                                             D.source, D.nst, (float2 *)work_terms.p);
    terms_h.resize(nt);
    cudaError_t e = cudaMemcpy(terms_h.data(), work_terms.p, nt * sizeof(float2), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { if (errmsg) snprintf(errmsg, 600, "CUDA error %s in CELL_SPLIT_TEST", cudaGetErrorString(e)); return 4; }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; k++) cell_criterion(&terms_h[(size_t)12 * k], &crit[k], &dir[k]);
    return 0;
}

// SPLIT_GRID.  On return `newrecs` holds, in creation order, the new points (0-based ids, SH offsets and lengths) that the
// caller has not yet given properties / radiance / source: interpolate_cb is called once per batch for them.
int AdaptGrid::split_grid(const AdaptDev &D, bool dosplit, bool &outofmem, float cursplitacc, float &splitcrit, int nphi0max, int nlm,
                          int (*interpolate_cb)(void *, const std::vector<NewPointRec> &, char *), void *cb_arg, char *errmsg)
{
    std::vector<float> crit;
    std::vector<int> dir, list;
    if (cellcrit.size() < (size_t)maxic) cellcrit.resize(maxic, -1.0f);
    if (!dosplit) {
        for (int ic = 1; ic <= ncells; ic++) if (tree(2, ic) == 0) list.push_back(ic);
        int rc = evaluate(list, D, crit, dir, errmsg);
        if (rc) return rc;
        splitcrit = 0.0f;
        for (float c : crit) splitcrit = c > splitcrit ? c : splitcrit;
        return 0;
    }
    bool outofmem0 = outofmem;
    int icell1 = 1;
    std::vector<NewPointRec> recs;
    while (icell1 <= ncells) {
        list.clear();
        for (int ic = icell1; ic <= ncells; ic++) if (tree(2, ic) == 0) list.push_back(ic);
        int rc = evaluate(list, D, crit, dir, errmsg);
        if (rc) return rc;
        const int n = (int)list.size();
        std::vector<float> key(crit);
        std::vector<int> ind(n);
        for (int k = 0; k < n; k++) { ind[k] = 4 * list[k] + dir[k]; cellcrit[list[k] - 1] = crit[k]; }
        icell1 = ncells + 1;
        ssort_desc(key, ind, n);
        const float frac = 0.03f;
        int maxcells = (int)(maxic - frac * (maxic - ncells) - 2);
        int maxpts = (int)(maxig - frac * (maxig - npts) - 4);
        int maxwork = (int)(maxido - frac * (maxido - nphi0max * npts) - 4 * nphi0max);
        int maxsh = (int)(maxiv - frac * (maxiv - shptr[npts]) - 4 * nlm);
        int maxrad = (int)(maxiv + maxig - frac * (maxiv + maxig - rshptr[npts]) - 4 * nlm);
        outofmem0 = outofmem;
        recs.clear();
        auto full = [&]() {
            return ncells > maxcells || npts > maxpts || npts * nphi0max > maxwork || shptr[npts] > maxsh || rshptr[npts] > maxrad;
        };
        auto split_one = [&](int icell, int idir) -> bool {
            int np[4][3];
            if (!divide_cell(icell, idir, np)) return false;
            for (int i = 0; i < 4; i++)
                if (np[i][2] > 0) {
                    const int ip1 = np[i][0], ip2 = np[i][1], ip = np[i][2];
                    const int nr = std::max(rshptr[ip1] - rshptr[ip1 - 1], rshptr[ip2] - rshptr[ip2 - 1]);
                    const int ns = std::max(shptr[ip1] - shptr[ip1 - 1], shptr[ip2] - shptr[ip2 - 1]);
                    NewPointRec r = {ip1 - 1, ip2 - 1, ip - 1, rshptr[ip - 1], nr, shptr[ip - 1], ns, 0};
                    rshptr[ip] = rshptr[ip - 1] + nr;
                    shptr[ip] = shptr[ip - 1] + ns;
                    recs.push_back(r);
                }
            return true;
        };
        int i = 0;
        while (i < n && key[i] > cursplitacc && !outofmem0) {
            if (full()) outofmem0 = true;
            else if (!split_one(ind[i] / 4, ind[i] & 3)) { if (errmsg) snprintf(errmsg, 600, "DIVIDE_CELL: Cannot divide already split cell."); return 1; }
            i++;
        }
        maxcells = maxic - 2; maxpts = maxig - 4; maxwork = maxido - 4 * nphi0max; maxsh = maxiv - 4 * nlm; maxrad = maxiv + maxig - 4 * nlm;
        while (!outofmem && i < n) {
            if (full()) outofmem = true;
            else {
                const int icell = ind[i] / 4, idir = grid_smooth_test(icell);
                if (idir > 0 && !split_one(icell, idir)) { if (errmsg) snprintf(errmsg, 600, "DIVIDE_CELL: Cannot divide already split cell."); return 1; }
            }
            i++;
        }
        if (!recs.empty()) {
            rc = interpolate_cb(cb_arg, recs, errmsg);
            if (rc) return rc;
        }
    }
    if (outofmem0) outofmem = true;
    // the largest criterion among the leaf cells: every one of them was evaluated in one of the batches above (or in an
    // earlier call, if it has not changed hands since -- but SOURCE has, so the unsplit old cells come from this call)
    splitcrit = 0.0f;
    for (int ic = 1; ic <= ncells; ic++)
        if (tree(2, ic) == 0 && cellcrit[ic - 1] > splitcrit) splitcrit = cellcrit[ic - 1];
    return 0;
}

// BOUNDARY_PNTS (shdomsub1.f:2173-2215)
int AdaptGrid::boundary_points(int nang, bool lambertian, int maxnbc, int maxbcrad, float zbot, float ztop, int *bcptr, int *ntop, int *nbot) const
{
    const int na = lambertian ? 1 : nang / 2 + 1;
    int it = 0, ib = 0;
    for (int i = 1; i <= npts; i++) {
        const float z = pos(3, i);
        if (z >= ztop) { if (++it > maxnbc || it > maxbcrad) return 1; bcptr[it - 1] = i; }
        if (z <= zbot) { if (++ib > maxnbc || it + ib * na > maxbcrad) return 1; bcptr[maxnbc + ib - 1] = i; }
    }
    *ntop = it; *nbot = ib;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// INIT_RADIANCE: Eddington two-stream first guess, thread = base-grid column
// ---------------------------------------------------------------------------------------------------------------------
struct EddArgs {
    int ncol, nz, ld, npart, nq, nstleg, nleg, ml, nstokes, interp_new, deltam, srctype, units;
    float phasemax, solarflux, solarmu, gndalbedo, gndtemp, skyrad, wavelen, surface_flux;
    const float *zgrid, *extinct, *albedo, *total_ext, *temp, *legen, *phaseinterpwt;
    const int *iphase;
    double *scratch;       // [nthreads][4][2*nz+2]
    float *radiance;       // [nstokes, 4*npts]
    int *bad;
};

__device__ __forceinline__ float edd_leg(const EddArgs &a, int l, int p, int ipa)
{
    // LEGEN(1,l,.) of the point's species ipa: dominant table or PHASEINTERPWT mixture (shdomsub2.f:676-700)
    const size_t po = (size_t)p + (size_t)a.ld * ipa;
    const int *iph = a.iphase + (size_t)a.nq * po;
    const float *pw = a.phaseinterpwt + (size_t)a.nq * po;
    const size_t nlt = (size_t)a.nstleg * (a.nleg + 1);
    if (!a.interp_new || a.phaseinterpwt[(size_t)a.nq * p] >= a.phasemax) return a.legen[nlt * (iph[0] - 1) + (size_t)a.nstleg * l];
    float v = 0.0f;
    for (int q = 0; q < a.nq; q++) v = v + a.legen[nlt * (iph[q] - 1) + (size_t)a.nstleg * l] * pw[q];
    return v;
}

__global__ void eddington_kernel(EddArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int nz = a.nz, nlayer = nz - 1, n = 2 * nlayer + 2;
    double *lower = a.scratch + (size_t)tid * 4 * (n + 1), *diag = lower + (n + 1), *upper = diag + (n + 1), *rhs = upper + (n + 1);
    const double pi = (double)3.1415926535f;
    const float pif = acosf(-1.0f), c0 = sqrtf(1.0f / pif), c1 = sqrtf(3.0f / (4 * pif));
    for (int col = tid; col < a.ncol; col += nthreads) {
        const int p0 = nz * col;          // first (bottom) point of the column
        double planck1 = 0.0, tau = 0.0;
        const double mu0 = fabsf(a.solarmu);
        if (a.srctype == 'T') planck1 = pi * dev_planck(a.temp ? a.temp[p0 + nz - 1] : 0.0f, a.units, a.wavelen);
        int i = 2;
        bool bad = false;
        for (int l = 1; l <= nlayer; l++) {
            const int iz = nz - l;        // layer between grid levels iz and iz+1 (1-based), counted from the top
            const int pa = p0 + iz - 1, pb = p0 + iz;
            const float ext0 = a.total_ext[pa], ext1 = a.total_ext[pb];
            float scat0 = 0.0f, scat1 = 0.0f, g0 = 0.0f, g1 = 0.0f;
            for (int ipa = 0; ipa < a.npart; ipa++) {
                scat0 = scat0 + a.albedo[pa + (size_t)a.ld * ipa] * a.extinct[pa + (size_t)a.ld * ipa];
                scat1 = scat1 + a.albedo[pb + (size_t)a.ld * ipa] * a.extinct[pb + (size_t)a.ld * ipa];
            }
            const float optdepth = (a.zgrid[iz] - a.zgrid[iz - 1]) * (ext0 + ext1) / 2;
            const float albedo = ext0 + ext1 > 0.0f ? (scat0 + scat1) / (ext0 + ext1) : 0.0f;
            for (int ipa = 0; ipa < a.npart; ipa++) {
                float lt0 = edd_leg(a, 1, pa, ipa), lt1 = edd_leg(a, 1, pb, ipa);
                if (a.interp_new && a.deltam) {
                    lt0 = lt0 / (1 - edd_leg(a, a.ml + 1, pa, ipa));
                    lt1 = lt1 / (1 - edd_leg(a, a.ml + 1, pb, ipa));
                }
                g0 = g0 + a.albedo[pa + (size_t)a.ld * ipa] * a.extinct[pa + (size_t)a.ld * ipa] * lt0;
                g1 = g1 + a.albedo[pb + (size_t)a.ld * ipa] * a.extinct[pb + (size_t)a.ld * ipa] * lt1;
            }
            const float asym = scat0 + scat1 > 0.0f ? (g0 + g1) / (scat0 + scat1) : 0.0f;
            // ---- EDDRTF layer coefficients (DELTAM=.FALSE.) ----
            double deltau = optdepth, trans, reflect, sourcep, sourcem;
            if (deltau < 0.0) bad = true;
            if (deltau == 0.0) { trans = 1.0; reflect = 0.0; sourcep = 0.0; sourcem = 0.0; }
            else {
                const double omega = albedo, g = asym;
                const double r = (1.0 - omega * (4.0 - 3.0 * g)) / 4.0, t = (7.0 - omega * (4.0 + 3.0 * g)) / 4.0;
                const double lambda = sqrt(3.0 * (1.0 - omega) * (1.0 - omega * g));
                double d, x1 = 0, x2 = 0, exlp = 0, exlm = 0;
                if (lambda == 0.0) { d = 1.0 / (1.0 + t * deltau); trans = d; reflect = -r * deltau * d; }
                else {
                    x1 = -r; x2 = lambda + t;
                    exlp = exp(fmin(lambda * deltau, 75.0)); exlm = 1.0 / exlp;
                    trans = 2. * lambda / (x2 * exlp + (lambda - t) * exlm);
                    reflect = x1 * (exlp - exlm) * trans / (2. * lambda);
                    d = 1.0 / (x2 * x2 * exlp - x1 * x1 * exlm);
                }
                double radp1p, radp1m, radp2p, radp2m;
                if (a.srctype == 'T') {
                    const double planck2 = pi * dev_planck(a.temp ? a.temp[pa] : 0.0f, a.units, a.wavelen);
                    const double v = 2.0 * (planck2 - planck1) / (3.0 * (1. - omega * g) * deltau);
                    radp1p = -v + planck1; radp2m = v + planck2; radp2p = -v + planck2; radp1m = v + planck1;
                    planck1 = planck2;
                } else {
                    const double ds = 1.0 / (lambda * lambda - 1.0 / (mu0 * mu0));
                    const double b1 = 0.5 * omega * (a.solarflux / mu0) * exp(-tau / mu0) * ds;
                    const double b2 = 0.5 * omega * (a.solarflux / mu0) * exp(-(tau + deltau) / mu0) * ds;
                    const double solpp = 1.0 + 1.5 * g * mu0, solpm = -1.0 + 1.5 * g * mu0;
                    radp1p = ((t + 1.0 / mu0) * solpp + r * solpm) * b1;
                    radp2m = ((-t + 1.0 / mu0) * solpm - r * solpp) * b2;
                    radp2p = ((t + 1.0 / mu0) * solpp + r * solpm) * b2;
                    radp1m = ((-t + 1.0 / mu0) * solpm - r * solpp) * b1;
                }
                if (lambda == 0.0) {
                    const double aa = (r * deltau * radp1p - radp2m) * d, bb = -(r * radp1p + t * radp2m) * d;
                    sourcep = (bb - t * (aa + bb * deltau)) / r + radp2p;
                    sourcem = aa + radp1m;
                } else {
                    const double cp = (x1 * exlm * radp1p - x2 * radp2m) * d, cm = (-x2 * exlp * radp1p + x1 * radp2m) * d;
                    sourcep = x1 * cp * exlp + x2 * cm * exlm + radp2p;
                    sourcem = x2 * cp + x1 * cm + radp1m;
                }
                if (a.srctype != 'T') tau = tau + deltau;
            }
            diag[i] = -reflect; diag[i + 1] = -reflect;
            lower[i] = 1.0; lower[i + 1] = -trans;
            upper[i] = -trans; upper[i + 1] = 1.0;
            rhs[i] = sourcem; rhs[i + 1] = sourcep;
            i += 2;
        }
        const float gndemis = 1.0f - a.gndalbedo;
        double gndflux, skyflux;
        if (a.srctype == 'S') {
            gndflux = (1.0f - gndemis) * a.solarflux * exp(-tau / mu0);
            skyflux = pi * a.skyrad;
        } else {
            gndflux = pi * dev_planck(a.gndtemp, a.units, a.wavelen) * gndemis;
            skyflux = pi * dev_planck(a.skyrad, a.units, a.wavelen);
        }
        gndflux = gndflux + a.surface_flux;
        rhs[1] = skyflux; diag[1] = 0.0; upper[1] = 1.0;
        diag[n] = -(1.0f - gndemis); lower[n] = 1.0; rhs[n] = gndflux;
        // ---- TRIDIAG (row interchanges for the largest pivot) ----
        lower[1] = diag[1]; diag[1] = upper[1]; upper[1] = 0.0; upper[n] = 0.0;
        for (int k = 1; k <= n - 1 && !bad; k++) {
            if (fabs(lower[k + 1]) >= fabs(lower[k])) {
                double t = lower[k + 1]; lower[k + 1] = lower[k]; lower[k] = t;
                t = diag[k + 1]; diag[k + 1] = diag[k]; diag[k] = t;
                t = upper[k + 1]; upper[k + 1] = upper[k]; upper[k] = t;
                t = rhs[k + 1]; rhs[k + 1] = rhs[k]; rhs[k] = t;
            }
            if (lower[k] == 0.0) { bad = true; break; }
            const double t = -lower[k + 1] / lower[k];
            lower[k + 1] = diag[k + 1] + t * diag[k];
            diag[k + 1] = upper[k + 1] + t * upper[k];
            upper[k + 1] = 0.0;
            rhs[k + 1] = rhs[k + 1] + t * rhs[k];
        }
        if (bad || lower[n] == 0.0) { atomicExch(a.bad, 1); continue; }
        rhs[n] = rhs[n] / lower[n];
        rhs[n - 1] = (rhs[n - 1] - diag[n - 1] * rhs[n]) / lower[n - 1];
        for (int k = n - 2; k >= 1; k--) rhs[k] = (rhs[k] - diag[k] * rhs[k + 1] - upper[k] * rhs[k + 2]) / lower[k];
        const double c = a.units == 'T' ? 1.0 / pi : 1.0;
        // fluxes of level L (1 = top) are rhs[2L-1] (up) and rhs[2L] (down); grid level iz is L = nz+1-iz
        for (int iz = 1; iz <= nz; iz++) {
            const int L = nz + 1 - iz;
            const float fup = (float)(c * rhs[2 * L - 1]), fdn = (float)(c * rhs[2 * L]);
            float *r = a.radiance + (size_t)a.nstokes * 4 * (p0 + iz - 1);
            for (int q = 0; q < 4 * a.nstokes; q++) r[q] = 0.0f;
            r[0] = c0 * (fup + fdn);
            r[(size_t)a.nstokes * 2] = c1 * (fup - fdn);
        }
    }
}

int adapt_init_radiance(const at3d_state_desc *d, int ld, int ncol, const float *zgrid_d, const float *extinct_d, const float *albedo_d,
                        const float *total_ext_d, const float *temp_d, const float *legen_d, const int *iphase_d,
                        const float *pwt_d, float skyradalb, float surface_flux, float *radiance_d, char *errmsg)
{
    EddArgs a;
    memset(&a, 0, sizeof(a));
    a.ncol = ncol; a.nz = d->nz; a.ld = ld; a.npart = d->npart; a.nq = 8 * d->maxnmicro; a.nstleg = d->nstleg; a.nleg = d->nleg;
    a.ml = d->ml; a.nstokes = d->nstokes; a.interp_new = d->interp_new; a.deltam = d->deltam; a.srctype = d->srctype; a.units = d->units;
    a.phasemax = d->phasemax; a.solarflux = d->solarflux; a.solarmu = d->solarmu; a.gndalbedo = d->gndalbedo; a.gndtemp = d->gndtemp;
    a.skyrad = skyradalb; a.wavelen = d->wavelen; a.surface_flux = surface_flux;
    a.zgrid = zgrid_d; a.extinct = extinct_d; a.albedo = albedo_d; a.total_ext = total_ext_d; a.temp = temp_d; a.legen = legen_d;
    a.phaseinterpwt = pwt_d; a.iphase = iphase_d; a.radiance = radiance_d;
    const int threads = 64, blocks = std::max(1, std::min((ncol + threads - 1) / threads, 256));
    const size_t per = (size_t)4 * (2 * (size_t)(d->nz - 1) + 3);
    double *scratch = nullptr;
    int *bad = nullptr;
    if (at3d_malloc(&scratch, (size_t)blocks * threads * per * sizeof(double)) != cudaSuccess || at3d_malloc(&bad, sizeof(int)) != cudaSuccess) {
        if (scratch) at3d_free(scratch);
        if (errmsg) snprintf(errmsg, 600, "INIT_RADIANCE: device allocation failure"); return 4;
    }
    cudaMemset(bad, 0, sizeof(int));
    a.scratch = scratch; a.bad = bad;
    eddington_kernel<<<blocks, threads>>>(a);
    int hbad = 0;
    cudaError_t e = cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost);
    at3d_free(scratch); at3d_free(bad);
    if (e != cudaSuccess) { if (errmsg) snprintf(errmsg, 600, "CUDA error %s in INIT_RADIANCE", cudaGetErrorString(e)); return 4; }
    if (hbad) { if (errmsg) snprintf(errmsg, 600, "EDDRTF: singular matrix in TRIDIAG or TAU<0"); return 1; }
    return 0;
}
