// at3d_adapt.h -- host view of the adaptive cell tree (at3d_adapt.cu), used by the solution iterations (at3d_solver.cu).
#pragma once
#include <vector>
#include "at3d_host.h"

// device arrays the splitting criterion reads (capacity-sized allocations of the solve: the pointers never change)
struct AdaptDev {
    const int *gridptr;      // [8, maxic]
    const float *gridpos;    // [3, maxig]
    const float *total_ext;  // [maxig]
    const int *shptr;        // current SHPTR [maxig+1]
    const float *source;     // current SOURCE [nst, maxiv]
    int nst;
};

struct AdaptGrid {
    // capacities (MAXIG, MAXIC, MAXIV, MAXIDO of RTE._setup_memory) and current sizes
    int maxig = 0, maxic = 0, maxiv = 0, maxido = 0, npts = 0, ncells = 0;
    // the caller's arrays in the reference layout (1-based contents)
    int *gridptr = nullptr, *neighptr = nullptr, *treeptr = nullptr;
    short *cellflags = nullptr;
    float *gridpos = nullptr;
    // host mirrors of SHPTR / RSHPTR (refreshed by the caller before split_grid, extended for new points here)
    std::vector<int> shptr, rshptr;
    std::vector<float> cellcrit;             // criterion of every cell evaluated so far (index cell-1)
    DevBuf work_cells, work_terms;
    std::vector<float2> terms_h;

    int &gp(int k, int ic) { return gridptr[(k - 1) + 8 * (size_t)(ic - 1)]; }
    int gp(int k, int ic) const { return gridptr[(k - 1) + 8 * (size_t)(ic - 1)]; }
    int &nb(int f, int ic) { return neighptr[(f - 1) + 6 * (size_t)(ic - 1)]; }
    int nb(int f, int ic) const { return neighptr[(f - 1) + 6 * (size_t)(ic - 1)]; }
    int &tree(int k, int ic) { return treeptr[(k - 1) + 2 * (size_t)(ic - 1)]; }
    int tree(int k, int ic) const { return treeptr[(k - 1) + 2 * (size_t)(ic - 1)]; }
    float &pos(int d, int ip) { return gridpos[(d - 1) + 3 * (size_t)(ip - 1)]; }
    float pos(int d, int ip) const { return gridpos[(d - 1) + 3 * (size_t)(ip - 1)]; }

    int split_dir(int ic) const;
    int next_cell(double xe, double ye, double ze, int iface, int jface, int icell) const;
    int match_grid_point(float xp, float yp, float zp, int icell, int iface) const;
    void inherit_neighbor(int icell, int iface, int in);
    void match_neighbor_face(int iface, int ic);
    bool divide_cell(int icell, int idir, int newpts[4][3]);
    int grid_smooth_test(int icell) const;
    int evaluate(const std::vector<int> &cells, const AdaptDev &D, std::vector<float> &crit, std::vector<int> &dir, char *errmsg);
    int split_grid(const AdaptDev &D, bool dosplit, bool &outofmem, float cursplitacc, float &splitcrit, int nphi0max, int nlm,
                   int (*interpolate_cb)(void *, const std::vector<NewPointRec> &, char *), void *cb_arg, char *errmsg);
    int boundary_points(int nang, bool lambertian, int maxnbc, int maxbcrad, float zbot, float ztop, int *bcptr, int *ntop, int *nbot) const;
    ~AdaptGrid() { work_cells.release(); work_terms.release(); }
};

int adapt_init_radiance(const at3d_state_desc *d, int ld, int ncol, const float *zgrid_d, const float *extinct_d, const float *albedo_d,
                        const float *total_ext_d, const float *temp_d, const float *legen_d, const int *iphase_d,
                        const float *pwt_d, float skyradalb, float surface_flux, float *radiance_d, char *errmsg);
