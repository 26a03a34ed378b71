#pragma once
// Bicubic patches of the local thin-plate interpolants (fp32 production build only).
//
// Inside one cell of a lookup grid the reference's interpolant is ONE smooth function - the 50-term
// thin-plate sum of the cell's neighbour set (aerodynamic_coefficients.py:57-66 through scipy's
// RBFInterpolator(neighbors=50)); a cell cut by a single order-50 Voronoi edge has two of them.
// For every sub-cell (sub_x x sub_y per lookup cell) the builder evaluates that exact sum at the
// 4 x 4 Chebyshev nodes, stores the interpolating bicubic in monomial form (16 floats, 64 B: the
// constant term as a float pair, the rest is the variation over the cell) and checks the stored
// form, evaluated as the kernel evaluates it, against the exact sum at 5 x 5 other points: a patch whose error exceeds `tol` is
// poisoned (first coefficient NaN) and the kernel evaluates the exact sum for queries that land on
// it, as it does for the cells that need the walk.  The fp64 parity build never reads patches.
#include <cuda_runtime.h>
#include <stdint.h>
namespace pd {
struct PatchGridIn {
    const double *rows;                    // device [n_sets][64]
    const double2 *points;                 // device [n_points]
    const int *cells;                      // device [nm*na]
    const unsigned long long *imp_hint;    // device
    const int *imp_id;                     // device
    const int32_t *cells_host;             // host copies for the index prefix
    const uint64_t *imp_hint_host;
    double m0, dm, a0, da;
    int nm, na;
    int sub_x, sub_y;
};
struct PatchGridOut {
    int2 *pbase = nullptr;                 // device [nm*na]: x = (first patch << 1) | two_sets or -1, y = bisector slots p << 8 | q
    float *patch = nullptr;                // device [n_patches][16] (64 B): C[q][p] of u^p v^q, constant hi / lo in [0] / [15]
    long long n_patches = 0;
    long long n_failed = 0;                // patches over the tolerance
    double max_err_kept = 0.0;             // largest validation error among the patches kept
};
int build_patch_grid(const PatchGridIn &in, double tol, PatchGridOut *out, cudaStream_t st);
void free_patch_grid(PatchGridOut *p);
}  // namespace pd
