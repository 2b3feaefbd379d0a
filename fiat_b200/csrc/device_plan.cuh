// Device-side view of the compiled element tables (see include/fiat_b200.h and fiat_b200/plan.py).
#pragma once
#include <stdint.h>
#include "../../include/fiat_b200.h"

#define FB_GEOM_DOUBLES 32
#define FB_NA_MAX 35          // C(3+4, 4): derivative order <= 4 in 3-D on the generic path
#define FB_MAX_LEAVES 4

// Recurrence program of one expansion set, small enough to travel as a kernel parameter
// (constant bank): the tile kernel dedicates all of shared memory and therefore all of L1 to the
// expansion table, so per-step tables read from global memory would pay an L2 round trip each.
#define FB_MAX_STEPS 455      // C(12+3,3) - 1: degree 12 on a tetrahedron
#define FB_MAX_LEVELS 30
#define FB_MAX_FIX 128
#define FB_MAX_RB 256        // row blocks of 8 (rows <= 2048; stacked per-alpha elements: nalpha * ndofs * components)

struct StepRec {
    short nxt, cur, prv, codim;     // member slots; prv < 0: first step of a chain
    double a, b, c;                 // Jacobi recurrence coefficients
};

struct RecTab {
    int nsteps, nlevels, nfix, nfixgrp;
    int start_slot;                 // slot of member 0
    short level_ptr[FB_MAX_LEVELS + 2];
    short fix_tgt[FB_MAX_FIX], fix_first[FB_MAX_FIX], fix_cnt[FB_MAX_FIX];   // per distinct target
    short fix_src[FB_MAX_FIX];
    double fix_w[FB_MAX_FIX];
    double geom0[FB_GEOM_DOUBLES];  // geometry of cell 0 (single-cell elements)
    int nrb;                        // block-sparse coefficient matrix: row blocks in hand-out order (plan.py: schedule_row_blocks)
    short rb_order[FB_MAX_RB];
    int blk_ptr[FB_MAX_RB + 1];
    short row_perm[FB_MAX_RB * 8];  // packed row -> table row (-1: padding)
    StepRec steps[FB_MAX_STEPS];    // sorted by total degree of the member produced
};

struct DevSimplex {
    int sd, degree, order, na, expansion, ncells, nslots, nrows, unique;
    int line_n, ncomp;
    const RecTab* tab;          // device copy of the recurrence program (tensor-product leaves)
    const double* geom;
    const double* bary;
    const double* ccell;
    const double* ccell_morton;
    const int* low1;
    const double* mul1;
    const int* low2;
    const double* mul2;
    const double* line_tab;
    int nrb, kpad, nblk;
    const int* blk_ptr;
    const int* blk_kb;
    const double* blk_frag;
    const int* rb_order;
    const double* cderiv;       // derivative-folded coefficients of the value-table kernel (ncp == 0: absent)
    int cderiv_len, ncp;
    int blk_cells;              // > 1: blk_ptr is blk_cells x (nrb + 1), one block-sparse matrix per subcell
    const double* cstream;      // fixed-k block stream of the register-operand split-cell kernel (cnsteps == 0: absent)
    const int* cstep_ptr;       // cnsteps + 1 offsets in doubles
    int cnsteps, crb, cmaxstep; // steps, row blocks per step, longest step in doubles
};

// Placement of a kernel's rows inside a larger table (wrapper elements: enriched, mixed, H(div)/H(curl)
// on tensor products -- FIAT/enriched.py:88-113, FIAT/mixed.py:61-92, FIAT/hdivcurl.py:43-108,165-254).
// Kernel row r = dof * nc_in + k  ->  output row (dof_base + dof) * nc_out + comp_out[k], value * sign[k].
struct DevRowMap {
    int identity;       // rows are written in place (no wrapper): skips the index arithmetic
    int nc_in, nc_out, dof_base, total_rows;
    int comp_out[9];
    double sign[9];
};

__device__ __forceinline__ size_t fb_map_row(const DevRowMap& M, int r, double& sgn) {
    if (M.identity) {
        sgn = 1.0;
        return (size_t)r;
    }
    int dof = r, k = 0;
    if (M.nc_in > 1) {
        dof = r / M.nc_in;
        k = r - dof * M.nc_in;
    }
    sgn = M.sign[k];
    return (size_t)(M.dof_base + dof) * M.nc_out + M.comp_out[k];
}

// A point that lies in no subcell of a split complex (NaN / Inf coordinates make every l1 distance NaN): the
// reference leaves the point's column of its zero-initialised tables untouched (FIAT/expansions.py:479-489).
__device__ __forceinline__ void fb_zero_column(const DevRowMap& M, double* __restrict__ out, long long ostride,
                                               long long p, int na, int nrows) {
    for (int a = 0; a < na; ++a)
        for (int r = 0; r < nrows; ++r) {
            double sgn;
            out[((size_t)a * M.total_rows + fb_map_row(M, r, sgn)) * ostride + p] = 0.0;
        }
}

struct DevEntity {
    int dim, identity;
    double C[9];
    double off[3];
};

struct DevTensorLeaf {
    DevSimplex prog;
    DevEntity ent;
    int point_offset;
    int table_off;      // offset (in doubles per point) of this leaf's table in shared memory
    int ndof, ncomp;    // leaf rows = ndof * ncomp
    int dof_stride;     // stride of this leaf's dof index in the product dof index
};

struct DevTensor {
    int nleaf, nalpha, nrows, order;
    int ncomp;                  // components of the product (that of its one vector-valued leaf, else 1)
    int scratch_doubles;        // per point
    int total_doubles;          // per point: scratch + all leaf tables
    DevTensorLeaf leaf[FB_MAX_LEAVES];
    const int* alpha_leaf;      // nalpha x FB_MAX_LEAVES: alpha index of every leaf for each product alpha
};

// register / value-table kernels: recurrence coefficients and barycentric rows as a kernel parameter
#define FB_SMALL_MAX_STEPS 35

struct SmallTab {
    double abc[FB_SMALL_MAX_STEPS][3];     // generation order: pass, sub-index (last entry outermost), i
    double bary[33 * 16];                  // rescaled barycentric rows of <= 32 subcells + parent (constant bank)
};

// product-form (lattice) plan
struct DevLattice {
    int sd, degree, order, na, ndofs;
    const int* rowmap;       // loop index (a0, a1[, a2]) -> dof row
    const double* recip;     // 1 / (k + 1), k = 0..degree-1
};

__host__ __device__ constexpr int fb_binom(int n, int k) {
    int r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}
