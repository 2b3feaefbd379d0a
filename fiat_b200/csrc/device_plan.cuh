// Device-side view of the compiled element tables (see include/fiat_b200.h and fiat_b200/plan.py).
#pragma once
#include <stdint.h>
#include "../../include/fiat_b200.h"

#define FB_STEP_DOUBLES 16
#define FB_GEOM_DOUBLES 16
#define FB_NA_MAX 35          // C(3+4, 4): derivative order <= 4 in 3-D on the generic path
#define FB_MAX_LEAVES 4

struct DevSimplex {
    int sd, degree, order, na, expansion, ncells, nslots, nrows, unique;
    int nsteps, nchains, nfix, line_n;
    int chain_ptr[4];
    const int4* step_idx;
    const double* step_dat;
    const int2* chains;
    const int2* fix_idx;
    const double* fix_w;
    const double* geom;
    const double* bary;
    const double* ccell;
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
};

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
};

struct DevTensor {
    int nleaf, nalpha, nrows, order;
    int scratch_doubles;        // per point
    int total_doubles;          // per point: scratch + all leaf tables
    DevTensorLeaf leaf[FB_MAX_LEAVES];
    const int* alpha_leaf;      // nalpha x FB_MAX_LEAVES: alpha index of every leaf for each product alpha
};

__host__ __device__ constexpr int fb_binom(int n, int k) {
    int r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}
