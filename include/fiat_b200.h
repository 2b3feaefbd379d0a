/* fiat_b200 -- C ABI of the B200-native basis-tabulation path.
 *
 * This is the drop-in boundary for the reference's `FiniteElement.tabulate(order, points, entity)`
 * (FIAT/finite_element.py:98-109,181-197).  The reference has no FFI of its own (it is pure
 * Python + numpy); a maintainer binds these entry points with ctypes, as INTEGRATION.md shows and
 * as fiat_b200/_lib.py does.  Plain pointers and sizes only; every function returns 0 on success
 * and a non-zero code otherwise (fiatb200_last_error() gives the text); nothing throws across the
 * boundary.  All device pointers refer to the CUDA device that was current when the plan was
 * created; `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 * asynchronous with respect to the host unless stated otherwise.
 *
 * Output layout (all kernels): out[(alpha_index * nrows + row) * out_row_stride + point], float64,
 * alpha_index running over mis(sd,0), mis(sd,1), ..., mis(sd,order) (FIAT/polynomial_set.py:23-32,
 * FIAT/expansions.py:427-432) and row = dof * prod(value_shape) + component, i.e. exactly the
 * C-ordered `(ndofs, *value_shape, npoints)` arrays of the reference's result dict, one after the
 * other, when out_row_stride == npoints.
 */
#ifndef FIAT_B200_H
#define FIAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIATB200_OK 0
#define FIATB200_ERR_CUDA 1
#define FIATB200_ERR_ARG 2
#define FIATB200_ERR_UNSUPPORTED 3

typedef struct fiatb200_plan fiatb200_plan;

/* Point-independent tables of one Ciarlet element on a (possibly split) simplex and one derivative
 * order; built on the host by fiat_b200/plan.py from the reference-constructed element.  Replaces the
 * per-call Python set-up of ExpansionSet._tabulate_on_cell / dubiner_recurrence
 * (FIAT/expansions.py:140-267,411-447), C0_basis (:270-322), the macro scatter (:449-490) and the
 * coefficient tensor of PolynomialSet.tabulate (FIAT/polynomial_set.py:68-72).  All pointers are HOST
 * pointers; the tables are copied to the device by fiatb200_simplex_plan_create. */
typedef struct {
    int32_t sd;            /* spatial dimension 1..3 */
    int32_t degree;        /* embedded degree n */
    int32_t order;         /* maximum derivative order */
    int32_t na;            /* number of derivative multi-indices C(sd+order, order) */
    int32_t expansion;     /* 0 Dubiner recurrence, 1 Legendre line (jacobi.py:47-74), 2 Lagrange line
                              (barycentric_interpolation.py:22-47) */
    int32_t ncells;        /* subcells of the split complex (1 = plain simplex), <= 32 */
    int32_t nslots;        /* expansion members per subcell */
    int32_t nrows;         /* ndofs * prod(value_shape) */
    int32_t ncomp;         /* prod(value_shape) */
    int32_t unique;        /* 1: first matching subcell wins (expansions.py:452,805-807) */
    int32_t nsteps, nlevels, nfix, nfixgrp, line_n;
    int32_t start_slot;    /* slot of expansion member 0, where the recurrence starts */
    const int32_t* step_idx;   /* nsteps x 4: next, cur, prev (-1 first of chain), codim; sorted by the total
                                  degree of the member produced (wavefront order) */
    const double* step_abc;    /* nsteps x 3 Jacobi recurrence coefficients (expansions.py:24-40) */
    const double* nat_abc;     /* nsteps x 3, the same in generation order (pass, sub-index, i) */
    const int32_t* level_ptr;  /* nlevels + 1: steps producing degree d+1 are [level_ptr[d], level_ptr[d+1]) */
    const int32_t* fix_idx;    /* nfix x 2: target, source slot (C0_basis fix-ups) */
    const double* fix_w;       /* nfix */
    const int32_t* fix_grp;    /* nfixgrp x 2: fix-ups sorted by target; first entry and count per target */
    const double* geom;        /* ncells x 32: A[9] (row-major sd x sd), b[3] at 9, start value at 12,
                                  grad fa per pass at 14 (3x3), grad fb per pass at 23 (3x3) */
    const double* bary;        /* (ncells+1) x 4 x 4: rescaled barycentric rows A_hat | b_hat, parent last
                                  (reference_element.py:616-644) */
    const double* ccell;       /* ncells x nrows x nslots folded coefficients coeffs[:, cell_node_map[c]] */
    const double* ccell_morton;/* same on Morton-numbered members with the C0 fix-ups folded in */
    const int32_t* low1;       /* na x 3, Leibniz index tables (expansions.py:66-137) */
    const double* mul1;        /* na x 3 */
    const int32_t* low2;       /* na x 6 */
    const double* mul2;        /* na x 6 */
    const double* line_tab;    /* expansion 1/2 tables */
    int64_t line_tab_len;
    /* 8x4 block-sparse "gather" packing of the (fix-up-folded) coefficient matrix in mma.m8n8k4 fragment order
     * (ncells == 1, or one stream per subcell when blk_cells > 1).  Row block rb = packed rows 8 rb .. 8 rb + 7;
     * its blocks are blk_ptr[rb] .. blk_ptr[rb + 1] - 1.  Block q multiplies the four member slots
     * blk_kb[4 q + t], t = 0..3 (ANY four slots: the kernel gathers the matching rows of the expansion table from
     * shared memory; slots that differ mod 4 are bank-conflict free), and blk_frag[32 q + 4 g + t] is the
     * coefficient of packed row 8 rb + g on member slot blk_kb[4 q + t]. */
    int32_t nrb, kpad, nblk;
    const int32_t* blk_ptr;    /* nrb + 1 */
    const int32_t* blk_kb;     /* 4 x nblk member slots */
    const double* blk_frag;    /* nblk x 32 */
    const int32_t* rb_order;   /* nrb: order in which row blocks are handed out to the warps (long and short
                                  blocks alternate, fiat_b200/plan.py: schedule_row_blocks) */
    const int32_t* row_perm;   /* nrows: packed row i holds table row row_perm[i] (rows are clustered by member
                                  support so that fewer blocks are stored) */
    /* Derivative-folded coefficients for the value-table kernel (optional, ncp == 0: absent).  D^alpha of an
     * expansion member of degree k lies in the span of the members of degree <= k - |alpha| (the fact behind
     * ExpansionSet.get_dmats, FIAT/expansions.py:577-599), so out_alpha = C_alpha[cell] . (member values):
     * cderiv[(off_alpha + row * nm_k + m) * ncp + cell], nm_k = C(degree - |alpha| + sd, sd), alphas in mis
     * order, Morton-numbered un-normalised members, C0 fix-ups / normalisation / chain rule folded in. */
    const double* cderiv;
    int64_t cderiv_len;
    int32_t ncp;               /* subcell stride: ncells padded to 1, 4 or 16 */
    /* Split-cell tile kernel (order-0 derived elements of fiat_b200.plan.macro_merged): blk_cells == ncells > 1
     * means blk_ptr holds one (nrb + 1)-entry row per subcell (offsets into blk_kb / blk_frag; every row block has
     * at least one, possibly zero, block), and all subcells share rb_order, row_perm and the member slots; 0: single
     * matrix as described above. */
    int32_t blk_cells;
    /* Fixed-k block stream for the register-operand split-cell kernel (optional; cstream_len == 0: absent).
     * k-block j = member slots 4 j .. 4 j + 3.  The packed row blocks are cut into steps of crb row blocks; step s is
     * the run of doubles cstream[cstep_ptr[s] .. cstep_ptr[s + 1]):  ncells * crb int32 records
     * (n | first block << 16, subcell-major; n = the PREFIX of k-blocks that (subcell, row block) stores: 0 .. n - 1,
     * everything up to the last k-block one of its rows touches), padded to a multiple of 16 bytes, then the step's
     * blocks (subcell, row block, k-block ascending), 32 doubles each in mma.m8n8k4 A-fragment order
     * (fiat_b200/plan.py: pack_fixed_stream; prefix_members orders the slots so that prefixes are short). */
    const double* cstream;
    int64_t cstream_len;
    const int32_t* cstep_ptr;  /* cnsteps + 1 offsets in doubles (even) */
    int32_t cnsteps, crb;
} fiatb200_simplex_program;

/* Entity transform x_cell = x_entity * C + offset (FIAT/reference_element.py:570-609);
 * identity != 0 skips it (default cell entity, :585-587). */
typedef struct {
    int32_t dim;           /* number of coordinates per input point */
    int32_t identity;
    double C[9];           /* dim x sd, row-major */
    double offset[3];
} fiatb200_entity_map;

/* One leaf factor of a (flattened) tensor-product element: a simplex plan evaluated on a slice of the
 * product point (FIAT/tensor_product.py:238-258). */
typedef struct {
    const fiatb200_plan* plan;     /* simplex plan of the factor, same order as the product */
    fiatb200_entity_map entity;    /* factor entity transform */
    int32_t point_offset;          /* first coordinate of the product point used by this factor */
} fiatb200_tensor_leaf;

/* Placement of a plan's result rows inside a larger table, for the wrapper elements that zero-pad,
 * stack, permute or sign-flip child tabulations (EnrichedElement FIAT/enriched.py:88-113, MixedElement
 * FIAT/mixed.py:61-92, Hdiv/Hcurl on tensor products FIAT/hdivcurl.py:43-108,165-254):
 * plan row dof * nc_in + k  ->  output row (dof_base + dof) * nc_out + comp_out[k], value * sign[k];
 * the output table has total_rows rows per derivative multi-index.  nc_in must equal the plan's
 * number of components. */
typedef struct {
    int32_t nc_in, nc_out, dof_base, total_rows;
    int32_t comp_out[9];
    double sign[9];
} fiatb200_row_map;

int fiatb200_version(void);
const char* fiatb200_last_error(void);

/* Upload the tables to the current device.  Replaces nothing in the reference (it re-derives them on
 * every call); corresponds to the per-element caches of FIAT/expansions.py:377-378. */
int fiatb200_simplex_plan_create(const fiatb200_simplex_program* prog, fiatb200_plan** plan);

/* Scalar x scalar tensor-product plan over nleaf <= 4 factor plans
 * (TensorProductElement.tabulate / FlattenedDimensions.tabulate, FIAT/tensor_product.py:231-292,396-407).
 * The factor plans must outlive the tensor plan. */
int fiatb200_tensor_plan_create(const fiatb200_tensor_leaf* leaves, int32_t nleaf, int32_t order,
                                fiatb200_plan** plan);

/* Product-form plan for the nodal basis of the principal lattice on the UFC simplex (equispaced
 * Lagrange elements, FIAT/lagrange.py:75-88): same results as the simplex plan of that element at a
 * fraction of the arithmetic.  rowmap[loop index of (a0, a1[, a2])] = dof index (fiat_b200/plan.py:
 * lattice_rowmap).  sd in {2, 3}, order <= 2. */
int fiatb200_lattice_plan_create(int32_t sd, int32_t degree, int32_t order, const int32_t* rowmap,
                                 int32_t ndofs, fiatb200_plan** plan);

int fiatb200_plan_destroy(fiatb200_plan* plan);

/* Which kernel fiatb200_tabulate would run for this plan and these flags: 1 thread-per-point, 2 DMMA tile,
 * 3 register (jet) kernel, 4 value-table kernel, 5 product-form (lattice), 6 tensor product, 7 split-cell DMMA
 * tile; 0 if the flags force a kernel that does not apply. */
int fiatb200_plan_kernel(const fiatb200_plan* plan, uint32_t flags);

/* Number of result rows per derivative multi-index, and number of multi-indices. */
int fiatb200_plan_shape(const fiatb200_plan* plan, int64_t* nrows, int64_t* nalpha);

/* Tabulate at npts device-resident points (row-major, leading dimension pts_ld doubles).
 * = CiarletElement.tabulate (FIAT/finite_element.py:181-197) / TensorProductElement.tabulate.
 * `entity` may be NULL for tensor plans (their leaves carry the factor entities).
 * flags: bit 0 forces the thread-per-point kernel, bit 1 forces the block-sparse DMMA kernel, bit 3
 *        disables the value-table kernel (derivative-folded coefficients)
 *        (testing / profiling); 0 lets the library choose. */
int fiatb200_tabulate(const fiatb200_plan* plan, const fiatb200_entity_map* entity,
                      const double* pts_dev, int64_t npts, int64_t pts_ld,
                      double* out_dev, int64_t out_row_stride, uint32_t flags, void* stream);

/* Same as fiatb200_tabulate, writing through a row placement (map == NULL: identity). */
int fiatb200_tabulate_mapped(const fiatb200_plan* plan, const fiatb200_entity_map* entity,
                             const double* pts_dev, int64_t npts, int64_t pts_ld,
                             double* out_dev, int64_t out_row_stride, const fiatb200_row_map* map,
                             uint32_t flags, void* stream);

/* Fused consumer of a scalar tensor-product tabulation (point evaluation / interpolation):
 * out[(alpha_index * nfunc + f) * out_row_stride + point] = sum_dof coef[f * ndofs + dof] * D^alpha phi_dof(point),
 * i.e. coef . TensorProductElement.tabulate(order, points) (FIAT/tensor_product.py:231-292) without forming the
 * (ndofs x npts) tables; the sum over the product dofs is nested over the factors.  coef_dev: nfunc x ndofs,
 * row-major, on the device. */
int fiatb200_evaluate_tensor(const fiatb200_plan* plan, const double* coef_dev, int32_t nfunc, const double* pts_dev,
                             int64_t npts, int64_t pts_ld, double* out_dev, int64_t out_row_stride, void* stream);

/* The same fused consumer for Ciarlet elements on (possibly split) simplices.  `plan` is the order-0 simplex plan of
 * the element's *stacked derived element* (fiat_b200/plan.py: stacked_derived), whose rows are
 * (derivative table j of nstack, dof i of ndofs, component c): D^alpha_j phi_i[c].  Then
 *   out[((j * nfunc + f) * ncomp + c) * out_row_stride + point] = sum_i coef[f * ndofs + i] * D^alpha_j phi_i[c](point)
 * = coef . CiarletElement.tabulate(order, points)[alpha_j] (FIAT/finite_element.py:181-197, polynomial_set.py:68-72)
 * without writing the (ndofs x npts) tables.  The weights coef . C are formed on the device in the same call, so a
 * new coefficient vector costs no re-planning.  coef_dev: nfunc x ndofs, row-major, on the device. */
int fiatb200_evaluate_simplex(const fiatb200_plan* plan, int32_t nstack, int32_t ndofs, const double* coef_dev,
                              int32_t nfunc, const fiatb200_entity_map* entity, const double* pts_dev, int64_t npts,
                              int64_t pts_ld, double* out_dev, int64_t out_row_stride, void* stream);

/* Zero-fill the listed rows (device array of nrows row numbers) of every derivative table: the
 * entries of a wrapper element's table that none of its parts writes. */
int fiatb200_zero_rows(double* out_dev, int64_t out_row_stride, int64_t npts, int64_t total_rows, int32_t nalpha,
                       const int32_t* rows_dev, int32_t nrows, void* stream);

/* Split-cell point location only: bitmask of the subcells each point is binned to
 * (= compute_cell_point_map, FIAT/expansions.py:771-811).  The object the parity tests compare
 * bit-for-bit with the reference. */
int fiatb200_locate_subcells(const fiatb200_plan* plan, const fiatb200_entity_map* entity,
                             const double* pts_dev, int64_t npts, int64_t pts_ld, int32_t unique,
                             uint32_t* mask_out_dev, void* stream);

/* End-to-end call with HOST buffers (what a numpy caller of the reference holds): points are staged
 * to the device, tabulated in chunks of at most chunk_pts points and copied back, with copies and
 * kernels overlapped on two internal streams.  out_host has the layout above with
 * out_row_stride == npts.  Synchronous. */
int fiatb200_tabulate_host(const fiatb200_plan* plan, const fiatb200_entity_map* entity,
                           const double* pts_host, int64_t npts, int64_t pts_ld,
                           double* out_host, int64_t chunk_pts, uint32_t flags);

/* One kernel launch of a tabulation that is made of several plans: the parts of a wrapper element
 * (FIAT/enriched.py:88-113, FIAT/mixed.py:61-92, FIAT/hdivcurl.py) and/or the per-derivative derived elements of
 * a split single-cell element.  plan == NULL zero-fills `zero_rows_dev` of derivative table `alpha_offset`. */
typedef struct {
    const fiatb200_plan* plan;
    const fiatb200_entity_map* entity;   /* NULL for tensor plans / the default cell entity */
    const fiatb200_row_map* map;         /* NULL: rows in place */
    int32_t alpha_offset;                /* derivative table that the plan's first table is written to */
    const int32_t* zero_rows_dev;        /* plan == NULL only: device array of row numbers */
    int32_t nzero_rows;
} fiatb200_launch;

/* fiatb200_tabulate_host for a list of launches that together fill a (nalpha x total_rows x npts) result:
 * every chunk of points is staged once, all launches of the list run on it, and the chunk's rows are copied
 * to their place in out_host.  zero_rows_dev / nzero_rows (may be NULL / 0): rows of every derivative table
 * that no launch writes.  Synchronous. */
int fiatb200_tabulate_host_list(const fiatb200_launch* launches, int32_t nlaunch, int32_t nalpha, int64_t total_rows,
                                const int32_t* zero_rows_dev, int32_t nzero_rows, const double* pts_host,
                                int64_t npts, int64_t pts_ld, double* out_host, int64_t chunk_pts, uint32_t flags);

/* Plan-compilation helpers (host only; no device is touched).  They optimise the two free choices of the gather
 * packing above for a boolean support matrix `support` (nrows x ncols, row-major, ncols = nseg segments of equal
 * width: one segment per subcell, the same members in each):
 *   fiatb200_cluster_rows    order[] (a permutation of the rows; with greedy_init != 0 it is built here, otherwise it
 *                            is the starting point) such that every run of 8 rows has a small union of members, by
 *                            `iters` rounds of best-swap local search between two random groups; *blocks_out =
 *                            sum over groups and segments of ceil(union / 4) = number of 8x4 blocks.
 *   fiatb200_colour_members  colour_out[member] in 0..3 = the member's slot number mod 4 (exactly as many members per
 *                            colour as there are slots of that residue), chosen so that the members each group
 *                            uses spread evenly over the colours; *conflicts_out = members that cannot be placed
 *                            in a block of four distinct colours.
 * They replace nothing in the reference (numpy.dot is dense, FIAT/polynomial_set.py:71). */
int fiatb200_cluster_rows(const uint8_t* support, int32_t nrows, int32_t ncols, int32_t nseg, int32_t greedy_init,
                          int32_t* order, int64_t iters, uint64_t seed, int32_t* blocks_out);
int fiatb200_colour_members(const uint8_t* support, int32_t nrows, int32_t ncols, int32_t nseg, const int32_t* order,
                            int64_t iters, uint64_t seed, int32_t* colour_out, int32_t* conflicts_out);

/* fiatb200_evaluate_simplex / fiatb200_evaluate_tensor with HOST buffers: coefficients and points are staged to the
 * device, the points are evaluated in chunks and the (nstack * nfunc * ncomp) x npts result is copied back, copies and
 * kernels overlapped on the plan's two internal streams.  Per point this moves 8 * dim bytes in and
 * 8 * nstack * nfunc * ncomp bytes out instead of the 8 * nstack * ndofs * ncomp bytes of the tables.  For
 * tensor-product plans nstack / ndofs must be the plan's number of derivative tables / rows.  Synchronous. */
int fiatb200_evaluate_host(const fiatb200_plan* plan, int32_t nstack, int32_t ndofs, const double* coef_host,
                           int32_t nfunc, const fiatb200_entity_map* entity, const double* pts_host, int64_t npts,
                           int64_t pts_ld, double* out_host, int64_t chunk_pts);

/* Number of kernel launches issued by this library in the calling process so far. */
int64_t fiatb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FIAT_B200_H */
