"""Compile an element description into the point-independent tables the CUDA kernels consume.

Everything that does not depend on the evaluation point is worked out here, once per element:

* the *recurrence program* of the Dubiner / integrated-Jacobi expansion
  (`FIAT/expansions.py:202-249`): one record per three-term step with the Jacobi coefficients
  (`jrc` / `integrated_jrc`, `:24-40`) and the point-independent parts of the recurrence-factor
  derivatives (`jacobi_factors`, `:54-63,205`), ordered by the total degree of the member they
  produce so that all steps of one degree can run concurrently;
* the per-pass normalisation (`:251-266`), accumulated per member and folded into the columns of
  the coefficient matrix together with the C0 entity reordering (`:297-322`), so that the device
  recurrence runs un-normalised and `PolynomialSet.tabulate`'s contraction
  (`FIAT/polynomial_set.py:71`) absorbs both for free;
* the C0 fix-ups (`:281-295`) as (target, source, weight) triples;
* per-subcell coefficient matrices `coeffs[:, cell_node_map[c]]` (`:477-490`);
* the 8x4 block-sparse packing of the coefficient matrix in `mma.m8n8k4.f64` fragment order;
* 1-D tables: Legendre/Jacobi recurrence constants and running derivative scales
  (`:659-678`, `FIAT/jacobi.py:47-74`), Lagrange nodes/weights/dmat
  (`FIAT/barycentric_interpolation.py:22-59`);
* for tensor-product elements the flattened list of leaf factors (`FIAT/tensor_product.py:231-292`).
"""
import math
import os
from dataclasses import dataclass, field

import numpy

__all__ = ["alpha_list", "compile_simplex", "SimplexProgram", "flatten_tensor", "TensorLeaf", "lattice_rowmap",
           "alpha_split", "merged_split", "macro_merged", "stacked_derived", "pullback_description",
           "resolve_parts", "Part", "value_shape_of", "num_dofs_of"]

EXPANSION_CODES = {"dubiner": 0, "legendre_line": 1, "lagrange_line": 2}
GEOM_DOUBLES = 32      # A[9], b[3] @9, start value @12, dfa[codim][d] @14, dfb[codim][d] @23


def multi_indices(m, n):
    """m-tuples summing to n, first entry descending (the order of FIAT's `mis`)."""
    if m == 1:
        return [(n,)]
    out = []
    for first in range(n, -1, -1):
        out.extend((first,) + rest for rest in multi_indices(m - 1, n - first))
    return out


def alpha_list(sd, order):
    """Dict-key order of a tabulation: mis(sd,0), mis(sd,1), ..., mis(sd,order)."""
    out = []
    for k in range(order + 1):
        out.extend(multi_indices(sd, k))
    return out


def _morton(index, sd):
    index = tuple(index) + (0,) * (sd - len(index))
    if sd == 1:
        return index[0]
    if sd == 2:
        p, q = index
        return (p + q) * (p + q + 1) // 2 + q
    p, q, r = index
    s = p + q + r
    return s * (s + 1) * (s + 2) // 6 + (q + r) * (q + r + 1) // 2 + r


def _sub_indices(n, length):
    """Index tuples of the given length with non-negative entries and sum < n."""
    if length == 0:
        return [()]
    out = []
    for last in range(n):
        out.extend(head + (last,) for head in _sub_indices(n - last, length - 1))
    return out


def _jacobi_abc(a, b, i):
    den = 2 * (i + 1) * (i + 1 + a + b)
    an = (2 * i + 1 + a + b) * (2 * i + 2 + a + b) / den
    bn = (a + b) * (a - b) * (2 * i + 1 + a + b) / (den * (2 * i + a + b))
    cn = (i + a) * (i + b) * (2 * i + 2 + a + b) / ((i + 1) * (i + 1 + a + b) * (2 * i + a + b))
    return an, bn, cn


def _chain_coefficients(variant, sub, length):
    """(a, b, c) for the `length` steps of the chain that extends sub-index `sub`."""
    ssum = sum(sub)
    beta = 1 if variant == "dual" else 0
    if variant == "bubble":
        alpha = 2 * ssum
        first = (-0.5, -0.5, 0.0)
    else:
        alpha = 2 * ssum + len(sub)
        if variant == "dual":
            alpha += 1 + len(sub)
        first = (0.5 * (alpha + beta) + 1.0, 0.5 * (alpha - beta), 0.0)
    out = [first]
    for i in range(1, length):
        if variant == "bubble":
            if i == 1:
                out.append(((alpha + beta + 2) / 2, (alpha - 3 * beta - 2) / 2, 0.0))
            else:
                out.append(_jacobi_abc(alpha - 1, beta + 1, i - 1))
        else:
            out.append(_jacobi_abc(alpha, beta, i))
    return out


def _normalisation(sd, n, variant):
    """Accumulated normalisation factor of every Morton-numbered member.

    Pass d rescales every member whose index has length d (trailing zeros implied); members
    created later by a chain inherit what their chain start had accumulated by then, because the
    recurrence is linear in its start value.
    """
    total = numpy.ones(math.comb(n + sd, sd))
    shift = 1 if variant == "dual" else 0
    for d in range(1, sd + 1):
        before = total.copy()
        for index in _sub_indices(n + 1, d):
            if variant == "none":
                norm2 = (2 * sum(index) + d) / d
            else:
                p = index[-1] + shift
                al = 2 * (sum(index[:-1]) + d * shift) - 1
                norm2 = (0.5 + d) / d
                if p > 0 and p + al > 0:
                    norm2 *= (p + al) * (2 * p + al) / p
            start = _morton(index[:-1], sd)
            total[_morton(index, sd)] = before[start] * math.sqrt(norm2)
    return total


def _c0_layout(sd, n):
    """(entity_order, fixups): gather list Morton -> entity order, and the in-place corrections
    [(target, source)] applied to the normalised hierarchical functions."""
    ix = lambda *idx: _morton(idx, sd)  # noqa: E731
    rng = range(2, n + 1)
    order = list(range(sd + 1))
    fix = []
    if sd == 1:
        order += list(rng)
    elif sd == 2:
        order += [ix(1, i - 1) for i in rng] + [ix(0, i) for i in rng] + [ix(i, 0) for i in rng]
        order += [ix(i, j) for j in range(1, n + 1) for i in range(2, n - j + 1)]
        fix += [(ix(0, i), ix(1, i - 1)) for i in rng]
    else:
        order += [ix(0, 1, i - 1) for i in rng] + [ix(1, 0, i - 1) for i in rng]
        order += [ix(1, i - 1, 0) for i in rng] + [ix(0, 0, i) for i in rng]
        order += [ix(0, i, 0) for i in rng] + [ix(i, 0, 0) for i in rng]
        inner = [(i, j) for j in range(1, n + 1) for i in range(2, n - j + 1)]
        order += [ix(1, i - 1, j) for i, j in inner] + [ix(0, i, j) for i, j in inner]
        order += [ix(i, 0, j) for i, j in inner] + [ix(i, j, 0) for i, j in inner]
        order += [ix(i, j, k) for k in range(1, n + 1) for j in range(1, n - k + 1)
                  for i in range(2, n - j - k + 1)]
        for i in rng:
            fix += [(ix(0, i, j), ix(1, i - 1, j)) for j in range(0, n + 1 - i)]
            fix += [(ix(0, 0, i), ix(0, 1, i - 1)), (ix(0, 0, i), ix(1, 0, i - 1))]
    return order, fix


def leibniz_tables(sd, order):
    """Index tables for D^alpha(F G) with F at most quadratic (`_product_derivative`, :66-137).

    low1[j, d]  = index of alpha_j - e_d (or -1), mul1[j, d] = alpha_d
    low2[j, k]  = index of alpha_j - e_d1 - e_d2 for the k-th pair d1<=d2 (or -1), mul2[j, k]
    """
    alphas = alpha_list(sd, order)
    pos = {a: j for j, a in enumerate(alphas)}
    pairs = [(d1, d2) for d1 in range(sd) for d2 in range(d1, sd)]
    low1 = -numpy.ones((len(alphas), 3), dtype=numpy.int32)
    mul1 = numpy.zeros((len(alphas), 3))
    low2 = -numpy.ones((len(alphas), 6), dtype=numpy.int32)
    mul2 = numpy.zeros((len(alphas), 6))
    for j, al in enumerate(alphas):
        for d in range(sd):
            if al[d] >= 1:
                lo = al[:d] + (al[d] - 1,) + al[d + 1:]
                low1[j, d], mul1[j, d] = pos[lo], al[d]
        for k, (d1, d2) in enumerate(pairs):
            need = 2 if d1 == d2 else 1
            if al[d1] >= need and al[d2] >= need:
                lo = list(al)
                lo[d1] -= 1
                lo[d2] -= 1
                low2[j, k] = pos[tuple(lo)]
                mul2[j, k] = al[d1] * (al[d1] - 1) // 2 if d1 == d2 else al[d1] * al[d2]
    return low1, mul1, low2, mul2


@dataclass
class SimplexProgram:
    """Host-side tables for one Ciarlet element and one derivative order."""
    sd: int
    degree: int
    order: int
    na: int
    expansion: int
    ncells: int
    nslots: int                 # expansion members per (sub)cell
    nrows: int                  # ndofs * prod(value_shape)
    ndofs: int
    value_shape: tuple
    unique: int                 # first-match binning (continuity is not None and order == 0)
    geom: numpy.ndarray         # (ncells, GEOM_DOUBLES)
    bary: numpy.ndarray         # (ncells + 1, 4, 4): rows of A_hat | b_hat
    step_idx: numpy.ndarray     # (nsteps, 4) int32: next, cur, prev(-1 = first of chain), codim; level order
    step_abc: numpy.ndarray     # (nsteps, 3) Jacobi recurrence coefficients a, b, c
    nat_abc: numpy.ndarray      # (nsteps, 3) the same coefficients in generation order (pass, sub-index, i)
    start_slot: int             # slot of member 0 (the constant function the recurrence starts from)
    ccell_morton: numpy.ndarray  # (ncells, nrows, nslots) coefficients on Morton-numbered members, fix-ups folded
    level_ptr: numpy.ndarray    # (degree + 1,) int32: steps producing degree d+1 are [level_ptr[d], level_ptr[d+1])
    fix_idx: numpy.ndarray      # (nfix, 2) int32 target, source slots
    fix_w: numpy.ndarray        # (nfix,)
    fix_grp: numpy.ndarray      # (ngroups, 2) int32: first fix-up and count per distinct target
    ccell: numpy.ndarray        # (ncells, nrows, nslots) folded coefficients
    low1: numpy.ndarray
    mul1: numpy.ndarray
    low2: numpy.ndarray
    mul2: numpy.ndarray
    line_tab: numpy.ndarray     # expansion-specific 1-D tables (see _line_tables)
    line_n: int = 0
    # block-sparse gather packing of the fix-up-folded coefficient matrix (tile kernels)
    blk_ptr: numpy.ndarray = field(default_factory=lambda: numpy.zeros(1, numpy.int32))
    blk_kb: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0, numpy.int32))      # (nblk, 4) member slots, flat
    blk_frag: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0))
    rb_order: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0, numpy.int32))
    row_perm: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0, numpy.int32))   # packed row -> table row
    kpad: int = 0
    # derivative-folded coefficients for the value-table kernel (see derivative_coefficients)
    cderiv: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0))
    ncp: int = 0                # subcell stride of cderiv (ncells padded to 1, 4 or 16); 0 = absent
    slot_of: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0, numpy.int64))    # Morton member -> slot
    blk_cells: int = 0          # > 1: blk_ptr holds one (nrb + 1)-entry row per subcell (split-cell tile kernel)
    # fixed-k block stream of the register-operand split-cell kernel (see pack_fixed_stream); empty = absent
    cstream: numpy.ndarray = field(default_factory=lambda: numpy.zeros(0))
    cstep_ptr: numpy.ndarray = field(default_factory=lambda: numpy.zeros(1, numpy.int32))
    crb: int = 0                # row blocks per step of the stream


def _dubiner_tables(desc, order, slot_perm=None):
    sd, n, variant = int(desc["sd"]), int(desc["degree"]), desc["variant"]
    ncells = int(desc["ncells"])
    c0 = bool(desc["c0"])
    nmem = math.comb(n + sd, sd)
    if c0:
        entity_order, fix_pairs = _c0_layout(sd, n)
    else:
        entity_order, fix_pairs = list(range(nmem)), []
    slot_of = numpy.empty(nmem, dtype=numpy.int64)       # Morton member -> slot (output position)
    slot_of[entity_order] = numpy.arange(nmem)
    if slot_perm is not None:
        # renumber the slots: new slot j holds what used to be slot slot_perm[j]
        inverse = numpy.empty(nmem, dtype=numpy.int64)
        inverse[numpy.asarray(slot_perm)] = numpy.arange(nmem)
        slot_of = inverse[slot_of]

    # steps, pass by pass (expansions.py:202-249)
    step_idx, step_abc, step_level = [], [], []
    if n > 0:
        for codim in range(sd):
            for sub in _sub_indices(n, codim):
                length = n - sum(sub)
                abc = _chain_coefficients(variant, sub, length)
                for i in range(length):
                    cur = slot_of[_morton(sub + (i,), sd)]
                    nxt = slot_of[_morton(sub + (i + 1,), sd)]
                    prv = slot_of[_morton(sub + (i - 1,), sd)] if i > 0 else -1
                    step_idx.append((nxt, cur, prv, codim))
                    step_abc.append(abc[i])
                    step_level.append(sum(sub) + i + 1)      # total degree of the member produced
    step_idx = numpy.array(step_idx, dtype=numpy.int32).reshape(-1, 4)
    step_abc = numpy.array(step_abc, dtype=float).reshape(-1, 3)
    nat_abc = step_abc.copy()
    # wavefront order: a member of total degree d only needs members of degree d-1 and d-2
    step_level = numpy.array(step_level, dtype=numpy.int64)
    by_level = numpy.argsort(step_level, kind="stable")
    step_idx, step_abc = step_idx[by_level], step_abc[by_level]
    level_ptr = numpy.searchsorted(step_level[by_level], numpy.arange(1, n + 2)).astype(numpy.int32)

    # per-cell geometry: affine map to the default simplex and the constant gradients of the
    # recurrence factors fa, fb of every collapsing pass (jacobi_factors, expansions.py:54-63)
    geom = numpy.zeros((ncells, GEOM_DOUBLES))
    for c in range(ncells):
        A = numpy.asarray(desc["cell_A"][c], dtype=float)
        geom[c, :sd * sd] = A.reshape(-1)
        geom[c, 9:9 + sd] = desc["cell_b"][c]
        scale = float(desc["cell_scale"][c])
        geom[c, 12] = -scale if variant == "bubble" else scale
        dX = [A[i] for i in range(sd)] + [numpy.zeros(sd), numpy.zeros(sd)]
        for codim in range(sd):
            dfb = 0.5 * (dX[codim + 1] + dX[codim + 2])
            dfa = dX[codim] + dfb
            geom[c, 14 + 3 * codim:14 + 3 * codim + sd] = dfa
            geom[c, 23 + 3 * codim:23 + 3 * codim + sd] = dfb

    # fold normalisation / sign / reordering into the coefficient columns ("raw_members": the coefficients
    # already refer to the un-normalised recurrence members -- derived elements of alpha_split)
    norm = _normalisation(sd, n, variant) if (n > 0 and not desc.get("raw_members")) else numpy.ones(nmem)
    fold = norm.copy()
    fix_idx, fix_w = [], []
    if c0:
        fold[0] = -norm[0]
        for m in range(1, sd + 1):
            fix_idx.append((slot_of[0], slot_of[m]))
            fix_w.append(-norm[m] / norm[0])
        for t, s in fix_pairs:
            fix_idx.append((slot_of[t], slot_of[s]))
            fix_w.append(norm[s] / norm[t])
    fold_by_slot = numpy.empty(nmem)
    fold_by_slot[slot_of] = fold
    pos_of_slot = numpy.arange(nmem) if slot_perm is None else numpy.asarray(slot_perm, dtype=numpy.int64)
    return dict(nslots=nmem, step_idx=step_idx, step_abc=step_abc, nat_abc=nat_abc, slot_of=slot_of,
                start_slot=int(slot_of[0]), pos_of_slot=pos_of_slot,
                geom=geom, level_ptr=level_ptr,
                fix_idx=numpy.array(fix_idx, dtype=numpy.int32).reshape(-1, 2),
                fix_w=numpy.array(fix_w, dtype=float), fold_by_slot=fold_by_slot)


def _line_tables(desc, order):
    """1-D sets that do not use the simplex recurrence."""
    n = int(desc["degree"])
    ncells = int(desc["ncells"])
    geom = numpy.zeros((ncells, GEOM_DOUBLES))
    if desc["expansion"] == "legendre_line":
        # per order k: Jacobi(k,k) recurrence constants; per cell running scales (expansions.py:669-676)
        rec = numpy.zeros((order + 1, n + 1, 4))
        for k in range(order + 1):
            a = b = float(k)
            apb = a + b
            if n >= 1:
                rec[k, 1, 0] = 0.5 * (a - b)
                rec[k, 1, 1] = 0.5 * (a + b + 2.0)
            for j in range(2, n + 1):
                a1 = 2.0 * j * (j + apb) * (2.0 * j + apb - 2.0)
                a2 = (2.0 * j + apb - 1.0) * (a * a - b * b)
                a3 = (2.0 * j + apb - 2.0) * (2.0 * j + apb - 1.0) * (2.0 * j + apb)
                a4 = 2.0 * (j + a - 1.0) * (j + b - 1.0) * (2.0 * j + apb)
                rec[k, j, :3] = (a2 / a1, a3 / a1, a4 / a1)
        scales = numpy.zeros((ncells, order + 1, n + 1))
        for c in range(ncells):
            A = numpy.asarray(desc["cell_A"][c], dtype=float)
            geom[c, 0] = A[0, 0]
            geom[c, 9] = desc["cell_b"][c][0]
            run = float(desc["cell_scale"][c]) * numpy.sqrt(2 * numpy.arange(n + 1) + 1)
            for k in range(order + 1):
                scales[c, k] = run
                run = run * (0.5 * (numpy.arange(n + 1) + k + 1) * A[0, 0])
        tab = numpy.concatenate([rec.reshape(-1), scales.reshape(-1)])
        return dict(nslots=n + 1, geom=geom, line_tab=tab, line_n=n)
    nodes, wts, dmat = desc["ll_nodes"], desc["ll_wts"], desc["ll_dmat"]
    nn = nodes.shape[1]
    tab = numpy.concatenate([numpy.concatenate([nodes[c], wts[c], dmat[c].reshape(-1)]) for c in range(ncells)])
    return dict(nslots=nn, geom=geom, line_tab=tab, line_n=nn)


def _support(C, tol, nseg=1):
    return numpy.ascontiguousarray(numpy.abs(C) > tol, dtype=numpy.uint8)


def cluster_rows(C, drop_tol=0.0, nseg=1, iters=None, seed=1):
    """Row order such that each run of 8 consecutive rows touches few expansion members (per column segment):
    the 8x4 blocks of the tile kernels gather any four members, so a row group costs ceil(|union| / 4) blocks.
    Greedy seed + swap local search, both in the library (csrc/cluster.cu).  -> (order, number of blocks)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    S = _support(C, drop_tol)
    nrows, ncols = S.shape
    order = numpy.zeros(nrows, dtype=numpy.int32)
    blocks = ctypes.c_int32()
    if iters is None:
        iters = CLUSTER_ITERS if nrows > 8 else 0
    _lib.check(lib.fiatb200_cluster_rows(S.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), nrows, ncols, nseg, 1,
                                         order.ctypes.data_as(_lib.p_i32), int(iters), seed, ctypes.byref(blocks)))
    return order.astype(numpy.int64), int(blocks.value)


def colour_members(C, order, drop_tol=0.0, nseg=1, seed=1):
    """colour[member] in 0..3: the member's slot number mod 4.  The four rows of the expansion table that one block
    reads are shared-memory bank-conflict free iff their slots differ mod 4; the colours spread the members each row
    group uses evenly.  -> (colours, members left in conflict)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    S = _support(C, drop_tol)
    nrows, ncols = S.shape
    colour = numpy.zeros(ncols // nseg, dtype=numpy.int32)
    conflicts = ctypes.c_int32()
    order32 = numpy.ascontiguousarray(order, dtype=numpy.int32)
    _lib.check(lib.fiatb200_colour_members(S.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), nrows, ncols, nseg,
                                           order32.ctypes.data_as(_lib.p_i32), 200000, seed,
                                           colour.ctypes.data_as(_lib.p_i32), ctypes.byref(conflicts)))
    return colour, int(conflicts.value)


def slots_from_colours(colour):
    """Slot permutation realising the colours: perm[new slot] = member (old slot), new slot % 4 == colour."""
    K = len(colour)
    perm = numpy.empty(K, dtype=numpy.int64)
    nxt = [c for c in range(4)]
    for m in range(K):
        c = int(colour[m])
        perm[nxt[c]] = m
        nxt[c] += 4
    assert sorted(perm.tolist()) == list(range(K))
    return perm


DROP_BUDGET = 1e-13         # the dropped entries of a row may together contribute this fraction of their table's largest entry


def significant_entries(desc, t, folded, nstack=1):
    """Boolean mask of the coefficient entries the block packing must keep, per subcell matrix `folded[c]`
    (rows x member slots).  The folded matrices carry entries that are rounding noise of the reference's own
    coefficient computation (exact zeros by symmetry that come out as 1e-16 of the largest coefficient); dropping
    them is what makes the packing sparse.  First guess: everything below 1e-14 of the largest coefficient.  The guess
    is then VERIFIED on sample points (a lattice of the default simplex plus random points): at every sample point the
    function that a row loses, |sum_dropped C[r, m] psi_m(x)|, must stay below DROP_BUDGET of the largest entry
    that the row's table has at that same point (rows come in `nstack` equal groups: the derivative tables of a
    stacked derived element) -- the north-star tolerance is 1e-12 of the table's maximum over the caller's points,
    whatever those are.  Rows that fail (members whose magnitudes differ by orders of magnitude: the stacked P12
    triangle lost 1e-12 of its first-derivative tables to the first guess alone) get their entries dropped smallest
    contribution first only as far as the same pointwise test allows."""
    sd, n = int(desc["sd"]), int(desc["degree"])
    lat = n + 2
    idx = numpy.array([i for i in numpy.ndindex(*([lat + 1] * sd)) if sum(i) <= lat], dtype=float)
    rng = numpy.random.default_rng(20261020)
    u = numpy.sort(rng.random((192, sd)), axis=1)
    rnd = numpy.diff(numpy.concatenate([numpy.zeros((192, 1)), u], axis=1), axis=1)
    x = numpy.concatenate([2.0 * idx / lat - 1.0, 2.0 * rnd - 1.0])
    if len(x) > 640:
        x = x[rng.choice(len(x), 640, replace=False)]
    x = x.T                                                                    # (sd, npts) on the default simplex
    V = numpy.empty((t["nslots"], x.shape[1]))
    V[numpy.asarray(t["slot_of"])] = _member_values(t, sd, x)                  # by slot
    bound = numpy.abs(V).max(axis=1)
    cmax = max((float(numpy.abs(F).max()) if F.size else 0.0) for F in folded)
    keep = []
    for c, F in enumerate(folded):
        nrows = F.shape[0]
        per = nrows // nstack
        tab = numpy.abs(F @ V)                                                  # (rows, points); the cell's scale cancels
        scale = numpy.maximum(tab.reshape(nstack, per, -1).max(axis=1), 1e-300)  # (nstack, points): table maximum at x
        drop = numpy.abs(F) <= 1e-14 * cmax
        lost = numpy.abs(numpy.where(drop, F, 0.0) @ V)                          # what every row loses, pointwise
        grp = numpy.arange(nrows) // per
        bad = numpy.flatnonzero((lost / scale[grp]).max(axis=1) > DROP_BUDGET)
        for r0 in range(0, len(bad), 32):
            rows = bad[r0:r0 + 32]
            A = F[rows]
            order = numpy.argsort(numpy.abs(A) * bound[None, :], axis=1)
            As = numpy.take_along_axis(A, order, axis=1)                        # ascending contribution
            cum = numpy.abs(numpy.cumsum(As[:, :, None] * V[order], axis=1))     # (R, K, points): lost function of a prefix
            ratio = numpy.maximum.accumulate((cum / scale[grp[rows]][:, None, :]).max(axis=2), axis=1)
            kr = numpy.ones(A.shape, dtype=bool)
            numpy.put_along_axis(kr, order, ratio > DROP_BUDGET, axis=1)
            drop[rows] = ~kr
        keep.append(~drop & (F != 0.0))
    return keep


CLUSTER_ITERS = 100000      # swap rounds of the row clustering (P8 tet order 2, 1650 x 165: 0.4 s, 2826 -> 2406 blocks)


def schedule_row_blocks(counts):
    """Order in which the tile kernels hand out row blocks (dynamic, one warp per item): long and short blocks
    alternate (longest, shortest, 2nd longest, 2nd shortest, ...).  With the classic longest-first rule consecutive
    items have nearly equal lengths, so the warps of an SM finish them -- and enter their store epilogues, during
    which they feed no DMMAs -- in lockstep; alternating lengths keeps them out of phase.  Measured (2^20 points):
    P8 tet order 2 5.43 -> 5.24 ms, Nedelec 2nd kind deg 4 order 1 2.18 -> 2.09 ms, P10 triangle 1.39 -> 1.31 ms;
    round robin over 4 / 8 / 16 length strata and a random order were all worse than this.  FIATB200_RB_ORDER=longest
    restores longest-first (experiments)."""
    by_len = numpy.argsort(-numpy.asarray(counts), kind="stable")
    if os.environ.get("FIATB200_RB_ORDER", "mixed") == "longest":
        return by_len.astype(numpy.int32)
    out, lo, hi = [], 0, len(by_len) - 1
    while lo <= hi:
        out.append(by_len[lo])
        lo += 1
        if lo <= hi:
            out.append(by_len[hi])
            hi -= 1
    return numpy.array(out, dtype=numpy.int32)


def prefix_members(C, nseg=1, iters=4000, seed=1):
    """Slot permutation (perm[new slot] = old slot) for PREFIX packing: k-block j = new slots 4j..4j+3, and a
    (row block, subcell) pair stores the k-blocks 0 .. n-1 up to the last one any of its rows touches.  Members are
    ordered by how many (row block, subcell) pairs use them -- the low-degree members that every derivative row
    uses come first -- then improved by swaps.  Walkington tet order 2: 2 350 blocks before the swaps (gather packing
    1 651, dense 4 592); Guzman-Neilan 661 (523, 1 800)."""
    nrows, ncols = C.shape
    K = ncols // nseg
    if K <= 4:
        return numpy.arange(K)
    nrb = -(-nrows // 8)
    Cp = numpy.zeros((nrb * 8, ncols), dtype=bool)
    Cp[:nrows] = C != 0.0
    U = Cp.reshape(nrb, 8, nseg, K).any(axis=1).reshape(nrb * nseg, K)
    U = U[U.any(axis=1)]
    perm = numpy.argsort(-U.sum(axis=0), kind="stable")

    def cost(pm):
        last = K - numpy.argmax(U[:, pm][:, ::-1], axis=1)          # prefix length in members (every row of U is used)
        return int(numpy.ceil(last / 4.0).sum())

    rng = numpy.random.default_rng(seed)
    best = cost(perm)
    for _ in range(iters):
        i, j = (int(v) for v in rng.integers(0, K, 2))
        if i // 4 == j // 4:
            continue
        perm[i], perm[j] = perm[j], perm[i]
        c = cost(perm)
        if c <= best:
            best = c
        else:
            perm[i], perm[j] = perm[j], perm[i]
    return perm.astype(numpy.int64)


CELLS_REG_MAX_MEMBERS = 20  # members per subcell up to which the register-operand split-cell kernel wins (cells_launch.cu)
CELLS_STEP_RB = 4           # row blocks per step of the register-operand split-cell kernel (two rows per warp; 2 was 3-6 % slower)


def pack_fixed_stream(C, nseg, rb_per_step=CELLS_STEP_RB):
    """Fixed-k 8x4 block stream of per-subcell matrices for the register-operand split-cell kernel
    (csrc/cells_reg.cuh): C is (nrows, nseg * K) in packed row order.  k-block j = member slots 4j..4j+3 (the kernel
    holds the matching B fragments of its columns in registers, so blocks cannot gather).  The stream is cut into
    steps of `rb_per_step` row blocks; one step is one contiguous run of doubles

        [ nseg * rb_per_step int32 records: n | first block << 16 ]  padded to a multiple of 16 bytes
        [ the step's blocks, subcell-major then row block then k-block 0 .. n - 1, 32 doubles each in
          mma.m8n8k4 A-fragment order: lane l holds C[8 rb + l // 4, 4 j + l % 4] ]

    where n is the PREFIX of k-blocks that (subcell, row block) stores: everything up to the last k-block one of its
    rows touches (plan.prefix_members orders the slots so that prefixes are short); the kernel then enters an
    unrolled run of n blocks with ONE jump and no per-block test.  -> (stream, step_ptr in doubles)."""
    nrows, ncols = C.shape
    K = ncols // nseg
    KB = -(-K // 4)
    if KB > 16:
        return numpy.zeros(0), numpy.zeros(1, numpy.int32)
    nrb = -(-nrows // 8)
    nstep = -(-nrb // rb_per_step)
    Cp = numpy.zeros((nstep * rb_per_step * 8, nseg, KB * 4))
    Cp[:nrows, :, :K] = C.reshape(nrows, nseg, K)
    blocks = Cp.reshape(nstep, rb_per_step, 8, nseg, KB, 4).transpose(0, 3, 1, 4, 2, 5)     # step, cell, rb, kb, 8, 4
    present = (blocks != 0.0).any(axis=(4, 5))                                                # step, cell, rb, kb
    hdr = -(-(nseg * rb_per_step) // 4) * 2                 # doubles: int32 records padded to 16 bytes
    out, ptr = [], [0]
    for s in range(nstep):
        meta = numpy.zeros(hdr * 2, dtype=numpy.int32)
        frs, nb = [], 0
        for c in range(nseg):
            for r in range(rb_per_step):
                used = numpy.flatnonzero(present[s, c, r])
                n = int(used[-1]) + 1 if len(used) else 0
                for kb in range(n):
                    frs.append(blocks[s, c, r, kb].reshape(32))
                meta[c * rb_per_step + r] = n | (nb << 16)
                nb += n
        if nb >= 1 << 15:
            return numpy.zeros(0), numpy.zeros(1, numpy.int32)
        out.append(meta.view(numpy.float64))
        if frs:
            out.append(numpy.concatenate(frs))
        ptr.append(ptr[-1] + hdr + 32 * nb)
    return numpy.concatenate(out), numpy.array(ptr, dtype=numpy.int32)


def pack_blocks(C, drop_tol=0.0, nseg=1, min_one=False):
    """8x4 block-sparse *gather* packing of a (nrows, nseg * K) matrix in mma.m8n8k4 A-fragment order.

    Rows are taken 8 at a time in the given order (the caller clusters them); within a row block and a column
    segment the members with a non-zero column are dealt into ceil(n / 4) blocks of four, position t of a block
    preferring a member whose slot is t mod 4 (then the kernel's gather of the four table rows is bank-conflict
    free); leftover positions hold coefficient zero on slot t.  Lane l of a warp holds
    C[8 rb + l // 4, blk_idx[q, l % 4]].
    Returns (blk_ptr (nseg, nrb + 1), blk_idx (nblk, 4), frags (nblk * 32), rb_order, Kpad); blk_ptr is flat for
    nseg == 1.  min_one: every (segment, row block) gets at least one (possibly all-zero) block."""
    nrows, ncols = C.shape
    K = ncols // nseg
    nrb = -(-nrows // 8)
    Cp = numpy.zeros((nrb * 8, ncols))
    Cp[:nrows] = C
    used = (numpy.abs(Cp) > drop_tol).reshape(nrb, 8, nseg, K).any(axis=1)          # (nrb, nseg, K)
    ptrs, idx, frags = numpy.zeros((nseg, nrb + 1), dtype=numpy.int64), [], []
    total = 0
    for sg in range(nseg):
        ptrs[sg, 0] = total
        for rb in range(nrb):
            members = numpy.flatnonzero(used[rb, sg])
            nb = -(-len(members) // 4)
            if min_one:
                nb = max(nb, 1)
            if nb:
                blk = -numpy.ones((nb, 4), dtype=numpy.int64)
                extra = []
                for c in range(4):
                    mine = members[members % 4 == c]
                    blk[:min(nb, len(mine)), c] = mine[:nb]
                    extra.extend(mine[nb:].tolist())
                holes = numpy.argwhere(blk < 0)
                for (bq, bt), m in zip(holes, extra):
                    blk[bq, bt] = m
                tile = Cp[rb * 8:rb * 8 + 8, sg * K:(sg + 1) * K]
                for bq in range(nb):
                    fr = numpy.zeros((8, 4))
                    for t in range(4):
                        if blk[bq, t] >= 0:
                            fr[:, t] = tile[:, blk[bq, t]]
                        else:
                            blk[bq, t] = t if t < K else 0
                    idx.append(blk[bq])
                    frags.append(fr.reshape(32))
                total += nb
            ptrs[sg, rb + 1] = total
    counts = numpy.diff(ptrs, axis=1).sum(axis=0)
    rb_order = schedule_row_blocks(counts)
    frags = numpy.array(frags, dtype=float).reshape(-1) if frags else numpy.zeros(0)
    blk_idx = numpy.array(idx, dtype=numpy.int32).reshape(-1, 4)
    blk_ptr = ptrs.astype(numpy.int32)
    return (blk_ptr[0] if nseg == 1 else blk_ptr), blk_idx, frags, rb_order, -(-K // 4) * 4


def _member_degree(m, sd):
    """Total degree of the Morton-numbered member m (members are numbered degree by degree)."""
    k = 0
    while math.comb(k + sd, sd) <= m:
        k += 1
    return k


def _member_values(t, sd, x):
    """Un-normalised, Morton-numbered members (start value 1) at default-simplex points x (sd, npts).
    Works on complex points too, which is how the gradients are taken (complex step)."""
    T = numpy.zeros((t["nslots"],) + x.shape[1:], dtype=x.dtype)
    T[t["start_slot"]] = 1.0
    minus = numpy.zeros_like(x[0]) - 1.0
    X = [x[i] for i in range(sd)] + [minus, minus]
    for (nxt, cur, prv, codim), (a, b, c) in zip(t["step_idx"], t["step_abc"]):
        fb = 0.5 * (X[codim + 1] + X[codim + 2])
        fa = X[codim] + (fb + 1.0)
        v = (a * fa - b * fb) * T[cur]
        if prv >= 0:
            v = v - c * (fb * fb) * T[prv]
        T[nxt] = v
    return T[t["slot_of"]]


DMAT_RESIDUAL_LIMIT = 2e-13    # fitted derivative matrices must reproduce exact gradients this well, else jets are kept


def _derivative_matrices(t, sd, n):
    """D[i][m, m'] with  d psi_m / d xi_i = sum_m' D[i][m, m'] psi_m'  for the un-normalised Morton-numbered
    members on the default simplex.  The members of degree <= k span P_k, so row m only involves members
    of lower degree; each degree level is fitted on its own (well-conditioned) lower-degree block.
    Gradients are exact to rounding (complex step on the polynomial recurrence).  Returns None if the fit fails
    its self-check at independent points (very high degrees); callers then keep the derivative jets."""
    nmem = t["nslots"]
    lat = n + 3
    idx = numpy.array([i for i in numpy.ndindex(*([lat + 1] * sd)) if sum(i) <= lat], dtype=float)
    pts = (2.0 * idx / lat - 1.0).T                                     # (sd, npts), vertices (-1,..), (1,-1,..)
    V = _member_values(t, sd, pts)
    scale = numpy.abs(V).max(axis=1)
    Vs = V / scale[:, None]
    deg = numpy.array([_member_degree(m, sd) for m in range(nmem)])
    D = numpy.zeros((sd, nmem, nmem))
    h = 1e-40
    for i in range(sd):
        xc = pts.astype(complex)
        xc[i] += 1j * h
        Gs = _member_values(t, sd, xc).imag / h / scale[:, None]
        for k in range(1, n + 1):
            rows = numpy.flatnonzero(deg == k)
            nlow = math.comb(k - 1 + sd, sd)
            sol = numpy.linalg.lstsq(Vs[:nlow].T, Gs[rows].T, rcond=None)[0]      # (nlow, len(rows))
            D[i][rows[:, None], numpy.arange(nlow)[None, :]] = sol.T * scale[rows][:, None] / scale[None, :nlow]
    # self-check at independent points: gradients reproduced from values, relative to each member's largest gradient
    rng = numpy.random.default_rng(20261018)
    u = numpy.sort(rng.random((64, sd)), axis=1)
    chk = (2.0 * numpy.diff(numpy.concatenate([numpy.zeros((64, 1)), u], axis=1), axis=1) - 1.0).T
    Vc = _member_values(t, sd, chk)
    residual = 0.0
    for i in range(sd):
        xc = chk.astype(complex)
        xc[i] += 1j * h
        G = _member_values(t, sd, xc).imag / h
        big = numpy.abs(G).max(axis=1)
        ok = big > 0
        if ok.any():
            residual = max(residual, float((numpy.abs(G - D[i] @ Vc).max(axis=1)[ok] / big[ok]).max()))
    if residual > DMAT_RESIDUAL_LIMIT:
        return None
    return D


def derivative_coefficients(desc, t, ccell_morton, order):
    """Coefficient tables of the value-table kernel: out_alpha = C_alpha[cell] . psi(x), where psi are the
    member VALUES only.  D^alpha of a member of degree k is a combination of the members of degree
    <= k - |alpha| (FIAT itself relies on this: ExpansionSet.get_dmats, expansions.py:577-599), so the
    derivative jets of the recurrence (expansions.py:66-137) are replaced by host-side matrix products
    and the contraction for |alpha| = k only runs over C(n - k + sd, sd) members.

    Returns (flat, ncp): flat[(off_a + r * nm_k + m) * ncp + cell], alphas in mis order (the subcell index
    fastest keeps the lanes of a warp, whose points lie in different subcells, on distinct shared-memory
    banks), or (None, 0)."""
    sd, n = int(desc["sd"]), int(desc["degree"])
    ncells, nrows, nmem = ccell_morton.shape
    if n < 1 or sd < 2 or order > 3 or ncells > 16 or n > (6 if sd == 2 else 4):
        return None, 0          # outside the kernel's instantiations (vals.cuh)
    ncp = 1 if ncells == 1 else (4 if ncells <= 4 else 16)
    mats = alpha_matrices(desc, t, ccell_morton, order)
    if mats is None:
        return None, 0
    blocks = []
    for per_cell in mats:
        blk = numpy.zeros((nrows, per_cell[0].shape[1], ncp))
        for c in range(ncells):
            blk[:, :, c] = per_cell[c]
        blocks.append(blk.reshape(-1))
    return numpy.concatenate(blocks) if blocks else numpy.zeros(0), ncp


def alpha_matrices(desc, t, ccell_morton, order):
    """[alpha][cell] -> (nrows, C(n - |alpha| + sd, sd)) matrix C_alpha with
    D^alpha(table row) = C_alpha . (values of the un-normalised Morton-numbered members of degree <= n - |alpha|),
    derivatives taken in the parent cell's coordinates (chain rule through the subcell map xi = A x + b)."""
    sd, n = int(desc["sd"]), int(desc["degree"])
    ncells, nrows, nmem = ccell_morton.shape
    D = _derivative_matrices(t, sd, n)
    if D is None:
        return None
    out = []
    for alpha in alpha_list(sd, order):
        k = sum(alpha)
        nm = math.comb(n - k + sd, sd) if k <= n else 0
        per_cell = []
        for c in range(ncells):
            A = numpy.asarray(desc["cell_A"][c], dtype=float)
            Dx = [sum(A[i, j] * D[i] for i in range(sd)) for j in range(sd)]
            Da = numpy.eye(nmem)
            for j in range(sd):
                for _ in range(alpha[j]):
                    Da = Da @ Dx[j]
            per_cell.append((ccell_morton[c] @ Da)[:, :nm])
        out.append(per_cell)
    return out


MAX_MERGED_ROWS = 2048        # row blocks of 8 the tile kernel's constant tables hold (FB_MAX_RB in device_plan.cuh)


def alpha_split(desc, order, prog=None):
    """Split the tabulation of a single-cell Dubiner element into one order-0 tabulation per derivative
    multi-index: D^alpha of the element's functions is itself a set of polynomials of degree n - |alpha|,
    i.e. a derived element whose coefficient matrix (on the un-normalised recurrence members) is
    C_alpha = C . D_alpha (alpha_matrices).  The derived elements need no derivative jets and their
    expansion shrinks with |alpha|; whether that beats the one-pass jet tabulation depends on how sparse
    C . D_alpha stays, so the split is only proposed when it stores fewer 8x4 blocks in total than
    (blocks of C) x (number of alphas), the tile kernel's cost measure.

    Returns [(alpha, derived description or None for an identically zero table)] or None.
    `merged_split` stacks the derived elements into one."""
    if desc.get("kind") != "simplex" or desc.get("expansion") != "dubiner" or int(desc["ncells"]) != 1:
        return None
    if order < 1 or desc.get("raw_members"):
        return None
    sd, n = int(desc["sd"]), int(desc["degree"])
    if sd < 2 or n < 2:
        return None
    if prog is None:
        prog = compile_simplex(desc, order)
    if len(prog.blk_kb) == 0:
        return None
    t = _dubiner_tables(desc, order)
    mats = alpha_matrices(desc, t, prog.ccell_morton, order)
    if mats is None:
        return None
    alphas = alpha_list(sd, order)
    coeffs = numpy.asarray(desc["coeffs"])
    ndofs, ncomp = coeffs.shape[0], coeffs.shape[1]

    def stored_blocks(mat):
        if mat.size == 0:
            return 0
        return cluster_rows(mat, 1e-14 * max(numpy.abs(mat).max(), 1e-300), iters=0)[1]      # greedy estimate

    # stacked into one launch (merged_split) the split also saves the jets of the recurrence, which is worth a few
    # more blocks; as separate launches it must store clearly fewer
    mergeable = len(alphas) * ndofs * ncomp <= MAX_MERGED_ROWS
    split_blocks = sum(stored_blocks(m[0]) for m in mats)
    limit = float(os.environ.get("FIATB200_SPLIT_RATIO", 1.3)) if mergeable else 0.9       # env: tuning override
    if split_blocks > limit * stored_blocks(prog.ccell_morton[0]) * len(alphas):
        return None
    out = []
    for alpha, per_cell in zip(alphas, mats):
        k, mat = sum(alpha), per_cell[0]
        if mat.shape[1] == 0 or not numpy.abs(mat).max() > 0.0:
            out.append((alpha, None))
            continue
        d = {key: val for key, val in desc.items() if key not in ("nodes", "coeffs", "cell_node_map", "degree", "c0")}
        d.update(degree=n - k, c0=False, raw_members=True,
                 coeffs=numpy.ascontiguousarray(mat.reshape(ndofs, ncomp, mat.shape[1])),
                 cell_node_map=numpy.arange(mat.shape[1], dtype=numpy.int64)[None, :])
        out.append((alpha, d))
    return out


def _nstack(desc, nrows):
    """Number of equal row groups (stacked derivative tables) of a derived description; 1 for ordinary elements."""
    ns = int(desc.get("nstack", 1))
    return ns if ns >= 1 and nrows % ns == 0 else 1


def compile_simplex(desc, order):
    """Build the SimplexProgram of a `kind == "simplex"` description for one derivative order."""
    sd, n = int(desc["sd"]), int(desc["degree"])
    ncells = int(desc["ncells"])
    coeffs = numpy.asarray(desc["coeffs"], dtype=float)
    ndofs, ncomp, nexp_total = coeffs.shape
    nrows = ndofs * ncomp
    C = coeffs.reshape(nrows, nexp_total)
    cnm = numpy.asarray(desc["cell_node_map"], dtype=numpy.int64)
    na = len(alpha_list(sd, order))
    low1, mul1, low2, mul2 = leibniz_tables(sd, order)

    if desc["expansion"] == "dubiner":
        t = _dubiner_tables(desc, order)
        tile_cells = ncells > 1 and bool(desc.get("raw_members"))       # split-cell tile kernel (cells.cuh)
        if (ncells == 1 or tile_cells) and nrows * nexp_total >= 1024 and not desc.get("dense_only"):
            # Large elements go to the tile kernels, whose 8x4 blocks gather any four member slots; the four rows of
            # the expansion table a block reads are bank-conflict free iff the slots differ mod 4.  Slot numbers are
            # a free choice: pick each member's slot mod 4 (its colour) so that the members every row group uses
            # spread evenly over the colours.  (Split cells: one column segment per subcell, common slots.)
            wide = []
            for c in range(ncells):
                base = C[:, cnm[c][t["pos_of_slot"]]] * t["fold_by_slot"][None, :]
                f0 = base.copy()
                for (tgt, src), w in zip(t["fix_idx"], t["fix_w"]):
                    f0[:, src] -= w * base[:, tgt]
                wide.append(f0)
            keep0 = significant_entries(desc, t, wide, _nstack(desc, nrows))
            wide = numpy.concatenate([numpy.where(k, f, 0.0) for k, f in zip(keep0, wide)], axis=1)
            packed_rows = cluster_rows(wide, 0.0, nseg=ncells)[0]
            if tile_cells and nexp_total // ncells <= CELLS_REG_MAX_MEMBERS:
                # split cells with few members per subcell go to the register-operand kernel, which multiplies a
                # PREFIX of fixed k-blocks (four consecutive slots)
                slot_perm = prefix_members(wide[packed_rows], nseg=ncells)
            else:
                colour, _ = colour_members(wide, packed_rows, 0.0, nseg=ncells)
                slot_perm = slots_from_colours(colour)
            t = _dubiner_tables(desc, order, slot_perm=slot_perm)
            t["packed_rows"] = packed_rows          # row supports do not depend on the slot numbering
        fold = t["fold_by_slot"]
        line_tab, line_n = numpy.zeros(0), 0
    else:
        t = _line_tables(desc, order)
        fold = numpy.ones(t["nslots"])
        t["pos_of_slot"] = numpy.arange(t["nslots"])
        line_tab, line_n = t["line_tab"], t["line_n"]
        t.update(step_idx=numpy.zeros((0, 4), numpy.int32), step_abc=numpy.zeros((0, 3)), nat_abc=numpy.zeros((0, 3)),
                 slot_of=numpy.arange(t["nslots"]),
                 level_ptr=numpy.zeros(1, numpy.int32),
                 fix_idx=numpy.zeros((0, 2), numpy.int32), fix_w=numpy.zeros(0))
    nslots = t["nslots"]
    if cnm.shape[1] != nslots:
        raise ValueError("cell -> member map does not match the expansion set")
    ccell = numpy.empty((ncells, nrows, nslots))
    for c in range(ncells):
        # slot s holds the member the reference lists at position pos_of_slot[s] of the cell
        ccell[c] = C[:, cnm[c][t["pos_of_slot"]]] * fold[None, :]

    # fix-ups sorted by target, with one (first, count) record per distinct target
    fix_idx, fix_w = t["fix_idx"], t["fix_w"]
    perm = numpy.argsort(fix_idx[:, 0], kind="stable") if len(fix_idx) else numpy.zeros(0, dtype=int)
    t["fix_idx"], t["fix_w"] = fix_idx[perm], fix_w[perm]
    targets, first, count = numpy.unique(t["fix_idx"][:, 0], return_index=True, return_counts=True)
    fix_grp = numpy.stack([first, count], axis=1).astype(numpy.int32).reshape(-1, 2)

    # register kernel: members stay Morton-numbered and no fix-up pass runs, so fold both into C
    ccell_morton = numpy.empty_like(ccell)
    for c in range(ncells):
        folded = ccell[c].copy()
        for (tgt, src), w in zip(t["fix_idx"], t["fix_w"]):
            folded[:, src] -= w * ccell[c][:, tgt]
        ccell_morton[c] = folded[:, t["slot_of"]]

    bary = numpy.zeros((ncells + 1, 4, 4))
    if ncells > 1:
        bary[:, :sd + 1, :sd] = desc["bary_A"]
        bary[:, :sd + 1, 3] = desc["bary_b"]

    prog = SimplexProgram(
        sd=sd, degree=n, order=order, na=na, expansion=EXPANSION_CODES[desc["expansion"]],
        ncells=ncells, nslots=nslots, nrows=nrows, ndofs=ndofs,
        value_shape=tuple(int(s) for s in desc["value_shape"]),
        unique=int(desc["unique"]) if "unique" in desc else int(bool(desc["c0"]) and order == 0),
        geom=t["geom"], bary=bary, step_idx=t["step_idx"], step_abc=t["step_abc"], level_ptr=t["level_ptr"],
        nat_abc=t["nat_abc"], ccell_morton=ccell_morton, start_slot=int(t.get("start_slot", 0)),
        fix_idx=t["fix_idx"], fix_w=t["fix_w"],
        fix_grp=fix_grp, ccell=ccell, low1=low1, mul1=mul1, low2=low2, mul2=mul2, line_tab=line_tab, line_n=line_n)
    if desc["expansion"] == "dubiner":
        cder, ncp = (None, 0) if desc.get("dense_only") else derivative_coefficients(desc, t, ccell_morton, order)
        if cder is not None:
            prog.cderiv, prog.ncp = cder, ncp
        prog.slot_of = numpy.asarray(t["slot_of"], dtype=numpy.int64)
    if desc["expansion"] == "dubiner" and (ncells == 1 or desc.get("raw_members")) and not desc.get("dense_only"):
        # Tile kernels (kernels.cuh: k_mma, cells.cuh: k_mma_cells) have no fix-up phase: T' = X T  =>  C T' = (C X) T.
        # Split cells: one block stream per subcell (blk_ptr holds one (nrb + 1)-entry row per subcell, every row
        # block has at least one, possibly zero, block), common row order and slots.
        folded = []
        for c in range(ncells):
            f = ccell[c].copy()
            for (tgt, src), w in zip(prog.fix_idx, prog.fix_w):
                f[:, src] -= w * ccell[c][:, tgt]
            folded.append(f)
        keep = significant_entries(desc, t, folded, _nstack(desc, nrows))
        wide = numpy.concatenate([numpy.where(k, f, 0.0) for k, f in zip(keep, folded)], axis=1)
        # row blocks of the packed matrix are clusters of rows with similar member support
        rows_order = t["packed_rows"] if "packed_rows" in t else cluster_rows(wide, 0.0, nseg=ncells)[0]
        bp, bi, prog.blk_frag, prog.rb_order, prog.kpad = pack_blocks(wide[rows_order], 0.0, nseg=ncells, min_one=ncells > 1)
        prog.blk_ptr = numpy.ascontiguousarray(bp.reshape(-1), dtype=numpy.int32)
        prog.blk_kb = numpy.ascontiguousarray(bi.reshape(-1), dtype=numpy.int32)
        prog.row_perm = numpy.asarray(rows_order, dtype=numpy.int32)
        prog.blk_cells = ncells if ncells > 1 else 0
        if ncells > 1 and nslots <= 64:
            rb_step = int(os.environ.get("FIATB200_CELLS_RB", CELLS_STEP_RB))         # (override: experiments)
            prog.cstream, prog.cstep_ptr = pack_fixed_stream(wide[rows_order], ncells, rb_step)
            prog.crb = rb_step if len(prog.cstream) else 0
    return prog


def lattice_rowmap(desc):
    """Recognise the nodal basis of the principal lattice on the UFC simplex.

    If the element is scalar, spans the full P_n (ndofs = C(n+sd, sd) = number of expansion
    members), and its dofs are point evaluations at exactly the points with barycentric
    coordinates alpha/n, |alpha| = n, then its basis is the Lagrange basis of that lattice, which
    has the closed form  phi_alpha(lambda) = prod_i l_{alpha_i}(lambda_i),
    l_k(t) = prod_{j<k} (n t - j)/(j + 1)  -- mathematically the same functions as
    coeffs . expansion (FIAT/finite_element.py:132-165 solves for exactly these), at a fraction of
    the arithmetic.  Returns rowmap[loop index] = dof index for the loop nest
    a0 = 0..n, a1 = 0..n-a0, (a2 = 0..n-a0-a1 in 3-D), or None if the element does not qualify.
    The caller still verifies the fast path against the general kernel on the device.
    """
    if desc.get("kind") != "simplex" or "nodes" not in desc or int(desc["ncells"]) != 1:
        return None
    sd, n = int(desc["sd"]), int(desc["degree"])
    if sd not in (2, 3) or n < 1 or len(desc["value_shape"]) != 0 or desc["expansion"] != "dubiner":
        return None
    nodes = numpy.asarray(desc["nodes"], dtype=float)
    ndofs = math.comb(n + sd, sd)
    if nodes.shape != (ndofs, sd) or desc["coeffs"].shape != (ndofs, 1, ndofs):
        return None
    verts = numpy.asarray(desc["vertices"], dtype=float)
    ufc = numpy.concatenate([numpy.zeros((1, sd)), numpy.eye(sd)])
    if verts.shape != ufc.shape or not numpy.array_equal(verts, ufc):
        return None
    lam = numpy.concatenate([1.0 - nodes.sum(axis=1, keepdims=True), nodes], axis=1) * n
    alpha = numpy.rint(lam).astype(numpy.int64)
    if numpy.abs(lam - alpha).max() > 1e-9 or (alpha < 0).any() or (alpha.sum(axis=1) != n).any():
        return None
    where = {tuple(a): i for i, a in enumerate(alpha)}
    if len(where) != ndofs:
        return None
    rowmap = []
    for a0 in range(n + 1):
        for a1 in range(n + 1 - a0):
            if sd == 2:
                rowmap.append(where[(a0, a1, n - a0 - a1)])
            else:
                for a2 in range(n + 1 - a0 - a1):
                    rowmap.append(where[(a0, a1, a2, n - a0 - a1 - a2)])
    return numpy.array(rowmap, dtype=numpy.int32)


def merged_split(desc, order, split):
    """The derived elements of `alpha_split` stacked into ONE order-0 element with nalpha * nrows rows: its table
    rows are exactly the element's derivative tables one after the other (the output layout), the columns a
    level does not use are zero and cost nothing in the block-sparse packing, and the value recurrence runs once
    for all alphas.  None if the stack exceeds the tile kernel's row tables."""
    coeffs = numpy.asarray(desc["coeffs"])
    ndofs, ncomp = coeffs.shape[0], coeffs.shape[1]
    if len(split) * ndofs * ncomp > MAX_MERGED_ROWS:
        return None
    width = max((d["coeffs"].shape[2] for _, d in split if d is not None), default=0)
    if width == 0:
        return None
    top = next(d for _, d in split if d is not None and d["coeffs"].shape[2] == width)
    stacked = numpy.zeros((len(split) * ndofs, ncomp, width))
    for j, (_, d) in enumerate(split):
        if d is not None:
            stacked[j * ndofs:(j + 1) * ndofs, :, :d["coeffs"].shape[2]] = d["coeffs"]
    out = dict(top)
    out["coeffs"] = stacked
    out["nstack"] = len(split)
    return out


def stacked_derived(desc, order, prog=None):
    """Order-0 element whose rows are ALL derivative tables of `desc` up to `order`, one after the other
    (row = (table j, dof i, component c)), with dense per-subcell coefficient matrices on the un-normalised
    recurrence members: the operand of the fused point evaluation (api.Tabulator.evaluate), whose weights
    coefficients . C are formed on the device.  Works for single-cell and split-cell Dubiner sets of any size (no
    block packing is built: "dense_only").  None when the derivative matrices cannot be trusted (very high degree)
    or the set is not a Dubiner set."""
    if desc.get("kind") != "simplex" or desc.get("expansion") != "dubiner" or desc.get("raw_members"):
        return None
    sd, n, ncells = int(desc["sd"]), int(desc["degree"]), int(desc["ncells"])
    if n < 1:
        return None
    coeffs = numpy.asarray(desc["coeffs"])
    ndofs, ncomp = coeffs.shape[0], coeffs.shape[1]
    alphas = alpha_list(sd, order)
    if prog is None:
        prog = compile_simplex(desc, order)
    mats = alpha_matrices(desc, _dubiner_tables(desc, order), prog.ccell_morton, order)
    if mats is None:
        return None
    nmem = prog.nslots
    nrows = ndofs * ncomp
    stacked = numpy.zeros((len(alphas) * nrows, ncells * nmem))
    for j, per_cell in enumerate(mats):
        for c in range(ncells):
            m = per_cell[c]
            stacked[j * nrows:(j + 1) * nrows, c * nmem:c * nmem + m.shape[1]] = m
    d = {key: val for key, val in desc.items() if key not in ("nodes", "coeffs", "cell_node_map", "c0")}
    d.update(c0=False, raw_members=True, dense_only=True, unique=int(bool(desc["c0"]) and order == 0),
             coeffs=numpy.ascontiguousarray(stacked.reshape(len(alphas) * ndofs, ncomp, ncells * nmem)),
             cell_node_map=(numpy.arange(nmem, dtype=numpy.int64)[None, :]
                            + nmem * numpy.arange(ncells, dtype=numpy.int64)[:, None]))
    return d


PULLBACK_FORM_DEGREES = {
    "affine": (0,), "covariant piola": (1,), "contravariant piola": (2,), "double covariant piola": (1, 1),
    "double contravariant piola": (2, 2), "covariant contravariant piola": (1, 2), "contravariant covariant piola": (2, 1),
}


def pullback_description(desc, mapping, J=None, Jinv=None, Jdet=None):
    """Description of the element whose tabulation is `pullback(element.tabulate(...), mapping, J, Jinv, Jdet)`
    (FIAT/macro.py:601-645): value axis i of every table is multiplied by Jinv^T (form degree 1) or J / det J (form
    degree 2).  The map acts on the value axes only and tabulation is linear in the coefficient tensor
    (FIAT/polynomial_set.py:71), so it is folded into the coefficients once, on the host: the mapped tables come out
    of the same kernels at no extra cost per point."""
    try:
        formdegree = PULLBACK_FORM_DEGREES[mapping]
    except KeyError:
        raise ValueError(f"Unrecognized mapping {mapping}")
    if desc.get("kind") != "simplex":
        raise NotImplementedError("pullbacks are folded into the coefficients of Ciarlet elements on simplices")
    if J is None and Jinv is None:
        raise ValueError("J or Jinv is required")
    if J is None:
        J = numpy.linalg.pinv(Jinv)
    if Jinv is None:
        Jinv = numpy.linalg.pinv(J)
    if Jdet is None:
        Jdet = numpy.linalg.det(J)
    J, Jinv = numpy.asarray(J, dtype=float), numpy.asarray(Jinv, dtype=float)
    F1, F2 = Jinv.T, J / Jdet
    vs = tuple(int(v) for v in desc["value_shape"])
    coeffs = numpy.asarray(desc["coeffs"], dtype=float)
    phi = coeffs.reshape((coeffs.shape[0],) + vs + (coeffs.shape[-1],))
    if len(formdegree) > len(vs) and any(formdegree):
        raise ValueError(f"{mapping} needs {len(formdegree)} value axes, the element has {len(vs)}")
    for i, k in enumerate(formdegree):
        if k == 0:
            continue
        F = F1 if k == 1 else F2
        phi = numpy.moveaxis(numpy.tensordot(phi, F, axes=([i + 1], [1])), -1, i + 1)
    out = {key: val for key, val in desc.items() if key != "nodes"}
    out["value_shape"] = numpy.array(phi.shape[1:-1], dtype=numpy.int64)
    out["coeffs"] = numpy.ascontiguousarray(phi.reshape(phi.shape[0], -1, phi.shape[-1]))
    return out


def macro_merged(desc, order, prog=None):
    """Split-cell counterpart of alpha_split + merged_split: ONE derived order-0 element on the same complex whose
    rows are the element's derivative tables one after the other and whose per-subcell coefficient matrices
    (on the un-normalised recurrence members of each subcell) are the stacked C_alpha[cell].  It is what the
    split-cell tile kernel (cells.cuh) tabulates: value recurrence only, points of a tile binned by subcell.
    None if the element does not qualify."""
    if desc.get("kind") != "simplex" or desc.get("expansion") != "dubiner" or int(desc["ncells"]) < 2:
        return None
    if desc.get("raw_members"):
        return None
    sd, n, ncells = int(desc["sd"]), int(desc["degree"]), int(desc["ncells"])
    coeffs = numpy.asarray(desc["coeffs"])
    ndofs, ncomp = coeffs.shape[0], coeffs.shape[1]
    alphas = alpha_list(sd, order)
    if sd < 2 or n < 1 or len(alphas) * ndofs * ncomp > MAX_MERGED_ROWS:
        return None
    if prog is None:
        prog = compile_simplex(desc, order)
    mats = alpha_matrices(desc, _dubiner_tables(desc, order), prog.ccell_morton, order)
    if mats is None:
        return None
    nmem = prog.nslots
    stacked = numpy.zeros((len(alphas) * ndofs * ncomp, ncells * nmem))
    nrows = ndofs * ncomp
    for j, per_cell in enumerate(mats):
        for c in range(ncells):
            m = per_cell[c]
            stacked[j * nrows:(j + 1) * nrows, c * nmem:c * nmem + m.shape[1]] = m
    d = {key: val for key, val in desc.items() if key not in ("nodes", "coeffs", "cell_node_map", "c0")}
    # first-match binning of the original tabulation (continuity is not None and order == 0, expansions.py:452)
    # must survive: the derived element is always tabulated at order 0 and is not a C0 set
    d.update(c0=False, raw_members=True, unique=int(bool(desc["c0"]) and order == 0), nstack=len(alphas),
             coeffs=numpy.ascontiguousarray(stacked.reshape(len(alphas) * ndofs, ncomp, ncells * nmem)),
             cell_node_map=(numpy.arange(nmem, dtype=numpy.int64)[None, :]
                            + nmem * numpy.arange(ncells, dtype=numpy.int64)[:, None]))
    return d


@dataclass
class TensorLeaf:
    desc: dict                  # simplex description of the factor
    entity: tuple               # (dim, id) on the factor's own cell
    point_offset: int           # first coordinate of the product point that belongs to this factor
    point_dim: int              # number of coordinates consumed (dimension of the entity)
    sd: int                     # spatial dimension of the factor cell (length of its alpha slice)


def _cell_dim(desc):
    if desc["kind"] in ("simplex", "trace", "quadrature"):
        return int(desc["sd"])
    if desc["kind"] == "flattened":
        return _cell_dim(desc["element"])
    if desc["kind"] == "composite":
        return _cell_dim(desc["parts"][0]["element"])
    return _cell_dim(desc["A"]) + _cell_dim(desc["B"])


def value_shape_of(desc):
    kind = desc["kind"]
    if kind in ("trace", "quadrature"):
        return ()
    if kind in ("simplex", "composite"):
        return tuple(int(v) for v in desc["value_shape"])
    if kind == "flattened":
        return value_shape_of(desc["element"])
    return value_shape_of(desc["A"]) + value_shape_of(desc["B"])       # at most one factor is vector valued


def num_dofs_of(desc):
    kind = desc["kind"]
    if kind == "trace":
        return int(desc["ndofs"])
    if kind == "quadrature":
        return int(len(desc["points"]))
    if kind == "simplex":
        return int(desc["coeffs"].shape[0])
    if kind == "composite":
        return int(desc["ndofs"])
    if kind == "flattened":
        return num_dofs_of(desc["element"])
    return num_dofs_of(desc["A"]) * num_dofs_of(desc["B"])


@dataclass
class Part:
    """One kernel launch of a (possibly wrapped) element: a simplex or tensor-product description
    evaluated on `entity`, whose rows land in the final table through a placement map
    (dof = dof_base + own dof, component = comp_out[own component], value * sign)."""
    desc: dict
    entity: object
    dof_base: int
    comp_out: list
    sign: list


def resolve_parts(desc, entity=None):
    """Flatten wrapper elements (composite / flattened) into kernel-level parts for one entity.

    Placement maps compose: a part of a part adds dof offsets, chains component maps and multiplies
    signs (EnrichedElement of Hdiv(TensorProductElement) etc.)."""
    kind = desc["kind"]
    if kind == "simplex":
        nc = max(1, int(numpy.prod(desc["value_shape"])) if len(desc["value_shape"]) else 1)
        return [Part(desc, entity, 0, list(range(nc)), [1.0] * nc)]
    if kind == "tensor":
        vs = value_shape_of(desc)
        nc = int(numpy.prod(vs)) if vs else 1
        return [Part(desc, entity, 0, list(range(nc)), [1.0] * nc)]
    if kind == "flattened":
        inner = desc["element"]
        if inner["kind"] == "tensor":
            vs = value_shape_of(desc)
            nc = int(numpy.prod(vs)) if vs else 1
            return [Part(desc, entity, 0, list(range(nc)), [1.0] * nc)]
        if entity is None:
            entity = (_cell_dim(desc), 0)
        for fdim, fent, pdim, pent in desc["unflatten"]:
            if fdim == entity[0] and fent == entity[1]:
                pdim = tuple(pdim) if isinstance(pdim, list) else pdim
                return resolve_parts(inner, (pdim, pent))
        raise KeyError(f"no entity {entity} on the flattened cell")
    if kind != "composite":
        raise ValueError(kind)
    out = []
    for part in desc["parts"]:
        comp_out = [int(c) for c in part["comp_out"]]
        sign = [float(v) for v in part["sign"]]
        for sub in resolve_parts(part["element"], entity):
            out.append(Part(sub.desc, sub.entity, int(part["dof_offset"]) + sub.dof_base,
                            [comp_out[c] for c in sub.comp_out],
                            [sign[c] * sg for c, sg in zip(sub.comp_out, sub.sign)]))
    return out


def _flat(key):
    return sum(_flat(k) for k in key) if isinstance(key, (list, tuple)) else int(key)


def _norm_key(key):
    return [_norm_key(k) for k in key] if isinstance(key, (list, tuple)) else int(key)


def _count(top, key):
    key = _norm_key(key)
    for k, cnt in top:
        if _norm_key(k) == key:
            return cnt
    raise KeyError(f"no entities of dimension {key}")


def dimension_key(desc):
    """ref_el.get_dimension() of the described element: an int on simplices and flattened cells, a (possibly
    nested) tuple on tensor-product cells (FIAT/reference_element.py TensorProductCell.get_dimension)."""
    kind = desc["kind"]
    if kind == "tensor":
        return (dimension_key(desc["A"]), dimension_key(desc["B"]))
    if kind == "composite":
        return dimension_key(desc["parts"][0]["element"])
    return _cell_dim(desc)


def flatten_tensor(desc, entity=None, point_offset=0):
    """Resolve a tensor-product / flattened element tree into its leaf factors for one entity.

    Mirrors the entity factorisation of TensorProductElement.tabulate (tensor_product.py:238-250)
    and FlattenedDimensions.tabulate (:399-407).  Leaves come out in dof-major order: the global
    dof index is ((i0 * n1) + i1) * n2 + ..., and the alpha of the product is the concatenation of
    the leaves' alphas.
    """
    kind = desc["kind"]
    if kind == "simplex":
        sd = int(desc["sd"])
        if entity is None:
            entity = (sd, 0)
        return [TensorLeaf(desc, (int(entity[0]), int(entity[1])), point_offset, int(entity[0]), sd)]
    if kind == "flattened":
        if entity is None:
            entity = (_cell_dim(desc), 0)
        for fdim, fent, pdim, pent in desc["unflatten"]:
            if fdim == entity[0] and fent == entity[1]:
                pdim = tuple(pdim) if isinstance(pdim, list) else pdim
                return flatten_tensor(desc["element"], (pdim, pent), point_offset)
        raise KeyError(f"no entity {entity} on the flattened cell")
    if kind != "tensor":
        raise ValueError(kind)
    if entity is None:
        entity = (dimension_key(desc), 0)
    (dA, dB), eid = entity
    shape = (_count(desc["topA"], dA), _count(desc["topB"], dB))
    if not 0 <= eid < shape[0] * shape[1]:
        raise KeyError(f"no entity {entity} on the product cell")
    idA, idB = divmod(int(eid), shape[1])
    dA = tuple(dA) if isinstance(dA, (list, tuple)) else dA
    dB = tuple(dB) if isinstance(dB, (list, tuple)) else dB
    left = flatten_tensor(desc["A"], (dA, idA), point_offset)
    right = flatten_tensor(desc["B"], (dB, idB), point_offset + _flat(dA))
    return left + right
