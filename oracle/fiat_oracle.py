"""TEST INFRASTRUCTURE -- not part of the product path.

CPU (numpy) restatement of the reference's basis-tabulation algorithm, working on the plain-data
element descriptions of `fiat_b200.extract.describe_element`.  Only `tests/`,
`__graft_entry__.smoke()` and the CPU legs of `bench.py` may import this module; the product
(`fiat_b200`) never does and has no CPU fallback.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this restatement against outputs of
the reference itself (`element.tabulate` of /root/reference, imported in the build container by
`tests/golden/gen/make_golden.py`, committed as fixtures under `tests/golden/`), including the
exact subcell assignment of split-cell elements.  The reference vendors no golden vectors of its
own for this path (its regression data lives in an external repository).

Each function cites the reference lines it follows (paths relative to /root/reference).
The arithmetic is vectorised over points with numpy, like the reference's, so it is also the
"port" CPU baseline timed by bench.py.
"""
import math

import numpy

__all__ = ["mis", "tabulate", "expansion_tabulate", "locate_cells", "resolve_entity"]


# ---------------------------------------------------------------------------------------------
# multi-indices -- FIAT/polynomial_set.py:23-32
# ---------------------------------------------------------------------------------------------
def mis(m, n):
    """All m-tuples of non-negative integers summing to n, in the reference's order."""
    if m == 1:
        return [(n,)]
    if n == 0:
        return [(0,) * m]
    out = []
    for i in range(n + 1):
        for tail in mis(m - 1, i):
            out.append((n - i,) + tail)
    return out


def all_alphas(sd, order):
    out = []
    for k in range(order + 1):
        out.extend(mis(sd, k))
    return out


# ---------------------------------------------------------------------------------------------
# member numbering -- FIAT/expansions.py:16-21, FIAT/reference_element.py:64-76
# ---------------------------------------------------------------------------------------------
def member_index(idx):
    if len(idx) == 1:
        return idx[0]
    if len(idx) == 2:
        p, q = idx
        return (p + q) * (p + q + 1) // 2 + q
    p, q, r = idx
    s = p + q + r
    return s * (s + 1) * (s + 2) // 6 + (q + r) * (q + r + 1) // 2 + r


def lattice(start, finish, depth):
    """Same visiting order as reference_element.lattice_iter (last entry is outermost)."""
    if depth == 0:
        yield ()
        return
    if depth == 1:
        for i in range(start, finish):
            yield (i,)
        return
    for i in range(start, finish):
        for head in lattice(start, finish - i, depth - 1):
            yield head + (i,)


# ---------------------------------------------------------------------------------------------
# Jacobi recurrence coefficients -- FIAT/expansions.py:24-40
# ---------------------------------------------------------------------------------------------
def jacobi_abc(a, b, n):
    an = (2 * n + 1 + a + b) * (2 * n + 2 + a + b) / (2 * (n + 1) * (n + 1 + a + b))
    bn = (a + b) * (a - b) * (2 * n + 1 + a + b) / (2 * (n + 1) * (n + 1 + a + b) * (2 * n + a + b))
    cn = (n + a) * (n + b) * (2 * n + 2 + a + b) / ((n + 1) * (n + 1 + a + b) * (2 * n + a + b))
    return an, bn, cn


def integrated_jacobi_abc(a, b, n):
    if n == 1:
        return (a + b + 2) / 2, (a - 3 * b - 2) / 2, 0.0
    return jacobi_abc(a - 1, b + 1, n - 1)


# ---------------------------------------------------------------------------------------------
# Leibniz rule for (affine or quadratic factor) x (jet) -- FIAT/expansions.py:66-137
# ---------------------------------------------------------------------------------------------
def _leibniz(F, dF, ddF, jet, alphas, index_of, k):
    """Order-k derivatives of F*G.  `jet[j]` holds D^{alphas[j]} G for all orders <= k.

    dF: list (per direction) of arrays/scalars or None; ddF: dict {(d1,d2): const} or None.
    Returns a list of arrays, one per multi-index of order k (reference order).
    """
    sd = len(alphas[0])
    out = []
    for alpha in alphas:
        if sum(alpha) != k:
            continue
        acc = F * jet[index_of[alpha]]
        if dF is not None and k >= 1:
            for d in range(sd):
                if alpha[d] >= 1:
                    lower = alpha[:d] + (alpha[d] - 1,) + alpha[d + 1:]
                    acc = acc + (alpha[d] * dF[d]) * jet[index_of[lower]]
        if ddF is not None and k >= 2:
            for d1 in range(sd):
                for d2 in range(d1, sd):
                    need = 2 if d1 == d2 else 1
                    if alpha[d1] < need or alpha[d2] < need:
                        continue
                    lower = list(alpha)
                    lower[d1] -= 1
                    lower[d2] -= 1
                    if d1 == d2:
                        mult = alpha[d1] * (alpha[d1] - 1) // 2
                    else:
                        mult = alpha[d1] * alpha[d2]
                    acc = acc + (mult * ddF[(d1, d2)]) * jet[index_of[tuple(lower)]]
        out.append(acc)
    return out


# ---------------------------------------------------------------------------------------------
# Dubiner / integrated-Jacobi recurrence -- FIAT/expansions.py:140-267
# ---------------------------------------------------------------------------------------------
def dubiner_table(sd, n, order, x, J, scale, variant):
    """Expansion members and derivatives on the (-1,1) simplex.

    x: (sd, npts) coordinates on the default simplex; J: (sd, sd) Jacobian of the map input ->
    default coordinates (rows = d x_i / d input).  Returns tab[member, alpha_index, point] with
    alpha_index running over all_alphas(sd, order), derivatives w.r.t. the *input* coordinates.
    """
    alphas = all_alphas(sd, order)
    index_of = {a: j for j, a in enumerate(alphas)}
    npts = x.shape[1]
    nmem = math.comb(n + sd, sd)
    tab = numpy.zeros((nmem, len(alphas), npts))
    if variant == "bubble":
        scale = -scale                                              # :176-177
    tab[0, 0, :] = scale                                            # :192
    if sd == 0 or n == 0:
        return tab
    beta = 1 if variant == "dual" else 0
    # padded coordinates and Jacobian rows -- :43-51, :181-182, :200
    X = [x[i] for i in range(sd)] + [-1.0, -1.0]
    dX = [J[i] for i in range(sd)] + [numpy.zeros(sd), numpy.zeros(sd)]
    orders_of = [sum(a) for a in alphas]

    def member(index):
        # idx(*index) with the trailing arguments defaulting to 0 -- :16-21, :201
        return member_index(tuple(index) + (0,) * (sd - len(index)))

    def jet_of(m):
        return tab[m]

    def store(m, k, comps):
        j0 = orders_of.index(k)
        for j, c in enumerate(comps):
            tab[m, j0 + j, :] = c

    for codim in range(sd):
        xc, yc, zc = X[codim:codim + 3]
        dx, dy, dz = dX[codim:codim + 3]
        # jacobi_factors -- :54-63
        fb = 0.5 * (yc + zc)
        fa = xc + (fb + 1.0)
        fc = fb ** 2
        dfb = 0.5 * (dy + dz)
        dfa = dx + dfb
        dfc = [2 * fb * dfb[d] for d in range(sd)]
        ddfc = {(d1, d2): 2 * dfb[d1] * dfb[d2] for d1 in range(sd) for d2 in range(d1, sd)}   # :205
        for sub in lattice(0, n, codim):
            ssum = sum(sub)
            icur = member(sub + (0,))
            inext = member(sub + (1,))
            if variant == "bubble":
                alpha = 2 * ssum
                a = b = -0.5
            else:
                alpha = 2 * ssum + len(sub)
                if variant == "dual":
                    alpha += 1 + len(sub)
                a = 0.5 * (alpha + beta) + 1.0
                b = 0.5 * (alpha - beta)
            # first step of the chain -- :221-228
            fcur = a * fa - b * fb
            tab[inext, 0, :] = tab[icur, 0, :] * fcur
            if order:
                dfcur = [a * dfa[d] - b * dfb[d] for d in range(sd)]
                deg = ssum + 1
                for k in range(1, min(order, deg) + 1):
                    store(inext, k, _leibniz(fcur, dfcur, None, jet_of(icur), alphas, index_of, k))
            # three-term steps -- :231-249
            for i in range(1, n - ssum):
                iprev, icur, inext = icur, inext, member(sub + (i + 1,))
                if variant == "bubble":
                    a, b, c = integrated_jacobi_abc(alpha, beta, i)
                else:
                    a, b, c = jacobi_abc(alpha, beta, i)
                fcur = a * fa - b * fb
                fprev = -c * fc
                v = tab[icur, 0, :] * fcur
                v += tab[iprev, 0, :] * fprev
                tab[inext, 0, :] = v
                if order:
                    dfcur = [a * dfa[d] - b * dfb[d] for d in range(sd)]
                    dfprev = [-c * dfc[d] for d in range(sd)]
                    ddfprev = {key: -c * val for key, val in ddfc.items()}
                    deg = ssum + 1 + i
                    for k in range(1, min(order, deg) + 1):
                        t1 = _leibniz(fcur, dfcur, None, jet_of(icur), alphas, index_of, k)
                        t2 = _leibniz(fprev, dfprev, ddfprev, jet_of(iprev), alphas, index_of, k)
                        store(inext, k, [u + w for u, w in zip(t1, t2)])
        # normalisation of this pass -- :251-266
        d = codim + 1
        shift = 1 if variant == "dual" else 0
        for index in lattice(0, n + 1, d):
            m = member(index)
            if variant != "none" and variant is not None:
                p = index[-1] + shift
                al = 2 * (sum(index[:-1]) + d * shift) - 1
                norm2 = (0.5 + d) / d
                if p > 0 and p + al > 0:
                    norm2 *= (p + al) * (2 * p + al) / p
            else:
                norm2 = (2 * sum(index) + d) / d
            tab[m] *= math.sqrt(norm2)
    return tab


# ---------------------------------------------------------------------------------------------
# C0 hierarchical basis -- FIAT/expansions.py:270-322
# ---------------------------------------------------------------------------------------------
def c0_entity_order(sd, n):
    """Gather list that reorders Morton-numbered members by (dimension, entity) -- :297-320."""
    ix = member_index
    dofs = list(range(sd + 1))
    rng = range(2, n + 1)
    if sd == 1:
        dofs += list(rng)
    elif sd == 2:
        dofs += [ix((1, i - 1)) for i in rng]
        dofs += [ix((0, i)) for i in rng]
        dofs += [ix((i, 0)) for i in rng]
        dofs += [ix((i, j)) for j in range(1, n + 1) for i in range(2, n - j + 1)]
    else:
        dofs += [ix((0, 1, i - 1)) for i in rng]
        dofs += [ix((1, 0, i - 1)) for i in rng]
        dofs += [ix((1, i - 1, 0)) for i in rng]
        dofs += [ix((0, 0, i)) for i in rng]
        dofs += [ix((0, i, 0)) for i in rng]
        dofs += [ix((i, 0, 0)) for i in rng]
        dofs += [ix((1, i - 1, j)) for j in range(1, n + 1) for i in range(2, n - j + 1)]
        dofs += [ix((0, i, j)) for j in range(1, n + 1) for i in range(2, n - j + 1)]
        dofs += [ix((i, 0, j)) for j in range(1, n + 1) for i in range(2, n - j + 1)]
        dofs += [ix((i, j, 0)) for j in range(1, n + 1) for i in range(2, n - j + 1)]
        dofs += [ix((i, j, k)) for k in range(1, n + 1) for j in range(1, n - k + 1)
                 for i in range(2, n - j - k + 1)]
    return dofs


def c0_fixups(sd, n):
    """[(target, [sources])]: target <- target - sum(sources), after target 0 is negated -- :281-295."""
    ix = member_index
    ops = []
    if sd == 2:
        for i in range(2, n + 1):
            ops.append((ix((0, i)), [ix((1, i - 1))]))
    elif sd == 3:
        for i in range(2, n + 1):
            for j in range(0, n + 1 - i):
                ops.append((ix((0, i, j)), [ix((1, i - 1, j))]))
            ops.append((ix((0, 0, i)), [ix((0, 1, i - 1)), ix((1, 0, i - 1))]))
    return ops


def apply_c0(sd, n, tab):
    tab = tab.copy()
    tab[0] *= -1.0
    for m in range(1, sd + 1):
        tab[0] -= tab[m]
    for target, sources in c0_fixups(sd, n):
        for s in sources:
            tab[target] -= tab[s]
    return tab[c0_entity_order(sd, n)]


# ---------------------------------------------------------------------------------------------
# 1-D sets
# ---------------------------------------------------------------------------------------------
def jacobi_batch(a, b, n, xs):
    """P_0..P_n^{(a,b)} at xs -- FIAT/jacobi.py:47-74."""
    out = numpy.zeros((n + 1, len(xs)))
    out[0] = 1.0
    if n > 0:
        out[1] = 0.5 * (a - b + (a + b + 2.0) * xs)
        apb = a + b
        for k in range(2, n + 1):
            a1 = 2.0 * k * (k + apb) * (2.0 * k + apb - 2.0)
            a2 = (2.0 * k + apb - 1.0) * (a * a - b * b)
            a3 = (2.0 * k + apb - 2.0) * (2.0 * k + apb - 1.0) * (2.0 * k + apb)
            a4 = 2.0 * (k + a - 1.0) * (k + b - 1.0) * (2.0 * k + apb)
            a2, a3, a4 = a2 / a1, a3 / a1, a4 / a1
            out[k] = (a2 + a3 * xs) * out[k - 1] - a4 * out[k - 2]
    return out


def legendre_line_table(n, order, pts, A, b, scale0):
    """LineExpansionSet._tabulate_on_cell for variant None -- FIAT/expansions.py:659-678."""
    Jinv = A[0, 0]
    xs = (pts @ A.T + b)[:, 0]
    tab = numpy.zeros((n + 1, order + 1, len(xs)))
    scale = scale0 * numpy.sqrt(2 * numpy.arange(n + 1) + 1)
    for k in range(order + 1):
        if n >= k:
            tab[k:, k, :] = jacobi_batch(k, k, n - k, xs)
        for p in range(n + 1):
            tab[p, k, :] *= scale[p]
            scale[p] *= 0.5 * (p + k + 1) * Jinv
    return tab


def lagrange_line_table(nodes, wts, dmat, order, pts):
    """Second barycentric formula -- FIAT/barycentric_interpolation.py:22-47."""
    x = pts[:, 0]
    with numpy.errstate(divide="ignore", invalid="ignore"):
        phi = 1.0 / (x[None, :] - nodes[:, None])
        phi *= wts[:, None]
        phi = (1.0 / numpy.sum(phi, axis=0)) * phi
    phi[phi != phi] = 1.0
    tab = numpy.zeros((len(nodes), order + 1, len(x)))
    tab[:, 0, :] = phi
    for r in range(1, order + 1):
        phi = dmat @ phi
        tab[:, r, :] = phi
    return tab


# ---------------------------------------------------------------------------------------------
# split-cell point location -- FIAT/expansions.py:771-811, FIAT/reference_element.py:616-644,779-780
# ---------------------------------------------------------------------------------------------
def l1_distance(pts, A_hat, b_hat):
    lam = numpy.dot(pts, A_hat.T)
    lam += b_hat
    return 0.5 * abs(numpy.sum(abs(lam) - lam, axis=-1))


def locate_cells(desc, pts, unique, tol=1e-12):
    """Boolean membership matrix near[cell, point] reproducing compute_cell_point_map."""
    ncells = int(desc["ncells"])
    npts = len(pts)
    if ncells == 1:
        return numpy.ones((1, npts), dtype=bool)
    best = l1_distance(pts, desc["bary_A"][ncells], desc["bary_b"][ncells])
    bound = best + tol
    near = numpy.zeros((ncells, npts), dtype=bool)
    taken = numpy.zeros(npts, dtype=bool)
    for c in range(ncells):
        hit = l1_distance(pts, desc["bary_A"][c], desc["bary_b"][c]) < bound
        if unique:
            hit &= ~taken
            taken |= hit
        near[c] = hit
    return near


# ---------------------------------------------------------------------------------------------
# expansion-set tabulation -- FIAT/expansions.py:411-490
# ---------------------------------------------------------------------------------------------
def tabulate_on_cell(desc, cell, pts, order):
    """tab[member_of_cell, alpha_index, point] on one (sub)cell -- :411-432."""
    sd = int(desc["sd"])
    n = int(desc["degree"])
    kind = desc["expansion"]
    if kind == "lagrange_line":
        return lagrange_line_table(desc["ll_nodes"][cell], desc["ll_wts"][cell], desc["ll_dmat"][cell],
                                   order, pts)
    A = desc["cell_A"][cell]
    b = desc["cell_b"][cell]
    scale = float(desc["cell_scale"][cell])
    if kind == "legendre_line":
        return legendre_line_table(n, order, pts, A, b, scale)
    x = numpy.add(numpy.dot(pts, A.T), b).T
    variant = desc["variant"]
    tab = dubiner_table(sd, n, order, x, A, scale, variant)
    if desc["c0"]:
        tab = apply_c0(sd, n, tab)
    return tab


def expansion_tabulate(desc, pts, order):
    """tab[member, alpha_index, point] over the whole complex -- ExpansionSet._tabulate :449-490."""
    sd = int(desc["sd"])
    ncells = int(desc["ncells"])
    nalpha = len(all_alphas(sd, order))
    npts = len(pts)
    if ncells == 1:
        return tabulate_on_cell(desc, 0, pts, order)
    unique = bool(desc["c0"]) and order == 0                 # :452
    near = locate_cells(desc, pts, unique)
    mult = near.sum(axis=0).astype(float)
    out = numpy.zeros((int(desc["nexp_total"]), nalpha, npts))
    cnm = desc["cell_node_map"]
    for c in range(ncells):
        ipts = numpy.where(near[c])[0]
        if len(ipts) == 0:
            continue
        tab = tabulate_on_cell(desc, c, pts[ipts], order)
        if not unique:
            tab = tab / mult[None, None, ipts]
        out[numpy.ix_(cnm[c], range(nalpha), ipts)] += tab
    return out


# ---------------------------------------------------------------------------------------------
# element level -- FIAT/finite_element.py:181-197, FIAT/polynomial_set.py:68-72
# ---------------------------------------------------------------------------------------------
def resolve_entity(desc, entity):
    """(C, offset) of the entity transform, or None for the identity (default entity)."""
    sd = int(desc["sd"])
    if entity is None:
        entity = (sd, 0)
    dim, ent = int(entity[0]), int(entity[1])
    keys = desc["ent_keys"]
    hit = numpy.where((keys[:, 0] == dim) & (keys[:, 1] == ent))[0]
    if len(hit) == 0:
        if dim == sd and ent == 0:
            return None
        raise KeyError(f"no entity {(dim, ent)} on this reference cell")
    j = int(hit[0])
    return desc["ent_C"][j][:dim], desc["ent_off"][j]


def _tabulate_simplex(desc, order, pts, entity):
    sd = int(desc["sd"])
    pts = numpy.asarray(pts, dtype=float)
    tr = resolve_entity(desc, entity)
    # one point given as a bare coordinate tuple: tables without a point axis (numpy broadcasting in
    # expansions.py:411-447; test/FIAT/unit/test_fiat.py test_single_point_tabulation)
    single = pts.ndim == 1 and pts.shape[0] > 0 and pts.shape[0] == (sd if tr is None else tr[0].shape[0])
    if single:
        pts = pts[None, :]
    if tr is not None:
        C, off = tr
        pts = pts.reshape(len(pts), C.shape[0])
        pts = numpy.add(numpy.dot(pts, C), off)
    pts = pts.reshape(-1, sd)
    base = expansion_tabulate(desc, pts, order)              # (nexp, nalpha, npts)
    coeffs = desc["coeffs"]                                   # (ndofs, ncomp, nexp)
    vs = tuple(int(s) for s in desc["value_shape"])
    flat = numpy.ascontiguousarray(coeffs.reshape(-1, coeffs.shape[-1]))
    result = {}
    for j, alpha in enumerate(all_alphas(sd, order)):
        # polynomial_set.py:71 -- same contraction, laid out so that numpy hands it to dgemm
        vals = numpy.dot(flat, numpy.ascontiguousarray(base[:, j, :]))
        result[alpha] = vals.reshape((coeffs.shape[0],) + vs + (() if single else (pts.shape[0],)))
    return result


def _norm_key(key):
    return [_norm_key(k) for k in key] if isinstance(key, (list, tuple)) else int(key)


def _key_count(top, key):
    for k, cnt in top:
        if _norm_key(k) == _norm_key(key):
            return cnt
    raise KeyError(key)


def _flat_dim(key):
    return sum(_flat_dim(k) for k in key) if isinstance(key, (list, tuple)) else int(key)


def dimension_key(desc):
    """ref_el.get_dimension(): nested tuples on (nested) tensor-product cells -- FIAT/tensor_product.py:234-235."""
    if desc["kind"] == "tensor":
        return (dimension_key(desc["A"]), dimension_key(desc["B"]))
    if desc["kind"] == "composite":
        return dimension_key(desc["parts"][0]["element"])
    return cell_dimension(desc)


def cell_dimension(desc):
    if desc["kind"] == "simplex":
        return int(desc["sd"])
    if desc["kind"] == "flattened":
        return cell_dimension(desc["element"])
    if desc["kind"] == "composite":
        return cell_dimension(desc["parts"][0]["element"])
    return cell_dimension(desc["A"]) + cell_dimension(desc["B"])


def _tabulate_tensor(desc, order, pts, entity):
    """TensorProductElement.tabulate, scalar x scalar -- FIAT/tensor_product.py:231-292."""
    sdA, sdB = cell_dimension(desc["A"]), cell_dimension(desc["B"])
    if entity is None:
        entity = (dimension_key(desc), 0)
    (dA, dB), eid = entity
    dA = tuple(dA) if isinstance(dA, (list, tuple)) else dA
    dB = tuple(dB) if isinstance(dB, (list, tuple)) else dB
    shape = (_key_count(desc["topA"], dA), _key_count(desc["topB"], dB))
    idA, idB = numpy.unravel_index(eid, shape)
    pts = numpy.asarray(pts, dtype=float)
    pA, pB = _flat_dim(dA), _flat_dim(dB)
    pts = pts.reshape(len(pts), pA + pB)
    Atab = tabulate(desc["A"], order, pts[:, :pA], (dA, int(idA)))
    Btab = tabulate(desc["B"], order, pts[:, pA:pA + pB], (dB, int(idB)))
    result = {}
    for alpha in all_alphas(sdA + sdB, order):
        a = Atab[alpha[:sdA]]
        b = Btab[alpha[sdA:]]
        if a.ndim == 2 and b.ndim == 2:                       # scalar x scalar -- :277-292
            result[alpha] = (a[:, None, :] * b[None, :, :]).reshape(a.shape[0] * b.shape[0], -1)
        elif a.ndim == 3 and b.ndim == 2:                     # vector x scalar -- :293-314
            prod = a[:, None, :, :] * b[None, :, None, :]     # (iA, iB, comp, pt)
            result[alpha] = prod.reshape(a.shape[0] * b.shape[0], a.shape[1], -1)
        elif a.ndim == 2 and b.ndim == 3:                     # scalar x vector -- :315-335
            prod = a[:, None, None, :] * b[None, :, :, :]
            result[alpha] = prod.reshape(a.shape[0] * b.shape[0], b.shape[1], -1)
        else:
            raise NotImplementedError("tabulate does not support two vector-valued inputs")
    return result


def _tabulate_composite(desc, order, pts, entity):
    """Wrapper elements: child tables placed into a zero-padded table --
    EnrichedElement FIAT/enriched.py:88-113, MixedElement FIAT/mixed.py:61-92,
    Hdiv/Hcurl FIAT/hdivcurl.py:43-108,165-254."""
    ndofs = int(desc["ndofs"])
    vs = tuple(int(v) for v in desc["value_shape"])
    nc_out = int(numpy.prod(vs)) if vs else 1
    result = {}
    for part in desc["parts"]:
        tab = tabulate(part["element"], order, pts, entity)
        off = int(part["dof_offset"])
        for alpha, arr in tab.items():
            npts = arr.shape[-1]
            arr3 = arr.reshape(arr.shape[0], -1, npts)
            out = result.get(alpha)
            if out is None:
                out = result[alpha] = numpy.zeros((ndofs, nc_out, npts))
            for k in range(arr3.shape[1]):
                out[off:off + arr3.shape[0], int(part["comp_out"][k]), :] = float(part["sign"][k]) * arr3[:, k, :]
    return {alpha: out.reshape((ndofs,) + vs + (out.shape[-1],)) for alpha, out in result.items()}


def tabulate(desc, order, pts, entity=None):
    """Drop-in restatement of `element.tabulate(order, points, entity)` on a description."""
    kind = desc["kind"]
    if kind == "simplex":
        return _tabulate_simplex(desc, order, pts, entity)
    if kind == "flattened":                                   # tensor_product.py:396-407
        if entity is None:
            entity = (cell_dimension(desc), 0)
        for fdim, fent, pdim, pent in desc["unflatten"]:
            if fdim == entity[0] and fent == entity[1]:
                return tabulate(desc["element"], order, pts, (pdim, pent))
        raise KeyError(entity)
    if kind == "tensor":
        return _tabulate_tensor(desc, order, pts, entity)
    if kind == "composite":
        return _tabulate_composite(desc, order, pts, entity)
    raise ValueError(kind)
