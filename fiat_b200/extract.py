"""Read a reference-constructed FIAT element into a plain-data *element description*.

The new tabulation path never rebuilds FIAT's element construction (dual sets, quadrature,
Vandermonde inversion -- `FIAT/finite_element.py:132-165`).  It takes the element object the
reference built and reads exactly the constants the per-point arithmetic needs:

* expansion-set metadata: spatial dimension, embedded degree, variant, continuity
  (`FIAT/expansions.py:360-378`), per-subcell affine maps (`:368-370`), per-subcell scale
  (`get_scale`, `:386-399`), cell -> member map (`get_cell_node_map`, `:404-409`);
* the coefficient tensor `(ndofs, *value_shape, nexp)` (`FIAT/polynomial_set.py:57-62`);
* split-cell location data: rescaled barycentric maps of every subcell and of the parent
  simplex, computed by the reference's own `make_affine_mapping`
  (`FIAT/reference_element.py:616-644,1621-1654`);
* entity transforms `x -> x C + offset` (`FIAT/reference_element.py:570-609`);
* for 1-D Lagrange sets the nodes, barycentric weights and differentiation matrices
  (`FIAT/barycentric_interpolation.py:62-75`);
* for tensor-product elements the factor tree and entity maps
  (`FIAT/tensor_product.py:231-258,396-407`).

The description is a nested dict of numpy arrays / scalars / strings, so that it can be stored
as a fixture (`fiat_b200.description.save/load`) and used on a machine where FIAT itself is not
installed (the GPU box).  Nothing here imports FIAT: elements are recognised by duck typing.
"""
import sys

import numpy

__all__ = ["describe_element", "describe_expansion_set", "UnsupportedElement"]


class UnsupportedElement(NotImplementedError):
    """Raised for element types the device path does not cover (no CPU fallback exists)."""


def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def _affine_transform_data(transform, dim_in, dim_out):
    """Recover (C, offset) of an entity transform `x -> x C + offset`.

    The reference returns a closure over C and offset (reference_element.py:605-609) or the
    identity lambda (:585-587).  Read the closure cells so the constants are bit-identical;
    fall back to probing the map.
    """
    code = getattr(transform, "__code__", None)
    closure = getattr(transform, "__closure__", None)
    if code is not None and closure is None and code.co_argcount == 1 and not code.co_freevars:
        probe = numpy.arange(1.0, dim_in + 1.0)[None, :]
        if dim_in == dim_out and numpy.array_equal(numpy.asarray(transform(probe)), probe):
            return numpy.eye(dim_in), numpy.zeros(dim_out)
    if code is not None and closure is not None:
        cells = dict(zip(code.co_freevars, (c.cell_contents for c in closure)))
        if "C" in cells and "offset" in cells:
            C = numpy.asarray(cells["C"], dtype=float).reshape(dim_in, dim_out)
            offset = numpy.asarray(cells["offset"], dtype=float).reshape(dim_out)
            return C, offset
    offset = numpy.asarray(transform(numpy.zeros((1, dim_in))), dtype=float).reshape(dim_out)
    C = numpy.zeros((dim_in, dim_out))
    for j in range(dim_in):
        e = numpy.zeros((1, dim_in))
        e[0, j] = 1.0
        C[j] = numpy.asarray(transform(e), dtype=float).reshape(dim_out) - offset
    return C, offset


def _entity_transforms(ref_el):
    """All entity transforms of a simplicial reference element, as stacked arrays."""
    sd = ref_el.get_spatial_dimension()
    top = ref_el.get_topology()
    keys, Cs, offs = [], [], []
    for dim in sorted(top):
        for ent in sorted(top[dim]):
            if dim == sd and len(top[sd]) != 1:
                continue  # CiarletElement.ref_el is the (unsplit) cell; nothing to do
            C, off = _affine_transform_data(ref_el.get_entity_transform(dim, ent), dim, sd)
            Cpad = numpy.zeros((sd, sd))
            Cpad[:dim] = C
            keys.append((dim, ent))
            Cs.append(Cpad)
            offs.append(off)
    return (numpy.array(keys, dtype=numpy.int64).reshape(-1, 2),
            numpy.array(Cs, dtype=float).reshape(-1, sd, sd),
            numpy.array(offs, dtype=float).reshape(-1, sd))


def _rescaled_barycentric_map(ref_module, verts, sd):
    """(A_hat, b_hat) with lambda = p A_hat^T + b_hat, rows scaled by 1/|A_row|.

    Same arithmetic as compute_barycentric_coordinates(..., rescale=True)
    (reference_element.py:635-642), through the reference's own make_affine_mapping.
    """
    A, b = ref_module.make_affine_mapping(verts, numpy.eye(sd + 1))
    A = numpy.array(A, dtype=float)
    b = numpy.array(b, dtype=float)
    h = 1 / numpy.linalg.norm(A, axis=1)
    b *= h
    A *= h[:, None]
    return A, b


def _describe_ciarlet(element):
    poly_set = element.get_nodal_basis()
    es = poly_set.get_expansion_set()
    n = int(poly_set.get_embedded_degree())
    coeffs = numpy.array(poly_set.get_coeffs(), dtype=float)
    desc = _describe_expansion(es, n, coeffs, element.get_reference_element())
    # Plain point-evaluation dual sets (Lagrange-type elements): keep the nodes.  The plan compiler
    # uses them to recognise the nodal basis of the principal lattice, which has a closed product
    # form (fiat_b200/plan.py: lattice_rowmap).
    sd = int(desc["sd"])
    nodes = _point_evaluation_nodes(element, sd)
    if nodes is not None and len(desc["value_shape"]) == 0:
        desc["nodes"] = nodes
    return desc


def describe_expansion_set(es, n):
    """Description of the expansion set itself up to degree n: the "element" whose coefficient tensor is the
    identity, i.e. what ExpansionSet.tabulate / tabulate_derivatives / tabulate_jet return
    (FIAT/expansions.py:601-637, = ExpansionSet._tabulate, :449-490)."""
    nexp = int(es.get_num_members(n))
    ref_el = es.ref_el.get_parent() or es.ref_el
    return _describe_expansion(es, int(n), numpy.eye(nexp).reshape(nexp, nexp), ref_el)


def _describe_expansion(es, n, coeffs, cell):
    es_names = _mro_names(es)
    complex_ = es.ref_el
    sd = complex_.get_spatial_dimension()
    if sd == 0 or "PointExpansionSet" in es_names:
        raise UnsupportedElement("elements on a point are not tabulated on the device")
    top = complex_.get_topology()
    cells = sorted(top[sd])
    ncells = len(cells)
    if cells != list(range(ncells)):
        raise UnsupportedElement("non-contiguous subcell numbering")

    if "LagrangeLineExpansionSet" in es_names:
        expansion = "lagrange_line"
    elif "LineExpansionSet" in es_names and es.variant is None:
        expansion = "legendre_line"
    else:
        expansion = "dubiner"
    variant = "none" if es.variant is None else str(es.variant)
    if variant not in ("none", "bubble", "dual"):
        raise UnsupportedElement(f"unknown expansion variant {variant!r}")
    continuity = es.continuity
    if continuity not in (None, "C0"):
        raise UnsupportedElement(f"unsupported expansion continuity {continuity!r}")

    value_shape = tuple(int(s) for s in coeffs.shape[1:-1])
    nexp_total = int(es.get_num_members(n))
    if coeffs.shape[-1] != nexp_total:
        raise UnsupportedElement("coefficient tensor does not match the expansion set")

    desc = {
        "kind": "simplex",
        "sd": sd,
        "degree": n,
        "expansion": expansion,
        "variant": variant,
        "c0": continuity == "C0",
        "ncells": ncells,
        "value_shape": numpy.array(value_shape, dtype=numpy.int64),
        "nexp_total": nexp_total,
        "coeffs": numpy.ascontiguousarray(coeffs.reshape(coeffs.shape[0], -1, coeffs.shape[-1])),
    }

    cell_A = numpy.zeros((ncells, sd, sd))
    cell_b = numpy.zeros((ncells, sd))
    cell_scale = numpy.zeros(ncells)
    for c in cells:
        A, b = es.affine_mappings[c]
        cell_A[c] = A
        cell_b[c] = b
        cell_scale[c] = float(es.get_scale(n, cell=c))
    desc["cell_A"], desc["cell_b"], desc["cell_scale"] = cell_A, cell_b, cell_scale

    if ncells == 1 and expansion != "lagrange_line":
        cnm = numpy.arange(nexp_total, dtype=numpy.int64)[None, :]
    else:
        cnm = es.get_cell_node_map(n)
    if isinstance(cnm, dict):
        rows = [numpy.asarray(range(nexp_total) if cnm[c] is Ellipsis else cnm[c], dtype=numpy.int64)
                for c in cells]
        if len({len(r) for r in rows}) != 1:
            raise UnsupportedElement("ragged cell -> member map")
        cnm = numpy.stack(rows)
    desc["cell_node_map"] = numpy.array(cnm, dtype=numpy.int64).reshape(ncells, -1)

    if expansion == "lagrange_line":
        nn = {len(es.nodes[c]) for c in cells}
        if len(nn) != 1:
            raise UnsupportedElement("ragged 1-D Lagrange node sets")
        desc["ll_nodes"] = numpy.array([es.nodes[c] for c in cells], dtype=float)
        desc["ll_wts"] = numpy.array([es.weights[c] for c in cells], dtype=float)
        desc["ll_dmat"] = numpy.array([es.dmats[c] for c in cells], dtype=float)

    if ncells > 1:
        ref_module = sys.modules[type(complex_).__module__]
        if not hasattr(ref_module, "make_affine_mapping"):
            ref_module = sys.modules[type(complex_.get_parent()).__module__]
        bary_A = numpy.zeros((ncells + 1, sd + 1, sd))
        bary_b = numpy.zeros((ncells + 1, sd + 1))
        for c in cells:
            verts = complex_.get_vertices_of_subcomplex(top[sd][c])
            bary_A[c], bary_b[c] = _rescaled_barycentric_map(ref_module, verts, sd)
        parent = complex_.get_parent()
        ptop = parent.get_topology()
        pverts = parent.get_vertices_of_subcomplex(ptop[sd][0])
        bary_A[ncells], bary_b[ncells] = _rescaled_barycentric_map(ref_module, pverts, sd)
        desc["bary_A"], desc["bary_b"] = bary_A, bary_b

    keys, Cs, offs = _entity_transforms(cell)
    desc["ent_keys"], desc["ent_C"], desc["ent_off"] = keys, Cs, offs
    desc["vertices"] = numpy.array(complex_.get_vertices(), dtype=float).reshape(-1, sd)
    return desc


def _point_evaluation_nodes(element, sd):
    """(ndofs, sd) evaluation points if every dof is u -> u(x) (functional.py:156-166), else None."""
    try:
        dual_nodes = element.dual_basis()
    except Exception:
        return None
    pts = []
    for node in dual_nodes:
        pt_dict = getattr(node, "pt_dict", None)
        if not pt_dict or getattr(node, "deriv_dict", None) or len(pt_dict) != 1:
            return None
        (x, terms), = pt_dict.items()
        if len(terms) != 1 or tuple(terms[0][1]) != () or float(terms[0][0]) != 1.0 or len(x) != sd:
            return None
        pts.append([float(v) for v in x])
    return numpy.array(pts, dtype=float).reshape(-1, sd)


def _topology_counts(cell):
    """[(dim_key, number of entities)] for a reference cell; dim keys may be tuples."""
    top = cell.get_topology()
    out = []
    for key in sorted(top, key=lambda k: (k if isinstance(k, tuple) else (k,))):
        out.append((list(key) if isinstance(key, tuple) else int(key), len(top[key])))
    return out


def _dim_sum(key):
    return sum(_dim_sum(k) for k in key) if isinstance(key, (tuple, list)) else int(key)


def _describe_tensor(element):
    A, B = element.A, element.B
    if len(A.value_shape()) + len(B.value_shape()) > 1:
        raise UnsupportedElement("two vector-valued tensor-product factors (tensor_product.py:271-272)")
    cellA, cellB = element.ref_el.cells
    return {
        "kind": "tensor",
        "A": describe_element(A),
        "B": describe_element(B),
        "sdA": int(A.get_reference_element().get_spatial_dimension()),
        "sdB": int(B.get_reference_element().get_spatial_dimension()),
        "topA": _topology_counts(cellA),
        "topB": _topology_counts(cellB),
    }


def _describe_flattened(element):
    inner = element.element
    table = []
    for (fdim, fent), (pdim, pent) in sorted(element.unflattening_map.items()):
        table.append([int(fdim), int(fent), list(pdim) if isinstance(pdim, tuple) else int(pdim), int(pent)])
    return {"kind": "flattened", "element": describe_element(inner), "unflatten": table}


def _ncomp(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def _composite(ndofs, value_shape, parts):
    return {"kind": "composite", "ndofs": int(ndofs), "value_shape": numpy.array(value_shape, dtype=numpy.int64),
            "parts": parts}


def _describe_enriched(element):
    """EnrichedElement: child tables stacked along the dof axis (FIAT/enriched.py:88-113)."""
    vs = tuple(int(v) for v in element.value_shape())
    nc = _ncomp(vs)
    if len(vs) > 1:
        vs = (nc,)          # the reference's table is (ndofs, prod(value_shape), npts) (enriched.py:93-94)
    parts, off = [], 0
    for sub in element.elements():
        parts.append({"element": describe_element(sub), "dof_offset": off,
                      "comp_out": list(range(nc)), "sign": [1.0] * nc})
        off += int(sub.space_dimension())
    return _composite(off, vs, parts)


def _describe_mixed(element):
    """MixedElement: block-diagonal stacking in dofs and components (FIAT/mixed.py:61-92)."""
    parts, doff, coff = [], 0, 0
    for sub in element.elements():
        nc = _ncomp(sub.value_shape())
        parts.append({"element": describe_element(sub), "dof_offset": doff,
                      "comp_out": list(range(coff, coff + nc)), "sign": [1.0] * nc})
        doff += int(sub.space_dimension())
        coff += nc
    return _composite(doff, (coff,), parts)


def _describe_hdivcurl(element):
    """Hdiv(...) / Hcurl(...) of a tensor-product element (FIAT/hdivcurl.py:43-108,165-254): the plain
    tensor-product table with its components placed (possibly rotated and sign-flipped) into a
    vector of the cell's dimension.  The placement is read off by comparing the wrapper's table with
    the wrapped one at a few points (the wrapper only copies, permutes and negates values, so the
    comparison is exact).  This covers every branch of the reference without re-deriving its case
    analysis -- including the branch that does not do what its comment says (Hdiv of a covariant-Piola
    second factor, hdivcurl.py:100-107, places the components un-rotated): the drop-in has to reproduce the
    reference's tables, whatever they are."""
    inner = _describe_tensor(element)
    cell = element.get_reference_element()
    sd = int(cell.get_spatial_dimension())
    verts = numpy.array(cell.get_vertices(), dtype=float)
    rng = numpy.random.default_rng(7)
    w = rng.random((7, len(verts))) + 0.1
    pts = (w / w.sum(axis=1, keepdims=True)) @ verts
    key = (0,) * sd
    old = numpy.asarray(element.old_tabulate(0, pts)[key], dtype=float)
    new = numpy.asarray(element.tabulate(0, pts)[key], dtype=float)
    nd, npts = old.shape[0], old.shape[-1]
    old = old.reshape(nd, -1, npts)
    nc_in = old.shape[1]
    comp_out, sign = [None] * nc_in, [1.0] * nc_in
    for c in range(sd):
        target = new[:, c, :]
        if not numpy.any(target):
            continue
        for k in range(nc_in):
            for sg in (1.0, -1.0):
                if comp_out[k] is None and numpy.array_equal(target, sg * old[:, k, :]):
                    comp_out[k], sign[k] = c, sg
    if any(c is None for c in comp_out):
        raise UnsupportedElement("could not identify the component placement of an Hdiv/Hcurl wrapper")
    return _composite(nd, (sd,), [{"element": inner, "dof_offset": 0, "comp_out": comp_out, "sign": sign}])


def _describe_quadrature(element):
    """QuadratureElement (FIAT/quadrature_element.py:19-66): a set of points pretending to be an element; its
    "tabulation" at exactly those points is the identity."""
    cell = element.get_reference_element()
    return {"kind": "quadrature", "sd": int(cell.get_spatial_dimension()), "dim": int(cell.get_dimension()),
            "points": numpy.array(element._points, dtype=float).reshape(len(element._points), -1)}


def _describe_trace(element):
    """HDivTrace (FIAT/hdiv_trace.py:35-233): one discontinuous element per facet dimension, tabulated on a facet
    and written into that facet's block of rows; everything else is zero."""
    cell = element.get_reference_element()
    sd = int(cell.get_spatial_dimension())
    top = cell.get_topology()
    facets = []
    for dim in sorted(element.dg_elements):
        dg = element.dg_elements[dim]
        if sd == 1:
            # facets of an interval are points: the "element" there is a constant (PointExpansionSet, expansions.py:646-656)
            sub = {"kind": "point", "values": numpy.array(dg.tabulate(0, numpy.zeros((1, 0)))[()], dtype=float).reshape(-1)}
        else:
            sub = describe_element(dg)
        facets.append({"dim": list(dim) if isinstance(dim, tuple) else int(dim), "count": len(top[dim]),
                       "nf": int(dg.space_dimension()), "element": sub})
    simplex = hasattr(cell, "vertices") and len(cell.get_vertices()) == sd + 1 and not hasattr(cell, "cells")
    return {"kind": "trace", "sd": sd, "ndofs": int(element.space_dimension()), "simplex": bool(simplex),
            "vertices": numpy.array(cell.get_vertices(), dtype=float) if simplex else numpy.zeros((0, sd)),
            "facets": facets}


def describe_element(element):
    """Return the plain-data description of a FIAT element (see module docstring)."""
    names = _mro_names(element)
    if "QuadratureElement" in names:
        return _describe_quadrature(element)
    if "HDivTrace" in names:
        return _describe_trace(element)
    if "FlattenedDimensions" in names:
        return _describe_flattened(element)
    if "TensorProductElement" in names:
        if hasattr(element, "old_tabulate"):
            return _describe_hdivcurl(element)
        return _describe_tensor(element)
    if "EnrichedElement" in names:
        return _describe_enriched(element)
    if "MixedElement" in names:
        return _describe_mixed(element)
    if "DiscontinuousElement" in names:            # FIAT/discontinuous.py:56: same tabulation
        return describe_element(element._element)
    if "CiarletElement" in names:
        return _describe_ciarlet(element)
    raise UnsupportedElement(
        f"{type(element).__name__}: only CiarletElement, TensorProductElement, FlattenedDimensions, "
        "EnrichedElement, MixedElement, DiscontinuousElement, Hdiv/Hcurl wrappers, HDivTrace and QuadratureElement "
        "are tabulated on the device (no CPU fallback)")
