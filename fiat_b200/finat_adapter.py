"""The FInAT caller's view of a tabulation, without `gem` or `ufl`.

`finat.FiatElement.basis_evaluation` (finat/fiat_elements.py:60-123) calls `fiat_element.tabulate(order, ps.points,
entity)` and re-shapes every table for code generation:

  * shape index_shape + value_shape + point_shape, the basis index first and the point index last (:117);
  * derivatives of total order == degree on a simplex are cell-wise constant: the point axis is dropped (:99-106);
  * derivatives of total order > degree are identically zero: a zero table without point axis (:107-111).

`basis_evaluation` below returns exactly those arrays as device tensors (what `gem.as_gem(fiat_table)` would wrap).
For tensor-product elements FInAT never forms the outer product: `TensorProductElement.basis_evaluation`
(finat/tensor_product.py:136-144) evaluates the factors on the factored point set and `_merge_evaluations` (:98-134)
multiplies them symbolically.  `factor_evaluations` returns the per-factor results (device tables that stay
unmultiplied -- nothing of size prod(n_l) is written) and `merge_evaluations` is the numeric counterpart of the
symbolic product, used by the tests and by callers that want the full table after all.
"""
import torch

from . import plan as planmod
from .api import get_tabulator

__all__ = ["basis_evaluation", "factor_evaluations", "merge_evaluations"]


def _degree_of(desc):
    if desc["kind"] == "simplex":
        return int(desc["degree"])
    if desc["kind"] == "flattened":
        return _degree_of(desc["element"])
    if desc["kind"] == "composite":
        return max(_degree_of(p["element"]) for p in desc["parts"])
    return max(_degree_of(desc["A"]), _degree_of(desc["B"]))


def basis_evaluation(element, order, points, entity=None, degree=None, point_shape=None, index_shape=None,
                     device=None, check=True):
    """{alpha: tensor of shape index_shape + value_shape + point_shape} for |alpha| <= order.

    element      FIAT element (or element description); `degree` defaults to its embedded degree, which is what
                 FInAT's `self.degree` is for the Lagrange-type families (pass it for the others)
    point_shape  extents of the point set's indices (default: one flat point index)
    index_shape  extents of the basis indices (default: (space_dimension,))
    check        verify the cell-wise constant / zero contracts like the reference's asserts (:104,109)"""
    tab = get_tabulator(element, device)
    desc = tab.desc
    simplex = desc["kind"] in ("simplex", "composite") and planmod._cell_dim(desc) >= 1 and \
        all(p.desc["kind"] == "simplex" for p in planmod.resolve_parts(desc, entity))
    degree = _degree_of(desc) if degree is None else int(degree)
    tables = tab.tabulate(order, points, entity)
    vs = planmod.value_shape_of(desc)
    ndofs = planmod.num_dofs_of(desc)
    index_shape = (ndofs,) if index_shape is None else tuple(index_shape)
    result = {}
    for alpha, table in tables.items():
        npts = table.shape[-1]
        pshape = (npts,) if point_shape is None else tuple(point_shape)
        derivative = sum(alpha)
        if derivative == degree and simplex:
            table = table.reshape(index_shape + vs + (-1,))
            if check and npts:
                scale = max(float(table.abs().max()), 1e-300)
                if float((table - table[..., :1]).abs().max()) > 1e-10 * scale:
                    raise AssertionError("tabulation of the top derivative is not cell-wise constant")
            result[alpha] = table[..., 0] if npts else table.new_zeros(index_shape + vs)
        elif derivative > degree:
            if check and npts and float(table.abs().max()) > 1e-8:        # numpy.allclose(table, 0.0), as the reference asserts
                raise AssertionError("tabulation above the degree is not zero")
            result[alpha] = table.new_zeros(index_shape + vs)
        else:
            result[alpha] = table.reshape(index_shape + vs + pshape)
    return result


def factor_evaluations(element, order, points, entity=None, device=None):
    """Per-factor results of a tensor-product element, the operands of finat/tensor_product.py:98-134:
    [(alpha slice, {delta: tensor (n_l, *value_shape_l, npts)})], factors in dof-major order."""
    tab = get_tabulator(element, device)
    out = []
    for aoff, sd, tables in tab.tabulate_factors(order, points, entity):
        out.append((slice(aoff, aoff + sd), tables))
    return out


def merge_evaluations(factors, order):
    """Numeric counterpart of `_merge_evaluations`: {Delta: tensor (n_0, ..., n_L, *value_shape, npts)} with
    result[Delta] = prod_l factor_l[Delta[slice_l]], basis indices of all factors first (:113-133)."""
    dimension = max(sl.stop for sl, _ in factors)
    result = {}
    for derivative in range(order + 1):
        for Delta in planmod.multi_indices(dimension, derivative):
            letters = iter("abcdefghijklmnopqrstuvw")
            sub_in, sub_basis, sub_value = [], [], []
            operands = []
            for sl, tables in factors:
                t = tables[tuple(Delta[sl])]
                basis = next(letters)
                value = "".join(next(letters) for _ in range(t.ndim - 2))
                sub_in.append(basis + value + "z")
                sub_basis.append(basis)
                sub_value.append(value)
                operands.append(t)
            spec = ",".join(sub_in) + "->" + "".join(sub_basis) + "".join(sub_value) + "z"
            result[Delta] = torch.einsum(spec, *operands)
    return result
