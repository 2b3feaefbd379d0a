"""Device versions of the reference's setup-path tabulations of an expansion set:

    ExpansionSet.tabulate / tabulate_derivatives / tabulate_jet   FIAT/expansions.py:601-637
    ExpansionSet.tabulate_normal_jumps                             FIAT/expansions.py:492-530
    ExpansionSet.tabulate_jumps                                    FIAT/expansions.py:532-575
    ExpansionSet.get_dmats                                         FIAT/expansions.py:577-599
    DualSet.to_riesz                                               FIAT/dual_set.py:86-206   (to_riesz below)

All of them are `ExpansionSet._tabulate` (:449-490) or `_tabulate_on_cell` (:411-447) of the set itself, i.e. the
tabulation of the "element" whose coefficient tensor is the identity (extract.describe_expansion_set), followed by
re-packing of the tables; they run through the same plans and kernels as `FiniteElement.tabulate`.  The reference uses
them while it CONSTRUCTS elements (macro.py, polynomial_set.py), where results feed linear solves whose output must
stay bit-compatible with numpy for the coefficients to be identical -- so these are offered beside the reference
(parity to the tabulation tolerance), not wired into element construction.
"""
import numpy
import torch

from . import plan as planmod
from .api import Tabulator
from .extract import describe_expansion_set

__all__ = ["ExpansionTabulator", "to_riesz"]


class ExpansionTabulator:
    """Device tabulation of one reference-constructed ExpansionSet up to degree n."""

    def __init__(self, expansion_set, n, device=None):
        self.es = expansion_set
        self.n = int(n)
        self.desc = describe_expansion_set(expansion_set, n)
        self.sd = int(self.desc["sd"])
        self.tab = Tabulator(self.desc, device)
        self.device = self.tab.device
        self._cells = {}

    # -- ExpansionSet.tabulate(n, pts): (num_members, npts) -------------------------------------------------------
    def tabulate(self, pts):
        if len(pts) == 0:
            return torch.zeros((0,), dtype=torch.float64, device=self.device)     # numpy.array([]) in the reference
        return self.tab.tabulate(0, pts)[(0,) * self.sd]

    # -- ExpansionSet.tabulate_derivatives(n, pts) ----------------------------------------------------------------
    def tabulate_derivatives(self, pts, nested=False):
        """(v, [dv_0, ..., dv_{sd-1}]) as device tensors (num_members, npts); nested=True re-packs them into the
        reference's list structure data[i][j] = (v[i, j], [dv_k[i, j] for k])."""
        vals = self.tab.tabulate(1, pts)
        v = vals[(0,) * self.sd]
        dv = [vals[a] for a in planmod.multi_indices(self.sd, 1)]
        if not nested:
            return v, dv
        vh, dvh = v.cpu().numpy(), [d.cpu().numpy() for d in dv]
        return [[(vh[i, j], [d[i, j] for d in dvh]) for j in range(vh.shape[1])] for i in range(vh.shape[0])]

    # -- ExpansionSet.tabulate_jet(n, pts, order) -----------------------------------------------------------------
    def tabulate_jet(self, pts, order=1):
        """[v0, v1, ..., v_order] with v_r of shape (num_members, npts) + (sd,) * r:
        v_r[i, j, k1, ..., kr] = d^r phi_i / dx_k1 ... dx_kr (pts[j])."""
        vals = self.tab.tabulate(order, pts)
        sd = self.sd
        v0 = vals[(0,) * sd]
        data = [v0]
        for r in range(1, order + 1):
            vr = torch.empty((sd,) * r + tuple(v0.shape), dtype=v0.dtype, device=v0.device)
            for index in numpy.ndindex(*((sd,) * r)):
                vr[index] = vals[tuple(index.count(k) for k in range(sd))]
            data.append(vr.permute((r, r + 1) + tuple(range(r))))
        return data

    # -- ExpansionSet.get_dmats(degree, cell) -----------------------------------------------------------------------
    def _cell_tabulator(self, cell):
        """_tabulate_on_cell(..., cell=cell): the single-cell set of one subcell (its affine map, scale, members)."""
        if cell not in self._cells:
            d = self.desc
            nmem = numpy.asarray(d["cell_node_map"]).shape[1]
            single = {k: v for k, v in d.items() if k not in ("bary_A", "bary_b", "nodes")}
            single.update(ncells=1, cell_A=d["cell_A"][cell:cell + 1], cell_b=d["cell_b"][cell:cell + 1],
                          cell_scale=d["cell_scale"][cell:cell + 1], nexp_total=nmem,
                          cell_node_map=numpy.arange(nmem, dtype=numpy.int64)[None, :],
                          coeffs=numpy.eye(nmem).reshape(nmem, 1, nmem))
            for key in ("ll_nodes", "ll_wts", "ll_dmat"):
                if key in d:
                    single[key] = d[key][cell:cell + 1]
            self._cells[cell] = Tabulator(single, self.device)
        return self._cells[cell]

    def get_dmats(self, lattice_points, cell=0):
        """dmat[k, j, i] with d/dx_k phi_j = sum_i dmat[k, j, i] phi_i, from the tabulation at `lattice_points`
        (the reference takes make_lattice(verts, degree, variant="gl")); the solve runs on the device."""
        sd = self.sd
        if self.n == 0:
            return torch.zeros((sd, 1, 1), dtype=torch.float64, device=self.device)
        v = self._cell_tabulator(cell).tabulate(1, lattice_points)
        rhs = torch.stack([v[a].T for a in planmod.multi_indices(sd, 1)])
        return torch.linalg.solve(v[(0,) * sd].T.unsqueeze(0).expand(sd, -1, -1), rhs)

    # -- ExpansionSet.tabulate_normal_jumps(n, ref_pts, facet, order) ----------------------------------------------
    def tabulate_normal_jumps(self, ref_pts, facet, order=0):
        """tensor (order + 1, num_members, npts): jumps of the r-th normal derivative (r = 0..order) across parent
        facet `facet` at points given on the reference facet (FIAT/expansions.py:492-530).  The r-th normal derivative
        of a member is sum_{|alpha| = r} r! / alpha! n^alpha D^alpha phi; per-subcell tables come from the device,
        subcell membership from the device binning."""
        import math
        es, sd = self.es, self.sd
        complex_ = es.ref_el
        transform = complex_.get_entity_transform(sd - 1, facet)
        pts = numpy.ascontiguousarray(numpy.asarray(transform(numpy.asarray(ref_pts, dtype=float)), dtype=numpy.float64)).reshape(-1, sd)
        cell_node_map = es.get_cell_node_map(self.n)
        ncells = int(self.desc["ncells"])
        if ncells > 1:
            mask = self.tab.locate_subcells(pts, unique=False).cpu().numpy().astype(numpy.int64)
        else:
            mask = numpy.ones(len(pts), dtype=numpy.int64)
        results = torch.zeros((order + 1, int(es.get_num_members(self.n)), len(pts)), dtype=torch.float64, device=self.device)
        facet_normal = numpy.asarray(complex_.compute_normal(facet), dtype=float)
        for cell in range(ncells):
            ipts = numpy.flatnonzero((mask >> cell) & 1)
            if len(ipts) == 0:
                continue
            normal = numpy.asarray(complex_.compute_normal(facet, cell=cell), dtype=float)
            side = float(numpy.dot(normal, facet_normal))
            phi = self._cell_tabulator(cell).tabulate(order, pts[ipts])
            rows = torch.as_tensor(numpy.asarray(cell_node_map[cell]), device=self.device, dtype=torch.long)
            cols = torch.as_tensor(ipts, device=self.device, dtype=torch.long)
            for r in range(order + 1):
                vr = None
                for alpha in planmod.multi_indices(sd, r):
                    w = math.factorial(r)
                    for k, a in enumerate(alpha):
                        w = w / math.factorial(a) * normal[k] ** a
                    vr = w * phi[alpha] if vr is None else vr + w * phi[alpha]
                sign = -1.0 if (r % 2 == 0 and side < 0) else 1.0
                results[r][rows[:, None], cols[None, :]] += sign * vr
        return results

    # -- ExpansionSet.tabulate_jumps(n, points, order) --------------------------------------------------------------
    def tabulate_jumps(self, points, order=0):
        """{r: tensor (num_members, len(mis(sd, r)) * num_jumps)}: jumps of the order-r derivatives across the
        interior facets of the complex at the points lying on them.  Point-to-subcell assignment: the device binning
        (bit-exact with compute_cell_point_map, unique=False)."""
        es, sd = self.es, self.sd
        complex_ = es.ref_el
        pts = numpy.ascontiguousarray(numpy.asarray(points, dtype=numpy.float64)).reshape(-1, sd)
        num_members = int(es.get_num_members(self.n))
        cell_node_map = es.get_cell_node_map(self.n)
        if int(self.desc["ncells"]) > 1:
            mask = self.tab.locate_subcells(pts, unique=False).cpu().numpy().astype(numpy.int64)
        else:
            mask = numpy.ones(len(pts), dtype=numpy.int64)
        ncells = int(self.desc["ncells"])
        cell_point_map = {c: numpy.flatnonzero((mask >> c) & 1) for c in range(ncells)}
        cell_point_map = {c: ip for c, ip in cell_point_map.items() if len(ip)}
        num_jumps, facet_point_map = 0, {}
        for facet in complex_.get_interior_facets(sd - 1):
            try:
                cells = complex_.connectivity[(sd - 1, sd)][facet]
                ipts = list(set.intersection(*(set(cell_point_map[c].tolist()) for c in cells)))
                if ipts != ():
                    facet_point_map[facet] = ipts
                    num_jumps += len(ipts)
            except KeyError:
                pass
        dpts = torch.as_tensor(pts, device=self.device)
        derivs = {c: self._cell_tabulator(c).tabulate(order, dpts) for c in cell_point_map}
        jumps = {}
        for r in range(order + 1):
            cur = 0
            alphas = planmod.multi_indices(sd, r)
            jr = torch.zeros((num_members, len(alphas) * num_jumps), dtype=torch.float64, device=self.device)
            for facet, ipts in facet_point_map.items():
                c0, c1 = complex_.connectivity[(sd - 1, sd)][facet]
                cols = torch.as_tensor(ipts, device=self.device, dtype=torch.long)
                for alpha in alphas:
                    rows1 = torch.as_tensor(numpy.asarray(cell_node_map[c1]), device=self.device, dtype=torch.long)
                    rows0 = torch.as_tensor(numpy.asarray(cell_node_map[c0]), device=self.device, dtype=torch.long)
                    block = slice(cur, cur + len(ipts))
                    jr[rows1, block] += derivs[c1][alpha][:, cols]
                    jr[rows0, block] -= derivs[c0][alpha][:, cols]
                    cur += len(ipts)
            jumps[r] = jr
        return jumps


def _is_moment(ell):
    """isinstance(ell, (functional.IntegralMoment, functional.IntegralMomentOfDerivative)) without importing the
    reference (FIAT/dual_set.py:128): these carry their own quadrature rule `Q`."""
    return any(c.__name__ in ("IntegralMoment", "IntegralMomentOfDerivative") for c in type(ell).__mro__)


def _quadratures_to_points(nodes, deriv):
    """FIAT/dual_set.py:121-149: functionals grouped by the quadrature rule they integrate with (None: point
    functionals), the points of every group, and the sorted union of all points."""
    from collections import defaultdict
    groups = defaultdict(list)
    for i, ell in enumerate(nodes):
        if len(ell.deriv_dict if deriv else ell.pt_dict) == 0:
            continue
        groups[ell.Q if _is_moment(ell) else None].append(i)
    pts, group_pts = set(), {}
    for Q, ells in groups.items():
        if Q is None:
            cur = set()
            for i in ells:
                cur.update((nodes[i].deriv_dict if deriv else nodes[i].pt_dict).keys())
            cur = tuple(cur)
        else:
            cur = tuple(map(tuple, Q.pts))
        group_pts[Q] = cur
        pts.update(cur)
    return groups, group_pts, sorted(pts)


def to_riesz(dual_set, poly_set, device=None):
    """Device version of `DualSet.to_riesz(poly_set)` (FIAT/dual_set.py:86-206): the action of every functional of
    the dual set on every member of the expansion set underlying `poly_set`, tensor
    (num_nodes, *target_shape, num_members).  The two tabulations of the expansion set (values at all evaluation /
    quadrature points, derivatives at the points of the derivative functionals) run through the device tabulator;
    the weight matrices are assembled on the host exactly like the reference's and contracted on the device."""
    nodes = dual_set.nodes
    tshape = tuple(nodes[0].target_shape)
    es = poly_set.get_expansion_set()
    ed = poly_set.get_embedded_degree()
    num_exp = int(es.get_num_members(ed))
    dev = ExpansionTabulator(es, ed, device)
    mat = torch.zeros((len(nodes),) + tshape + (num_exp,), dtype=torch.float64, device=dev.device)

    def accumulate(ells, wts, values):
        # mat[ells] += wts . values   (wts: (len(ells), *tshape, npts), values: (npts, num_exp))
        w = torch.as_tensor(wts, device=dev.device)
        idx = torch.as_tensor(ells, device=dev.device, dtype=torch.long)
        mat.index_add_(0, idx, (w.reshape(-1, w.shape[-1]) @ values).reshape(w.shape[:-1] + (num_exp,)))

    groups, group_pts, pts = _quadratures_to_points(nodes, deriv=False)
    if pts:
        values = dev.tabulate(numpy.asarray(pts, dtype=float)).T.contiguous()              # (npts, num_exp)
        where = {pt: j for j, pt in enumerate(pts)}
        for Q, ells in groups.items():
            cur = group_pts[Q]
            wts = numpy.zeros((len(ells),) + tshape + (len(cur),))
            if Q is None:
                col = {pt: j for j, pt in enumerate(cur)}
                for i, k in enumerate(ells):
                    for pt, wc_list in nodes[k].pt_dict.items():
                        for w, c in wc_list:
                            wts[i][c][col[pt]] = w
            else:
                for i, k in enumerate(ells):
                    wts[i][nodes[k].comp][:] = nodes[k].f_at_qpts
                wts *= Q.get_weights()
            rows = torch.as_tensor([where[pt] for pt in cur], device=dev.device, dtype=torch.long)
            accumulate(ells, wts, values[rows])

    max_deriv_order = max(ell.max_deriv_order for ell in nodes)
    if max_deriv_order > 0:
        groups, group_pts, pts = _quadratures_to_points(nodes, deriv=True)
        if pts:
            dvals = dev.tab.tabulate(max_deriv_order, numpy.asarray(pts, dtype=float))      # alpha -> (num_exp, npts)
            where = {pt: j for j, pt in enumerate(pts)}
            for Q, ells in groups.items():
                cur = group_pts[Q]
                dwts = {alpha: numpy.zeros((len(ells),) + tshape + (len(cur),)) for alpha in dvals if sum(alpha) > 0}
                if Q is None:
                    col = {pt: j for j, pt in enumerate(cur)}
                    for i, k in enumerate(ells):
                        for pt, wac_list in nodes[k].deriv_dict.items():
                            for w, alpha, c in wac_list:
                                dwts[tuple(alpha)][i][c][col[pt]] = w
                else:
                    for i, k in enumerate(ells):
                        for alpha in nodes[k].weights:
                            dwts[tuple(alpha)][i][nodes[k].comp][:] = nodes[k].weights[alpha]
                rows = torch.as_tensor([where[pt] for pt in cur], device=dev.device, dtype=torch.long)
                for alpha, wts in dwts.items():
                    if wts.any():
                        accumulate(ells, wts, dvals[alpha].T[rows])
    return mat
