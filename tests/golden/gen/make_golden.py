"""Generate the golden fixtures under tests/golden/ by running the *reference* itself.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/gen/make_golden.py

The reference is imported from /root/reference with the `recursivenodes` stand-in that sits next
to this script (the real package is not installed here; it is only used at element-construction
time).  For every case we store
  * the element description (`fiat_b200.extract.describe_element`),
  * the evaluation points, derivative order and entity,
  * the reference's `element.tabulate(order, points, entity)` result,
  * for split-cell elements, the reference's `compute_cell_point_map` membership matrices.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import FIAT  # noqa: E402
from FIAT import expansions, polynomial_set  # noqa: E402
from FIAT.reference_element import ufc_simplex, UFCInterval  # noqa: E402
from FIAT.tensor_product import FlattenedDimensions  # noqa: E402

from fiat_b200 import description  # noqa: E402
from fiat_b200.extract import describe_element  # noqa: E402

OUT = os.path.abspath(os.path.join(HERE, ".."))


def simplex_points(rng, n, sd):
    u = numpy.sort(rng.random((n, sd)), axis=1)
    return numpy.diff(numpy.concatenate([numpy.zeros((n, 1)), u], axis=1), axis=1)


def adversarial_triangle_points(complex_):
    """Vertices, edge midpoints, barycentres of every subcell, offsets across interior facets,
    and a ring of exterior points."""
    top = complex_.get_topology()
    verts = numpy.array(complex_.get_vertices())
    pts = [v for v in verts]
    for e in top[1].values():
        a, b = verts[list(e)]
        mid = 0.5 * (a + b)
        nrm = numpy.array([-(b - a)[1], (b - a)[0]])
        nrm = nrm / numpy.linalg.norm(nrm)
        pts.append(mid)
        pts.append(0.25 * a + 0.75 * b)
        for eps in (2e-16, 1e-15, 1e-13, 1e-11, 1e-9):
            pts.append(mid + eps * nrm)
            pts.append(mid - eps * nrm)
    for c in top[2].values():
        pts.append(verts[list(c)].mean(axis=0))
    pts.append(numpy.array([1.0 / 3.0, 1.0 / 3.0]))
    for t in numpy.linspace(0, 2 * numpy.pi, 13)[:-1]:
        pts.append(numpy.array([1 / 3 + 0.9 * numpy.cos(t), 1 / 3 + 0.9 * numpy.sin(t)]))
    pts.append(numpy.array([0.2, 0.2 + 1e-13]))
    pts.append(numpy.array([0.2, 0.2 + 1e-11]))
    pts.append(numpy.array([-1e-14, 0.5]))
    pts.append(numpy.array([0.5, -1e-10]))
    return numpy.array(pts)


def _ulp_steps(x, k):
    """x moved by k representable numbers (k may be negative), coordinate-wise."""
    x = numpy.array(x, dtype=float)
    for _ in range(abs(k)):
        x = numpy.nextafter(x, numpy.inf if k > 0 else -numpy.inf)
    return x


OFFSETS = (1e-13, 1e-11, 1e-9, 3e-13, 5e-12)        # across a facet; the binning tolerance is 1e-12 (never ON it:
ULPS = (1, 2, 8)                                    # there the reference's own BLAS rounding decides)


def adversarial_points(complex_, rng, per_facet=4, n_random=0):
    """Adversarial point set for split-cell location in any dimension (FIAT/expansions.py:771-811): every subcell
    vertex, the barycentre of every edge / face / subcell, and for every interior facet its barycentre and
    `per_facet` random points on it, each also moved by +-{1,2,8} ulp in every coordinate and by
    +-{1e-13, 3e-13, 5e-12, 1e-11, 1e-9} along the facet normal; points near the facet's own boundary (where three
    or more subcells meet); a ring of exterior points and points just outside every parent facet; uniform points."""
    sd = complex_.get_spatial_dimension()
    top = complex_.get_topology()
    verts = numpy.array(complex_.get_vertices(), dtype=float)
    pts = [v for v in verts]
    for dim in range(1, sd + 1):
        for ent in top[dim].values():
            pts.append(verts[list(ent)].mean(axis=0))
    centre = verts[:sd + 1].mean(axis=0)

    def with_offsets(x, nrm):
        out = [x]
        for k in ULPS:
            out += [_ulp_steps(x, k), _ulp_steps(x, -k)]
        for eps in OFFSETS:
            out += [x + eps * nrm, x - eps * nrm]
        return out

    def normal_of(fv):
        # unit normal of the facet spanned by the sd vertices fv (rows)
        edges = fv[1:] - fv[0]
        _, _, vh = numpy.linalg.svd(edges)
        return vh[-1]

    for f in complex_.get_interior_facets(sd - 1):
        fv = verts[list(top[sd - 1][f])]
        nrm = normal_of(fv)
        lams = [numpy.full(sd, 1.0 / sd)]
        lams += list(rng.dirichlet(numpy.ones(sd), size=per_facet))
        # near a vertex / an edge of the facet: several subcells meet there
        e = numpy.zeros(sd)
        e[0] = 1.0
        lams += [(1 - 1e-9) * e + 1e-9 / sd, (1 - 1e-13) * e + 1e-13 / sd]
        for lam in lams:
            pts += with_offsets(lam @ fv, nrm)
    # exterior: a ring / sphere around the cell and points just outside every facet of the parent
    dirs = rng.standard_normal((12 * sd, sd))
    dirs /= numpy.linalg.norm(dirs, axis=1, keepdims=True)
    for r in (0.9, 2.5):
        pts += list(centre + r * dirs)
    parent = complex_.get_parent() or complex_
    ptop = parent.get_topology()
    pverts = numpy.array(parent.get_vertices(), dtype=float)
    for ent in ptop[sd - 1].values():
        fv = pverts[list(ent)]
        nrm = normal_of(fv)
        if numpy.dot(nrm, fv.mean(axis=0) - centre) < 0:
            nrm = -nrm
        for lam in [numpy.full(sd, 1.0 / sd)] + list(rng.dirichlet(numpy.ones(sd), size=2)):
            x = lam @ fv
            pts += [x + eps * nrm for eps in (1e-14, 1e-13, 1e-11, 1e-9, 1e-3, -1e-14)]
    if n_random:
        pts += list(simplex_points(rng, n_random, sd) @ (pverts[1:sd + 1] - pverts[0]) + pverts[0])
    return numpy.array(pts)


class CiarletElement:
    """Holder that presents a bare PolynomialSet through the element interface
    (used to pin expansion variants that no shipped element family exposes directly)."""

    def __init__(self, poly_set):
        self.poly_set = poly_set
        self.ref_el = poly_set.get_reference_element()

    def get_nodal_basis(self):
        return self.poly_set

    def get_reference_element(self):
        ref_el = self.ref_el
        return ref_el.get_parent() or ref_el

    def tabulate(self, order, points, entity=None):
        return self.poly_set.tabulate(numpy.asarray(points), order)


def membership(complex_, pts, unique):
    cpm = expansions.compute_cell_point_map(complex_, pts, unique=unique)
    ncells = len(complex_.get_topology()[complex_.get_spatial_dimension()])
    near = numpy.zeros((ncells, len(pts)), dtype=bool)
    for c, ipts in cpm.items():
        near[c, ipts] = True
    return near


ONLY = None        # set by `--only name1,name2`: write just these cases (the rng stream is still consumed in order)


def _as_lists(key):
    return [_as_lists(k) for k in key] if isinstance(key, tuple) else int(key)


def write_case(name, element, order, pts, entity=None, with_cells=False, table_stride=1):
    """table_stride > 1: the reference tables are stored for pts[::table_stride] only (large adversarial sets);
    the subcell membership matrices always cover all points (`mask_points`)."""
    if ONLY is not None and name not in ONLY:
        return
    desc = describe_element(element)
    all_pts = numpy.asarray(pts, dtype=float)
    pts = all_pts[::table_stride]
    ref = element.tabulate(order, pts, entity)
    # (trace elements put exception objects into the slots that are not defined, hdiv_trace.py:150-158)
    errors = [list(k) for k, v in ref.items() if isinstance(v, Exception)]
    ref = {k: v for k, v in ref.items() if not isinstance(v, Exception)}
    case = {
        "name": name,
        "desc": desc,
        "order": order,
        "points": numpy.asarray(pts, dtype=float),
        "entity": "none" if entity is None else [_as_lists(entity[0]), int(entity[1])],
        "keys": [list(k) for k in ref.keys()],
        "values": [numpy.asarray(v, dtype=float) for v in ref.values()],
    }
    if errors:
        case["error_keys"] = errors
    if with_cells:
        complex_ = element.get_nodal_basis().get_expansion_set().ref_el
        mpts = all_pts
        if entity is not None:      # binning sees the points on the cell (FIAT/finite_element.py:190-196)
            mpts = element.get_reference_element().get_entity_transform(*entity)(all_pts)
        case["near_unique"] = membership(complex_, mpts, True)
        case["near_all"] = membership(complex_, mpts, False)
        if table_stride > 1 or entity is not None:
            case["mask_points"] = numpy.asarray(mpts, dtype=float)
    path = os.path.join(OUT, f"case_{name}.npz")
    description.save(path, case)
    shape = next(iter(ref.values())).shape if ref else None
    print(f"{name:32s} {os.path.getsize(path) / 1024:8.1f} KiB  keys={len(ref)}  shape={shape}")


def main():
    rng = numpy.random.default_rng(20261018)
    T1, T2, T3 = UFCInterval(), ufc_simplex(2), ufc_simplex(3)

    # --- BASELINE configs at fixture sizes ---
    write_case("p3_tri_o1", FIAT.Lagrange(T2, 3), 1, simplex_points(rng, 200, 2))
    write_case("p8_tet_o2", FIAT.Lagrange(T3, 8), 2, simplex_points(rng, 32, 3))
    write_case("n2curl4_tet_o1", FIAT.NedelecSecondKind(T3, 4), 1, simplex_points(rng, 32, 3))
    for nm, el in (("hct", FIAT.HsiehCloughTocher(T2)),
                   ("ps6", FIAT.QuadraticPowellSabin6(T2)),
                   ("ps12", FIAT.QuadraticPowellSabin12(T2))):
        complex_ = el.get_nodal_basis().get_expansion_set().ref_el
        pts = numpy.concatenate([simplex_points(rng, 150, 2), adversarial_triangle_points(complex_)])
        write_case(f"{nm}_o2", el, 2, pts, with_cells=True)
        write_case(f"{nm}_o0", el, 0, pts, with_cells=True)
    G = FIAT.GaussLobattoLegendre(T1, 10)
    quad = FlattenedDimensions(FIAT.TensorProductElement(G, G))
    hexa = FlattenedDimensions(FIAT.TensorProductElement(quad, G))
    hpts = rng.random((8, 3))
    nodes = numpy.array([list(nd.get_point_dict().keys())[0][0] for nd in G.dual_basis()])
    hpts[0] = (nodes[3], 0.3, nodes[7])        # coordinates that coincide with GLL nodes
    hpts[1] = (0.0, 1.0, nodes[5])
    write_case("gll_q10_hex_o1", hexa, 1, hpts)

    # --- wider coverage of the same path ---
    write_case("p1_tri_o2", FIAT.Lagrange(T2, 1), 2, simplex_points(rng, 40, 2))
    write_case("p2_tri_facet1_o1", FIAT.Lagrange(T2, 2), 1, rng.random((17, 1)), entity=(1, 1))
    write_case("p2_tet_vertex_o1", FIAT.Lagrange(T3, 2), 1, numpy.zeros((1, 0)), entity=(0, 2))
    write_case("p4_tet_face2_o2", FIAT.Lagrange(T3, 4), 2, simplex_points(rng, 20, 2), entity=(2, 2))
    write_case("p5_tet_o3", FIAT.Lagrange(T3, 5), 3, simplex_points(rng, 24, 3))
    write_case("p6_tri_o4", FIAT.Lagrange(T2, 6), 4, simplex_points(rng, 24, 2))
    write_case("p4_line_o2", FIAT.Lagrange(T1, 4), 2, rng.random((33, 1)))
    write_case("gll7_line_o3", FIAT.GaussLobattoLegendre(T1, 7), 3,
               numpy.concatenate([rng.random((20, 1)), [[0.0], [1.0], [0.5]]]))
    write_case("legendre5_line_o3", FIAT.Legendre(T1, 5), 3, rng.random((25, 1)))
    write_case("dg3_tri_o1", FIAT.DiscontinuousLagrange(T2, 3), 1, simplex_points(rng, 30, 2))
    write_case("dp0_tet_o1", FIAT.DiscontinuousLagrange(T3, 0), 1, simplex_points(rng, 9, 3))
    write_case("cr_tri_o1", FIAT.CrouzeixRaviart(T2, 1), 1, simplex_points(rng, 30, 2))
    write_case("rt3_tri_o1", FIAT.RaviartThomas(T2, 3), 1, simplex_points(rng, 30, 2))
    write_case("bdm2_tet_o1", FIAT.BrezziDouglasMarini(T3, 2), 1, simplex_points(rng, 30, 3))
    write_case("ned2_tet_o2", FIAT.Nedelec(T3, 2), 2, simplex_points(rng, 30, 3))
    write_case("regge1_tri_o1", FIAT.Regge(T2, 1), 1, simplex_points(rng, 30, 2))
    write_case("argyris_tri_o2", FIAT.Argyris(T2, 5), 2, simplex_points(rng, 30, 2))
    write_case("morley_tri_o2", FIAT.Morley(T2), 2, simplex_points(rng, 30, 2))
    write_case("bubble4_tet_o1", FIAT.Bubble(T3, 4), 1, simplex_points(rng, 30, 3))
    write_case("intleg4_tri_o2", FIAT.IntegratedLegendre(T2, 4), 2, simplex_points(rng, 30, 2))

    for nm, el in (("p2_alfeld_tri", FIAT.Lagrange(T2, 2, variant="alfeld")),
                   ("p1_iso_tri", FIAT.Lagrange(T2, 1, variant="iso")),
                   ("p2_iso_line", FIAT.Lagrange(T1, 2, variant="iso"))):
        complex_ = el.get_nodal_basis().get_expansion_set().ref_el
        if complex_.get_spatial_dimension() == 2:
            pts = numpy.concatenate([simplex_points(rng, 60, 2), adversarial_triangle_points(complex_)])
        else:
            pts = numpy.concatenate([rng.random((30, 1)), [[0.5], [0.5 + 1e-13], [0.5 - 1e-11], [0.0], [1.0], [-0.1], [1.2]]])
        write_case(f"{nm}_o1", el, 1, pts, with_cells=True)
        write_case(f"{nm}_o0", el, 0, pts, with_cells=True)
    al3 = FIAT.Lagrange(T3, 2, variant="alfeld")
    write_case("p2_alfeld_tet_o2", al3, 2,
               numpy.concatenate([simplex_points(rng, 40, 3), [[0.25, 0.25, 0.25], [0.1, 0.1, 0.1], [0.2, 0.2, 0.3]]]),
               with_cells=True)

    # expansion variants with identity coefficients ("dual" is not exposed by an element family)
    for variant in (None, "bubble", "dual"):
        for sd, cell in ((1, T1), (2, T2), (3, T3)):
            P = polynomial_set.ONPolynomialSet(cell, 4, variant=variant)
            pts = rng.random((12, 1)) if sd == 1 else simplex_points(rng, 12, sd)
            write_case(f"on4_{variant or 'none'}_{sd}d_o2", CiarletElement(P), 2, pts)

    # tensor-product elements: quad with entities, prism, hex facet
    Q2 = FlattenedDimensions(FIAT.TensorProductElement(FIAT.Lagrange(T1, 2), FIAT.Lagrange(T1, 2)))
    write_case("q2_quad_o2", Q2, 2, rng.random((15, 2)))
    write_case("q2_quad_edge2_o1", Q2, 1, rng.random((9, 1)), entity=(1, 2))
    write_case("q2_quad_vertex3_o1", Q2, 1, numpy.zeros((1, 0)), entity=(0, 3))
    prism = FIAT.TensorProductElement(FIAT.Lagrange(T2, 2), FIAT.Lagrange(T1, 1))
    write_case("p2xp1_prism_o1", prism, 1, numpy.concatenate([simplex_points(rng, 11, 2), rng.random((11, 1))], axis=1))
    write_case("p2xp1_prism_facet_o1", prism, 1, numpy.concatenate([rng.random((7, 1)), rng.random((7, 1))], axis=1),
               entity=((1, 1), 2))
    G3 = FIAT.GaussLobattoLegendre(T1, 3)
    hex3 = FlattenedDimensions(FIAT.TensorProductElement(
        FlattenedDimensions(FIAT.TensorProductElement(G3, G3)), G3))
    write_case("gll_q3_hex_face4_o2", hex3, 2, rng.random((10, 2)), entity=(2, 4))
    dq = FlattenedDimensions(FIAT.TensorProductElement(FIAT.GaussLegendre(T1, 3), FIAT.GaussLegendre(T1, 2)))
    write_case("dq32_quad_o2", dq, 2, rng.random((10, 2)))

    # more element families through the same path (single-cell, split-cell, vector / tensor valued)
    from FIAT.reference_element import UFCQuadrilateral
    extra = [
        ("hermite3_tet_o2", FIAT.CubicHermite(T3), 2, 3), ("bell_tri_o2", FIAT.Bell(T2), 2, 2),
        ("ned1_3_tet_o1", FIAT.Nedelec(T3, 3), 1, 3), ("regge2_tet_o1", FIAT.Regge(T3, 2), 1, 3),
        ("aw_tri_o2", FIAT.ArnoldWinther(T2, 3), 2, 2), ("hz3_tri_o1", FIAT.HuZhang(T2, 3), 1, 2),
        ("jm_tri_o2", FIAT.JohnsonMercier(T2, 1), 2, 2), ("gn_tet_o2", FIAT.GuzmanNeilanFirstKindH1(T3, 1), 2, 3),
        ("alfeld_sorokina_tri_o2", FIAT.AlfeldSorokina(T2, 2), 2, 2), ("christiansen_hu_tri_o1", FIAT.ChristiansenHu(T2, 1), 1, 2),
        ("kmv2_tri_o2", FIAT.KongMulderVeldhuizen(T2, 2), 2, 2), ("p4_spectral_tri_o2", FIAT.Lagrange(T2, 4, variant="spectral"), 2, 2),
        ("hct4_tri_o2", FIAT.HsiehCloughTocher(T2, 4), 2, 2), ("hct_red_tri_o2", FIAT.HsiehCloughTocher(T2, reduced=True), 2, 2),
        ("walkington_tet_o2", FIAT.Walkington(T3), 2, 3),
        ("restricted_p3_tri_o1", FIAT.RestrictedElement(FIAT.Lagrange(T2, 3), restriction_domain="facet"), 1, 2),
        ("nodal_enriched_tri_o2", FIAT.NodalEnrichedElement(FIAT.Lagrange(T2, 2), FIAT.Bubble(T2, 3)), 2, 2),
        ("gauss_radau3_line_o2", FIAT.GaussRadau(T1, 3), 2, 1), ("fdm3_line_o2", FIAT.FDMLagrange(T1, 3), 2, 1),
        ("histopolation3_line_o1", FIAT.Histopolation(T1, 3), 1, 1), ("p10_tri_o2", FIAT.Lagrange(T2, 10), 2, 2),
        ("p6_tet_o1", FIAT.Lagrange(T3, 6, variant="spectral"), 1, 3),
    ]
    for nm, el, order, sd in extra:
        pts = rng.random((14, 1)) if sd == 1 else simplex_points(rng, 14, sd)
        macro = el.get_nodal_basis().get_expansion_set().ref_el.is_macrocell()
        write_case(nm, el, order, pts, with_cells=macro)
    write_case("dpc2_quad_o2", FIAT.DPC(UFCQuadrilateral(), 2), 2, rng.random((14, 2)))

    # wrapper elements: enriched, mixed, discontinuous, Hdiv/Hcurl on tensor products,
    # vector-valued tensor-product factors
    from FIAT.hdivcurl import Hdiv, Hcurl
    P1, P2 = FIAT.Lagrange(T1, 1), FIAT.Lagrange(T1, 2)
    DP0, DP1 = FIAT.DiscontinuousLagrange(T1, 0), FIAT.DiscontinuousLagrange(T1, 1)
    TPE = FIAT.TensorProductElement

    def prism_points(n):
        return numpy.concatenate([simplex_points(rng, n, 2), rng.random((n, 1))], axis=1)

    rtcf1 = FlattenedDimensions(FIAT.EnrichedElement(Hdiv(TPE(P1, DP0)), Hdiv(TPE(DP0, P1))))
    rtce2 = FlattenedDimensions(FIAT.EnrichedElement(Hcurl(TPE(DP1, P2)), Hcurl(TPE(P2, DP1))))
    nchex = FlattenedDimensions(FIAT.EnrichedElement(
        Hcurl(TPE(FlattenedDimensions(TPE(DP0, P1)), P1)), Hcurl(TPE(FlattenedDimensions(TPE(P1, DP0)), P1)),
        Hcurl(TPE(FlattenedDimensions(TPE(P1, P1)), DP0))))
    wrappers = [
        ("rtcf1_quad_o1", rtcf1, 1, rng.random((11, 2)), None),
        ("rtcf1_quad_edge1_o1", rtcf1, 1, rng.random((5, 1)), (1, 1)),
        ("rtce2_quad_o2", rtce2, 2, rng.random((11, 2)), None),
        ("nce1_hex_o1", nchex, 1, rng.random((11, 3)), None),
        ("mini_tri_o2", FIAT.EnrichedElement(FIAT.Lagrange(T2, 1), FIAT.Bubble(T2, 3)), 2, simplex_points(rng, 11, 2), None),
        ("taylor_hood_tri_o1", FIAT.MixedElement([FIAT.Lagrange(T2, 2), FIAT.Lagrange(T2, 2), FIAT.Lagrange(T2, 1)]), 1,
         simplex_points(rng, 11, 2), None),
        ("rt1_dg0_mixed_tri_o1", FIAT.MixedElement([FIAT.RaviartThomas(T2, 1), FIAT.DiscontinuousLagrange(T2, 0)]), 1,
         simplex_points(rng, 11, 2), None),
        ("rt1xdp0_prism_hdiv_o1", Hdiv(TPE(FIAT.RaviartThomas(T2, 1), DP0)), 1, prism_points(11), None),
        ("ned1xp1_prism_hcurl_o1", Hcurl(TPE(FIAT.Nedelec(T2, 1), P1)), 1, prism_points(11), None),
        ("rt1xp1_prism_hcurl_rot_o1", Hcurl(TPE(FIAT.RaviartThomas(T2, 1), P1)), 1, prism_points(11), None),
        ("dp1xp2_prism_hdiv_o2", Hdiv(TPE(FIAT.DiscontinuousLagrange(T2, 1), P2)), 2, prism_points(11), None),
        ("rt2xp1_prism_vector_o1", TPE(FIAT.RaviartThomas(T2, 2), P1), 1, prism_points(11), None),
        ("p1xrt1_vector_b_o1", TPE(P1, FIAT.RaviartThomas(T2, 1)), 1,
         numpy.concatenate([rng.random((11, 1)), simplex_points(rng, 11, 2)], axis=1), None),
        ("dg_wrapped_p2_tri_o2", FIAT.DiscontinuousElement(FIAT.Lagrange(T2, 2)), 2, simplex_points(rng, 11, 2), None),
        # wrapper elements whose parts are mid-size single-cell elements (per-alpha split + row placement)
        ("n2curl3_p3_mixed_tet_o1", FIAT.MixedElement([FIAT.NedelecSecondKind(T3, 3), FIAT.Lagrange(T3, 3, variant="spectral")]), 1,
         simplex_points(rng, 11, 3), None),
        ("enriched_p4s_bubble5_tet_o2", FIAT.EnrichedElement(FIAT.Lagrange(T3, 4, variant="spectral"), FIAT.Bubble(T3, 5)), 2,
         simplex_points(rng, 11, 3), None),
    ]
    for nm, el, order, pts, ent in wrappers:
        write_case(nm, el, order, pts, entity=ent)

    # ---- round 2: appended with their own generator so that every earlier file still reproduces bit for bit ----
    rng2 = numpy.random.default_rng(20261019)
    # (a) 3-D split complexes with an adversarial generator: Alfeld (4 subcells), Worsey-Farin (12), Powell-Sabin (24)
    for nm, el, order in (("p2_alfeld_tet_adv", FIAT.Lagrange(T3, 2, variant="alfeld"), 2),
                          ("p2_wf_tet_adv", FIAT.Lagrange(T3, 2, variant="worsey-farin"), 2),
                          ("ch_wf_tet_adv", FIAT.ChristiansenHu(T3, 1), 1),
                          ("alfeld_sorokina_tet_adv", FIAT.AlfeldSorokina(T3, 2), 2),
                          ("p1_ps_tet_adv", FIAT.Lagrange(T3, 1, variant="powell-sabin"), 1),
                          ("gn_tet_adv", FIAT.GuzmanNeilanFirstKindH1(T3, 1), 2),
                          ("walkington_tet_adv", FIAT.Walkington(T3), 2)):
        complex_ = el.get_nodal_basis().get_expansion_set().ref_el
        pts = adversarial_points(complex_, rng2, per_facet=3, n_random=200)
        stride = max(1, len(pts) // 160)
        write_case(f"{nm}_o{order}", el, order, pts, with_cells=True, table_stride=stride)
        write_case(f"{nm}_o0", el, 0, pts, with_cells=True, table_stride=stride)
    # (b) >= 10^4 adversarial points in 2-D for the BASELINE macro elements
    for nm, el in (("hct", FIAT.HsiehCloughTocher(T2)), ("ps6", FIAT.QuadraticPowellSabin6(T2)),
                   ("ps12", FIAT.QuadraticPowellSabin12(T2))):
        complex_ = el.get_nodal_basis().get_expansion_set().ref_el
        nfac = len(complex_.get_interior_facets(1))
        pts = adversarial_points(complex_, rng2, per_facet=-(-9000 // (17 * nfac)), n_random=1500)
        assert len(pts) >= 10000, len(pts)
        write_case(f"{nm}_adv10k_o2", el, 2, pts, with_cells=True, table_stride=16)
        write_case(f"{nm}_adv10k_o0", el, 0, pts, with_cells=True, table_stride=16)
    # (c) macro elements on a facet entity (points of a parent edge, incl. the split points of Powell-Sabin)
    edge_pts = numpy.concatenate([rng2.random((21, 1)), [[0.0], [1.0], [0.5], [0.5 + 1e-13], [0.5 - 1e-11], [0.25], [1.0 / 3.0]]])
    write_case("hct_edge2_o2", FIAT.HsiehCloughTocher(T2), 2, edge_pts, entity=(1, 2), with_cells=True)
    write_case("ps12_edge0_o2", FIAT.QuadraticPowellSabin12(T2), 2, edge_pts, entity=(1, 0), with_cells=True)
    write_case("ps6_edge1_o0", FIAT.QuadraticPowellSabin6(T2), 0, edge_pts, entity=(1, 1), with_cells=True)
    face_pts = numpy.concatenate([simplex_points(rng2, 17, 2), [[1.0 / 3.0, 1.0 / 3.0], [0.5, 0.5], [0.0, 0.0], [0.25, 0.25]]])
    write_case("p2_wf_tet_face1_o2", FIAT.Lagrange(T3, 2, variant="worsey-farin"), 2, face_pts, entity=(2, 1), with_cells=True)
    # (d) the highest degrees the reference's own tests construct (test_gauss_lobatto_legendre.py:111-138,
    # test_hct.py:79-82 go further on quadrature only)
    write_case("p12_tri_o2", FIAT.Lagrange(T2, 12), 2, simplex_points(rng2, 14, 2))
    write_case("p12_spectral_tri_o2", FIAT.Lagrange(T2, 12, variant="spectral"), 2, simplex_points(rng2, 14, 2))
    write_case("p10_spectral_tet_o2", FIAT.Lagrange(T3, 10, variant="spectral"), 2, simplex_points(rng2, 10, 3))
    write_case("p8_spectral_tet_o2", FIAT.Lagrange(T3, 8, variant="spectral"), 2, simplex_points(rng2, 14, 3))
    write_case("gll16_line_o3", FIAT.GaussLobattoLegendre(T1, 16), 3,
               numpy.concatenate([rng2.random((20, 1)), [[0.0], [1.0], [0.5]]]))
    for deg in (5, 6):
        el = FIAT.HsiehCloughTocher(T2, deg)
        complex_ = el.get_nodal_basis().get_expansion_set().ref_el
        pts = numpy.concatenate([simplex_points(rng2, 24, 2), adversarial_triangle_points(complex_)[:40]])
        write_case(f"hct{deg}_tri_o2", el, 2, pts, with_cells=True)

    # (e) elements that are not polynomial tabulations: HDivTrace (values on facets only, TraceError objects in the
    # derivative slots, geometric facet identification for entity=None) and QuadratureElement (identity at its points)
    from FIAT.reference_element import UFCQuadrilateral as UQ, TensorProductCell
    tr2, tr3 = FIAT.HDivTrace(T2, 2), FIAT.HDivTrace(T3, 1)
    write_case("hdivtrace2_tri_facet1_o0", tr2, 0, rng2.random((9, 1)), entity=(1, 1))
    write_case("hdivtrace2_tri_facet2_o1", tr2, 1, rng2.random((9, 1)), entity=(1, 2))
    write_case("hdivtrace2_tri_cell_o0", tr2, 0, simplex_points(rng2, 5, 2), entity=(2, 0))      # not on facets
    tverts = numpy.array(T3.get_vertices(), dtype=float)
    on_facets = []
    for f in range(4):
        fv = numpy.delete(tverts, f, axis=0)
        on_facets += list(rng2.dirichlet(numpy.ones(3), size=4) @ fv)
    write_case("hdivtrace1_tet_none_o0", tr3, 0, numpy.array(on_facets)[rng2.permutation(16)])
    write_case("hdivtrace1_tet_none_o1", tr3, 1, numpy.array(on_facets))
    write_case("hdivtrace1_tet_interior_o0", tr3, 0, numpy.array(on_facets[:3] + [[0.2, 0.2, 0.2]]))   # -> NaN tables
    write_case("hdivtrace1_tet_face3_o0", tr3, 0, simplex_points(rng2, 7, 2), entity=(2, 3))
    write_case("hdivtrace0_line_none_o0", FIAT.HDivTrace(T1, 0), 0, numpy.array([[0.0], [1.0], [1.0]]))
    write_case("hdivtrace0_line_vertex1_o0", FIAT.HDivTrace(T1, 0), 0, numpy.zeros((2, 0)), entity=(0, 1))
    write_case("hdivtrace1_quad_edge3_o0", FIAT.HDivTrace(UQ(), 1), 0, rng2.random((6, 1)), entity=(1, 3))
    trp = FIAT.HDivTrace(TensorProductCell(T2, T1), (1, 2))
    write_case("hdivtrace_prism_side1_o0", trp, 0, rng2.random((6, 2)), entity=((1, 1), 1))
    write_case("hdivtrace_prism_top_o0", trp, 0, simplex_points(rng2, 6, 2), entity=((2, 0), 1))
    qpts = simplex_points(rng2, 6, 2)
    write_case("quadrature6_tri_o0", FIAT.QuadratureElement(T2, qpts), 0, qpts)

    # (round 2, found by running the reference's own unit tests over the drop-in) tensor products of tensor products
    # that are NOT flattened: the default entity and entity keys are nested tuples (tensor_product.py:234-250)
    rng3 = numpy.random.default_rng(20261020)
    nested = TPE(TPE(P1, DP1), P2)
    write_case("nested_tpe_o1", nested, 1, rng3.random((7, 3)))
    write_case("nested_tpe_o2", TPE(TPE(P2, P1), TPE(DP1, P2)), 2, rng3.random((7, 4)))
    write_case("nested_tpe_face_o1", nested, 1, rng3.random((7, 2)), entity=(((1, 0), 1), 1))
    write_case("nested_tpe_edge_o1", nested, 1, rng3.random((7, 1)), entity=(((0, 0), 1), 3))
    write_case("nested_prism_x_interval_o1", TPE(TPE(FIAT.Lagrange(T2, 2), P1), DP1), 1,
               numpy.concatenate([simplex_points(rng3, 7, 2), rng3.random((7, 2))], axis=1))

    # element descriptions alone, for bench.py and full-size GPU tests
    for nm, el in () if ONLY is not None else (("p8_tet", FIAT.Lagrange(T3, 8)), ("n2curl4_tet", FIAT.NedelecSecondKind(T3, 4)),
                   ("hct", FIAT.HsiehCloughTocher(T2)), ("ps6", FIAT.QuadraticPowellSabin6(T2)),
                   ("ps12", FIAT.QuadraticPowellSabin12(T2)), ("gll_q10_hex", hexa),
                   ("p3_tri", FIAT.Lagrange(T2, 3)), ("p8_spectral_tet", FIAT.Lagrange(T3, 8, variant="spectral"))):
        path = os.path.join(OUT, f"desc_{nm}.npz")
        description.save(path, describe_element(el))
        print(f"desc {nm:27s} {os.path.getsize(path) / 1024:8.1f} KiB")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--only":
        ONLY = set(sys.argv[2].split(","))
    main()
