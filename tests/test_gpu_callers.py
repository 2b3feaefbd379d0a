"""Callers either side of the hot path: the FInAT caller's re-shaped tables (rank 2) and the
reference's setup-path tabulations of an expansion set (rank 4), device results against the reference."""
import numpy
import pytest
import torch

from conftest import load_case

pytestmark = pytest.mark.gpu


def _reference():
    from oracle.make_ref import import_reference
    FIAT = import_reference()
    if FIAT is None:
        pytest.skip("oracle/_ref has not been materialised")
    return FIAT


@pytest.mark.parametrize("name,order", [("p3_tri_o1", 5), ("p1_tri_o2", 2), ("n2curl4_tet_o1", 1), ("dg3_tri_o1", 4),
                                        ("p4_line_o2", 5), ("bdm2_tet_o1", 3)])
def test_basis_evaluation_shapes_and_contracts(name, order, cuda_device):
    """finat/fiat_elements.py:60-123 restated with numpy on the oracle's tables: point axis dropped for |alpha| ==
    degree on a simplex (cell-wise constant), zero table above the degree, index + value + point shape below."""
    from fiat_b200.finat_adapter import basis_evaluation
    from oracle import fiat_oracle
    case = load_case(name)
    desc = case["desc"]
    pts = numpy.asarray(case["points"], dtype=float)[:12]
    want = fiat_oracle.tabulate(desc, order, pts)
    degree = int(desc["degree"])
    got = basis_evaluation(desc, order, pts, device=cuda_device, point_shape=(3, 4))
    assert list(got) == list(want)
    for alpha, table in want.items():
        g = got[alpha].cpu().numpy()
        if sum(alpha) == degree:
            assert numpy.allclose(table, table[..., 0, None], atol=1e-9 * max(abs(table).max(), 1.0))
            expect = table[..., 0]
        elif sum(alpha) > degree:
            assert abs(table).max() <= 1e-8          # numpy.allclose(table, 0.0), the reference's own check
            expect = numpy.zeros(table.shape[:-1])
        else:
            expect = table.reshape(table.shape[:-1] + (3, 4))
        assert g.shape == expect.shape, (alpha, g.shape, expect.shape)
        assert abs(g - expect).max() <= 1e-11 * max(abs(expect).max(), 1e-300)


@pytest.mark.parametrize("name", ["gll_q10_hex_o1", "q2_quad_o2", "p2xp1_prism_o1", "rt2xp1_prism_vector_o1"])
def test_factor_evaluations_merge_to_the_reference_table(name, cuda_device):
    """finat/tensor_product.py:98-144: factors stay unmultiplied; their numeric product is the reference's table
    with the factors' basis indices flattened dof-major."""
    from fiat_b200.finat_adapter import factor_evaluations, merge_evaluations
    case = load_case(name)
    factors = factor_evaluations(case["desc"], case["order"], case["points"], case["entity"], device=cuda_device)
    merged = merge_evaluations(factors, case["order"])
    assert [tuple(k) for k in merged] == [tuple(k) for k in case["ref"]]
    nbasis = len(factors)
    for alpha, ref in case["ref"].items():
        m = merged[alpha]
        flat = m.reshape((-1,) + tuple(m.shape[nbasis:])).cpu().numpy()
        assert flat.shape == ref.shape
        assert abs(flat - ref).max() <= 1e-12 * max(abs(ref).max(), 1e-300)


def _sets(FIAT):
    from FIAT import expansions, macro
    from FIAT.reference_element import ufc_simplex, UFCInterval
    T1, T2, T3 = UFCInterval(), ufc_simplex(2), ufc_simplex(3)
    return [
        ("tri_none", expansions.ExpansionSet(T2), 5),
        ("tet_bubble_c0", expansions.ExpansionSet(T3, variant="bubble"), 4),
        ("line_legendre", expansions.ExpansionSet(T1), 6),
        ("alfeld_tri_c0", expansions.ExpansionSet(macro.AlfeldSplit(T2), variant="bubble"), 3),
        ("ps12_dg", expansions.ExpansionSet(macro.PowellSabin12Split(T2)), 2),
        ("alfeld_tet_c0", expansions.ExpansionSet(macro.AlfeldSplit(T3), variant="bubble"), 3),
    ]


def _points(es, rng, n=60):
    sd = es.ref_el.get_spatial_dimension()
    if sd == 1:
        return rng.random((n, 1))
    u = numpy.sort(rng.random((n, sd)), axis=1)
    return numpy.diff(numpy.concatenate([numpy.zeros((n, 1)), u], axis=1), axis=1)


def test_expansion_set_tabulations_match_the_reference(cuda_device):
    """ExpansionSet.tabulate / tabulate_derivatives / tabulate_jet (FIAT/expansions.py:601-637) on the device."""
    from fiat_b200.setup_path import ExpansionTabulator
    FIAT = _reference()
    rng = numpy.random.default_rng(21)
    for label, es, n in _sets(FIAT):
        pts = _points(es, rng)
        dev = ExpansionTabulator(es, n, cuda_device)
        v = es.tabulate(n, pts)
        assert abs(dev.tabulate(pts).cpu().numpy() - v).max() <= 1e-12 * abs(v).max(), label
        assert dev.tabulate(pts[:0]).numel() == 0
        nested = es.tabulate_derivatives(n, pts)
        got = dev.tabulate_derivatives(pts, nested=True)
        scale = max(abs(numpy.array([[d for _, d in row] for row in nested])).max(), 1.0)
        for i in range(len(nested)):
            for j in range(len(nested[0])):
                assert abs(got[i][j][0] - nested[i][j][0]) <= 1e-12 * scale
                assert numpy.allclose(got[i][j][1], nested[i][j][1], rtol=0, atol=1e-12 * scale)
        for order in (1, 2):
            want = es.tabulate_jet(n, pts, order=order)
            jet = dev.tabulate_jet(pts, order=order)
            assert len(jet) == len(want)
            for w, g in zip(want, jet):
                assert tuple(g.shape) == w.shape, label
                assert abs(g.cpu().numpy() - w).max() <= 1e-12 * max(abs(w).max(), 1e-300), label


def test_expansion_set_jumps_match_the_reference(cuda_device):
    """ExpansionSet.tabulate_jumps (FIAT/expansions.py:532-575): derivative jumps across interior facets at points on
    them (the reference's C^k macro constructions integrate these), device binning and per-subcell tabulation."""
    from fiat_b200.setup_path import ExpansionTabulator
    FIAT = _reference()
    rng = numpy.random.default_rng(22)
    for label, es, n in _sets(FIAT):
        complex_ = es.ref_el
        if not complex_.is_macrocell():
            continue
        sd = complex_.get_spatial_dimension()
        top = complex_.get_topology()
        verts = numpy.array(complex_.get_vertices())
        pts = []
        for f in complex_.get_interior_facets(sd - 1):
            fv = verts[list(top[sd - 1][f])]
            pts += list(rng.dirichlet(numpy.ones(sd), size=5) @ fv)
        pts = numpy.array(pts + list(_points(es, rng, 7)))
        want = es.tabulate_jumps(n, pts, order=2)
        got = ExpansionTabulator(es, n, cuda_device).tabulate_jumps(pts, order=2)
        assert sorted(got) == sorted(want)
        for r in want:
            assert tuple(got[r].shape) == want[r].shape, (label, r)
            assert abs(got[r].cpu().numpy() - want[r]).max() <= 1e-11 * max(abs(want[r]).max(), 1.0), (label, r)


def test_dmats_match_the_reference(cuda_device):
    """ExpansionSet.get_dmats (FIAT/expansions.py:577-599) from the device tabulation at the same lattice."""
    from fiat_b200.setup_path import ExpansionTabulator
    FIAT = _reference()
    from FIAT import reference_element
    for label, es, n in _sets(FIAT)[:3]:
        D = es.ref_el.get_dimension()
        verts = es.ref_el.get_vertices_of_subcomplex(es.ref_el.get_topology()[D][0])
        lattice = numpy.array(reference_element.make_lattice(verts, n, variant="gl"))
        want = es.get_dmats(n)
        got = ExpansionTabulator(es, n, cuda_device).get_dmats(lattice).cpu().numpy()
        assert got.shape == want.shape
        assert abs(got - want).max() <= 1e-9 * max(abs(want).max(), 1.0), label


@pytest.mark.parametrize("name,mapping", [("n2curl4_tet_o1", "covariant piola"), ("rt3_tri_o1", "contravariant piola"),
                                          ("bdm2_tet_o1", "contravariant piola"), ("regge2_tet_o1", "double covariant piola"),
                                          ("hz3_tri_o1", "double contravariant piola"), ("aw_tri_o2", "covariant contravariant piola"),
                                          ("p3_tri_o1", "affine")])
def test_pullback_folded_into_the_coefficients(name, mapping, cuda_device):
    """Tabulator.mapped(): tables equal the reference's pullback(phi, mapping, J) (FIAT/macro.py:601-645) applied to
    the golden tabulation, for every mapping type the reference knows."""
    from fiat_b200.api import Tabulator
    FIAT = _reference()
    from FIAT.macro import pullback
    case = load_case(name)
    sd = int(case["desc"]["sd"])
    rng = numpy.random.default_rng(5)
    J = numpy.eye(sd) + 0.3 * rng.standard_normal((sd, sd))
    tab = Tabulator(case["desc"], cuda_device).mapped(mapping, J=J)
    got = tab.tabulate(case["order"], case["points"], case["entity"])
    for alpha, ref in case["ref"].items():
        want = pullback(ref, mapping, J=J)
        g = got[alpha].cpu().numpy()
        assert g.shape == want.shape
        assert abs(g - want).max() <= 1e-12 * max(abs(want).max(), 1e-300), (name, alpha)
    with pytest.raises(ValueError):
        Tabulator(case["desc"], cuda_device).mapped("no such piola", J=J)


def test_expansion_set_normal_jumps_match_the_reference(cuda_device):
    """ExpansionSet.tabulate_normal_jumps (FIAT/expansions.py:492-530): jumps of normal derivatives across a facet of
    the split complex (interior facets have a subcell on either side) at points given on the reference facet."""
    from fiat_b200.setup_path import ExpansionTabulator
    FIAT = _reference()
    rng = numpy.random.default_rng(23)
    for label, es, n in _sets(FIAT):
        sd = es.ref_el.get_spatial_dimension()
        if sd == 1 or not es.ref_el.is_macrocell():      # the reference needs the complex's cell connectivity
            continue
        dev = ExpansionTabulator(es, n, cuda_device)
        nfacets = len(es.ref_el.get_topology()[sd - 1])
        for facet in range(0, nfacets, max(1, nfacets // 6)):
            ref_pts = rng.random((9, 1)) if sd == 2 else _points(es, rng, 9)[:, :2]
            want = es.tabulate_normal_jumps(n, ref_pts, facet, order=2)
            got = dev.tabulate_normal_jumps(ref_pts, facet, order=2).cpu().numpy()
            assert got.shape == want.shape, (label, facet)
            assert abs(got - want).max() <= 1e-11 * max(abs(want).max(), 1.0), (label, facet)


def test_to_riesz_matches_the_reference(cuda_device):
    """DualSet.to_riesz (FIAT/dual_set.py:86-206): point evaluations, derivatives (Hermite, Argyris, Morley, Bell),
    integral moments against quadrature rules (RT, Nedelec, BDM, Regge, MTW), macro elements (HCT, Guzman-Neilan):
    the expansion-set tabulations run on the device, the result equals the reference's matrix."""
    _reference()
    import FIAT
    from fiat_b200.setup_path import to_riesz
    from test_setup_path_host import riesz_elements
    for element in riesz_elements(FIAT):
        poly_set, dual = element.get_nodal_basis(), element.dual
        want = dual.to_riesz(poly_set)
        got = to_riesz(dual, poly_set, device=cuda_device)
        assert got.is_cuda and tuple(got.shape) == want.shape
        assert abs(got.cpu().numpy() - want).max() <= 1e-12 * abs(want).max(), type(element).__name__
