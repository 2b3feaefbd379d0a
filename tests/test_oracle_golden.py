"""Pin the CPU oracle against outputs of the reference itself (tests/golden/case_*.npz)."""
import numpy
import pytest

from conftest import golden_case_names, is_polynomial_case, load_case, tolerance
from oracle import fiat_oracle


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_matches_reference(name):
    case = load_case(name)
    if not is_polynomial_case(case):
        pytest.skip("not a polynomial tabulation (pinned by the golden file itself)")
    got = fiat_oracle.tabulate(case["desc"], case["order"], case["points"], case["entity"])
    ref = case["ref"]
    assert list(got.keys()) == list(ref.keys())          # same keys in the same (mis) order
    for alpha, expect in ref.items():
        assert got[alpha].shape == expect.shape
        scale = max(abs(expect).max(), 1e-300) if expect.size else 1.0
        err = abs(got[alpha] - expect).max() if expect.size else 0.0
        assert err <= tolerance(case["desc"], alpha) * scale, (alpha, err / scale)


@pytest.mark.parametrize("name", [n for n in golden_case_names()])
def test_oracle_subcell_assignment_bit_exact(name):
    case = load_case(name)
    if "near_all" not in case:
        pytest.skip("single-cell element")
    # mask_points: all points of a large adversarial set (tables are stored for a subset), already on the cell
    desc, pts = case["desc"], case.get("mask_points", case["points"])
    assert numpy.array_equal(fiat_oracle.locate_cells(desc, pts, unique=False), case["near_all"])
    assert numpy.array_equal(fiat_oracle.locate_cells(desc, pts, unique=True), case["near_unique"])


@pytest.mark.parametrize("name,entity", [("p3_tri_o1", None), ("regge2_tet_o1", None), ("p2_tri_facet1_o1", (1, 1))])
def test_oracle_single_point_has_no_point_axis(name, entity):
    """test/FIAT/unit/test_fiat.py test_single_point_tabulation: a bare coordinate tuple is one point and the
    tables have shape (ndofs,) + value_shape."""
    case = load_case(name)
    p = tuple(float(x) for x in numpy.asarray(case["points"])[3])
    one = fiat_oracle.tabulate(case["desc"], 1, p, entity)
    batched = fiat_oracle.tabulate(case["desc"], 1, [p], entity)
    for alpha, v in one.items():
        assert v.shape == batched[alpha].shape[:-1]
        assert numpy.array_equal(v, batched[alpha][..., 0])
        assert numpy.allclose(v, case["ref"][alpha][..., 3], rtol=0, atol=1e-12 * max(abs(case["ref"][alpha]).max(), 1.0))
