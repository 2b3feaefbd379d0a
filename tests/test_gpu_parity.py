"""Parity of the CUDA path (through the C ABI) against the reference's golden outputs and the
CPU oracle, on the same points.  Tolerance: north_star's 1e-12 * max|ref| per derivative component
(1e-10 for order >= 2 at degree >= 8), see conftest.tolerance."""
import numpy
import pytest
import torch

from conftest import golden_case_names, load_case, tolerance
from oracle import fiat_oracle

pytestmark = pytest.mark.gpu


def _compare(desc, got, ref):
    assert [tuple(k) for k in got.keys()] == [tuple(k) for k in ref.keys()]
    for alpha, expect in ref.items():
        g = got[alpha]
        g = g.cpu().numpy() if isinstance(g, torch.Tensor) else g
        assert g.shape == expect.shape and g.dtype == numpy.float64
        if expect.size == 0:
            continue
        scale = max(abs(expect).max(), 1e-300)
        err = abs(g - expect).max()
        assert err <= tolerance(desc, alpha) * scale, (alpha, err / scale)


@pytest.mark.parametrize("name", golden_case_names())
def test_matches_reference_golden(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    got = tab.tabulate(case["order"], case["points"], case["entity"])
    _compare(case["desc"], got, case["ref"])
    for v in got.values():
        assert v.is_cuda and v.dtype == torch.float64


@pytest.mark.parametrize("name", golden_case_names())
def test_both_kernels_agree_with_oracle(name, cuda_device):
    """Single-cell Dubiner elements run on the thread-per-point kernel and on the DMMA tile kernel."""
    from fiat_b200.api import Tabulator, FORCE_THREAD_PER_POINT, FORCE_DMMA
    case = load_case(name)
    desc = case["desc"]
    if desc["kind"] != "simplex" or desc["expansion"] != "dubiner" or int(desc["ncells"]) != 1 or case["order"] > 2:
        pytest.skip("DMMA kernel not applicable")
    tab = Tabulator(desc, cuda_device)
    want = fiat_oracle.tabulate(desc, case["order"], case["points"], case["entity"])
    for flags in (FORCE_THREAD_PER_POINT, FORCE_DMMA):
        got = tab.tabulate(case["order"], case["points"], case["entity"], flags=flags)
        _compare(desc, got, want)


@pytest.mark.parametrize("name", [n for n in golden_case_names()])
def test_subcell_assignment_bit_exact(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    if "near_all" not in case:
        pytest.skip("single-cell element")
    tab = Tabulator(case["desc"], cuda_device)
    for unique, key in ((False, "near_all"), (True, "near_unique")):
        mask = tab.locate_subcells(case["points"], unique).cpu().numpy().astype(numpy.int64)
        near = case[key]
        want = sum((near[c].astype(numpy.int64) << c) for c in range(near.shape[0]))
        assert numpy.array_equal(mask, want)


@pytest.mark.parametrize("name", ["p3_tri_o1", "hct_o2", "gll_q3_hex_face4_o2", "n2curl4_tet_o1"])
def test_host_buffer_call(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    got = tab.tabulate_host(case["order"], case["points"], case["entity"], chunk_pts=16)
    _compare(case["desc"], got, case["ref"])


def test_empty_point_set(cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case("p3_tri_o1")
    got = Tabulator(case["desc"], cuda_device).tabulate(1, numpy.zeros((0, 2)))
    assert [v.shape for v in got.values()] == [(10, 0)] * 3


def test_tabulate_into_streaming(cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case("p3_tri_o1")
    tab = Tabulator(case["desc"], cuda_device)
    pts = torch.as_tensor(case["points"], device=cuda_device)
    buf = torch.full((3, 10, 256), float("nan"), dtype=torch.float64, device=cuda_device)
    n = tab.tabulate_into(buf, 1, pts)
    assert n == 200
    for j, alpha in enumerate(case["ref"]):
        ref = case["ref"][alpha]
        assert abs(buf[j, :, :200].cpu().numpy() - ref).max() <= 1e-12 * abs(ref).max()
    assert torch.isnan(buf[:, :, 200:]).all()
