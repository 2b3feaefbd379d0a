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


@pytest.mark.parametrize("name,expect", [("p8_tet_o2", "lattice"), ("p3_tri_o1", "lattice"), ("p1_tri_o2", "lattice"),
                                         ("p4_tet_face2_o2", "lattice"), ("p5_tet_o3", "simplex"),
                                         ("dg3_tri_o1", "lattice"), ("cr_tri_o1", "simplex"), ("n2curl4_tet_o1", "simplex"),
                                         ("hct_o2", "simplex"), ("gll_q10_hex_o1", "tensor")])
def test_kernel_selection_and_parity(name, expect, cuda_device):
    """Equispaced Lagrange elements take the product-form kernel (after its on-device self-check);
    everything else the general kernels.  Both must match the reference."""
    from fiat_b200.api import Tabulator, FORCE_GENERAL
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    assert tab.kernel_path(case["order"]) == expect
    _compare(case["desc"], tab.tabulate(case["order"], case["points"], case["entity"]), case["ref"])
    if expect == "lattice":
        _compare(case["desc"], tab.tabulate(case["order"], case["points"], case["entity"], flags=FORCE_GENERAL),
                 case["ref"])


def test_full_size_properties_p8(cuda_device):
    """Size-independent properties at bench scale: partition of unity (values sum to 1, every
    derivative of the sum vanishes) for 2^17 points on both kernel paths."""
    from conftest import load_desc
    from fiat_b200.api import Tabulator, FORCE_GENERAL
    desc = load_desc("p8_tet")
    tab = Tabulator(desc, cuda_device)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(5)
    u, _ = torch.sort(torch.rand((1 << 17, 3), generator=g, device=cuda_device, dtype=torch.float64), dim=1)
    pts = torch.diff(torch.cat([torch.zeros((1 << 17, 1), device=cuda_device, dtype=torch.float64), u], dim=1), dim=1)
    for flags in (0, FORCE_GENERAL):
        got = tab.tabulate(2, pts, flags=flags)
        for alpha, v in got.items():
            total = v.sum(dim=0)
            target = 1.0 if sum(alpha) == 0 else 0.0
            assert (total - target).abs().max().item() <= 1e-9 * max(1.0, v.abs().max().item())
    a, b = tab.tabulate(2, pts), tab.tabulate(2, pts, flags=FORCE_GENERAL)
    for alpha in a:
        assert (a[alpha] - b[alpha]).abs().max().item() <= 1e-12 * b[alpha].abs().max().item()
