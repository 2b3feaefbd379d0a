"""Parity of the CUDA path (through the C ABI) against the reference's golden outputs and the
CPU oracle, on the same points.  Tolerance: north_star's 1e-12 * max|ref| per derivative component
(1e-10 for order >= 2 at degree >= 8), see conftest.tolerance."""
import numpy
import pytest
import torch

from conftest import golden_case_names, load_case, tolerance
from oracle import fiat_oracle

pytestmark = pytest.mark.gpu


def _compare(desc, got, ref, error_keys=()):
    # trace elements: slots that are not defined hold exception objects, in the reference and here
    assert [tuple(k) for k, v in got.items() if isinstance(v, Exception)] == list(error_keys)
    got = {k: v for k, v in got.items() if not isinstance(v, Exception)}
    assert [tuple(k) for k in got.keys()] == [tuple(k) for k in ref.keys()]
    for alpha, expect in ref.items():
        g = got[alpha]
        g = g.cpu().numpy() if isinstance(g, torch.Tensor) else g
        assert g.shape == expect.shape and g.dtype == numpy.float64
        if expect.size == 0:
            continue
        assert numpy.array_equal(numpy.isnan(g), numpy.isnan(expect))       # NaN tables of a failed trace tabulation
        g, expect = numpy.nan_to_num(g), numpy.nan_to_num(expect)
        scale = max(abs(expect).max(), 1e-300)
        err = abs(g - expect).max()
        assert err <= tolerance(desc, alpha) * scale, (alpha, err / scale)


@pytest.mark.parametrize("name", golden_case_names())
def test_matches_reference_golden(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    got = tab.tabulate(case["order"], case["points"], case["entity"])
    _compare(case["desc"], got, case["ref"], case["error_keys"])
    for v in got.values():
        assert isinstance(v, Exception) or (v.is_cuda and v.dtype == torch.float64)


@pytest.mark.parametrize("name", [n for n in golden_case_names() if "hdivtrace" in n or "quadrature" in n])
def test_trace_and_quadrature_elements_with_host_buffers(name, cuda_device):
    """HDivTrace / QuadratureElement through the numpy-in / numpy-out call (what the reference-side binding uses)."""
    from fiat_b200.api import Tabulator
    case = load_case(name)
    got = Tabulator(case["desc"], cuda_device).tabulate_host(case["order"], case["points"], case["entity"])
    _compare(case["desc"], got, case["ref"], case["error_keys"])
    assert all(isinstance(v, (Exception, numpy.ndarray)) for v in got.values())


@pytest.mark.parametrize("name", golden_case_names())
def test_both_kernels_agree_with_oracle(name, cuda_device):
    """Single-cell Dubiner elements run on the thread-per-point kernel and on the DMMA tile kernel."""
    from fiat_b200.api import Tabulator, FORCE_THREAD_PER_POINT, FORCE_DMMA
    case = load_case(name)
    desc = case["desc"]
    if desc["kind"] != "simplex" or desc["expansion"] != "dubiner" or int(desc["ncells"]) != 1 or case["order"] > 3:
        pytest.skip("DMMA kernel not applicable")
    tab = Tabulator(desc, cuda_device)
    want = fiat_oracle.tabulate(desc, case["order"], case["points"], case["entity"])
    for flags in (FORCE_THREAD_PER_POINT, FORCE_DMMA):
        got = tab.tabulate(case["order"], case["points"], case["entity"], flags=flags)
        _compare(desc, got, want)


@pytest.mark.parametrize("name", golden_case_names())
def test_value_table_and_jet_kernels_agree_with_golden(name, cuda_device):
    """Low-degree (incl. split-cell) Dubiner elements: the value-table kernel (derivative-folded coefficients)
    and the kernels that propagate derivative jets through the recurrence must both match the reference."""
    from fiat_b200.api import Tabulator, FORCE_GENERAL, NO_VALUE_TABLE
    case = load_case(name)
    desc = case["desc"]
    if desc["kind"] != "simplex" or desc["expansion"] != "dubiner" or case["order"] > 3:
        pytest.skip("not a Dubiner simplex element")
    tab = Tabulator(desc, cuda_device)
    for flags in (FORCE_GENERAL, FORCE_GENERAL | NO_VALUE_TABLE):
        got = tab.tabulate(case["order"], case["points"], case["entity"], flags=flags)
        _compare(desc, got, case["ref"])


@pytest.mark.parametrize("name", ["n2curl4_tet_o1", "ned2_tet_o2", "p5_tet_o3", "hermite3_tet_o2", "p6_tri_o4",
                                  "p4_tet_face2_o2", "regge2_tet_o1", "argyris_tri_o2"])
def test_alpha_split_matches_golden_and_jets(name, cuda_device):
    """Mid-size single-cell elements are tabulated as one derived order-0 element per derivative multi-index
    (plan.alpha_split); the one-pass jet tabulation must give the same tables."""
    from fiat_b200.api import Tabulator, FORCE_GENERAL, NO_ALPHA_SPLIT, NO_MERGED_SPLIT
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    split = tab.tabulate(case["order"], case["points"], case["entity"], flags=FORCE_GENERAL)
    _compare(case["desc"], split, case["ref"])
    separate = tab.tabulate(case["order"], case["points"], case["entity"], flags=FORCE_GENERAL | NO_MERGED_SPLIT)
    _compare(case["desc"], separate, case["ref"])
    whole = tab.tabulate(case["order"], case["points"], case["entity"], flags=FORCE_GENERAL | NO_ALPHA_SPLIT)
    _compare(case["desc"], whole, case["ref"])
    assert len(tab._resolve(case["order"], case["entity"], FORCE_GENERAL | NO_ALPHA_SPLIT)[0]) == 1


@pytest.mark.parametrize("name", ["gn_tet_o2", "walkington_tet_o2"])
def test_split_cell_tile_kernel_matches_golden(name, cuda_device):
    """Split-cell elements too large for the value-table kernel: points binned by subcell, DMMA contraction with
    per-subcell matrices (cells.cuh) against the thread-per-point kernel and the reference, including the fixture's
    points on interior facets (several subcells per point)."""
    from fiat_b200.api import Tabulator, NO_MACRO_MERGED
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    assert tab.kernel_names(case["order"], case["entity"]) == ["mma_cells"]
    _compare(case["desc"], tab.tabulate(case["order"], case["points"], case["entity"]), case["ref"])
    assert tab.kernel_names(case["order"], case["entity"], NO_MACRO_MERGED) == ["cellwise"]
    _compare(case["desc"], tab.tabulate(case["order"], case["points"], case["entity"], flags=NO_MACRO_MERGED), case["ref"])
    # order 0 keeps the reference's first-match binning (exterior points near several subcells)
    ext = numpy.asarray(case["points"], dtype=float) + 0.3
    for pts in (ext, numpy.asarray(case["points"], dtype=float)):
        _compare(case["desc"], tab.tabulate(0, pts), fiat_oracle.tabulate(case["desc"], 0, pts))
        _compare(case["desc"], tab.tabulate(case["order"], pts), fiat_oracle.tabulate(case["desc"], case["order"], pts))
    # uniform points at a size with partial tiles, against the oracle
    rng = numpy.random.default_rng(3)
    lam = numpy.diff(numpy.concatenate([numpy.zeros((1000, 1)), numpy.sort(rng.random((1000, 3)), axis=1)], axis=1), axis=1)
    want = fiat_oracle.tabulate(case["desc"], case["order"], lam)
    _compare(case["desc"], tab.tabulate(case["order"], lam), want)


@pytest.mark.parametrize("name", ["gll_q10_hex_o1", "gll_q3_hex_face4_o2", "q2_quad_o2", "p2xp1_prism_o1", "dq32_quad_o2"])
def test_fused_evaluation_on_tensor_products(name, cuda_device):
    """evaluate() on scalar tensor-product elements (nested sums over the factors, no table) against
    coefficients . reference table."""
    from fiat_b200.api import Tabulator
    case = load_case(name)
    ref0 = next(iter(case["ref"].values()))
    rng = numpy.random.default_rng(11)
    u = rng.standard_normal((3, ref0.shape[0]))
    got = Tabulator(case["desc"], cuda_device).evaluate(u, case["order"], case["points"], case["entity"])
    assert [tuple(k) for k in got.keys()] == [tuple(k) for k in case["ref"].keys()]
    for alpha, ref in case["ref"].items():
        want = u @ ref
        bound = (abs(u) @ abs(ref)).max()
        assert got[alpha].shape == want.shape
        assert abs(got[alpha].cpu().numpy() - want).max() <= 1e-12 * max(bound, 1e-300), alpha


@pytest.mark.parametrize("name", [n for n in golden_case_names()])
def test_subcell_assignment_bit_exact(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    if "near_all" not in case:
        pytest.skip("single-cell element")
    tab = Tabulator(case["desc"], cuda_device)
    pts = case.get("mask_points", case["points"])      # all points of a large adversarial set, already on the cell
    for unique, key in ((False, "near_all"), (True, "near_unique")):
        mask = tab.locate_subcells(pts, unique).cpu().numpy().astype(numpy.int64)
        near = case[key]
        want = sum((near[c].astype(numpy.int64) << c) for c in range(near.shape[0]))
        assert numpy.array_equal(mask, want)


@pytest.mark.parametrize("name", ["p3_tri_o1", "hct_o2", "gll_q3_hex_face4_o2", "n2curl4_tet_o1", "rtcf1_quad_o1",
                                  "mini_tri_o2", "n2curl3_p3_mixed_tet_o1", "enriched_p4s_bubble5_tet_o2", "p6_tri_o4"])
def test_host_buffer_call(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    got = tab.tabulate_host(case["order"], case["points"], case["entity"], chunk_pts=16)
    _compare(case["desc"], got, case["ref"])


def test_point_containers_and_error_contract(cuda_device):
    """The reference accepts any iterable of coordinate tuples or an ndarray (float32 is promoted), returns zero
    tables for derivative orders above the degree, and raises for bad orders / entities."""
    from fiat_b200.api import Tabulator
    case = load_case("p3_tri_o1")
    tab = Tabulator(case["desc"], cuda_device)
    pts = numpy.asarray(case["points"], dtype=float)[:16]
    base = tab.tabulate(1, pts)
    as_list = tab.tabulate(1, [tuple(p) for p in pts])
    strided = torch.as_tensor(numpy.concatenate([pts, pts], axis=1), device=cuda_device)[:, :2]     # row stride 4
    from_strided = tab.tabulate(1, strided)
    pts32 = pts.astype(numpy.float32)
    from32 = tab.tabulate(1, pts32)
    want32 = fiat_oracle.tabulate(case["desc"], 1, pts32.astype(numpy.float64))
    for alpha in base:
        assert torch.equal(base[alpha], as_list[alpha]) and torch.equal(base[alpha], from_strided[alpha])
    _compare(case["desc"], from32, want32)
    high = tab.tabulate(5, pts)                      # cubic element: orders 4 and 5 vanish identically
    assert len(high) == 21
    for alpha, v in high.items():
        if sum(alpha) > 3:
            assert not v.any()
    _compare(case["desc"], {a: v for a, v in high.items() if sum(a) <= 1}, {a: case["ref"][a][:, :16] for a in base})
    with pytest.raises(ValueError):
        tab.tabulate(-1, pts)
    with pytest.raises(KeyError):
        tab.tabulate(1, pts[:, :1], entity=(1, 7))
    with pytest.raises(NotImplementedError):
        tab.tabulate(1, numpy.array([[object(), object()]], dtype=object))


@pytest.mark.parametrize("name,entity", [("p3_tri_o1", None), ("regge2_tet_o1", None), ("p2_tri_facet1_o1", (1, 1)),
                                         ("hct_o2", None)])
def test_single_point_without_point_axis(name, entity, cuda_device):
    """A bare coordinate tuple of shape (sd,) is ONE point and the tables lose their point axis, like the reference's
    (test/FIAT/unit/test_fiat.py test_single_point_tabulation, test_regge_hhj.py); found by running the reference's
    unit tests over the drop-in (tests/test_reference_suite.py)."""
    from fiat_b200.api import Tabulator
    case = load_case(name)
    tab = Tabulator(case["desc"], cuda_device)
    p = tuple(float(x) for x in numpy.asarray(case["points"])[3])
    batched = tab.tabulate(1, [p], entity)
    want = fiat_oracle.tabulate(case["desc"], 1, p, entity)
    for got in (tab.tabulate(1, p, entity), tab.tabulate(1, numpy.array(p), entity), tab.tabulate_host(1, p, entity)):
        assert list(got) == list(batched)
        for alpha, v in got.items():
            assert tuple(v.shape) == tuple(batched[alpha].shape[:-1]) == want[alpha].shape
            assert numpy.array_equal(numpy.asarray(v.cpu() if isinstance(v, torch.Tensor) else v),
                                     batched[alpha][..., 0].cpu().numpy())
    _compare(case["desc"], tab.tabulate(1, p, entity), want)


@pytest.mark.parametrize("name", ["p8_tet_o2", "n2curl4_tet_o1", "hct_o2", "gn_tet_o2", "gll_q3_hex_face4_o2", "rtcf1_quad_o1",
                                  "nested_tpe_o1", "p2_tri_facet1_o1"])
def test_quick_plan_for_small_calls(name, cuda_device, monkeypatch):
    """Small calls (<= api.QUICK_NPTS points) before any large one run on a quick plan -- thread-per-point kernels on a
    description marked dense_only, no clustering / packing / derived elements / self-checks -- and match the
    reference; a large call then builds the streaming plan, which small calls use from then on."""
    from fiat_b200 import api
    monkeypatch.setattr(api, "QUICK_NPTS", 4096)
    case = load_case(name)
    tab = api.Tabulator(case["desc"], cuda_device)
    got = tab.tabulate(case["order"], case["points"], case["entity"])
    assert tab._quick is not None and not any(k[0] == "resolved" for k in tab._plans)
    _compare(case["desc"], got, case["ref"])
    _compare(case["desc"], tab.tabulate_host(case["order"], case["points"], case["entity"]), case["ref"])
    pts = numpy.asarray(case["points"], dtype=float)
    big = numpy.tile(pts, (-(-5000 // max(len(pts), 1)), 1))
    wide = tab.tabulate(case["order"], big, case["entity"])
    assert any(k[0] == "resolved" for k in tab._plans)
    again = tab.tabulate(case["order"], case["points"], case["entity"])
    for alpha in got:
        scale = max(float(wide[alpha].abs().max()), 1e-300)
        assert float((again[alpha] - wide[alpha][..., :len(pts)]).abs().max()) <= 1e-14 * scale
    _compare(case["desc"], again, case["ref"])


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


@pytest.mark.parametrize("name", ["hct_o2", "ps12_o2", "n2curl4_tet_o1", "p8_tet_o2", "regge2_tet_o1", "gn_tet_o2", "walkington_tet_o2",
                                  "gll_q3_hex_face4_o2", "p5_tet_o3", "mini_tri_o2", "n2curl3_p3_mixed_tet_o1",
                                  "p4_line_o2", "argyris_tri_o2"])
@pytest.mark.parametrize("npts", [1, 37, 203])
def test_no_write_outside_the_table(name, npts, cuda_device):
    """Guard bands around and inside the caller's buffer (row stride > npts, odd point counts that leave partial
    warps, octets and vector stores): every kernel path must leave them untouched and fill exactly its table."""
    from fiat_b200.api import Tabulator, FORCE_GENERAL
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    pts0 = numpy.asarray(case["points"], dtype=float)
    pts = pts0[numpy.arange(npts) % len(pts0)]
    want = fiat_oracle.tabulate(desc, order, pts, case["entity"])
    na = len(want)
    nrows = int(numpy.prod(next(iter(want.values())).shape[:-1]))
    stride, pad = 256, 64
    for flags in (0, FORCE_GENERAL):
        tab = Tabulator(desc, cuda_device)
        raw = torch.full((pad + na * nrows * stride + pad,), float("nan"), dtype=torch.float64, device=cuda_device)
        buf = raw[pad:pad + na * nrows * stride].view(na, nrows, stride)
        assert tab.tabulate_into(buf, order, torch.as_tensor(pts, device=cuda_device), case["entity"], flags=flags) == npts
        assert torch.isnan(raw[:pad]).all() and torch.isnan(raw[-pad:]).all()
        assert torch.isnan(buf[:, :, npts:]).all()
        got = {alpha: buf[j, :, :npts].reshape(want[alpha].shape) for j, alpha in enumerate(want)}
        _compare(desc, got, want)


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


def _device_points(kind, n, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd = int(kind[-1])
    u = torch.rand((n, sd), generator=g, device=device, dtype=torch.float64)
    if kind.startswith("cube"):
        return u
    u, _ = torch.sort(u, dim=1)
    return torch.diff(torch.cat([torch.zeros((n, 1), device=device, dtype=torch.float64), u], dim=1), dim=1).contiguous()


@pytest.mark.parametrize("name,order,kind,npts", [("gll_q10_hex", 1, "cube3", 1 << 14), ("p3_tri", 1, "simplex2", 1 << 18)])
def test_full_size_partition_of_unity(name, order, kind, npts, cuda_device):
    """Nodal Lagrange-type bases sum to one and every derivative of the sum vanishes
    (size-independent property, checked at bench-like point counts)."""
    from conftest import load_desc
    from fiat_b200.api import Tabulator
    tab = Tabulator(load_desc(name), cuda_device)
    got = tab.tabulate(order, _device_points(kind, npts, cuda_device, 11))
    for alpha, v in got.items():
        total = v.sum(dim=0)
        target = 1.0 if sum(alpha) == 0 else 0.0
        assert (total - target).abs().max().item() <= 1e-10 * max(1.0, v.abs().max().item())


@pytest.mark.parametrize("name", ["hct", "ps6", "ps12"])
def test_full_size_macro_vertex_functions_sum_to_one(name, cuda_device):
    """test_hct.py:33-52 / test_powell_sabin.py:14-32 at 2^18 points: the three vertex-value basis
    functions sum to 1 with vanishing derivatives; also every point is binned to >= 1 subcell and the
    same subcell mask comes back from the locate entry point as the tabulation used."""
    from conftest import load_desc
    from fiat_b200.api import Tabulator
    desc = load_desc(name)
    tab = Tabulator(desc, cuda_device)
    pts = _device_points("simplex2", 1 << 18, cuda_device, 13)
    got = tab.tabulate(2, pts)
    ndofs = next(iter(got.values())).shape[0]
    # vertex point-evaluation dofs come first for each vertex: (value, d/dx, d/dy) triples
    stride = ndofs // 3 if name != "ps12" else 3
    vertex_value_dofs = [0, 3, 6]
    for alpha, v in got.items():
        total = v[vertex_value_dofs].sum(dim=0)
        target = 1.0 if sum(alpha) == 0 else 0.0
        assert (total - target).abs().max().item() <= 1e-9 * max(1.0, v.abs().max().item()), (alpha, stride)
    mask = tab.locate_subcells(pts, unique=False)
    assert int((mask == 0).sum().item()) == 0
    assert int((mask >= (1 << int(desc["ncells"]))).sum().item()) == 0


def test_nodality_of_tensor_product_dof_order(cuda_device):
    """test_fiat.py:560-581: tabulating a nodal tensor-product element at its own nodes gives the
    identity -- pins the (iA * nB + iB) dof ordering.  The GLL Q10 hex nodes are the products of the
    1-D node table stored in the description."""
    from conftest import load_desc
    from fiat_b200.api import Tabulator
    desc = load_desc("gll_q10_hex")
    leaf = desc["element"]["B"]                       # the innermost 1-D GLL factor
    x = numpy.sort(numpy.asarray(leaf["ll_nodes"][0]))
    # 1-D dof order of the factor: ask the factor itself
    one_d = Tabulator(leaf, cuda_device).tabulate(0, x[:, None])[(0,)].cpu().numpy()
    order_1d = one_d.argmax(axis=0)                   # dof index that is 1 at node j
    assert numpy.allclose(one_d[order_1d, numpy.arange(len(x))], 1.0)
    node_of_dof = numpy.empty(len(x), dtype=int)
    node_of_dof[order_1d] = numpy.arange(len(x))
    xs = x[node_of_dof]
    n = len(xs)
    pts = numpy.stack(numpy.meshgrid(xs, xs, xs, indexing="ij"), axis=-1).reshape(-1, 3)
    vals = Tabulator(desc, cuda_device).tabulate(0, pts)[(0, 0, 0)]
    eye = torch.eye(n ** 3, dtype=torch.float64, device=cuda_device)
    assert (vals - eye).abs().max().item() <= 1e-12


def _check_evaluation(desc, got, u, ref_tables):
    ndofs = u.shape[1]
    for alpha, ref in ref_tables.items():
        want = numpy.tensordot(u, ref, axes=(1, 0))
        g = got[alpha]
        g = g.cpu().numpy() if isinstance(g, torch.Tensor) else g
        assert g.shape == want.shape, (alpha, g.shape, want.shape)
        bound = (abs(u)[:, :, None] * abs(ref.reshape(ndofs, -1))[None]).sum(axis=1).max()
        assert abs(g - want).max() <= tolerance(desc, alpha) * max(bound, 1e-300), alpha


@pytest.mark.parametrize("name", ["p8_tet_o2", "hct_o2", "hct_o0", "ps12_o2", "n2curl4_tet_o1", "p2_tri_facet1_o1", "gn_tet_o2",
                                  "walkington_tet_o2", "p2_wf_tet_adv_o2", "hct_edge2_o2", "regge2_tet_o1", "p6_tri_o4",
                                  "p4_line_o2", "gll7_line_o3", "legendre5_line_o3", "dp0_tet_o1", "p10_spectral_tet_o2",
                                  # wrapper elements: parts' dof slices, component placement and signs
                                  "mini_tri_o2", "taylor_hood_tri_o1", "rt1_dg0_mixed_tri_o1", "rtcf1_quad_o1", "rtce2_quad_o2",
                                  "nce1_hex_o1", "rt2xp1_prism_vector_o1", "dp1xp2_prism_hdiv_o2", "n2curl3_p3_mixed_tet_o1",
                                  # scalar tensor products
                                  "gll_q3_hex_face4_o2", "q2_quad_o2", "p2xp1_prism_o1"])
def test_fused_point_evaluation(name, cuda_device):
    """evaluate(): sum_i c[f, i] D^alpha phi_i without materialising the tables, against coefficients . reference
    table (incl. the fixtures' points on interior facets); a second coefficient set reuses every plan (the weights
    are formed on the device), coefficients may already live on the device, and the host-buffer form agrees."""
    from fiat_b200 import plan as planmod
    from fiat_b200.api import Tabulator
    case = load_case(name)
    desc = case["desc"]
    rng = numpy.random.default_rng(3)
    ndofs = planmod.num_dofs_of(desc)
    tab = Tabulator(desc, cuda_device)
    u = rng.standard_normal((3, ndofs))
    _check_evaluation(desc, tab.evaluate(u, case["order"], case["points"], case["entity"]), u, case["ref"])
    plans = len(tab._plans)
    u2 = rng.standard_normal((5, ndofs))
    _check_evaluation(desc, tab.evaluate(torch.as_tensor(u2, device=cuda_device), case["order"], case["points"], case["entity"]),
                      u2, case["ref"])
    if desc["kind"] == "simplex" and desc["expansion"] == "dubiner" and int(desc["degree"]) >= 1:
        assert len(tab._plans) == plans          # new coefficients: no new plan
    u1 = rng.standard_normal(ndofs)              # a single function as a 1-D vector
    _check_evaluation(desc, tab.evaluate(u1, case["order"], case["points"], case["entity"]), u1[None, :], case["ref"])
    host = tab.evaluate_host(u, case["order"], case["points"], case["entity"], chunk_pts=8)
    assert all(isinstance(v, numpy.ndarray) for v in host.values())
    _check_evaluation(desc, host, u, case["ref"])


def test_fused_point_evaluation_at_scale(cuda_device):
    """2^16 bench points of P8 (tile kernel with device-built weight fragments) and HCT (thread per point with
    per-subcell weights) against coefficients . tabulate()."""
    import bench
    from fiat_b200.api import Tabulator
    for workload in ("p8_tet_o2", "hct_o2", "n2curl4_tet_o1"):
        dname, order, kind, _ = bench.WORKLOADS[workload]
        desc = bench.load_desc(dname)
        tab = Tabulator(desc, cuda_device)
        pts = bench.device_points(kind, 1 << 16, 99, cuda_device)
        u = torch.as_tensor(numpy.random.default_rng(5).standard_normal((2, desc["coeffs"].shape[0])), device=cuda_device)
        got = tab.evaluate(u, order, pts)
        full = tab.tabulate(order, pts)
        for alpha, table in full.items():
            want = torch.einsum("fd,d...->f...", u, table)
            bound = torch.einsum("fd,d...->f...", u.abs(), table.abs()).max().item()
            assert got[alpha].shape == want.shape
            assert (got[alpha] - want).abs().max().item() <= tolerance(desc, alpha) * max(bound, 1e-300), (workload, alpha)


@pytest.mark.parametrize("name", ["gll_q10_hex_o1", "q2_quad_edge2_o1", "p2xp1_prism_o1"])
def test_factored_tensor_tables(name, cuda_device):
    """tabulate_factors(): the per-factor tables multiply back to the reference's full table."""
    from fiat_b200.api import Tabulator
    case = load_case(name)
    factors = Tabulator(case["desc"], cuda_device).tabulate_factors(case["order"], case["points"], case["entity"])
    for alpha, ref in case["ref"].items():
        prod = None
        for aoff, sd, tab in factors:
            t = tab[tuple(alpha[aoff:aoff + sd])].cpu().numpy()
            prod = t if prod is None else (prod[:, None, :] * t[None, :, :]).reshape(-1, t.shape[-1])
        assert prod.shape == ref.shape
        assert abs(prod - ref).max() <= 1e-12 * max(abs(ref).max(), 1e-300)


@pytest.mark.parametrize("name", ["p8_tet", "p3_tri"])
def test_nodality_at_lattice_nodes(name, cuda_device):
    """test_fiat.py:446-470 in tabulated form: at its own nodes a Lagrange basis is the identity, on
    the product-form kernel and on the general kernels (points on vertices, edges and faces)."""
    from conftest import load_desc
    from fiat_b200.api import Tabulator, FORCE_GENERAL
    desc = load_desc(name)
    tab = Tabulator(desc, cuda_device)
    nodes = numpy.asarray(desc["nodes"])
    sd = nodes.shape[1]
    eye = torch.eye(len(nodes), dtype=torch.float64, device=cuda_device)
    for flags, tol in ((0, 1e-13), (FORCE_GENERAL, 1e-10)):
        vals = tab.tabulate(0, nodes, flags=flags)[(0,) * sd]
        assert (vals - eye).abs().max().item() <= tol


@pytest.mark.parametrize("name", ["hct_o2", "hct_o0", "ps12_o2", "ps12_o0", "ps6_o2", "p2_alfeld_tet_o2", "gn_tet_o2",
                                  "walkington_tet_o2", "hct4_tri_o2"])
def test_points_in_no_subcell_give_zero_columns(name, cuda_device):
    """NaN / Inf coordinates make every l1 distance NaN, so the point is binned to no subcell and the reference
    leaves its column of the zero-initialised tables untouched (FIAT/expansions.py:479-489).  Every kernel that
    locates subcells must write zeros there (never stale memory), at the fixture's order and at order 0."""
    from fiat_b200.api import Tabulator, FORCE_GENERAL, FORCE_THREAD_PER_POINT, NO_VALUE_TABLE, NO_MACRO_MERGED
    case = load_case(name)
    desc = case["desc"]
    sd = int(desc["sd"])
    pts0 = numpy.asarray(case["points"], dtype=float)
    pts = pts0[numpy.arange(40) % len(pts0)].copy()
    bad = [1, 5, 17, 18, 33]
    pts[1, 0] = numpy.nan
    pts[5, sd - 1] = numpy.inf
    pts[17, 0] = -numpy.inf
    pts[18, :] = numpy.nan
    pts[33, 1] = numpy.nan
    tab = Tabulator(desc, cuda_device)
    with numpy.errstate(all="ignore"):
        want = fiat_oracle.tabulate(desc, case["order"], pts)
    assert all(not v[..., bad].any() for v in want.values())
    for flags in (0, FORCE_GENERAL, FORCE_GENERAL | NO_VALUE_TABLE, FORCE_THREAD_PER_POINT, NO_MACRO_MERGED):
        out = torch.full((len(want), int(numpy.prod(next(iter(want.values())).shape[:-1])), len(pts)), 7.0,
                         dtype=torch.float64, device=cuda_device)
        tab.tabulate_into(out, case["order"], torch.as_tensor(pts, device=cuda_device), flags=flags)
        got = {a: out[j].reshape(want[a].shape) for j, a in enumerate(want)}
        _compare(desc, got, want)
        assert int(tab.locate_subcells(pts, unique=False)[bad].abs().sum().item()) == 0


def test_host_buffer_out_is_validated(cuda_device):
    """tabulate_host(out=...) hands the raw pointer to the library: dtype, shape, contiguity and writeability
    are checked first (a float32 / transposed / short buffer would be overrun)."""
    from fiat_b200.api import Tabulator
    case = load_case("p3_tri_o1")
    tab = Tabulator(case["desc"], cuda_device)
    pts = numpy.asarray(case["points"], dtype=float)[:32]
    good = numpy.empty((3, 10, 32))
    _compare(case["desc"], tab.tabulate_host(1, pts, out=good), {a: v[:, :32] for a, v in case["ref"].items()})
    readonly = numpy.empty((3, 10, 32))
    readonly.flags.writeable = False
    for bad in (numpy.empty((3, 10, 32), dtype=numpy.float32), numpy.empty((3, 10, 31)), numpy.empty((32, 10, 3)).T,
                numpy.empty((3, 10, 64))[:, :, ::2], readonly, [[0.0]]):
        with pytest.raises(ValueError):
            tab.tabulate_host(1, pts, out=bad)


@pytest.mark.parametrize("name", ["walkington_tet_o2", "gn_tet_o2", "hct5_tri_o2", "hct6_tri_o2", "alfeld_sorokina_tet_adv_o2"])
def test_split_cell_tile_kernel_many_tiles(name, cuda_device):
    """The split-cell tile kernels (cells_reg.cuh: expansion values in registers, coefficient steps streamed through
    shared memory) over many tiles and a ragged last tile, against the thread-per-point jet kernel on the same points;
    the set includes points on interior facets / vertices of the split (several subcells) and NaN points (none)."""
    from fiat_b200 import api
    case = load_case(name)
    desc, order = case["desc"], case["order"]
    sd = int(desc["sd"])
    tab = api.Tabulator(desc, cuda_device)
    assert "mma_cells" in tab.kernel_names(order, None)
    assert tab._self_check_flags(desc, order) == 0
    rng = numpy.random.default_rng(7)
    lam = rng.dirichlet(numpy.ones(sd + 1), size=5003)
    verts = numpy.asarray(desc["vertices"], dtype=float)[:sd + 1]
    pts = lam @ verts
    pts[100] = verts.mean(axis=0)                      # the split point of Alfeld-type complexes: every subcell
    pts[101] = 0.5 * (verts[0] + verts.mean(axis=0))   # on an interior edge
    pts[4000] = numpy.nan
    pts[5002] = verts[1]
    got = tab.tabulate(order, pts)
    want = tab.tabulate(order, pts, flags=api.FORCE_THREAD_PER_POINT)
    for alpha in got:
        g, w = got[alpha].cpu().numpy(), want[alpha].cpu().numpy()
        assert not numpy.isnan(g).any() and not g[..., 4000].any()
        assert abs(g - w).max() <= 2e-13 * max(abs(w).max(), 1e-300), alpha


@pytest.mark.parametrize("name", ["hct_o2", "n2curl4_tet_o1", "gn_tet_o2", "p4_spectral_tri_o2", "p8_spectral_tet_o2"])
def test_derived_paths_are_self_checked_on_the_device(name, cuda_device, monkeypatch):
    """Paths that replace the derivative jets by host-folded derivative matrices (value table, stacked derived element,
    split-cell tile kernel) are compared with the jet kernel on 96 points at plan time (api._self_check_flags).  With
    the real tolerance they are accepted; with an impossible tolerance they are switched off, the jet kernels run, and
    the tables still match the reference."""
    from fiat_b200 import api
    case = load_case(name)
    tab = api.Tabulator(case["desc"], cuda_device)
    accepted = tab.kernel_names(case["order"], case["entity"])
    assert tab._self_check_flags(case["desc"], case["order"]) == 0
    _compare(case["desc"], tab.tabulate(case["order"], case["points"], case["entity"]), case["ref"])
    monkeypatch.setattr(api, "SELF_CHECK_TOL", -1.0)
    strict = api.Tabulator(case["desc"], cuda_device)
    assert strict._self_check_flags(case["desc"], case["order"]) == api.NO_VALUE_TABLE | api.NO_ALPHA_SPLIT | api.NO_MACRO_MERGED
    rejected = strict.kernel_names(case["order"], case["entity"])
    assert all(k in ("cellwise", "mma", "small") for k in rejected), (accepted, rejected)
    # the launch is the element's own plan at the requested order (jets), not a derived order-0 plan
    launches = strict._resolve(case["order"], case["entity"])[0]
    assert len(launches) == 1 and launches[0][0] is strict._simplex_plan(case["desc"], case["order"])[0]
    _compare(case["desc"], strict.tabulate(case["order"], case["points"], case["entity"]), case["ref"])
    _compare(case["desc"], strict.tabulate_host(case["order"], case["points"], case["entity"]), case["ref"])
