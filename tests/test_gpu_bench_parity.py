"""Parity at BASELINE scale: for every BASELINE configuration the first 2^16 points of the bench's
own shard 0 (same generator and seed as bench.py) go through the CUDA path and through the CPU oracle; tables must
agree to the north-star tolerance and the subcell bitmasks bit for bit.  Also the live reference (oracle/_ref,
materialised by __graft_entry__.build()) called next to the drop-in on the same element object."""
import numpy
import pytest
import torch

from conftest import tolerance
from oracle import fiat_oracle

pytestmark = pytest.mark.gpu

NPTS = 1 << 16


def _bench_points(workload, device, n=NPTS):
    import bench
    dname, order, kind, _ = bench.WORKLOADS[workload]
    pts = bench.device_points(kind, bench.DEFAULT_BATCH.get(workload, 1 << 20), 1234, device)[:n].contiguous()
    return bench.load_desc(dname), order, pts


@pytest.mark.parametrize("workload", ["p3_tri_o1", "p8_tet_o2", "n2curl4_tet_o1", "hct_o2", "ps6_o2", "ps12_o2",
                                      "gll_q10_hex_o1"])
@pytest.mark.parametrize("general", [False, True])
def test_first_2_16_bench_points_match_the_oracle(workload, general, cuda_device):
    from fiat_b200.api import Tabulator, FORCE_GENERAL
    desc, order, pts = _bench_points(workload, cuda_device)
    tab = Tabulator(desc, cuda_device)
    if general and tab.kernel_path(order) != "lattice":
        pytest.skip("the default path is the general one")
    got = tab.tabulate(order, pts, flags=FORCE_GENERAL if general else 0)
    host = pts.cpu().numpy()
    worst = 0.0
    chunk = 1 << 13                     # bounds the oracle's temporaries (hexahedron: 5324 values per point)
    scale = {a: 0.0 for a in got}
    err = {a: 0.0 for a in got}
    for s in range(0, NPTS, chunk):
        want = fiat_oracle.tabulate(desc, order, host[s:s + chunk])
        assert list(want) == list(got)
        for a, w in want.items():
            g = got[a][..., s:s + chunk].cpu().numpy()
            assert g.shape == w.shape
            scale[a] = max(scale[a], float(abs(w).max()))
            err[a] = max(err[a], float(abs(g - w).max()))
    for a in got:
        assert err[a] <= tolerance(desc, a) * max(scale[a], 1e-300), (workload, a, err[a] / scale[a])
        worst = max(worst, err[a] / max(scale[a], 1e-300))
    print(workload, "general" if general else "default", "worst relative error", worst)


@pytest.mark.parametrize("workload", ["hct_o2", "ps6_o2", "ps12_o2"])
def test_first_2_16_bench_points_subcell_masks_bit_exact(workload, cuda_device):
    from fiat_b200.api import Tabulator
    desc, order, pts = _bench_points(workload, cuda_device)
    tab = Tabulator(desc, cuda_device)
    host = pts.cpu().numpy()
    for unique in (False, True):
        near = fiat_oracle.locate_cells(desc, host, unique=unique)
        want = sum(near[c].astype(numpy.int64) << c for c in range(near.shape[0]))
        mask = tab.locate_subcells(pts, unique).cpu().numpy().astype(numpy.int64)
        assert numpy.array_equal(mask, want)
        assert (mask != 0).all()


# ---- the live reference next to the drop-in ------------------------------------------------------------------

def _reference():
    from oracle.make_ref import import_reference
    FIAT = import_reference()
    if FIAT is None:
        pytest.skip("oracle/_ref has not been materialised (run __graft_entry__.build() where /root/reference exists)")
    return FIAT


def _live_elements(FIAT):
    from FIAT.reference_element import ufc_simplex, UFCInterval
    from FIAT.tensor_product import FlattenedDimensions
    T1, T2, T3 = UFCInterval(), ufc_simplex(2), ufc_simplex(3)
    G = FIAT.GaussLobattoLegendre(T1, 10)
    hexa = FlattenedDimensions(FIAT.TensorProductElement(FlattenedDimensions(FIAT.TensorProductElement(G, G)), G))
    return {
        "p3_tri_o1": (lambda: FIAT.Lagrange(T2, 3), 1, "simplex2", 4096),
        "p8_tet_o2": (lambda: FIAT.Lagrange(T3, 8), 2, "simplex3", 2048),
        "p8_spectral_tet_o2": (lambda: FIAT.Lagrange(T3, 8, variant="spectral"), 2, "simplex3", 1024),
        "n2curl4_tet_o1": (lambda: FIAT.NedelecSecondKind(T3, 4), 1, "simplex3", 2048),
        "hct_o2": (lambda: FIAT.HsiehCloughTocher(T2), 2, "simplex2", 8192),
        "ps6_o2": (lambda: FIAT.QuadraticPowellSabin6(T2), 2, "simplex2", 8192),
        "ps12_o2": (lambda: FIAT.QuadraticPowellSabin12(T2), 2, "simplex2", 8192),
        "gll_q10_hex_o1": (lambda: hexa, 1, "cube3", 64),       # the reference loops over points in Python here
    }


@pytest.mark.parametrize("name", ["p3_tri_o1", "p8_tet_o2", "p8_spectral_tet_o2", "n2curl4_tet_o1", "hct_o2", "ps6_o2",
                                  "ps12_o2", "gll_q10_hex_o1"])
def test_drop_in_with_a_live_reference_element(name, cuda_device):
    """fiat_b200.tabulate(element, order, points) == element.tabulate(order, points) with `element` constructed by
    the reference itself in this process (FIAT/finite_element.py:181-197, FIAT/tensor_product.py:231-336): the
    drop-in's duck-typed reader (extract.describe_element) runs on a live object, on the GPU box."""
    import bench
    import fiat_b200
    FIAT = _reference()
    make, order, kind, npts = _live_elements(FIAT)[name]
    element = make()
    pts = bench.host_points(kind, npts, 4321)
    want = element.tabulate(order, pts)
    got = fiat_b200.tabulate(element, order, pts)
    assert [tuple(k) for k in got] == [tuple(k) for k in want]
    desc = fiat_b200.api.get_tabulator(element).desc
    for a, w in want.items():
        g = got[a].cpu().numpy()
        assert g.shape == w.shape and g.dtype == numpy.float64
        assert abs(g - w).max() <= tolerance(desc, a) * abs(w).max(), (name, a)
    # second call: cached tabulator, same tables; host-buffer entry point returns numpy like the reference
    again = fiat_b200.tabulate_host(element, order, pts[:128])
    for a, w in want.items():
        assert isinstance(again[a], numpy.ndarray)
        assert abs(again[a] - w[..., :128]).max() <= tolerance(desc, a) * abs(w).max()


def test_live_reference_masks_on_fresh_adversarial_points(cuda_device):
    """compute_cell_point_map of the live reference (FIAT/expansions.py:771-811) against the device binning on a
    FRESH adversarial set (other seed than the committed fixtures), 2-D and 3-D split complexes."""
    import sys, os
    FIAT = _reference()
    from FIAT import expansions
    from FIAT.reference_element import ufc_simplex
    import fiat_b200
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden", "gen"))
    # the generator module imports /root/reference at import time; only its point generator is needed here
    gen = {}
    src = open(os.path.join(os.path.dirname(__file__), "golden", "gen", "make_golden.py")).read()
    start, stop = src.index("def _ulp_steps"), src.index("class CiarletElement")
    exec("import numpy\n" + src[src.index("def simplex_points"):src.index("def adversarial_triangle_points")]
         + src[start:stop], gen)
    rng = numpy.random.default_rng(777)
    T2, T3 = ufc_simplex(2), ufc_simplex(3)
    for element in (FIAT.HsiehCloughTocher(T2), FIAT.QuadraticPowellSabin12(T2),
                    FIAT.Lagrange(T3, 2, variant="worsey-farin"), FIAT.Walkington(T3)):
        complex_ = element.get_nodal_basis().get_expansion_set().ref_el
        pts = gen["adversarial_points"](complex_, rng, per_facet=6, n_random=500)
        tab = fiat_b200.api.get_tabulator(element)
        ncells = len(complex_.get_topology()[complex_.get_spatial_dimension()])
        for unique in (False, True):
            cpm = expansions.compute_cell_point_map(complex_, pts, unique=unique)
            want = numpy.zeros(len(pts), dtype=numpy.int64)
            for c, ipts in cpm.items():
                want[ipts] |= 1 << c
            mask = tab.locate_subcells(pts, unique).cpu().numpy().astype(numpy.int64)
            assert numpy.array_equal(mask, want), (type(element).__name__, unique, ncells)


@pytest.mark.parametrize("workload", ["p8_tet_o2", "hct_o2", "gll_q10_hex_o1"])
def test_one_point_array_sharded_over_the_devices(workload, cuda_device):
    """SURVEY 8e on hardware: ONE point array split contiguously over all visible GPUs of this process
    (fiat_b200.tabulate_sharded), every shard tabulated on its own device with its own plan; the gathered blocks equal
    the single-device tables bit for bit and the shard boundaries follow shard_range.  Runs with the devices the box
    has (one device: a single shard)."""
    import bench
    import fiat_b200
    dname, order, kind, _ = bench.WORKLOADS[workload]
    desc = bench.load_desc(dname)
    ndev = torch.cuda.device_count()
    npts = 4099 if workload != "gll_q10_hex_o1" else 1031          # not divisible by the device count
    pts = bench.host_points(kind, npts, 31)
    shards = fiat_b200.tabulate_sharded(desc, order, pts)
    assert len(shards) == ndev
    assert [(s, e) for _, s, e, _ in shards] == [fiat_b200.shard_range(npts, g, ndev) for g in range(ndev)]
    for g, (dev, s, e, tab) in enumerate(shards):
        assert dev.index == g and all(v.device == dev and v.shape[-1] == e - s for v in tab.values())
    whole = fiat_b200.api.get_tabulator(desc, cuda_device).tabulate(order, pts)
    gathered = fiat_b200.gather_shards(shards, cuda_device)
    assert list(gathered) == list(whole)
    for a in whole:
        assert torch.equal(gathered[a], whole[a])
