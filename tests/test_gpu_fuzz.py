"""Randomised parity sweep on the GPU: every golden element description, several point counts (odd
sizes, tails, single points), orders 0..order, points inside and slightly outside the cell, every
entity the fixture names -- CUDA path (through the C ABI) against the CPU oracle."""
import os
import zlib

import numpy
import pytest
import torch

from conftest import golden_case_names, load_case, tolerance
from oracle import fiat_oracle

pytestmark = pytest.mark.gpu


def _points_like(case, n, rng):
    """Fresh points with the same dimension and rough location as the fixture's."""
    pts = numpy.asarray(case["points"], dtype=float)
    dim = pts.shape[1] if pts.ndim == 2 else 0
    if dim == 0:
        return numpy.zeros((n, 0))
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    base = pts[rng.integers(0, len(pts), size=n)]
    jitter = (rng.random((n, dim)) - 0.5) * 0.2 * numpy.maximum(hi - lo, 0.1)
    out = base + jitter
    # a few points clearly outside the cell
    k = max(1, n // 10)
    out[:k] += 0.3
    return out


@pytest.mark.parametrize("name", golden_case_names())
def test_random_points_against_oracle(name, cuda_device):
    from fiat_b200.api import Tabulator
    case = load_case(name)
    desc = case["desc"]
    if desc["kind"] in ("trace", "quadrature"):
        pytest.skip("not a polynomial tabulation (pinned by its golden file)")
    tab = Tabulator(desc, cuda_device)
    # stable across processes (hash() is salted); FIATB200_FUZZ_SEED varies the sweep
    rng = numpy.random.default_rng(zlib.crc32(name.encode()) + int(os.environ.get("FIATB200_FUZZ_SEED", "0")))
    for n in (1, 7, 33, 257):
        pts = _points_like(case, n, rng)
        for order in sorted({0, case["order"]}):
            got = tab.tabulate(order, pts, case["entity"])
            want = fiat_oracle.tabulate(desc, order, pts, case["entity"])
            assert list(got.keys()) == list(want.keys())
            for alpha, ref in want.items():
                g = got[alpha].cpu().numpy()
                assert g.shape == ref.shape
                scale = max(abs(ref).max(), 1e-300)
                err = abs(g - ref).max()
                assert err <= tolerance(desc, alpha) * scale, (n, order, alpha, err / scale)
