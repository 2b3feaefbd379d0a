"""Every golden case through the QUICK plan (api.QUICK_NPTS: what a small first call gets by default -- thread-per-point
kernels on a description marked dense_only) against the reference's tables.  (Named to run last: the other GPU tests
are about the streaming kernels and switch the quick plan off, tests/conftest.py.)"""
import numpy
import pytest
import torch

from conftest import golden_case_names, load_case, tolerance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_case_names())
def test_quick_plan_matches_reference_golden(name, cuda_device, monkeypatch):
    from fiat_b200 import api
    case = load_case(name)
    if len(case["points"]) > 4096:
        pytest.skip("large adversarial set: takes the streaming plan by design")
    monkeypatch.setattr(api, "QUICK_NPTS", 4096)
    tab = api.Tabulator(case["desc"], cuda_device)
    got = tab.tabulate_host(case["order"], case["points"], case["entity"])
    if case["desc"]["kind"] not in ("trace", "quadrature"):
        assert tab._quick is not None and not any(k[0] == "resolved" for k in tab._plans)
    assert [tuple(k) for k, v in got.items() if isinstance(v, Exception)] == list(case["error_keys"])
    got = {k: v for k, v in got.items() if not isinstance(v, Exception)}
    assert [tuple(k) for k in got] == [tuple(k) for k in case["ref"]]
    for alpha, expect in case["ref"].items():
        g = numpy.asarray(got[alpha])
        assert g.shape == expect.shape
        if expect.size == 0:
            continue
        assert numpy.array_equal(numpy.isnan(g), numpy.isnan(expect))
        g, e = numpy.nan_to_num(g), numpy.nan_to_num(expect)
        assert abs(g - e).max() <= tolerance(case["desc"], alpha) * max(abs(e).max(), 1e-300), alpha
