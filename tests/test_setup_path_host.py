"""Host logic of the setup-path callers with the device tabulator replaced by the CPU oracle:
the bookkeeping of `to_riesz` (FIAT/dual_set.py:86-206) against the live reference's `DualSet.to_riesz`."""
import numpy
import pytest
import torch


class _OracleTabulator:
    """Stand-in for setup_path.ExpansionTabulator: same interface, tables from oracle.fiat_oracle."""

    def __init__(self, expansion_set, n, device=None):
        from fiat_b200.extract import describe_expansion_set
        from oracle import fiat_oracle
        self.desc = describe_expansion_set(expansion_set, n)
        self.sd = int(self.desc["sd"])
        self.device = torch.device("cpu")
        desc = self.desc

        class _Tab:
            def tabulate(self, order, pts):
                tabs = fiat_oracle.tabulate(desc, order, numpy.asarray(pts, dtype=float))
                return {a: torch.as_tensor(v) for a, v in tabs.items()}
        self.tab = _Tab()

    def tabulate(self, pts):
        return self.tab.tabulate(0, pts)[(0,) * self.sd]


def riesz_elements(FIAT):
    from FIAT.reference_element import ufc_simplex
    T2, T3 = ufc_simplex(2), ufc_simplex(3)
    return [FIAT.Lagrange(T3, 4), FIAT.CubicHermite(T2), FIAT.Argyris(T2, 5), FIAT.RaviartThomas(T3, 3),
            FIAT.Nedelec(T3, 2), FIAT.Regge(T2, 1), FIAT.BrezziDouglasMarini(T2, 2), FIAT.HsiehCloughTocher(T2),
            FIAT.Morley(T2), FIAT.GuzmanNeilanFirstKindH1(T3, 1), FIAT.MardalTaiWinther(T2), FIAT.Bell(T2)]


def test_to_riesz_bookkeeping_against_the_reference(monkeypatch):
    from oracle.make_ref import import_reference
    FIAT = import_reference()
    if FIAT is None:
        pytest.skip("oracle/_ref has not been materialised")
    from fiat_b200 import setup_path
    monkeypatch.setattr(setup_path, "ExpansionTabulator", _OracleTabulator)
    for element in riesz_elements(FIAT):
        poly_set, dual = element.get_nodal_basis(), element.dual
        want = dual.to_riesz(poly_set)
        got = setup_path.to_riesz(dual, poly_set).numpy()
        assert got.shape == want.shape
        assert abs(got - want).max() <= 1e-13 * abs(want).max(), type(element).__name__
