"""pytest plugin (TEST INFRASTRUCTURE): run the reference's OWN unit tests with the tabulation path swapped out.

    PYTHONPATH=oracle/_ref:. FIATB200_DROPIN=device python -m pytest -p oracle.dropin_plugin oracle/_ref/ref_tests/...

Before the reference's tests are collected, every `tabulate` method on the hot path -- `CiarletElement.tabulate`
(FIAT/finite_element.py:181-197), `TensorProductElement.tabulate` / `FlattenedDimensions.tabulate`
(FIAT/tensor_product.py:231-336,396-407), the wrappers' (enriched.py:88, mixed.py:61, discontinuous.py:56,
hdiv_trace.py:133, quadrature_element.py:43) -- is replaced by a call into

    FIATB200_DROPIN=device   the CUDA drop-in (`fiat_b200.tabulate_host`, numpy in / numpy out through the C ABI)
    FIATB200_DROPIN=oracle   the CPU oracle (`oracle.fiat_oracle.tabulate` on `describe_element(element)`)
    FIATB200_DROPIN=emulate  the device mode's HOST logic on a machine without a GPU: `fiat_b200.api.Tabulator` with
                             the kernel launches of polynomial elements answered by the oracle, so that what the
                             device mode adds in Python -- trace / quadrature elements, the single-point form, the
                             exception types handed back -- is judged by the reference's tests on the CPU too

so that the reference's known-answer tests (exact Dubiner values, nodality, partition of unity, macro-element
continuity, tensor-product dof ordering, ...) judge the replacement directly.  What neither path takes on --
sympy-based elements, symbolic (object-dtype) points -- falls through to the reference's own method and is counted
(`fiat_b200_dropin_stats`); element CONSTRUCTION is untouched except where the reference itself calls
`element.tabulate` while building another element, which then also goes through the replacement.
"""
import os

import numpy

MODE = os.environ.get("FIATB200_DROPIN", "")
stats = {"replaced": 0, "fallback": 0, "fallback_reasons": {}}


def _note_fallback(exc):
    stats["fallback"] += 1
    key = f"{type(exc).__name__}: {str(exc)[:80]}"
    stats["fallback_reasons"][key] = stats["fallback_reasons"].get(key, 0) + 1


def _wrap(original):
    def tabulate(self, order, points, entity=None):
        if "old_tabulate" in self.__dict__:
            # an Hdiv / Hcurl wrapper instance calling the tabulation of the tensor-product element it wraps
            # (hdivcurl.py:40,162): its description is the wrapper's, so this inner call stays with the reference
            stats["fallback"] += 1
            return original(self, order, points, entity)
        try:
            pts = numpy.asarray(points)
            if pts.dtype == object:
                raise NotImplementedError("symbolic points")
            if MODE in ("device", "emulate"):
                import fiat_b200
                from FIAT.hdiv_trace import TraceError as ReferenceTraceError
                out = fiat_b200.tabulate_host(self, order, points, entity, device="cpu" if MODE == "emulate" else None)
                # (the binding layer hands the reference's own exception type to the reference's callers)
                out = {k: (ReferenceTraceError(getattr(v, "msg", str(v))) if isinstance(v, Exception) else numpy.array(v))
                       for k, v in out.items()}
            else:
                from fiat_b200.extract import describe_element
                from oracle import fiat_oracle
                desc = self.__dict__.get("_fiat_b200_desc")
                if desc is None:
                    desc = describe_element(self)
                    try:
                        self.__dict__["_fiat_b200_desc"] = desc
                    except Exception:
                        pass
                out = fiat_oracle.tabulate(desc, order, numpy.asarray(points, dtype=float), entity)
            stats["replaced"] += 1
            return out
        except (NotImplementedError, KeyError) as exc:       # UnsupportedElement is a NotImplementedError
            _note_fallback(exc)
            return original(self, order, points, entity)
    tabulate._fiat_b200_original = original
    return tabulate


def _install_cpu_tabulator():
    """emulate mode: a Tabulator without a device whose polynomial tabulations come from the oracle."""
    import threading
    import torch
    from fiat_b200 import api
    from oracle import fiat_oracle

    Base = api.Tabulator

    class CpuTabulator(Base):
        def __init__(self, desc, device=None):
            self.lib, self.desc, self.device, self.kind = None, desc, torch.device("cpu"), desc["kind"]
            self._plans, self._lock, self._quick = {}, threading.Lock(), None

        def tabulate(self, order, points, entity=None, flags=0):
            if self.kind in ("trace", "quadrature"):
                return Base.tabulate(self, order, points, entity, flags)
            pts = numpy.asarray(points.cpu() if isinstance(points, torch.Tensor) else points, dtype=float)
            return {a: torch.as_tensor(v) for a, v in fiat_oracle.tabulate(self.desc, order, pts, entity).items()}

        def tabulate_host(self, order, points, entity=None, chunk_pts=1 << 16, flags=0, out=None):
            if self.kind in ("trace", "quadrature"):
                return Base.tabulate_host(self, order, points, entity, chunk_pts, flags, out)
            return fiat_oracle.tabulate(self.desc, order, numpy.asarray(points, dtype=float), entity)

    api.Tabulator = CpuTabulator


def pytest_configure(config):
    if MODE not in ("device", "oracle", "emulate"):
        return
    if MODE == "emulate":
        _install_cpu_tabulator()
    import FIAT  # noqa: F401  (the live reference, from oracle/_ref on sys.path)
    from FIAT import finite_element, tensor_product, enriched, mixed, discontinuous, hdiv_trace, quadrature_element
    targets = [(finite_element.CiarletElement, "tabulate"), (tensor_product.TensorProductElement, "tabulate"),
               (tensor_product.FlattenedDimensions, "tabulate"), (enriched.EnrichedElement, "tabulate"),
               (mixed.MixedElement, "tabulate"), (discontinuous.DiscontinuousElement, "tabulate")]
    if MODE in ("device", "emulate"):    # the oracle has no trace / quadrature elements (pinned by golden files instead)
        targets += [(hdiv_trace.HDivTrace, "tabulate"), (quadrature_element.QuadratureElement, "tabulate")]
    for cls, name in targets:
        method = cls.__dict__.get(name)
        if method is not None and not hasattr(method, "_fiat_b200_original"):
            setattr(cls, name, _wrap(method))


def pytest_sessionfinish(session):
    """Under pytest-xdist every worker has its own counters: each appends them to $FIATB200_DROPIN_STATS."""
    path = os.environ.get("FIATB200_DROPIN_STATS")
    if path and MODE in ("device", "oracle", "emulate") and (stats["replaced"] or stats["fallback"]):
        import json
        with open(path, "a") as f:
            f.write(json.dumps(stats) + "\n")


def pytest_terminal_summary(terminalreporter):
    if MODE in ("device", "oracle", "emulate"):
        path = os.environ.get("FIATB200_DROPIN_STATS")
        if path and os.path.exists(path):
            import json
            total = {"replaced": 0, "fallback": 0, "fallback_reasons": {}}
            with open(path) as f:
                for line in f:
                    rec = json.loads(line)
                    total["replaced"] += rec["replaced"]
                    total["fallback"] += rec["fallback"]
                    for k, v in rec["fallback_reasons"].items():
                        total["fallback_reasons"][k] = total["fallback_reasons"].get(k, 0) + v
            stats.update(total)
        terminalreporter.write_line(f"fiat_b200_dropin_stats mode={MODE} replaced={stats['replaced']} "
                                    f"fallback={stats['fallback']} reasons={stats['fallback_reasons']}")
