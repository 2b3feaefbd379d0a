"""Throughput of the device path for the element of any golden case (random points like the fixture's)."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy, torch
from conftest import load_case
from fiat_b200.api import Tabulator
name = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
case = load_case(name)
tab = Tabulator(case["desc"], torch.device("cuda:0"))
pts0 = numpy.asarray(case["points"]); rng = numpy.random.default_rng(0)
desc = case["desc"]
if desc["kind"] == "simplex" and case["entity"] is None and "vertices" in desc:
    # uniform points of the cell (the fixture points of split-cell elements sit on interior facets)
    verts = numpy.asarray(desc["vertices"], dtype=float)[:int(desc["sd"]) + 1]      # parent cell's vertices come first
    lam = numpy.diff(numpy.concatenate([numpy.zeros((n, 1)), numpy.sort(rng.random((n, len(verts) - 1)), axis=1),
                                        numpy.ones((n, 1))], axis=1), axis=1)
    pts = torch.as_tensor(lam @ verts, device="cuda:0")
else:
    pts = torch.as_tensor(pts0[rng.integers(0, len(pts0), size=n)] * (1 - 1e-3 * rng.random((n, 1))), device="cuda:0")
order = case["order"]
flags = int(os.environ.get("FIATB200_FLAGS", "0"))
out = tab.tabulate(order, pts, case["entity"], flags=flags)
vals = sum(v.numel() for v in out.values())
na = len(out); nrows = vals // (na * n)
buf = torch.empty((na, nrows, n), dtype=torch.float64, device="cuda:0")
for _ in range(3): tab.tabulate_into(buf, order, pts, case["entity"], flags=flags)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): tab.tabulate_into(buf, order, pts, case["entity"], flags=flags)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{name}: {vals / n} values/pt, {ms:.3f} ms per {n} pts, {vals / ms / 1e6:.1f} Gval/s, {vals * 8 / ms / 1e6:.0f} GB/s ({vals * 8 / ms / 1e6 / 6553:.2f} of HBM peak), path {tab.kernel_path(order, flags)} flags {flags} kernels {'+'.join(tab.kernel_names(order, case['entity'], flags))}")
