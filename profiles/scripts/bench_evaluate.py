"""Throughput of the fused point evaluation on the GLL Q10 hexahedron (order 1): points/s and the tabulated
values/s it stands for (values that tabulate + contraction would have had to write and read back)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy, torch
from fiat_b200 import description
from fiat_b200.api import Tabulator
desc = description.load("tests/golden/desc_gll_q10_hex.npz")
tab = Tabulator(desc, torch.device("cuda:0"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
nfunc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pts = torch.rand((n, 3), dtype=torch.float64, device="cuda:0")
u = numpy.random.default_rng(0).standard_normal((nfunc, 1331))
for _ in range(2): out = tab.evaluate(u, 1, pts)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = tab.evaluate(u, 1, pts)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"evaluate GLL Q10 hex order 1, {nfunc} function(s): {ms:.3f} ms per {n} pts = {n / ms / 1e6:.2f} Gpt/s "
      f"(the tabulation it replaces: {n * 5324 / ms / 1e6:.0f} Gval/s-equivalent; tabulate alone runs at 0.14 Gpt/s)")
