"""Per-path error against the golden reference output for high-degree elements (diagnostic)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy, torch
from conftest import load_case
from fiat_b200 import api
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["p12_tri_o2", "p10_tri_o2", "p8_tet_o2", "p12_spectral_tri_o2"]:
    case = load_case(name)
    tab = api.Tabulator(case["desc"], dev)
    print(name, "default path:", tab.kernel_path(case["order"]), tab.kernel_names(case["order"]),
          "self-check flags", tab._self_check_flags(case["desc"], case["order"]))
    for label, flags in (("default", 0), ("general", 4), ("thread-per-point jets", 1), ("DMMA jets", 2),
                         ("general, no derived", 4 | 16), ("general, per-alpha launches", 4 | 32)):
        try:
            got = tab.tabulate(case["order"], case["points"], case["entity"], flags=flags)
        except Exception as exc:
            print("   ", label, "->", type(exc).__name__, exc)
            continue
        errs = []
        for a, ref in case["ref"].items():
            e = abs(got[a].cpu().numpy() - ref)
            errs.append("%.1e" % (e.max() / abs(ref).max()))
        print("   %-28s" % label, tab.kernel_names(case["order"], None, flags), errs)
