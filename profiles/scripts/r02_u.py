"""Last GPU seconds of round 2: the bound drop-in (oracle/dropin_plugin.py, device mode) on trace elements --
the interval trace that dead-locked before get_tabulator described elements outside its cache lock."""
import os, sys, time
t0 = time.time()
sys.path[:0] = [os.path.join(os.getcwd(), "oracle", "_ref"), os.getcwd()]
os.environ["FIATB200_DROPIN"] = "device"
import numpy
from oracle import dropin_plugin as dp
dp.MODE = "device"
dp.pytest_configure(None)
import FIAT
from FIAT.hdiv_trace import TraceError
from FIAT.reference_element import ufc_simplex
for dim, deg in ((1, 0), (2, 1), (3, 1)):
    el = FIAT.HDivTrace(ufc_simplex(dim), deg)
    verts = numpy.array(ufc_simplex(dim).get_vertices(), dtype=float)
    pts = verts[:dim].mean(axis=0, keepdims=True)      # barycentre of the facet opposite the last vertex: on ONE facet
    tab = el.tabulate(1, pts)
    zero = (0,) * dim
    assert isinstance(tab[zero], numpy.ndarray) and all(isinstance(v, TraceError) for k, v in tab.items() if k != zero)
    want = type(el).tabulate._fiat_b200_original(el, 0, pts)[zero]      # (the reference's interval trace stops at order 0)
    assert numpy.allclose(tab[zero], want, atol=1e-13), (dim, deg)
    print("HDivTrace", dim, deg, "ok", tab[zero].shape, round(time.time() - t0, 1), "s")
print(dp.stats)
