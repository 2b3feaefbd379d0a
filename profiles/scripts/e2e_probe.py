import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy, torch
from fiat_b200 import description
from fiat_b200.api import Tabulator
desc = description.load("tests/golden/desc_p8_tet.npz")
tab = Tabulator(desc, torch.device("cuda:0"))
rng = numpy.random.default_rng(0)
ne = 65536
u = numpy.sort(rng.random((ne,3)),axis=1); pts = numpy.diff(numpy.concatenate([numpy.zeros((ne,1)),u],axis=1),axis=1)
hp = torch.empty((ne,3), dtype=torch.float64, pin_memory=True); hp.copy_(torch.from_numpy(pts))
ho = torch.empty((10,165,ne), dtype=torch.float64, pin_memory=True)
for flags in (0, 4, 0, 4):
    for chunk in (1<<14, 1<<16):
        ts=[]
        for i in range(4):
            t0=time.perf_counter(); tab.tabulate_host(2, hp.numpy(), out=ho.numpy(), chunk_pts=chunk, flags=flags); ts.append(time.perf_counter()-t0)
        print("flags",flags,"chunk",chunk,["%.1f ms"%(t*1e3) for t in ts], "GB/s d2h", ne*13200/min(ts)/1e9)
