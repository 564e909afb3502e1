import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from swmhd_b200 import abi
from swmhd_b200.context import Context
from swmhd_b200.distributed import SlabModel
from cases import make_case
g, cfg, U = make_case("BJ", 96, Ny=80, arith=abi.ARITH_FAST, perturb=17)
for nst in (1,):
    c = Context(cfg); c.set_state(U); c.fill_halos()
    sm = SlabModel(cfg, rank=0, world=1, device=0); sm.set_state(U); sm.fill_halos()
    for stage in (1,2,3):
        c.substage(0.004, stage)
        sm.substage(0.004, stage); sm.synchronize()
        a = c.get_state(); b = sm.get_state()
        for k in range(4):
            d = np.abs(a[k]-b[k]).max(axis=1)
            rows = np.nonzero(d)[0]
            print("stage",stage,"field",k,"rows differing:",rows.tolist()[:20], d.max())
print("---- detail stage 1")
from oracle import pyoracle as O
c = Context(cfg); c.set_state(U); c.fill_halos()
sm = SlabModel(cfg, rank=0, world=1, device=0); sm.set_state(U); sm.fill_halos()
c.substage(0.004, 1); sm.substage(0.004, 1); sm.synchronize()
a = c.get_state(); b = sm.get_state()
Uo = [u.copy() for u in U]; O.fill_halos(cfg, Uo)
Gn=[np.zeros_like(x) for x in Uo]; Gm=[np.zeros_like(x) for x in Uo]
O.substage(cfg, Uo, Gn, Gm, 0.004, 1)
for r in (2,3,82,83):
    da = np.nonzero(a[0][r]-Uo[0][r])[0]; db = np.nonzero(b[0][r]-Uo[0][r])[0]
    print("row",r,"ctx-vs-oracle cols",da.tolist()[:12],len(da),"slab-vs-oracle cols",db.tolist()[:12],len(db), np.abs(b[0][r]-Uo[0][r]).max(), np.abs(a[0][r]-Uo[0][r]).max())
