import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from swmhd_b200 import abi
from swmhd_b200.context import Context
from cases import make_case
from oracle import pyoracle as O
for kind in ("BJ","J","BD"):
  for arith in (abi.ARITH_FAST, abi.ARITH_STRICT):
    g, cfg, U = make_case(kind, 96, Ny=80, arith=arith, perturb=17)
    Uo = [u.copy() for u in U]; O.fill_halos(cfg, Uo)
    inp = [u.copy() for u in Uo]
    Gn=[np.zeros_like(x) for x in Uo]; Gm=[np.zeros_like(x) for x in Uo]
    O.substage(cfg, Uo, Gn, Gm, 0.004, 1)
    bad = {}
    for rep in range(6):
        c = Context(cfg); c.set_state(inp)     # halos already filled by the oracle: no fill_halos call
        c.substage(0.004, 1); a = c.get_state(); c.close()
        for k in range(4):
            rr, cc = np.nonzero(np.abs(a[k]-Uo[k]) > (0 if arith else 1e-13))
            if len(rr): bad.setdefault(k, set()).update(zip(rr.tolist(), cc.tolist()))
    print(kind, "strict" if arith else "fast", {k: sorted(v)[:10] for k, v in bad.items()})
