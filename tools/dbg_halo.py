import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from swmhd_b200 import abi
from swmhd_b200.context import Context
from cases import make_case
from oracle import pyoracle as O
for kind in ("BJ",):
    g, cfg, U = make_case(kind, 96, Ny=80, perturb=17)
    Uo = [u.copy() for u in U]; O.fill_halos(cfg, Uo)
    for rep in range(4):
        c = Context(cfg); c.set_state(U); c.fill_halos(); a = c.get_state(); c.close()
        for k in range(4):
            rr, cc = np.nonzero(a[k] != Uo[k])
            print(rep, k, sorted(set(zip(rr.tolist(), cc.tolist())))[:12], len(rr))
