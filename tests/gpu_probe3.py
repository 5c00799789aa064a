import sys, time
import numpy as np
from common import *
from oracle.oracle import Oracle
PASSIVE = dict(HTurbOn=0, VTurbOn=0, Behavior=0, settlementon=0, mortality=0)
def pair(name, n=2000, nint=4, **kw):
    w = World(**SMALL)
    g = LtransLib(); o = Oracle()
    prm = make_params(w, n, **kw)
    x0,y0,z0 = setup(g, w, prm, n); setup(o, w, prm, n)
    for s in range(1, nint+1):
        g.step(1, s); o.step(1, s)
        fg, fo = g.fetch(), o.fetch()
        dz = np.abs(fg['z']-fo['z']); dx = np.abs(fg['x']-fo['x'])
        idx = np.argsort(-dz)[:4]
        print(name, 'step', s, 'max dz', dz.max(), 'n>1e-10', (dz>1e-10).sum(), 'top', idx.tolist(), dz[idx].tolist(), 'z', fo['z'][idx].tolist(), fg['z'][idx].tolist(), 'dx', dx[idx].tolist(), flush=True)
    g.destroy(); o.destroy()
pair('vturb2000', **dict(PASSIVE, VTurbOn=1))
pair('both2000', **dict(PASSIVE, VTurbOn=1, HTurbOn=1))
